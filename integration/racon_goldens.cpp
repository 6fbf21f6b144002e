/*
 * integration/racon_goldens.cpp — the reference's CUDA test cases (test/racon_test.cpp:297-507: same files, same
 * parameters, cudapoa batches = 1) run through the UNMODIFIED reference host (createPolisher -> CUDAPolisher) linked
 * against libracon_b200.so, printing one JSON line per case.  tests/test_integration.py asserts the reference's CPU
 * goldens (the "// CPU ..." values of those very tests, i.e. test/racon_test.cpp:86-295): the reference's own CUDA
 * build does not reproduce them (it carries separate goldens, 1385 / 1607 / ...), this backend does.
 *   racon_goldens <test data dir> [--banded] [--aligner-batches N] [--only NAME]
 */
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "bioparser/fasta_parser.hpp"
#include "edlib.h"
#include "polisher.hpp"
#include "sequence.hpp"

namespace {

uint32_t edit_distance(const std::string& query, const std::string& target) {  // test/racon_test.cpp:14-23
    EdlibAlignResult r = edlibAlign(query.c_str(), query.size(), target.c_str(), target.size(), edlibDefaultAlignConfig());
    uint32_t d = r.editDistance;
    edlibFreeAlignResult(r);
    return d;
}

struct Case {
    const char* name;
    const char* reads;
    const char* overlaps;
    const char* targets;
    racon::PolisherType type;
    uint32_t window;
    int8_t m, x, g;
    bool drop_unpolished;
    bool vs_reference;  // edit distance to sample_reference (contig polishing) or count/total length (fragment correction)
};

const Case kCases[] = {
    // test/racon_test.cpp:298-319, 321-342, 344-365, 367-388, 390-411, 413-434
    {"ConsensusWithQualitiesCUDA", "sample_reads.fastq.gz", "sample_overlaps.paf.gz", "sample_layout.fasta.gz",
     racon::PolisherType::kC, 500, 5, -4, -8, true, true},
    {"ConsensusWithoutQualitiesCUDA", "sample_reads.fasta.gz", "sample_overlaps.paf.gz", "sample_layout.fasta.gz",
     racon::PolisherType::kC, 500, 5, -4, -8, true, true},
    {"ConsensusWithQualitiesAndAlignmentsCUDA", "sample_reads.fastq.gz", "sample_overlaps.sam.gz", "sample_layout.fasta.gz",
     racon::PolisherType::kC, 500, 5, -4, -8, true, true},
    {"ConsensusWithoutQualitiesAndWithAlignmentsCUDA", "sample_reads.fasta.gz", "sample_overlaps.sam.gz",
     "sample_layout.fasta.gz", racon::PolisherType::kC, 500, 5, -4, -8, true, true},
    {"ConsensusWithQualitiesLargerWindowCUDA", "sample_reads.fastq.gz", "sample_overlaps.paf.gz", "sample_layout.fasta.gz",
     racon::PolisherType::kC, 1000, 5, -4, -8, true, true},
    {"ConsensusWithQualitiesEditDistanceCUDA", "sample_reads.fastq.gz", "sample_overlaps.paf.gz", "sample_layout.fasta.gz",
     racon::PolisherType::kC, 500, 1, -1, -1, true, true},
    // test/racon_test.cpp:436-452, 454-470, 472-488, 490-506
    {"FragmentCorrectionWithQualitiesCUDA", "sample_reads.fastq.gz", "sample_ava_overlaps.paf.gz", "sample_reads.fastq.gz",
     racon::PolisherType::kC, 500, 1, -1, -1, true, false},
    {"FragmentCorrectionWithQualitiesFullCUDA", "sample_reads.fastq.gz", "sample_ava_overlaps.paf.gz",
     "sample_reads.fastq.gz", racon::PolisherType::kF, 500, 1, -1, -1, false, false},
    {"FragmentCorrectionWithoutQualitiesFullCUDA", "sample_reads.fasta.gz", "sample_ava_overlaps.paf.gz",
     "sample_reads.fasta.gz", racon::PolisherType::kF, 500, 1, -1, -1, false, false},
    {"FragmentCorrectionWithQualitiesFullMhapCUDA", "sample_reads.fastq.gz", "sample_ava_overlaps.mhap.gz",
     "sample_reads.fastq.gz", racon::PolisherType::kF, 500, 1, -1, -1, false, false},
};

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: racon_goldens <test data dir> [--banded] [--aligner-batches N] [--only NAME]\n");
        return 2;
    }
    std::string dir = argv[1];
    if (!dir.empty() && dir.back() != '/') dir += '/';
    bool banded = false;
    uint32_t aligner_batches = 0;
    const char* only = nullptr;
    for (int i = 2; i < argc; ++i) {
        if (!strcmp(argv[i], "--banded")) banded = true;
        else if (!strcmp(argv[i], "--aligner-batches") && i + 1 < argc) aligner_batches = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--only") && i + 1 < argc) only = argv[++i];
    }
    for (const Case& c : kCases) {
        if (only && strcmp(only, c.name)) continue;
        /* RaconPolishingTest::SetUp (test/racon_test.cpp:27-37): quality 10, error 0.3, trim, 4 threads, 1 cudapoa batch */
        auto polisher = racon::createPolisher(dir + c.reads, dir + c.overlaps, dir + c.targets, c.type, c.window, 10, 0.3,
                                              true, c.m, c.x, c.g, 4, 1, banded, aligner_batches);
        polisher->initialize();
        std::vector<std::unique_ptr<racon::Sequence>> polished;
        polisher->polish(polished, c.drop_unpolished);
        uint64_t total = 0;
        for (const auto& s : polished) total += s->data().size();
        long long dist = -1;
        if (c.vs_reference && polished.size() == 1) {
            polished[0]->create_reverse_complement();
            auto parser = bioparser::Parser<racon::Sequence>::Create<bioparser::FastaParser>(dir + "sample_reference.fasta.gz");
            auto reference = parser->Parse(-1);
            dist = edit_distance(polished[0]->reverse_complement(), reference[0]->data());
        }
        printf("{\"case\": \"%s\", \"banded\": %s, \"aligner_batches\": %u, \"sequences\": %zu, \"total_length\": %llu, "
               "\"edit_distance\": %lld}\n",
               c.name, banded ? "true" : "false", aligner_batches, polished.size(), (unsigned long long)total, dist);
        fflush(stdout);
    }
    return 0;
}
