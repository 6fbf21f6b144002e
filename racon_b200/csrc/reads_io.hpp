/*
 * reads_io.hpp — input side of the host layer: sequence and overlap files -> the in-memory state
 * racon::Polisher::initialize() holds just before it looks for breaking points (/root/reference/src/polisher.cpp:200-358):
 * targets first, then the reads that are not a target, overlaps transmuted to indices and filtered.
 *
 *   formats, chosen by file extension like createPolisher (polisher.cpp:84-141):
 *     sequences  .fasta .fa .fna | .fastq .fq          (each also .gz)
 *     overlaps   .mhap | .paf | .sam                   (each also .gz)
 *   record rules follow bioparser 3.0.x as racon uses it (names cut at the first white space, data lines right-stripped and
 *   joined, SAM header lines skipped) and racon's record constructors (sequence.cpp:20-45: bases upper-cased, an all-'!'
 *   quality string dropped; overlap.cpp:15-118: span, error and — for SAM — coordinates from the CIGAR).
 *
 * Everything here is host code over zlib; the byte-heavy work of the pipeline (alignment, breaking points, consensus)
 * stays on the device behind the C ABI.
 */
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "host_mirror.hpp"

namespace racon_b200 {

struct OwnedSequence {   // racon::Sequence after its constructor (sequence.cpp:20-45)
    std::string name, data, quality;   // quality empty when the file had none or only '!'
};

enum class SequenceFormat { kFasta, kFastq };
enum class OverlapFormat { kMhap, kPaf, kSam };

/* false when the extension is none of the accepted ones */
bool sequence_format_of(const std::string& path, SequenceFormat* fmt);
bool overlap_format_of(const std::string& path, OverlapFormat* fmt);

/* Whole file -> records, in file order.  Throw std::runtime_error on an unreadable file or a malformed record. */
std::vector<OwnedSequence> parse_sequences(const std::string& path);

struct RawOverlap {      // racon::Overlap after its format-specific constructor, before transmute()
    std::string q_name, t_name;      // empty for MHAP (ids instead)
    uint64_t q_id = 0, t_id = 0;     // MHAP: 0-based ordinal in the reads / target file
    uint32_t q_begin = 0, q_end = 0, q_length = 0, t_begin = 0, t_end = 0, t_length = 0, strand = 0;
    uint32_t length = 0;             // max of the two spans
    double error = 0;                // 1 - min span / max span
    std::string cigar;               // SAM only
    bool valid = true;               // SAM: false for an unmapped record (flag 0x4)
};
std::vector<RawOverlap> parse_overlaps(const std::string& path);

struct InputSet {
    std::vector<OwnedSequence> sequences;   // targets_size targets, then the reads that are not one of the targets
    uint64_t targets_size = 0;
    WindowType window_type = WindowType::kTGS;   // kNGS when the reads average <= 1000 bases (polisher.cpp:276-277)
    std::vector<Overlap> overlaps;          // transmuted + filtered, file order
    std::vector<SequenceView> views() const;
};

/* polisher.cpp:200-358.  error_threshold = racon -e; fragment_correction = racon -f (keeps every overlap of a read
 * instead of its longest one).  Throws std::runtime_error where the reference prints an error and exits. */
InputSet load_input(const std::string& reads_path, const std::string& overlaps_path, const std::string& targets_path,
                    bool fragment_correction, double error_threshold);

/* breaking_points_from_cigar(Overlap&, window_length) — declared next to Overlap in host_mirror.hpp, defined here. */

}  // namespace racon_b200
