/*
 * integration/cudaaligner.cpp — racon::CUDABatchAligner implemented on the racon_b200 C ABI.
 *
 * Drop-in for /root/reference/src/cuda/cudaaligner.cpp, compiled against the reference's own unmodified
 * src/cuda/cudaaligner.hpp; cudaaligner::Aligner (integration/gw_stub) owns an rp_aln object.
 * Overlap::cigar_ receives the byte-identical CIGAR edlib would have produced (Overlap::align_overlaps,
 * src/overlap.cpp:205-224), so the caller's CIGAR walk and everything after it equal the CPU run.  An overlap the device
 * cannot take (soft RP_ALN_* status) keeps an empty cigar_ and is aligned by the caller's CPU pass exactly like a
 * cudaaligner failure (cudapolisher.cpp:199-213).  `max_bandwidth` has no counterpart: the kernel finds the exact
 * edit distance by band doubling like edlib (k = -1), it never returns an approximate alignment.
 */
#include <cstdio>
#include <cstdlib>

#include <claraparabricks/genomeworks/utils/cudautils.hpp>

#include "cuda/cudaaligner.hpp"

namespace racon {

using claraparabricks::genomeworks::cudaaligner::Aligner;

std::atomic<uint32_t> CUDABatchAligner::batches;

std::unique_ptr<CUDABatchAligner> createCUDABatchAligner(uint32_t max_bandwidth, uint32_t device_id,
                                                         int64_t max_gpu_memory) {
    return std::unique_ptr<CUDABatchAligner>(new CUDABatchAligner(max_bandwidth, device_id, max_gpu_memory));
}

CUDABatchAligner::CUDABatchAligner(uint32_t max_bandwidth, uint32_t device_id, int64_t max_gpu_memory)
    : overlaps_(), stream_(0) {
    (void)max_bandwidth;
    bid_ = CUDABatchAligner::batches++;
    GW_CU_CHECK_ERR(cudaSetDevice(device_id));
    GW_CU_CHECK_ERR(cudaStreamCreate(&stream_));
    aligner_.reset(new Aligner(static_cast<int>(device_id), max_gpu_memory > 0 ? static_cast<size_t>(max_gpu_memory) : 0, 0));
    rp_aln_set_stream(aligner_->handle(), stream_);
}

CUDABatchAligner::~CUDABatchAligner() {
    aligner_.reset();
    GW_CU_CHECK_ERR(cudaStreamDestroy(stream_));
}

bool CUDABatchAligner::addOverlap(Overlap* overlap, std::vector<std::unique_ptr<Sequence>>& sequences) {
    /* an overlap that came with its alignment (SAM input) keeps it, as on the CPU path (Overlap::find_breaking_points,
     * overlap.cpp:179-203 aligns only when cigar_ is empty); the reference's CUDA shim re-aligns it, which is one of the
     * reasons its results differ from the CPU build's */
    if (!overlap->cigar_.empty()) return true;
    /* same spans as Overlap::find_breaking_points (overlap.cpp:193-197) and the reference shim (cudaaligner.cpp:53-57) */
    const char* q = !overlap->strand_ ? &(sequences[overlap->q_id_]->data()[overlap->q_begin_])
                                      : &(sequences[overlap->q_id_]->reverse_complement()[overlap->q_length_ - overlap->q_end_]);
    const uint32_t q_len = overlap->q_end_ - overlap->q_begin_;
    const char* t = &(sequences[overlap->t_id_]->data()[overlap->t_begin_]);
    const uint32_t t_len = overlap->t_end_ - overlap->t_begin_;
    const rp_status s = rp_aln_add(aligner_->handle(), q, q_len, t, t_len);  // edlib's roles: query = read, target = contig
    if (s == RP_BATCH_FULL) return false;
    if (s != RP_OK) {
        fprintf(stderr, "[racon::CUDABatchAligner::addOverlap] error: %s (%s)\n", rp_strerror(s), rp_last_error());
        exit(1);
    }
    overlaps_.push_back(overlap);
    return true;
}

void CUDABatchAligner::alignAll() {
    const rp_status s = rp_aln_run(aligner_->handle());
    if (s != RP_OK) {
        fprintf(stderr, "[racon::CUDABatchAligner::alignAll] error: %s (%s)\n", rp_strerror(s), rp_last_error());
        exit(1);
    }
}

void CUDABatchAligner::generate_cigar_strings() {
    rp_aln* h = aligner_->handle();
    if (rp_aln_sync(h) != RP_OK) {
        fprintf(stderr, "[racon::CUDABatchAligner::generate_cigar_strings] error: %s\n", rp_last_error());
        exit(1);
    }
    if (overlaps_.size() != rp_aln_size(h))
        throw std::runtime_error("Number of alignments doesn't match number of overlaps in cudaaligner.");
    for (uint32_t a = 0; a < overlaps_.size(); ++a) {
        const char* cigar = nullptr;
        uint32_t len = 0, status = RP_ALN_OK;
        rp_aln_fetch_cigar(h, a, &cigar, &len, nullptr, &status);
        if (status == RP_ALN_OK) overlaps_[a]->cigar_.assign(cigar, len);  // else: stays empty => CPU edlib (cudapolisher.cpp:213)
    }
}

void CUDABatchAligner::reset() {
    overlaps_.clear();
    cpu_overlap_data_.clear();
    rp_aln_reset(aligner_->handle());
}

}  // namespace racon
