/*
 * integration/cudabatch.cpp — racon::CUDABatchProcessor implemented on the racon_b200 C ABI.
 *
 * Drop-in for /root/reference/src/cuda/cudabatch.cpp: it is compiled against the reference's OWN, unmodified
 * src/cuda/cudabatch.hpp (same class, same members, same factory) together with the unmodified
 * src/cuda/cudapolisher.cpp, src/polisher.cpp, src/main.cpp ... (integration/Makefile); only the GenomeWorks headers
 * that cudabatch.hpp includes are replaced by the stand-ins under integration/gw_stub/, where cudapoa::Batch is the owner
 * of an rp_poa object.
 *
 * Differences to the reference's CUDA implementation (each one moves the result TO the CPU path's result):
 *   - the layers' positions_ are passed on (rp_poa_add_window begin/end), so partial-span layers are aligned to the
 *     subgraph exactly as Window::generate_consensus does (the reference only uses them to order the layers,
 *     cudabatch.cpp:103-104);
 *   - no depth cap (cudapolisher.cpp:226 MAX_DEPTH_PER_WINDOW), no 1023-base cap (cudabatch.cpp:56);
 *   - kNGS windows are polished on the device (the reference marks them failed and re-runs them on the CPU,
 *     cudabatch.cpp:229-256); trimming of kTGS windows is done on the device with spoa's coverage definition.
 * What is kept: add-until-full back-pressure (addWindow == false), one status per accepted window, a window the device
 * could not finish (soft RP_WIN_* status) is reported as false with an empty consensus, so the caller's CPU pass
 * (cudapolisher.cpp:354-370) produces it.
 */
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "cuda/cudabatch.hpp"
#include <claraparabricks/genomeworks/utils/cudautils.hpp>

namespace racon {

using claraparabricks::genomeworks::cudapoa::Batch;

std::atomic<uint32_t> CUDABatchProcessor::batches;

std::unique_ptr<CUDABatchProcessor> createCUDABatch(uint32_t max_window_depth, uint32_t device, size_t avail_mem,
                                                    int8_t gap, int8_t mismatch, int8_t match,
                                                    bool cuda_banded_alignment) {
    return std::unique_ptr<CUDABatchProcessor>(
        new CUDABatchProcessor(max_window_depth, device, avail_mem, gap, mismatch, match, cuda_banded_alignment));
}

CUDABatchProcessor::CUDABatchProcessor(uint32_t max_window_depth, uint32_t device, size_t avail_mem, int8_t gap,
                                       int8_t mismatch, int8_t match, bool cuda_banded_alignment)
    : windows_(), seqs_added_per_window_() {
    (void)max_window_depth;  // no layer is dropped
    bid_ = CUDABatchProcessor::batches++;
    GW_CU_CHECK_ERR(cudaSetDevice(device));
    GW_CU_CHECK_ERR(cudaStreamCreate(&stream_));
    /* window length 0: the per-window scratch is sized from the first batch (createCUDABatch is not told -w) */
    cudapoa_batch_.reset(new Batch(static_cast<int>(device), avail_mem, match, mismatch, gap, cuda_banded_alignment, 0));
    rp_poa_set_stream(cudapoa_batch_->handle(), stream_);
}

CUDABatchProcessor::~CUDABatchProcessor() {
    cudapoa_batch_.reset();
    GW_CU_CHECK_ERR(cudaStreamDestroy(stream_));
}

bool CUDABatchProcessor::addWindow(std::shared_ptr<Window> window) {
    const uint32_t n = static_cast<uint32_t>(window->sequences_.size());
    std::vector<const char*> seq(n), qual(n);
    std::vector<uint32_t> len(n), begin(n), end(n);
    for (uint32_t i = 0; i < n; ++i) {
        seq[i] = window->sequences_[i].first;
        len[i] = window->sequences_[i].second;
        qual[i] = window->qualities_[i].first;  // nullptr => weight 1 (graph.cpp:137-146)
        begin[i] = window->positions_[i].first;
        end[i] = window->positions_[i].second;
    }
    /* the reference's CUDA path always trims kTGS windows (cudabatch.cpp:231); so does the CPU path by default */
    const rp_status s = rp_poa_add_window(cudapoa_batch_->handle(), n, seq.data(), len.data(), qual.data(), begin.data(),
                                          end.data(), window->type_ == WindowType::kTGS ? RP_WINDOW_TGS : RP_WINDOW_NGS, 1);
    if (s == RP_BATCH_FULL) return false;
    if (s != RP_OK) {
        fprintf(stderr, "[racon::CUDABatchProcessor::addWindow] error: %s (%s)\n", rp_strerror(s), rp_last_error());
        exit(1);
    }
    windows_.push_back(window);
    seqs_added_per_window_.push_back(n ? n - 1 : 0);
    return true;
}

bool CUDABatchProcessor::hasWindows() const { return rp_poa_size(cudapoa_batch_->handle()) > 0; }

void CUDABatchProcessor::convertPhredQualityToWeights(const char*, uint32_t, std::vector<int8_t>& weights) {
    weights.clear();  // done inside the library while it copies a window into its pinned staging (poa_pack.hpp)
}

void CUDABatchProcessor::generatePOA() {
    const rp_status s = rp_poa_run(cudapoa_batch_->handle());  // H2D + kernel + D2H, asynchronous on stream_
    if (s != RP_OK) {
        fprintf(stderr, "[racon::CUDABatchProcessor::generatePOA] error: %s (%s)\n", rp_strerror(s), rp_last_error());
        exit(1);
    }
}

void CUDABatchProcessor::getConsensus() {
    rp_poa* h = cudapoa_batch_->handle();
    rp_status s = rp_poa_sync(h);
    if (s != RP_OK) {
        fprintf(stderr, "[racon::CUDABatchProcessor::getConsensus] error: %s (%s)\n", rp_strerror(s), rp_last_error());
        exit(1);
    }
    for (uint32_t i = 0; i < windows_.size(); ++i) {
        const char* c = nullptr;
        uint32_t len = 0, status = RP_WIN_OK;
        int polished = 0;
        rp_poa_fetch(h, i, &c, &len, nullptr, &polished);
        rp_poa_window_status(h, i, &status);
        if (status != RP_WIN_OK) {
            /* a device limit even the escalation pass could not lift: leave it to the caller's CPU pass */
            window_consensus_status_.emplace_back(false);
            continue;
        }
        windows_[i]->consensus_.assign(c, len);
        window_consensus_status_.emplace_back(polished != 0);
    }
}

const std::vector<bool>& CUDABatchProcessor::generateConsensus() {
    generatePOA();
    getConsensus();
    return window_consensus_status_;
}

void CUDABatchProcessor::reset() {
    windows_.clear();
    window_consensus_status_.clear();
    seqs_added_per_window_.clear();
    rp_poa_reset(cudapoa_batch_->handle());
}

}  // namespace racon
