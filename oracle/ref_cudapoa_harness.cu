/*
 * ref_cudapoa_harness.cu — TEST/BENCH INFRASTRUCTURE ONLY: the reference's OWN GPU path, unmodified.
 *
 * Drives GenomeWorks cudapoa (vendor/GenomeWorks/cudapoa, the library racon links with -Dracon_enable_cuda=ON) the
 * way racon::CUDABatchProcessor does (/root/reference/src/cuda/cudabatch.cpp:41-68 create_batch with
 * BatchConfig(1023, max_window_depth, 256, band); :77-150 addWindow -> add_poa_group, layers sorted by begin,
 * weights from qualities; :153-175 generate_poa; :177-270 get_consensus, reversed strings, status handling), over
 * the flat window-set arrays used everywhere in this repo.  It exists to time the GPU baseline SURVEY §8(d) names
 * ("reference cudapoa ... same windows, on the same box") beside our kernel; cudapoa's consensus is NOT spoa's
 * (Kahn topological order, full-layer alignment), so its output is never used as a parity oracle.
 * Built by `make -C oracle refcuda` from the sources where they lie; nothing is copied.
 */
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <atomic>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include <cuda_runtime_api.h>

#include <claraparabricks/genomeworks/cudapoa/batch.hpp>
#include <claraparabricks/genomeworks/cudapoa/cudapoa.hpp>

using namespace claraparabricks::genomeworks;
using namespace claraparabricks::genomeworks::cudapoa;

extern "C" {

/* Returns the number of windows cudapoa reported success for, <0 on a hard error.
 * times[0] = wall seconds of the whole loop (add + generate + fetch), times[1] = seconds inside
 * generate_poa + get_consensus only.  out: n_windows x stride, out_len per window (0 when cudapoa rejected it). */
int64_t ref_cudapoa_consensus(uint32_t n_windows, const char* bases, const char* quals, const uint64_t* seq_off,
                              const uint8_t* seq_has_qual, const uint32_t* seq_begin, const uint32_t* win_first,
                              int8_t match, int8_t mismatch, int8_t gap, int banded, uint32_t max_depth, int device,
                              double mem_fraction, char* out, uint32_t stride, uint32_t* out_len, double* times) {
    if (cudapoa::Init() != StatusType::success) return -1;
    cudaSetDevice(device);
    cudaFree(0);
    cudaStream_t stream;
    if (cudaStreamCreate(&stream) != cudaSuccess) return -2;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const size_t mem = static_cast<size_t>(free_b * mem_fraction);
    BatchConfig cfg(1023, max_depth, 256, banded ? BandMode::static_band : BandMode::full_band);
    std::unique_ptr<Batch> batch = create_batch(device, stream, mem, OutputType::consensus, cfg, gap, mismatch, match);
    int64_t ok = 0;
    double t_gpu = 0;
    const auto t0 = std::chrono::steady_clock::now();
    uint32_t w = 0;
    std::vector<std::vector<std::vector<int8_t>>> keep;  // weights must outlive add_poa_group? (copied inside) kept anyway
    while (w < n_windows) {
        batch->reset();
        keep.clear();
        const uint32_t first = w;
        for (; w < n_windows; ++w) {
            const uint32_t s0 = win_first[w], s1 = win_first[w + 1];
            const uint32_t ns = s1 - s0;
            keep.emplace_back(ns);
            auto& weights = keep.back();
            std::vector<uint32_t> rank(ns);
            for (uint32_t i = 0; i < ns; ++i) rank[i] = i;
            std::sort(rank.begin() + 1, rank.end(),
                      [&](uint32_t l, uint32_t r) { return seq_begin[s0 + l] < seq_begin[s0 + r]; });
            Group group;
            for (uint32_t j = 0; j < ns; ++j) {
                const uint32_t s = s0 + rank[j];
                const uint32_t len = static_cast<uint32_t>(seq_off[s + 1] - seq_off[s]);
                auto& wt = weights[j];
                if (quals && seq_has_qual && seq_has_qual[s]) {  // convertPhredQualityToWeights, cudabatch.cpp:272-286
                    wt.resize(len);
                    for (uint32_t k = 0; k < len; ++k) wt[k] = static_cast<int8_t>(quals[seq_off[s] + k] - 33);
                }
                Entry e = {bases + seq_off[s], wt.empty() ? nullptr : wt.data(), static_cast<int32_t>(len)};
                group.push_back(e);
            }
            std::vector<StatusType> entry_status;
            StatusType st = batch->add_poa_group(entry_status, group);
            if (st != StatusType::success) {
                keep.pop_back();
                break;
            }
        }
        if (w == first) {  // a window that does not fit an empty batch: cudapoa cannot take it (racon -> CPU)
            out_len[w] = 0;
            ++w;
            continue;
        }
        const auto g0 = std::chrono::steady_clock::now();
        batch->generate_poa();
        std::vector<std::string> consensuses;
        std::vector<std::vector<uint16_t>> coverages;
        std::vector<StatusType> output_status;
        batch->get_consensus(consensuses, coverages, output_status);
        t_gpu += std::chrono::duration<double>(std::chrono::steady_clock::now() - g0).count();
        for (uint32_t i = 0; i < consensuses.size(); ++i) {
            const uint32_t wi = first + i;
            if (output_status[i] != StatusType::success) {
                out_len[wi] = 0;
                continue;
            }
            std::string c(consensuses[i].rbegin(), consensuses[i].rend());  // cudabatch.cpp:222: reverse
            const uint32_t n = static_cast<uint32_t>(std::min<size_t>(c.size(), stride));
            std::memcpy(out + static_cast<uint64_t>(wi) * stride, c.data(), n);
            out_len[wi] = n;
            ++ok;
        }
    }
    times[0] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    times[1] = t_gpu;
    batch.reset();
    cudaStreamDestroy(stream);
    return ok;
}

/* The reference's multi-GPU, multi-batch driver restated as a harness: CUDAPolisher::polish creates `batches_per_device`
 * batch objects on EVERY visible device, each with 0.9 * free / batches bytes (/root/reference/src/cuda/cudapolisher.cpp:
 * 226-240), and runs one host thread per batch object that, under a mutex, fills its batch with the next windows until
 * add_poa_group refuses one (:254-276), then generate_poa + get_consensus, until the window list is exhausted
 * (:284-345).  Same here, over the first n_devices devices.  times[0] = wall seconds from the first fill to the last
 * result (batch creation excluded, as bench.py excludes it for our arm too).  Returns windows reported successful. */
int64_t ref_cudapoa_multi(uint32_t n_windows, const char* bases, const char* quals, const uint64_t* seq_off,
                          const uint8_t* seq_has_qual, const uint32_t* seq_begin, const uint32_t* win_first,
                          int8_t match, int8_t mismatch, int8_t gap, int banded, uint32_t max_depth, int n_devices,
                          int batches_per_device, double* times) {
    if (cudapoa::Init() != StatusType::success) return -1;
    struct Obj {
        int device;
        cudaStream_t stream;
        std::unique_ptr<Batch> batch;
    };
    std::vector<Obj> objs;
    for (int d = 0; d < n_devices; ++d) {
        if (cudaSetDevice(d) != cudaSuccess) return -2;
        cudaFree(0);
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const size_t mem = static_cast<size_t>(0.9 * free_b / batches_per_device);
        for (int b = 0; b < batches_per_device; ++b) {
            Obj o;
            o.device = d;
            if (cudaStreamCreate(&o.stream) != cudaSuccess) return -2;
            BatchConfig cfg(1023, max_depth, 256, banded ? BandMode::static_band : BandMode::full_band);
            o.batch = create_batch(d, o.stream, mem, OutputType::consensus, cfg, gap, mismatch, match);
            objs.push_back(std::move(o));
        }
    }
    std::mutex mtx;
    uint32_t next = 0;
    std::atomic<int64_t> ok{0};
    auto worker = [&](Obj* o) {
        cudaSetDevice(o->device);
        std::vector<std::vector<std::vector<int8_t>>> keep;
        for (;;) {
            o->batch->reset();
            keep.clear();
            uint32_t first, last;
            {
                std::lock_guard<std::mutex> guard(mtx);
                first = next;
                while (next < n_windows) {
                    const uint32_t w = next;
                    const uint32_t s0 = win_first[w], s1 = win_first[w + 1];
                    const uint32_t ns = s1 - s0;
                    keep.emplace_back(ns);
                    auto& weights = keep.back();
                    std::vector<uint32_t> rank(ns);
                    for (uint32_t i = 0; i < ns; ++i) rank[i] = i;
                    std::sort(rank.begin() + 1, rank.end(),
                              [&](uint32_t l, uint32_t r) { return seq_begin[s0 + l] < seq_begin[s0 + r]; });
                    Group group;
                    for (uint32_t j = 0; j < ns; ++j) {
                        const uint32_t s = s0 + rank[j];
                        const uint32_t len = static_cast<uint32_t>(seq_off[s + 1] - seq_off[s]);
                        auto& wt = weights[j];
                        if (quals && seq_has_qual && seq_has_qual[s]) {
                            wt.resize(len);
                            for (uint32_t k = 0; k < len; ++k) wt[k] = static_cast<int8_t>(quals[seq_off[s] + k] - 33);
                        }
                        Entry e = {bases + seq_off[s], wt.empty() ? nullptr : wt.data(), static_cast<int32_t>(len)};
                        group.push_back(e);
                    }
                    std::vector<StatusType> entry_status;
                    if (o->batch->add_poa_group(entry_status, group) != StatusType::success) {
                        keep.pop_back();
                        if (next == first) ++next;  // does not fit an empty batch: the reference hands it to the CPU
                        break;
                    }
                    ++next;
                }
                last = next;
            }
            if (o->batch->get_total_poas() == 0) {
                if (first >= n_windows) break;
                continue;
            }
            o->batch->generate_poa();
            std::vector<std::string> consensuses;
            std::vector<std::vector<uint16_t>> coverages;
            std::vector<StatusType> output_status;
            o->batch->get_consensus(consensuses, coverages, output_status);
            int64_t good = 0;
            for (auto st : output_status) good += st == StatusType::success ? 1 : 0;
            ok += good;
            (void)last;
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> threads;
    for (auto& o : objs) threads.emplace_back(worker, &o);
    for (auto& t : threads) t.join();
    times[0] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (auto& o : objs) {
        cudaSetDevice(o.device);
        o.batch.reset();
        cudaStreamDestroy(o.stream);
    }
    return ok.load();
}

}  // extern "C"
