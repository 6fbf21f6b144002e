/*
 * Stand-in for <claraparabricks/genomeworks/cudapoa/batch.hpp>: racon's unmodified src/cuda/cudabatch.hpp keeps a
 * std::unique_ptr<claraparabricks::genomeworks::cudapoa::Batch> (cudabatch.hpp:105-106) and cudapolisher.cpp:44 calls
 * cudapoa::Init().  Here Batch is nothing but the owner of an rp_poa object of the C ABI (include/racon_b200.h);
 * integration/cudabatch.cpp implements racon::CUDABatchProcessor on top of it.
 */
#pragma once
#include <stdexcept>
#include <string>

#include "racon_b200.h"

namespace claraparabricks {
namespace genomeworks {
namespace cudapoa {

inline void Init() {}

class Batch {
public:
    Batch(int device, size_t mem_bytes, int8_t match, int8_t mismatch, int8_t gap, bool banded, uint32_t window_len) {
        rp_status s = rp_poa_create(&h_, device, mem_bytes, match, mismatch, gap, banded ? 1 : 0, window_len, 0);
        if (s != RP_OK) throw std::runtime_error(std::string("rp_poa_create: ") + rp_strerror(s) + " (" + rp_last_error() + ")");
    }
    ~Batch() { rp_poa_destroy(h_); }
    Batch(const Batch&) = delete;
    Batch& operator=(const Batch&) = delete;
    rp_poa* handle() const { return h_; }

private:
    rp_poa* h_ = nullptr;
};

}  // namespace cudapoa
}  // namespace genomeworks
}  // namespace claraparabricks
