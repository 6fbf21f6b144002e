/*
 * Stand-in for <claraparabricks/genomeworks/cudaaligner/aligner.hpp>: racon's unmodified src/cuda/cudaaligner.hpp keeps
 * a std::unique_ptr<claraparabricks::genomeworks::cudaaligner::Aligner> (cudaaligner.hpp:77).  Here Aligner is the owner
 * of an rp_aln object of the C ABI; integration/cudaaligner.cpp implements racon::CUDABatchAligner on top of it.
 */
#pragma once
#include <stdexcept>
#include <string>

#include "racon_b200.h"

namespace claraparabricks {
namespace genomeworks {
namespace cudaaligner {

class Aligner {
public:
    Aligner(int device, size_t mem_bytes, uint32_t max_len) {
        rp_status s = rp_aln_create(&h_, device, mem_bytes, max_len);
        if (s != RP_OK) throw std::runtime_error(std::string("rp_aln_create: ") + rp_strerror(s) + " (" + rp_last_error() + ")");
    }
    ~Aligner() { rp_aln_destroy(h_); }
    Aligner(const Aligner&) = delete;
    Aligner& operator=(const Aligner&) = delete;
    rp_aln* handle() const { return h_; }

private:
    rp_aln* h_ = nullptr;
};

}  // namespace cudaaligner
}  // namespace genomeworks
}  // namespace claraparabricks
