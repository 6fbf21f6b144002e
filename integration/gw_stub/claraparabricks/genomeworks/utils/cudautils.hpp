/*
 * Stand-in for GenomeWorks' <claraparabricks/genomeworks/utils/cudautils.hpp>, found first on the include path when the
 * UNMODIFIED reference host (src/cuda/cudapolisher.cpp, src/cuda/cudabatch.hpp, src/cuda/cudaaligner.hpp) is compiled
 * against libracon_b200.so instead of GenomeWorks (integration/Makefile).  Only what those files use: GW_CU_CHECK_ERR
 * (cudapolisher.cpp:47-63,165-171,230-231; original: vendor/GenomeWorks/common/base/include/.../utils/cudautils.hpp).
 */
#pragma once
#include <cuda_runtime_api.h>

#include <cstdio>
#include <cstdlib>

#define GW_CU_CHECK_ERR(ans)                                                                                   \
    do {                                                                                                       \
        cudaError_t gw_err_ = (ans);                                                                           \
        if (gw_err_ != cudaSuccess) {                                                                          \
            std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(gw_err_), __FILE__, __LINE__); \
            std::abort();                                                                                      \
        }                                                                                                      \
    } while (0)
