// Status / error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mwa_b200.h"

namespace b200 {

void record_cuda_error(cudaError_t e, const char* where);   // abi.cu

inline int check_launch(const char* where) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        record_cuda_error(e, where);
        return MWA_ERR_CUDA;
    }
    return MWA_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

}  // namespace b200

#define MWA_TRY_CUDA(expr, where)                      \
    do {                                               \
        cudaError_t e__ = (expr);                      \
        if (e__ != cudaSuccess) {                      \
            ::b200::record_cuda_error(e__, where);     \
            return MWA_ERR_CUDA;                       \
        }                                              \
    } while (0)
