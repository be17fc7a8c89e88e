// Mode-pair dispatch shared by the kernel TUs: picks the (A loader, B loader) instantiation.
#pragma once

#include "mtm_kernels.h"
#include "mtm_simt.cuh"

namespace b200 {

inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// KERNEL(AMODE, BMODE) must expand to a __global__ function pointer expression.
#define B200_DISPATCH_MODES(KERNEL, BM, BN, THREADS)                                              \
    do {                                                                                          \
        int64_t const tiles_m = ceil_div64(s.M, BM), tiles_n = ceil_div64(s.N, BN);               \
        int64_t const grid = tiles_m * tiles_n;                                                   \
        if (grid <= 0) return cudaSuccess;                                                        \
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;                            \
        dim3 const g((unsigned)grid), b(THREADS);                                                 \
        if (amode == LOAD_GENERIC || bmode == LOAD_GENERIC)                                       \
            KERNEL(LOAD_GENERIC, LOAD_GENERIC)<<<g, b, 0, stream>>>(C, A, B, s, tiles_m, tiles_n, vec_c); \
        else if (amode == LOAD_MN_VEC && bmode == LOAD_MN_VEC)                                    \
            KERNEL(LOAD_MN_VEC, LOAD_MN_VEC)<<<g, b, 0, stream>>>(C, A, B, s, tiles_m, tiles_n, vec_c);   \
        else if (amode == LOAD_MN_VEC && bmode == LOAD_K_VEC)                                     \
            KERNEL(LOAD_MN_VEC, LOAD_K_VEC)<<<g, b, 0, stream>>>(C, A, B, s, tiles_m, tiles_n, vec_c);    \
        else if (amode == LOAD_K_VEC && bmode == LOAD_MN_VEC)                                     \
            KERNEL(LOAD_K_VEC, LOAD_MN_VEC)<<<g, b, 0, stream>>>(C, A, B, s, tiles_m, tiles_n, vec_c);    \
        else                                                                                      \
            KERNEL(LOAD_K_VEC, LOAD_K_VEC)<<<g, b, 0, stream>>>(C, A, B, s, tiles_m, tiles_n, vec_c);     \
        return cudaGetLastError();                                                                \
    } while (0)

}  // namespace b200
