// Shared definitions for the sm_100a mtm kernels.
//
// Every kernel works on ONE canonical problem:
//
//     C[m, n] += sum_k A(m, k) * B(k, n),   C row-contiguous:  C(m, n) = c[m * ldc + n]
//     A(m, k) = a[m * a_sm + k * a_sk],     B(k, n) = b[k * b_sk + n * b_sn]
//
// The reference's eight (C, A, B) layout combinations (test/test.mtm.cpp:30-460) and its
// arbitrary two-stride operands (include/utils.hpp:99-141 honours both strides) reduce to this
// form on the host: a column-major C is handled through C^T += B^T * A^T, i.e. by swapping the
// operands and their strides (mtm_api.cu: canonicalise()).  There is no transpose pass and no
// packing pass (the reference's amt::pack, utils.hpp:99-141): strides go straight into the
// tile loaders.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace b200 {

struct MtmShape {
    int64_t M, N, K;
    int64_t a_sm, a_sk;  // element strides of A along m and k
    int64_t b_sk, b_sn;  // element strides of B along k and n
    int64_t ldc;         // element stride of C along m (n is unit stride)
};

// How a tile loader walks global memory for one operand.
enum LoadMode : int {
    LOAD_MN_VEC = 0,   // unit stride along m (A) / n (B): 16-byte loads along the smem-fast dim
    LOAD_K_VEC = 1,    // unit stride along k: 16-byte loads along k, transposed on the smem store
    LOAD_GENERIC = 2,  // any strides / any alignment: scalar loads
};

// Kernel families (values are the C-ABI `flags` variant codes, include/b200_mtm.h).
enum Variant : int {
    VAR_AUTO = 0,
    VAR_SIMT = 1,     // fp32 FFMA / fp64 DFMA on the CUDA cores
    VAR_3XTF32 = 2,   // fp32 via three tcgen05 kind::tf32 MMAs (hi*hi + hi*lo + lo*hi), TMEM accumulators
    VAR_DFMA = 3,     // fp64 DFMA (alias of SIMT for double)
    VAR_DMMA = 4,     // fp64 mma.sync m8n8k4 (DMMA)
};

template <typename T>
struct VecOf;
template <>
struct VecOf<float> {
    using type = float4;
    static constexpr int N = 4;
};
template <>
struct VecOf<double> {
    using type = double2;
    static constexpr int N = 2;
};

__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }

// Grouped tile order: consecutive CTAs walk down GROUP tile-rows before moving to the next
// tile-column, so the CTAs resident at any time share A row-panels and B column-panels in L2.
template <int GROUP>
__device__ __forceinline__ void tile_coords(int64_t pid, int64_t tiles_m, int64_t tiles_n,
                                            int64_t& pid_m, int64_t& pid_n) {
    int64_t const width = GROUP * tiles_n;
    int64_t const group_id = pid / width;
    int64_t const first_m = group_id * GROUP;
    int64_t const gsz = (tiles_m - first_m) < GROUP ? (tiles_m - first_m) : GROUP;
    int64_t const r = pid - group_id * width;
    pid_m = first_m + r % gsz;
    pid_n = r / gsz;
}

// Same walk with a run-time group height (the gated tensor-core product widens the group so that the first
// wave of tiles touches few column panels of B: those are still arriving).
__device__ __forceinline__ void tile_coords_rt(int64_t pid, int64_t tiles_m, int64_t tiles_n, int group,
                                               int64_t& pid_m, int64_t& pid_n) {
    int64_t const width = (int64_t)group * tiles_n;
    int64_t const group_id = pid / width;
    int64_t const first_m = group_id * group;
    int64_t const gsz = (tiles_m - first_m) < group ? (tiles_m - first_m) : group;
    int64_t const r = pid - group_id * width;
    pid_m = first_m + r % gsz;
    pid_n = r / gsz;
}

}  // namespace b200
