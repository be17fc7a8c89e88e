// CUDA-core mtm kernels for sm_100a: fp32 FFMA and fp64 DFMA (mtm_simt_kernel), and the fp64
// tensor-core DMMA variant (mtm_dmma_kernel).
//
// These replace, on the GPU, the reference's packing + macro-kernel + SIMD micro-kernel
// (include/mtm.hpp:83-206, include/simd_loop.hpp:80-192, include/utils.hpp:99-141):
//   * the 5-loop cache blocking becomes a grid of BM x BN CTA tiles walked in an L2-friendly
//     grouped order, each CTA looping over K in BK slices;
//   * amt::pack becomes the tile loaders below: strides are applied while staging a slice into
//     shared memory ([k][m] for A, [k][n] for B), 16-byte vector loads whenever a unit stride
//     and alignment allow, transposing on the shared-memory store when k is the unit stride;
//   * the MR x NR register tile (16x6 AVX2 floats in the reference) becomes a TM x TN per-thread
//     accumulator block fed by 16-byte shared-memory reads;
//   * copy_from_buff's `C += buff` (simd_loop.hpp:160-190) is the epilogue: read-modify-write of C.
//
// Pipeline: global -> registers (prefetch of slice t+1 is in flight while slice t is multiplied)
// -> shared memory, two shared-memory stages, one __syncthreads per slice.
#pragma once

#include "mtm_common.cuh"

namespace b200 {

constexpr int SMEM_PAD = 4;  // elements; keeps rows 16-byte aligned and DMMA fragment reads conflict-free

template <typename T, int BMN, int BK, int NT, int LD, int MODE>
struct TileLoader {
    using Vec = typename VecOf<T>::type;
    static constexpr int V = VecOf<T>::N;
    static constexpr int NELEM = BMN * BK / NT;  // elements this thread stages per slice
    static constexpr int NVEC = NELEM / V;
    static_assert(BMN * BK % NT == 0, "tile must divide evenly over the CTA");
    static_assert(NELEM % V == 0 && NELEM >= V, "need at least one 16-byte vector per thread");
    static_assert(MODE != LOAD_MN_VEC || NT % (BMN / V) == 0, "MN_VEC thread map");
    static_assert(MODE != LOAD_K_VEC || NT % (BK / V) == 0, "K_VEC thread map");
    static_assert(MODE != LOAD_GENERIC || NT % BMN == 0, "GENERIC thread map");

    const T* ptr;       // this thread's first element of the current slice
    int64_t step_item;  // element offset between this thread's successive vectors/elements
    int64_t step_slice; // element offset from one K-slice to the next
    int mn_left;        // extent_mn - (tile origin + this thread's first mn index); <= 0: out of range
    int k_left;         // K - (slice origin + this thread's first k index)
    int smem_off;       // this thread's first smem element
    T regs[NELEM];

    __device__ __forceinline__ void init(const T* base, int64_t stride_mn, int64_t stride_k,
                                         int64_t origin_mn, int64_t extent_mn, int64_t K, int tid) {
        int mn, k;
        if constexpr (MODE == LOAD_MN_VEC) {
            mn = (tid % (BMN / V)) * V;
            k = tid / (BMN / V);
            step_item = (int64_t)(NT * V / BMN) * stride_k;
        } else if constexpr (MODE == LOAD_K_VEC) {
            k = (tid % (BK / V)) * V;
            mn = tid / (BK / V);
            step_item = (int64_t)(NT * V / BK) * stride_mn;
        } else {
            mn = tid % BMN;
            k = tid / BMN;
            step_item = (int64_t)(NT / BMN) * stride_k;
        }
        int64_t const left = extent_mn - origin_mn - mn;
        mn_left = left > 0 ? (left > BMN ? BMN : (int)left) : 0;
        k_left = (int)K - k;
        smem_off = k * LD + mn;
        step_slice = (int64_t)BK * stride_k;
        // Keep the pointer inside the allocation for fully out-of-range threads.
        ptr = base + (mn_left > 0 ? (origin_mn + mn) * stride_mn : 0) + (k_left > 0 ? k * stride_k : 0);
    }

    // Issue this slice's global loads into registers (zero-filled outside the matrix).
    __device__ __forceinline__ void load() {
        if constexpr (MODE == LOAD_MN_VEC) {
            constexpr int KSTEP = NT * V / BMN;
#pragma unroll
            for (int i = 0; i < NVEC; ++i) {
                const T* p = ptr + i * step_item;
                bool const k_ok = i * KSTEP < k_left;
                if (k_ok && mn_left >= V) {
                    Vec v = *reinterpret_cast<const Vec*>(p);
                    *reinterpret_cast<Vec*>(&regs[i * V]) = v;
                } else {
#pragma unroll
                    for (int j = 0; j < V; ++j) regs[i * V + j] = (k_ok && j < mn_left) ? p[j] : T(0);
                }
            }
        } else if constexpr (MODE == LOAD_K_VEC) {
            constexpr int MSTEP = NT * V / BK;
#pragma unroll
            for (int i = 0; i < NVEC; ++i) {
                const T* p = ptr + i * step_item;
                bool const mn_ok = i * MSTEP < mn_left;
                if (mn_ok && k_left >= V) {
                    Vec v = *reinterpret_cast<const Vec*>(p);
                    *reinterpret_cast<Vec*>(&regs[i * V]) = v;
                } else {
#pragma unroll
                    for (int j = 0; j < V; ++j) regs[i * V + j] = (mn_ok && j < k_left) ? p[j] : T(0);
                }
            }
        } else {
            constexpr int KSTEP = NT / BMN;
#pragma unroll
            for (int i = 0; i < NELEM; ++i)
                regs[i] = (mn_left > 0 && i * KSTEP < k_left) ? ptr[i * step_item] : T(0);
        }
    }

    __device__ __forceinline__ void next_slice() {
        k_left -= BK;
        if (k_left > 0) ptr += step_slice;  // never step past the last slice that has data
    }

    // Registers -> shared memory tile laid out [k][mn] with row stride LD.
    __device__ __forceinline__ void store(T* smem) const {
        T* s = smem + smem_off;
        if constexpr (MODE == LOAD_MN_VEC) {
            constexpr int KSTEP = NT * V / BMN;
#pragma unroll
            for (int i = 0; i < NVEC; ++i)
                *reinterpret_cast<Vec*>(s + i * KSTEP * LD) = *reinterpret_cast<const Vec*>(&regs[i * V]);
        } else if constexpr (MODE == LOAD_K_VEC) {
            constexpr int MSTEP = NT * V / BK;
#pragma unroll
            for (int i = 0; i < NVEC; ++i)
#pragma unroll
                for (int j = 0; j < V; ++j) s[j * LD + i * MSTEP] = regs[i * V + j];
        } else {
            constexpr int KSTEP = NT / BMN;
#pragma unroll
            for (int i = 0; i < NELEM; ++i) s[i * KSTEP * LD] = regs[i];
        }
    }
};

// ------------------------------------------------------------------------------------------
// FFMA / DFMA kernel.  Thread (tx, ty) owns rows  ty*V + i*(TY*V) + ii  and columns
// tx*V + j*(TX*V) + jj  (i < TM/V, j < TN/V, ii,jj < V): every fragment read is a 16-byte LDS,
// a warp is 8 (n) x 4 (m) threads so an A read touches 4 and a B read 8 distinct 16-byte words
// (one wavefront each), and C rows are written in 128-byte runs.
// ------------------------------------------------------------------------------------------
template <typename T, int BM, int BN, int BK, int TM, int TN, int MINB, int AMODE, int BMODE>
__global__ void __launch_bounds__((BM / TM) * (BN / TN), MINB)
mtm_simt_kernel(T* __restrict__ C, const T* __restrict__ A, const T* __restrict__ B, MtmShape s,
                int64_t tiles_m, int64_t tiles_n, int vec_c) {
    using Vec = typename VecOf<T>::type;
    constexpr int V = VecOf<T>::N;
    constexpr int TX = BN / TN, TY = BM / TM, NT = TX * TY;
    constexpr int LDA = BM + SMEM_PAD, LDB = BN + SMEM_PAD;
    static_assert(TX % 8 == 0 && TY % 4 == 0, "warp is 8 x 4 threads");
    static_assert(TM % V == 0 && TN % V == 0, "fragments are 16-byte vectors");

    __shared__ __align__(16) T As[2][BK * LDA];
    __shared__ __align__(16) T Bs[2][BK * LDB];

    int const tid = threadIdx.x;
    int const lane = tid & 31, warp = tid >> 5;
    constexpr int WX = TX / 8;
    int const tx = (warp % WX) * 8 + (lane & 7);
    int const ty = (warp / WX) * 4 + (lane >> 3);

    int64_t pid_m, pid_n;
    tile_coords<8>(blockIdx.x, tiles_m, tiles_n, pid_m, pid_n);
    int64_t const m0 = pid_m * BM, n0 = pid_n * BN;

    TileLoader<T, BM, BK, NT, LDA, AMODE> la;
    TileLoader<T, BN, BK, NT, LDB, BMODE> lb;
    la.init(A, s.a_sm, s.a_sk, m0, s.M, s.K, tid);
    lb.init(B, s.b_sn, s.b_sk, n0, s.N, s.K, tid);

    T acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = T(0);

    int const num_slices = (int)((s.K + BK - 1) / BK);

    la.load();
    lb.load();
    la.store(As[0]);
    lb.store(Bs[0]);
    __syncthreads();

    for (int t = 0; t < num_slices; ++t) {
        int const cur = t & 1;
        bool const more = t + 1 < num_slices;
        if (more) {
            la.next_slice();
            lb.next_slice();
            la.load();
            lb.load();
        }
        const T* as = As[cur] + ty * V;
        const T* bs = Bs[cur] + tx * V;
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            T af[TM], bf[TN];
#pragma unroll
            for (int i = 0; i < TM / V; ++i)
                *reinterpret_cast<Vec*>(&af[i * V]) = *reinterpret_cast<const Vec*>(as + k * LDA + i * TY * V);
#pragma unroll
            for (int j = 0; j < TN / V; ++j)
                *reinterpret_cast<Vec*>(&bf[j * V]) = *reinterpret_cast<const Vec*>(bs + k * LDB + j * TX * V);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fma_t(af[i], bf[j], acc[i][j]);
        }
        if (more) {
            la.store(As[cur ^ 1]);
            lb.store(Bs[cur ^ 1]);
        }
        __syncthreads();
    }

    // Epilogue: C += acc   (reference: copy_from_buff, simd_loop.hpp:160-190).
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int64_t const m = m0 + ty * V + (i / V) * (TY * V) + (i % V);
        if (m >= s.M) continue;
        T* crow = C + m * s.ldc;
#pragma unroll
        for (int j = 0; j < TN / V; ++j) {
            int64_t const n = n0 + tx * V + j * (TX * V);
            if (vec_c && n + V <= s.N) {
                Vec c = *reinterpret_cast<Vec*>(crow + n);
                T* cv = reinterpret_cast<T*>(&c);
#pragma unroll
                for (int jj = 0; jj < V; ++jj) cv[jj] += acc[i][j * V + jj];
                *reinterpret_cast<Vec*>(crow + n) = c;
            } else {
#pragma unroll
                for (int jj = 0; jj < V; ++jj)
                    if (n + jj < s.N) crow[n + jj] += acc[i][j * V + jj];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// fp64 DMMA kernel: mma.sync.aligned.m8n8k4 (SASS DMMA.8x8x4, the only native FP64 tensor shape
// on sm_100a — the m16n8k{4,8,16} PTX forms are split into it by ptxas).  Same loaders and
// shared-memory tiles as the DFMA kernel; each warp owns a WM x WN block of 8x8 MMA tiles.
//   A fragment (8x4, "row"): lane l holds A[m = l/4][k = l%4]
//   B fragment (4x8, "col"): lane l holds B[k = l%4][n = l/4]
//   C fragment (8x8):        lane l holds C[m = l/4][n = 2*(l%4) + {0,1}]
// ------------------------------------------------------------------------------------------
template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int MINB, int AMODE, int BMODE>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32, MINB)
mtm_dmma_kernel(double* __restrict__ C, const double* __restrict__ A, const double* __restrict__ B,
                MtmShape s, int64_t tiles_m, int64_t tiles_n, int vec_c) {
    constexpr int NT = WARPS_M * WARPS_N * 32;
    constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;  // warp tile
    constexpr int MT = WM / 8, NTL = WN / 8;             // 8x8 MMA tiles per warp
    constexpr int LDA = BM + SMEM_PAD, LDB = BN + SMEM_PAD;
    static_assert(BK % 4 == 0 && WM % 8 == 0 && WN % 8 == 0, "DMMA tiling");

    __shared__ __align__(16) double As[2][BK * LDA];
    __shared__ __align__(16) double Bs[2][BK * LDB];

    int const tid = threadIdx.x;
    int const lane = tid & 31, warp = tid >> 5;
    int const wm0 = (warp / WARPS_N) * WM, wn0 = (warp % WARPS_N) * WN;
    int const lq = lane >> 2, lr = lane & 3;

    int64_t pid_m, pid_n;
    tile_coords<8>(blockIdx.x, tiles_m, tiles_n, pid_m, pid_n);
    int64_t const m0 = pid_m * BM, n0 = pid_n * BN;

    TileLoader<double, BM, BK, NT, LDA, AMODE> la;
    TileLoader<double, BN, BK, NT, LDB, BMODE> lb;
    la.init(A, s.a_sm, s.a_sk, m0, s.M, s.K, tid);
    lb.init(B, s.b_sn, s.b_sk, n0, s.N, s.K, tid);

    double acc[MT][NTL][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTL; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    int const num_slices = (int)((s.K + BK - 1) / BK);

    la.load();
    lb.load();
    la.store(As[0]);
    lb.store(Bs[0]);
    __syncthreads();

    for (int t = 0; t < num_slices; ++t) {
        int const cur = t & 1;
        bool const more = t + 1 < num_slices;
        if (more) {
            la.next_slice();
            lb.next_slice();
            la.load();
            lb.load();
        }
        const double* as = As[cur] + lr * LDA + wm0 + lq;
        const double* bs = Bs[cur] + lr * LDB + wn0 + lq;
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            double af[MT], bf[NTL];
#pragma unroll
            for (int i = 0; i < MT; ++i) af[i] = as[k4 * LDA + i * 8];
#pragma unroll
            for (int j = 0; j < NTL; ++j) bf[j] = bs[k4 * LDB + j * 8];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NTL; ++j)
                    asm(
                        "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                        : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                        : "d"(af[i]), "d"(bf[j]));
        }
        if (more) {
            la.store(As[cur ^ 1]);
            lb.store(Bs[cur ^ 1]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < MT; ++i) {
        int64_t const m = m0 + wm0 + i * 8 + lq;
        if (m >= s.M) continue;
        double* crow = C + m * s.ldc;
#pragma unroll
        for (int j = 0; j < NTL; ++j) {
            int64_t const n = n0 + wn0 + j * 8 + 2 * lr;
            if (vec_c && n + 2 <= s.N) {
                double2 c = *reinterpret_cast<double2*>(crow + n);
                c.x += acc[i][j][0];
                c.y += acc[i][j][1];
                *reinterpret_cast<double2*>(crow + n) = c;
            } else {
                if (n < s.N) crow[n] += acc[i][j][0];
                if (n + 1 < s.N) crow[n + 1] += acc[i][j][1];
            }
        }
    }
}

}  // namespace b200
