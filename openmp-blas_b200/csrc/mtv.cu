// Matrix-times-vector kernels for sm_100a:  c[i] (op)= sum_k A(i,k) * b[k].
//
// Replaces the reference's amt::mtv_helper (include/mtv.hpp:15-100) and, through the host layer's
// layout flip, amt::vtm (mtv.hpp:170-236).  The operation streams A once and does 2 flops per
// element: it is HBM-bound (roofline = measured copy bandwidth), so the design is about keeping
// enough 16-byte loads in flight on every SM, not about the math pipes.
//
//   i-contiguous A (uBLAS first_order, the reference's blocked column sweep, mtv.hpp:39-70):
//     one thread owns V consecutive rows (one 16-byte load per k), the K range is cut into chunks
//     across blockIdx.y so that even a single tall slab fills the machine; chunk partials go to a
//     workspace and a second kernel adds them into c in chunk order (deterministic, no atomics).
//   k-contiguous A (last_order, the reference's dot-per-row, mtv.hpp:90-99):
//     one warp (or, for few long rows, one CTA) per row, lanes stride along k with 16-byte loads,
//     fixed-shape shuffle reduction.
// ACC selects the reference's per-layout semantics: first_order accumulates (c += A b,
// simd_loop.hpp:58-75), last_order assigns (c = A b, mtv.hpp:98).
#include "mtm_kernels.h"

namespace b200 {
namespace {

template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
    using type = float4;
    static constexpr int N = 4;
};
template <>
struct Vec16<double> {
    using type = double2;
    static constexpr int N = 2;
};

constexpr int MTV_THREADS = 256;
constexpr int KTILE = 512;  // b[] staged in shared memory per k-tile (i-contiguous kernel)

// ---- i-contiguous (V > 1: 16-byte loads, needs s_i == 1, lda % V == 0, aligned a) or generic strides (V == 1)
template <typename T, int V, bool ACC>
__global__ void __launch_bounds__(MTV_THREADS)
mtv_icontig_kernel(T* __restrict__ c, const T* __restrict__ a, int64_t s_i, int64_t s_k, const T* __restrict__ b,
                   int M, int K, int k_per_block, T* __restrict__ partial) {
    using VT = typename Vec16<T>::type;
    __shared__ T bs[KTILE];
    int64_t const i0 = ((int64_t)blockIdx.x * MTV_THREADS + threadIdx.x) * V;
    int const k_begin = blockIdx.y * k_per_block;
    int const k_end = min(K, k_begin + k_per_block);
    bool const full = i0 + V <= M;
    T acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = T(0);
    const T* ap = a + i0 * s_i + (int64_t)k_begin * s_k;

    for (int kt = k_begin; kt < k_end; kt += KTILE) {
        int const kn = min(KTILE, k_end - kt);
        __syncthreads();
        for (int k = threadIdx.x; k < kn; k += MTV_THREADS) bs[k] = b[kt + k];
        __syncthreads();
        if (full) {
            int k = 0;
            if constexpr (V > 1) {
                for (; k + 8 <= kn; k += 8) {   // 8 independent 16-byte loads in flight per thread
                    VT r[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) r[u] = *reinterpret_cast<const VT*>(ap + (int64_t)(k + u) * s_k);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const T* rv = reinterpret_cast<const T*>(&r[u]);
                        T const bk = bs[k + u];
#pragma unroll
                        for (int v = 0; v < V; ++v) acc[v] = fma_t(rv[v], bk, acc[v]);
                    }
                }
                for (; k < kn; ++k) {
                    VT r = *reinterpret_cast<const VT*>(ap + (int64_t)k * s_k);
                    const T* rv = reinterpret_cast<const T*>(&r);
#pragma unroll
                    for (int v = 0; v < V; ++v) acc[v] = fma_t(rv[v], bs[k], acc[v]);
                }
            } else {
                for (; k + 8 <= kn; k += 8) {
                    T r[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) r[u] = ap[(int64_t)(k + u) * s_k];
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc[0] = fma_t(r[u], bs[k + u], acc[0]);
                }
                for (; k < kn; ++k) acc[0] = fma_t(ap[(int64_t)k * s_k], bs[k], acc[0]);
            }
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (i0 + v < M)
                    for (int k = 0; k < kn; ++k) acc[v] = fma_t(ap[v * s_i + (int64_t)k * s_k], bs[k], acc[v]);
        }
        ap += (int64_t)kn * s_k;
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
        if (i0 + v >= M) continue;
        if (gridDim.y == 1) {
            if (ACC) c[i0 + v] += acc[v];
            else c[i0 + v] = acc[v];
        } else {
            partial[(int64_t)blockIdx.y * M + i0 + v] = acc[v];
        }
    }
}

// c[i] (op)= partial[0][i] + partial[1][i] + ... in chunk order.
template <typename T, bool ACC>
__global__ void __launch_bounds__(MTV_THREADS)
mtv_reduce_partials_kernel(T* __restrict__ c, const T* __restrict__ partial, int M, int chunks) {
    int64_t const i = (int64_t)blockIdx.x * MTV_THREADS + threadIdx.x;
    if (i >= M) return;
    T s = partial[i];
    for (int j = 1; j < chunks; ++j) s += partial[(int64_t)j * M + i];
    if (ACC) c[i] += s;
    else c[i] = s;
}

// ---- k-contiguous: WPR warps cooperate on one row (1 = warp per row, 8 = CTA per row)
template <typename T, int V, int WPR, bool ACC>
__global__ void __launch_bounds__(MTV_THREADS)
mtv_kcontig_kernel(T* __restrict__ c, const T* __restrict__ a, int64_t lda, const T* __restrict__ b, int M, int K) {
    using VT = typename Vec16<T>::type;
    constexpr int WARPS = MTV_THREADS / 32;
    constexpr int ROWS_PER_BLOCK = WARPS / WPR;
    __shared__ T red[WARPS];
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int const sub = warp % WPR;                 // which slice of the row this warp takes
    int const lanes_total = 32 * WPR;
    for (int64_t row = (int64_t)blockIdx.x * ROWS_PER_BLOCK + warp / WPR; row < M;
         row += (int64_t)gridDim.x * ROWS_PER_BLOCK) {
        const T* ar = a + row * lda;
        T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
        int const t = sub * 32 + lane;
        int k = t * V;
        int const stride = lanes_total * V;
        if constexpr (V > 1) {
            for (; k + 3 * stride + V <= K; k += 4 * stride) {   // 4 independent 16-byte loads per operand
                VT a0 = *reinterpret_cast<const VT*>(ar + k), a1 = *reinterpret_cast<const VT*>(ar + k + stride);
                VT a2 = *reinterpret_cast<const VT*>(ar + k + 2 * stride), a3 = *reinterpret_cast<const VT*>(ar + k + 3 * stride);
                VT b0 = *reinterpret_cast<const VT*>(b + k), b1 = *reinterpret_cast<const VT*>(b + k + stride);
                VT b2 = *reinterpret_cast<const VT*>(b + k + 2 * stride), b3 = *reinterpret_cast<const VT*>(b + k + 3 * stride);
                const T *pa0 = reinterpret_cast<const T*>(&a0), *pa1 = reinterpret_cast<const T*>(&a1);
                const T *pa2 = reinterpret_cast<const T*>(&a2), *pa3 = reinterpret_cast<const T*>(&a3);
                const T *pb0 = reinterpret_cast<const T*>(&b0), *pb1 = reinterpret_cast<const T*>(&b1);
                const T *pb2 = reinterpret_cast<const T*>(&b2), *pb3 = reinterpret_cast<const T*>(&b3);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    acc0 = fma_t(pa0[v], pb0[v], acc0);
                    acc1 = fma_t(pa1[v], pb1[v], acc1);
                    acc2 = fma_t(pa2[v], pb2[v], acc2);
                    acc3 = fma_t(pa3[v], pb3[v], acc3);
                }
            }
            for (; k + V <= K; k += stride) {
                VT a0 = *reinterpret_cast<const VT*>(ar + k);
                VT b0 = *reinterpret_cast<const VT*>(b + k);
                const T *pa0 = reinterpret_cast<const T*>(&a0), *pb0 = reinterpret_cast<const T*>(&b0);
#pragma unroll
                for (int v = 0; v < V; ++v) acc0 = fma_t(pa0[v], pb0[v], acc0);
            }
            // ragged end of the row (K % V elements), taken by the lane whose next vector would start there
            if (k < K)
                for (int kk = k; kk < K && kk < k + V; ++kk) acc1 = fma_t(ar[kk], b[kk], acc1);
        } else {
            for (; k < K; k += stride) acc0 = fma_t(ar[k], b[k], acc0);
        }
        T s = (acc0 + acc1) + (acc2 + acc3);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if constexpr (WPR == 1) {
            if (lane == 0) {
                if (ACC) c[row] += s;
                else c[row] = s;
            }
        } else {
            __syncthreads();
            if (lane == 0) red[warp] = s;
            __syncthreads();
            if (threadIdx.x == 0) {
                T tot = red[0];
#pragma unroll
                for (int w = 1; w < WARPS; ++w) tot += red[w];
                if (ACC) c[row] += tot;
                else c[row] = tot;
            }
        }
    }
}

template <typename T>
bool aligned16(const T* p) {
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

template <typename T, bool ACC>
cudaError_t launch_mtv_t(T* c, const T* a, int64_t M, int64_t K, int64_t s_i, int64_t s_k, const T* b, void* ws,
                         size_t ws_bytes, int sm_count, cudaStream_t stream, int* launches, const char** name) {
    constexpr int V = Vec16<T>::N;
    *launches = 0;
    if (M <= 0) return cudaSuccess;
    if (K <= 0) {   // empty sum: assign 0 / accumulate nothing
        if (!ACC) {
            cudaError_t e = cudaMemsetAsync(c, 0, (size_t)M * sizeof(T), stream);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    if (s_k == 1 && s_i != 1) {
        // k-contiguous rows
        bool const vec = aligned16(a) && aligned16(b) && s_i % V == 0;
        bool const few_rows = M < (int64_t)sm_count * 8 * 2 && K >= 4096;
        int64_t const rows_per_block = few_rows ? 1 : MTV_THREADS / 32;
        int64_t blocks = (M + rows_per_block - 1) / rows_per_block;
        int64_t const cap = (int64_t)sm_count * 32;
        if (blocks > cap) blocks = cap;
        dim3 const g((unsigned)blocks), t(MTV_THREADS);
        if (vec) {
            if (few_rows) mtv_kcontig_kernel<T, V, 8, ACC><<<g, t, 0, stream>>>(c, a, s_i, b, (int)M, (int)K);
            else mtv_kcontig_kernel<T, V, 1, ACC><<<g, t, 0, stream>>>(c, a, s_i, b, (int)M, (int)K);
            *name = few_rows ? "mtv_kcontig_v16_cta_per_row" : "mtv_kcontig_v16_warp_per_row";
        } else {
            if (few_rows) mtv_kcontig_kernel<T, 1, 8, ACC><<<g, t, 0, stream>>>(c, a, s_i, b, (int)M, (int)K);
            else mtv_kcontig_kernel<T, 1, 1, ACC><<<g, t, 0, stream>>>(c, a, s_i, b, (int)M, (int)K);
            *name = few_rows ? "mtv_kcontig_scalar_cta_per_row" : "mtv_kcontig_scalar_warp_per_row";
        }
        *launches = 1;
        return cudaGetLastError();
    }
    // i-contiguous (or arbitrary strides through the scalar instantiation)
    bool const vec = s_i == 1 && aligned16(a) && s_k % V == 0 && M >= V;
    int const v = vec ? V : 1;
    int64_t const blocks_m = (M + (int64_t)MTV_THREADS * v - 1) / ((int64_t)MTV_THREADS * v);
    // enough CTAs to fill the machine ~8 times over, each chunk at least one k-tile long
    int64_t chunks = ((int64_t)sm_count * 8 + blocks_m - 1) / blocks_m;
    int64_t const max_chunks = (K + KTILE - 1) / KTILE;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks > 1 && (size_t)chunks * (size_t)M * sizeof(T) > ws_bytes) chunks = (int64_t)(ws_bytes / ((size_t)M * sizeof(T)));
    if (chunks < 1) chunks = 1;
    if (chunks > 65535) chunks = 65535;
    int k_per_block = (int)((K + chunks - 1) / chunks);
    k_per_block = (k_per_block + KTILE - 1) / KTILE * KTILE;
    chunks = (K + k_per_block - 1) / k_per_block;
    dim3 const g((unsigned)blocks_m, (unsigned)chunks), t(MTV_THREADS);
    T* partial = static_cast<T*>(ws);
    if (vec) mtv_icontig_kernel<T, V, ACC><<<g, t, 0, stream>>>(c, a, s_i, s_k, b, (int)M, (int)K, k_per_block, partial);
    else mtv_icontig_kernel<T, 1, ACC><<<g, t, 0, stream>>>(c, a, s_i, s_k, b, (int)M, (int)K, k_per_block, partial);
    *name = vec ? "mtv_icontig_v16" : "mtv_icontig_scalar";
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    *launches = 1;
    if (chunks > 1) {
        dim3 const gr((unsigned)((M + MTV_THREADS - 1) / MTV_THREADS));
        mtv_reduce_partials_kernel<T, ACC><<<gr, t, 0, stream>>>(c, partial, (int)M, (int)chunks);
        *launches = 2;
        e = cudaGetLastError();
    }
    return e;
}

}  // namespace

size_t mtv_workspace_bytes(int64_t M, int elem_size, int sm_count) {
    // room for up to 8 * sm_count chunk partials of length M (capped at 256 MiB)
    size_t const want = (size_t)M * (size_t)elem_size * (size_t)sm_count * 8;
    size_t const cap = (size_t)256 << 20;
    size_t const floor_ = (size_t)M * (size_t)elem_size * 2;
    return want > cap ? (cap > floor_ ? cap : floor_) : want;
}

cudaError_t launch_mtv_f32(float* c, const float* a, int64_t M, int64_t K, int64_t s_i, int64_t s_k, const float* b,
                           int accumulate, void* ws, size_t ws_bytes, int sm_count, cudaStream_t stream, int* launches,
                           const char** name) {
    return accumulate ? launch_mtv_t<float, true>(c, a, M, K, s_i, s_k, b, ws, ws_bytes, sm_count, stream, launches, name)
                      : launch_mtv_t<float, false>(c, a, M, K, s_i, s_k, b, ws, ws_bytes, sm_count, stream, launches, name);
}
cudaError_t launch_mtv_f64(double* c, const double* a, int64_t M, int64_t K, int64_t s_i, int64_t s_k, const double* b,
                           int accumulate, void* ws, size_t ws_bytes, int sm_count, cudaStream_t stream, int* launches,
                           const char** name) {
    return accumulate ? launch_mtv_t<double, true>(c, a, M, K, s_i, s_k, b, ws, ws_bytes, sm_count, stream, launches, name)
                      : launch_mtv_t<double, false>(c, a, M, K, s_i, s_k, b, ws, ws_bytes, sm_count, stream, launches, name);
}

}  // namespace b200
