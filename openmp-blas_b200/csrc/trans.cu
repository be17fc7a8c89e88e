// Matrix transpose kernels for sm_100a: c(j, i) = a(i, j) out of place, and the square in-place swap.
//
// Replaces the reference's amt::transpose_helper (include/trans.hpp:33-92) and its element loops
// (simd_loop<TRANS>, simd_loop.hpp:196-238).  Pure data movement, so the roofline is the measured
// HBM copy bandwidth (bytes = 2 * M * N * sizeof(T)).  Tiles of 64 x 64 elements go through padded
// shared memory so that both the global reads and the global writes run along whichever index has
// the smaller stride in a and in c respectively (any combination of first_order / last_order /
// strided views), each thread moving 16 elements per tile.
#include "mtm_kernels.h"

namespace b200 {
namespace {

constexpr int TT = 64;            // tile edge
constexpr int TROWS = 16;         // blockDim = (64, 16): each thread covers TT / TROWS = 4 rows of the tile

// a(i, j) = a[i * sa_i + j * sa_j], i < M, j < N;  c(j, i) = c[j * sc_j + i * sc_i].
// READ_J: the warp's x index runs along j when reading a (a is j-contiguous), else along i.
// WRITE_I: the warp's x index runs along i when writing c (c is i-contiguous), else along j.
template <typename T, bool READ_J, bool WRITE_I>
__global__ void __launch_bounds__(TT * TROWS)
transpose_kernel(T* __restrict__ c, const T* __restrict__ a, int M, int N, int64_t sa_i, int64_t sa_j,
                 int64_t sc_j, int64_t sc_i) {
    __shared__ T tile[TT][TT + 1];   // tile[i_local][j_local]
    int const tx = threadIdx.x, ty = threadIdx.y;
    int64_t const i0 = (int64_t)blockIdx.y * TT, j0 = (int64_t)blockIdx.x * TT;
#pragma unroll
    for (int r = 0; r < TT; r += TROWS) {
        int const il = READ_J ? ty + r : tx, jl = READ_J ? tx : ty + r;
        int64_t const i = i0 + il, j = j0 + jl;
        if (i < M && j < N) tile[il][jl] = a[i * sa_i + j * sa_j];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < TT; r += TROWS) {
        int const il = WRITE_I ? tx : ty + r, jl = WRITE_I ? ty + r : tx;
        int64_t const i = i0 + il, j = j0 + jl;
        if (i < M && j < N) c[j * sc_j + i * sc_i] = tile[il][jl];
    }
}

// 16-byte variant: both operands have a unit stride (along j or i for a, along i or j for c) and
// 16-byte aligned rows, so every global access moves a float4 / double2 (4-byte accesses cap the
// fp32 transpose at ~62% of the copy bandwidth: too many requests per byte).  256 threads per
// 64 x 64 tile; the shared-memory side stays scalar (padded rows, at most 2-way conflicts).
template <typename T>
struct TVec;
template <>
struct TVec<float> {
    using type = float4;
    static constexpr int N = 4;
};
template <>
struct TVec<double> {
    using type = double2;
    static constexpr int N = 2;
};

template <typename T, bool READ_J, bool WRITE_I>
__global__ void __launch_bounds__(256)
transpose_vec_kernel(T* __restrict__ c, const T* __restrict__ a, int M, int N, int64_t lda, int64_t ldc) {
    using VT = typename TVec<T>::type;
    constexpr int V = TVec<T>::N;
    constexpr int VPR = TT / V;                  // vectors per tile row
    constexpr int ITERS = TT * VPR / 256;
    __shared__ T tile[TT][TT + 1];               // tile[i_local][j_local]
    int64_t const i0 = (int64_t)blockIdx.y * TT, j0 = (int64_t)blockIdx.x * TT;
#pragma unroll
    for (int r = 0; r < ITERS; ++r) {
        int const idx = threadIdx.x + 256 * r;
        int const outer = idx / VPR, inner = (idx % VPR) * V;
        // READ_J: rows of a are i, vectors run along j.  Otherwise rows are j, vectors run along i.
        int64_t const i = READ_J ? i0 + outer : i0 + inner, j = READ_J ? j0 + inner : j0 + outer;
        const T* src = READ_J ? a + i * lda + j : a + j * lda + i;
        int const lim = READ_J ? N : M;
        int64_t const pos = READ_J ? j : i;
        if ((READ_J ? i < M : j < N) && pos < lim) {
            T e[V];
            if (pos + V <= lim) {
                *reinterpret_cast<VT*>(e) = *reinterpret_cast<const VT*>(src);
            } else {
#pragma unroll
                for (int q = 0; q < V; ++q) e[q] = pos + q < lim ? src[q] : T(0);
            }
#pragma unroll
            for (int q = 0; q < V; ++q) {
                if (READ_J) tile[outer][inner + q] = e[q];
                else tile[inner + q][outer] = e[q];
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ITERS; ++r) {
        int const idx = threadIdx.x + 256 * r;
        int const outer = idx / VPR, inner = (idx % VPR) * V;
        // WRITE_I: rows of c are j, vectors run along i.  Otherwise rows of c are i, vectors run along j.
        int64_t const i = WRITE_I ? i0 + inner : i0 + outer, j = WRITE_I ? j0 + outer : j0 + inner;
        T* dst = WRITE_I ? c + j * ldc + i : c + i * ldc + j;
        int const lim = WRITE_I ? M : N;
        int64_t const pos = WRITE_I ? i : j;
        if ((WRITE_I ? j < N : i < M) && pos < lim) {
            T e[V];
#pragma unroll
            for (int q = 0; q < V; ++q) e[q] = WRITE_I ? tile[inner + q][outer] : tile[outer][inner + q];
            if (pos + V <= lim) {
                *reinterpret_cast<VT*>(dst) = *reinterpret_cast<const VT*>(e);
            } else {
#pragma unroll
                for (int q = 0; q < V; ++q)
                    if (pos + q < lim) dst[q] = e[q];
            }
        }
    }
}

// In place, n x n, element (i, j) at a[i + j * n] (the reference builds exactly these strides,
// trans.hpp:163-165).  One CTA per tile pair (bi <= bj): both tiles are read, then written swapped.
constexpr int IT = 32, IROWS = 8;   // in-place: two 32 x 32 tiles per CTA

template <typename T>
__global__ void __launch_bounds__(IT * IROWS)
transpose_inplace_kernel(T* __restrict__ a, int n) {
    __shared__ T t0[IT][IT + 1], t1[IT][IT + 1];
    int const bi = blockIdx.y, bj = blockIdx.x;
    if (bi > bj) return;
    int const tx = threadIdx.x, ty = threadIdx.y;
    int64_t const i0 = (int64_t)bi * IT, j0 = (int64_t)bj * IT;
    // x runs along the contiguous index i of the stored element (i, j)
#pragma unroll
    for (int r = 0; r < IT; r += IROWS) {
        int64_t const i = i0 + tx, j = j0 + ty + r;        // tile (bi, bj)
        if (i < n && j < n) t0[tx][ty + r] = a[i + j * (int64_t)n];
        int64_t const p = j0 + tx, q = i0 + ty + r;        // tile (bj, bi): element (p, q)
        if (bi != bj && p < n && q < n) t1[tx][ty + r] = a[p + q * (int64_t)n];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < IT; r += IROWS) {
        // element (p, q) of tile (bj, bi) receives a(q, p), which is t0[q - i0][p - j0]
        int64_t const p = j0 + tx, q = i0 + ty + r;
        if (p < n && q < n) a[p + q * (int64_t)n] = t0[ty + r][tx];
        // element (i, j) of tile (bi, bj) receives a(j, i) = t1[j - j0][i - i0]
        int64_t const i = i0 + tx, j = j0 + ty + r;
        if (bi != bj && i < n && j < n) a[i + j * (int64_t)n] = t1[ty + r][tx];
    }
}

template <typename T>
cudaError_t launch_transpose_t(T* c, const T* a, int64_t M, int64_t N, int64_t sa_i, int64_t sa_j, int64_t sc_j,
                               int64_t sc_i, cudaStream_t stream) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    int64_t const gx = (N + TT - 1) / TT, gy = (M + TT - 1) / TT;
    if (gy > 65535) {
        // gridDim.y holds the row tiles: more than 65535 of them (M > ~4.19 M rows) go in bands of rows
        int64_t const band = 65535 * (int64_t)TT;
        for (int64_t i0 = 0; i0 < M; i0 += band) {
            cudaError_t e = launch_transpose_t<T>(c + i0 * sc_i, a + i0 * sa_i, (M - i0 < band ? M - i0 : band), N, sa_i, sa_j,
                                                  sc_j, sc_i, stream);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    dim3 const g((unsigned)gx, (unsigned)gy), b(TT, TROWS);
    bool const read_j = sa_j <= sa_i, write_i = sc_i <= sc_j;
    {   // 16-byte path: unit stride on both sides, aligned bases and pitches
        constexpr int V = TVec<T>::N;
        int64_t const lda = read_j ? sa_i : sa_j, ldc = write_i ? sc_j : sc_i;
        bool const unit = (read_j ? sa_j : sa_i) == 1 && (write_i ? sc_i : sc_j) == 1;
        bool const aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(c)) & 15u) == 0 &&
                             lda % V == 0 && ldc % V == 0;
        // measured (16384^2): fp32 5.8-6.1 TB/s with 16-byte accesses vs 4.0-4.3 with 4-byte ones; fp64 is
        // already at 5.7-6.2 TB/s with the 8-byte tile kernel (5.4-5.7 with double2), so only fp32 takes this path
        if (unit && aligned && sizeof(T) == 4) {
            dim3 const bv(256);
            if (read_j && write_i) transpose_vec_kernel<T, true, true><<<g, bv, 0, stream>>>(c, a, (int)M, (int)N, lda, ldc);
            else if (read_j) transpose_vec_kernel<T, true, false><<<g, bv, 0, stream>>>(c, a, (int)M, (int)N, lda, ldc);
            else if (write_i) transpose_vec_kernel<T, false, true><<<g, bv, 0, stream>>>(c, a, (int)M, (int)N, lda, ldc);
            else transpose_vec_kernel<T, false, false><<<g, bv, 0, stream>>>(c, a, (int)M, (int)N, lda, ldc);
            return cudaGetLastError();
        }
    }
    if (read_j && write_i) transpose_kernel<T, true, true><<<g, b, 0, stream>>>(c, a, (int)M, (int)N, sa_i, sa_j, sc_j, sc_i);
    else if (read_j) transpose_kernel<T, true, false><<<g, b, 0, stream>>>(c, a, (int)M, (int)N, sa_i, sa_j, sc_j, sc_i);
    else if (write_i) transpose_kernel<T, false, true><<<g, b, 0, stream>>>(c, a, (int)M, (int)N, sa_i, sa_j, sc_j, sc_i);
    else transpose_kernel<T, false, false><<<g, b, 0, stream>>>(c, a, (int)M, (int)N, sa_i, sa_j, sc_j, sc_i);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_transpose_inplace_t(T* a, int64_t n, cudaStream_t stream) {
    if (n <= 1) return cudaSuccess;
    int64_t const gt = (n + IT - 1) / IT;
    if (gt > 65535) return cudaErrorInvalidConfiguration;
    transpose_inplace_kernel<T><<<dim3((unsigned)gt, (unsigned)gt), dim3(IT, IROWS), 0, stream>>>(a, (int)n);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_transpose_f32(float* c, const float* a, int64_t M, int64_t N, int64_t sa_i, int64_t sa_j,
                                 int64_t sc_j, int64_t sc_i, cudaStream_t stream) {
    return launch_transpose_t<float>(c, a, M, N, sa_i, sa_j, sc_j, sc_i, stream);
}
cudaError_t launch_transpose_f64(double* c, const double* a, int64_t M, int64_t N, int64_t sa_i, int64_t sa_j,
                                 int64_t sc_j, int64_t sc_i, cudaStream_t stream) {
    return launch_transpose_t<double>(c, a, M, N, sa_i, sa_j, sc_j, sc_i, stream);
}
cudaError_t launch_transpose_inplace_f32(float* a, int64_t n, cudaStream_t stream) {
    return launch_transpose_inplace_t<float>(a, n, stream);
}
cudaError_t launch_transpose_inplace_f64(double* a, int64_t n, cudaStream_t stream) {
    return launch_transpose_inplace_t<double>(a, n, stream);
}

}  // namespace b200
