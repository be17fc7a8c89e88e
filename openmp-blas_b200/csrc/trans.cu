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
    if (gy > 65535) return cudaErrorInvalidConfiguration;
    dim3 const g((unsigned)gx, (unsigned)gy), b(TT, TROWS);
    bool const read_j = sa_j <= sa_i, write_i = sc_i <= sc_j;
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
