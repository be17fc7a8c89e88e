/* b200_trans.h — C ABI of the matrix-transpose path of libb200mtm.so (sm_100a).
 *
 * Replaces amt::transpose_helper(c, nc, wc, a, na, wa, tag::outplace / tag::inplace) —
 * include/trans.hpp:33-92 — behind amt::transpose(c, a, n) and amt::transpose(a, n) (trans.hpp:94-168).
 *
 *   out of place:  c(j, i) = a(i, j)   for i < na[0], j < na[1];  nc must be {na[1], na[0]};
 *                  a(i,j) = a[i*wa[0] + j*wa[1]], c(j,i) = c[j*wc[0] + i*wc[1]] — any strides on both
 *                  (every first_order / last_order pairing, sub-views).
 *   in place:      square na[0] == na[1], contiguous first_order storage (element (i,j) at a[i + j*n]),
 *                  the only form the reference supports (it builds these strides itself, trans.hpp:163-165).
 * Pure data movement: results are bit-identical to the reference's on any data.
 * Status codes / error string / devices: see b200_mtm.h.
 */
#ifndef B200_TRANS_H
#define B200_TRANS_H

#include "b200_mtm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Host pointers, synchronous. */
int b200_transpose_f32(float* c, const size_t nc[2], const size_t wc[2], const float* a, const size_t na[2],
                       const size_t wa[2], int flags);
int b200_transpose_f64(double* c, const size_t nc[2], const size_t wc[2], const double* a, const size_t na[2],
                       const size_t wa[2], int flags);
int b200_transpose_inplace_f32(float* a, const size_t na[2], int flags);
int b200_transpose_inplace_f64(double* a, const size_t na[2], int flags);

/* Device pointers, asynchronous on `stream`. */
int b200_transpose_f32_dev(float* c, const size_t nc[2], const size_t wc[2], const float* a, const size_t na[2],
                           const size_t wa[2], int flags, void* stream);
int b200_transpose_f64_dev(double* c, const size_t nc[2], const size_t wc[2], const double* a, const size_t na[2],
                           const size_t wa[2], int flags, void* stream);
int b200_transpose_inplace_f32_dev(float* a, const size_t na[2], int flags, void* stream);
int b200_transpose_inplace_f64_dev(double* a, const size_t na[2], int flags, void* stream);

/* Mean ms per out-of-place call over `iters` device calls after `warmup` (CUDA events on `stream`). */
int b200_transpose_bench_f32_dev(float* c, const size_t nc[2], const size_t wc[2], const float* a, const size_t na[2],
                                 const size_t wa[2], int flags, void* stream, int warmup, int iters, double* mean_ms);
int b200_transpose_bench_f64_dev(double* c, const size_t nc[2], const size_t wc[2], const double* a, const size_t na[2],
                                 const size_t wa[2], int flags, void* stream, int warmup, int iters, double* mean_ms);

#ifdef __cplusplus
}
#endif
#endif /* B200_TRANS_H */
