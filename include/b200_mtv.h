/* b200_mtv.h — C ABI of the matrix-times-vector path of libb200mtm.so (sm_100a).
 *
 * Replaces amt::mtv_helper(c, nc, a, na, wa, b, nb, max_threads, layout) — include/mtv.hpp:15-100 —
 * the routine behind both amt::mtv (mtv.hpp:102-168) and amt::vtm (mtv.hpp:170-236; vtm is mtv on
 * the flipped extents/strides with the other layout's path, done by the host layer exactly as the
 * reference does).
 *
 *   c[i] (op)= sum_k A(i,k) * b[k],   A(i,k) = a[i*wa[0] + k*wa[1]],  i < na[0], k < na[1]
 *   b: na[1] contiguous elements, c: na[0] contiguous elements.
 *   a_last_order = 0  the reference's first_order path:  c += A b   (ACCUMULATES, simd_loop.hpp:58-75)
 *   a_last_order = 1  the reference's last_order path:   c  = A b   (ASSIGNS,     mtv.hpp:94-99)
 * The tag only selects accumulate/assign; both strides of A are honoured in either case
 * (16-byte loads when one of them is 1 and the pointer/pitch are aligned).
 *
 * Status codes, error string, device selection and shutdown are those of b200_mtm.h.
 */
#ifndef B200_MTV_H
#define B200_MTV_H

#include "b200_mtm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Host pointers, synchronous (stages A, b, c; copies c back). */
int b200_mtv_f32(float* c, const float* a, const size_t na[2], const size_t wa[2], const float* b,
                 int a_last_order, int flags);
int b200_mtv_f64(double* c, const double* a, const size_t na[2], const size_t wa[2], const double* b,
                 int a_last_order, int flags);

/* Device pointers, asynchronous on `stream`. */
int b200_mtv_f32_dev(float* c, const float* a, const size_t na[2], const size_t wa[2], const float* b,
                     int a_last_order, int flags, void* stream);
int b200_mtv_f64_dev(double* c, const double* a, const size_t na[2], const size_t wa[2], const double* b,
                     int a_last_order, int flags, void* stream);

/* Mean ms per call over `iters` back-to-back device calls after `warmup` (CUDA events on `stream`). */
int b200_mtv_bench_f32_dev(float* c, const float* a, const size_t na[2], const size_t wa[2], const float* b,
                           int a_last_order, int flags, void* stream, int warmup, int iters, double* mean_ms);
int b200_mtv_bench_f64_dev(double* c, const double* a, const size_t na[2], const size_t wa[2], const double* b,
                           int a_last_order, int flags, void* stream, int warmup, int iters, double* mean_ms);

#ifdef __cplusplus
}
#endif
#endif /* B200_MTV_H */
