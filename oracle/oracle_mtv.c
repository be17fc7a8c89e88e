/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU oracle for matrix-times-vector / vector-times-matrix (amt::mtv, amt::vtm of
 * amitsingh19975/OpenMP-BLAS): plain-C restatement of `amt::mtv_helper` (include/mtv.hpp:15-100).
 * Same rules as oracle_mtm.c: only tests/, smoke() and bench.py's CPU legs may load it.
 *
 * Pinned (tests/test_oracle_mtv.py) against the reference's own cases — test/test.mtv.cpp and
 * test/test.vtm.cpp: first/last order x {f32,f64} x sizes 2..511, rand()%100 inputs, exact integer
 * comparator — and against the reference itself (oracle/_ref) on random data within rounding
 * (the reference vectorises its k-loops with `omp simd reduction`, so its summation order is the
 * compiler's, not pinned bit-wise).
 *
 * One semantic detail the reference's tests do not show but its code does, and which is kept:
 *   first_order A (column-major): c[i] += sum_k A(i,k) b[k]   — ACCUMULATES (simd_loop.hpp:58-75)
 *   last_order  A (row-major)   : c[i]  = sum_k A(i,k) b[k]   — ASSIGNS      (mtv.hpp:94-99)
 * vtm(c, A, v) is mtv on the flipped layout (mtv.hpp:176-236): c = v*A computed as A^T v, so a
 * first_order A takes the assigning path and a last_order A the accumulating one.
 */
#include <math.h>
#include <stddef.h>

#define ORACLE_MTV_DEFINE(T, SFX, FMA)                                                           \
    /* mtv_helper(c,nc,a,na,wa,b,nb,max_threads,layout): rows = na[0] (length of c), cols =     \
     * na[1] (length of b); A(i,k) = a[i*wa[0] + k*wa[1]].  kb is the K-block of the           \
     * first_order path (mtv.hpp:35: each block's products are added into c in k order).       \
     * a_last_order selects the path; the restatement honours both strides in both paths.      */\
    void oracle_mtv_##SFX(T* c, const T* a, const size_t* na, const size_t* wa, const T* b,     \
                          int a_last_order, size_t kb) {                                        \
        size_t const M = na[0], K = na[1];                                                      \
        if (kb == 0) kb = K ? K : 1;                                                            \
        if (a_last_order) {                                                                     \
            for (size_t i = 0; i < M; ++i) { /* simd_loop<INNER>, simd_loop.hpp:21-35 */        \
                T sum = (T)0;                                                                   \
                for (size_t k = 0; k < K; ++k) sum = FMA(a[i * wa[0] + k * wa[1]], b[k], sum);  \
                c[i] = sum;                                                                     \
            }                                                                                   \
        } else {                                                                                \
            for (size_t k0 = 0; k0 < K; k0 += kb) { /* mtv.hpp:45-69 */                         \
                size_t const k1 = k0 + kb < K ? k0 + kb : K;                                    \
                for (size_t i = 0; i < M; ++i) {                                                \
                    T acc = c[i];                                                               \
                    for (size_t k = k0; k < k1; ++k) acc = FMA(a[i * wa[0] + k * wa[1]], b[k], acc); \
                    c[i] = acc;                                                                 \
                }                                                                               \
            }                                                                                   \
        }                                                                                       \
    }                                                                                           \
    /* amt::vtm(c, a, b): c (length na[1]) from b (length na[0]); mtv on the transposed view    \
     * with the OTHER layout's path (mtv.hpp:206-236).                                          */\
    void oracle_vtm_##SFX(T* c, const T* a, const size_t* na, const size_t* wa, const T* b,     \
                          int a_last_order, size_t kb) {                                        \
        size_t const nt[2] = {na[1], na[0]};                                                    \
        size_t const wt[2] = {wa[1], wa[0]};                                                    \
        oracle_mtv_##SFX(c, a, nt, wt, b, !a_last_order, kb);                                   \
    }

ORACLE_MTV_DEFINE(float, f32, fmaf)
ORACLE_MTV_DEFINE(double, f64, fma)
