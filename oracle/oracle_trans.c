/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU oracle for amt::transpose (amitsingh19975/OpenMP-BLAS, include/trans.hpp): plain-C restatement
 * of transpose_helper (out-of-place trans.hpp:33-61, in-place trans.hpp:63-92) and its element loops
 * (simd_loop<TRANS>, simd_loop.hpp:196-238).  Pure data movement: parity is bit-exact on any data.
 * Pinned by tests/test_oracle_trans.py against the reference's cases (test/test.trans.cpp: sizes
 * 2..31, in- and out-of-place) and against the reference itself (oracle/_ref).
 */
#include <stddef.h>

#define ORACLE_TRANS_DEFINE(T, SFX)                                                               \
    /* out-of-place: c(j,i) = a(i,j);  a is na[0] x na[1] with strides wa, c has strides wc      \
     * (c[i*wc[1] + j*wc[0]] = a[i*wa[0] + j*wa[1]], simd_loop.hpp:222-236).  The reference's    \
     * cache blocking (trans.hpp:47-58) only reorders independent copies.                        */\
    void oracle_transpose_##SFX(T* c, const size_t* wc, const T* a, const size_t* na, const size_t* wa) { \
        for (size_t i = 0; i < na[0]; ++i)                                                        \
            for (size_t j = 0; j < na[1]; ++j) c[i * wc[1] + j * wc[0]] = a[i * wa[0] + j * wa[1]]; \
    }                                                                                             \
    /* in-place (trans.hpp:147-168 builds wc = {1, na[0]}, wa = {1, na[1]} and swaps the upper    \
     * triangle with the lower, simd_loop.hpp:197-219).  Meaningful for square matrices, which is \
     * all the reference tests; restated for the square case.                                     */\
    void oracle_transpose_inplace_##SFX(T* a, size_t n) {                                         \
        for (size_t i = 0; i < n; ++i)                                                            \
            for (size_t j = i + 1; j < n; ++j) {                                                  \
                T const t = a[j + i * n];                                                         \
                a[j + i * n] = a[i + j * n];                                                      \
                a[i + j * n] = t;                                                                 \
            }                                                                                     \
    }

ORACLE_TRANS_DEFINE(float, f32)
ORACLE_TRANS_DEFINE(double, f64)
