/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU oracle for the matrix-times-matrix (mtm) hot path: a plain-C restatement of the
 * algorithm in amitsingh19975/OpenMP-BLAS (`amt::mtm_helper` and what it calls).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may
 * load this library, and only as the checker / the timed CPU baseline.  The product
 * (include/, openmp-blas_b200/) never links it and has no CPU fallback.
 *
 * Parity is PINNED, three ways (tests/test_oracle.py):
 *   1. the packed-panel known-answer vectors of the reference's test/test.pack.cpp:35-43, :74-82;
 *   2. the reference's own test/test.mtm.cpp cases (8 layout combos x {f32,f64} x sz 2..31,
 *      rand()%100 integer inputs) against an exact integer triple loop (stricter than the
 *      BLIS comparator the reference uses, which is not installed here);
 *   3. bit-for-bit against the reference itself, compiled unmodified into
 *      oracle/_ref/libref_mtm_*.so (oracle/ref_mtm.cpp), on random non-integer data with
 *      the reference's own block sizes.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference root).  Nothing here is copied: the reference is C++ templates over layout
 * tags; this is C over runtime strides.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct oracle_blocks {
    size_t mr, nr; /* register tile    (mtm.hpp:27-38, cpuinfo.hpp:201-223) */
    size_t kb;     /* K block, "KC"    (mtm.hpp:64-80)  */
    size_t mb;     /* M block, "MC"    (mtm.hpp:41-50)  */
    size_t nb;     /* N block, "NC"    (mtm.hpp:52-62)  */
} oracle_blocks;

/* Block sizes the reference derives for a Skylake-class core with 256-bit vectors
 * (doc/matrix_times_matrix.tex:290-302 for MR x NR; KB/MB/NB as probed on an 8-core
 * AVX-512 Xeon).  No result depends on them except the order in which
 * K-blocks are added into C (rounding only). */
void oracle_default_blocks(int is_f64, int c_last_order, oracle_blocks* out) {
    if (!is_f64) {
        out->mr = 16; out->nr = 6; out->kb = 512; out->mb = 448; out->nb = 67200;
    } else if (!c_last_order) {
        out->mr = 8; out->nr = 6; out->kb = 448; out->mb = 256; out->nb = 38400;
    } else { /* double + last_order swaps MR/NR: mtm.hpp:27-38 */
        out->mr = 6; out->nr = 8; out->kb = 426; out->mb = 276; out->nb = 40376;
    }
}

static size_t min_sz(size_t a, size_t b) { return a < b ? a : b; }

#define ORACLE_DEFINE(T, SFX, FMA)                                                              \
                                                                                                \
    /* amt::pack, utils.hpp:99-118 (trans == 0): out[i*wo + j] = in[j*wi[0] + i*wi[1]],         \
     * and amt::pack(..., tag::trans), utils.hpp:120-141: out[i*wo + j] = in[i*wi[0] + j*wi[1]];\
     * j < m (panel width), i < n (panel depth). */                                             \
    void oracle_pack_##SFX(T* out, size_t wo, const T* in, const size_t* wi, size_t m,          \
                           size_t n, int trans) {                                               \
        size_t const s_inner = trans ? wi[1] : wi[0];                                           \
        size_t const s_outer = trans ? wi[0] : wi[1];                                           \
        for (size_t i = 0; i < n; ++i)                                                          \
            for (size_t j = 0; j < m; ++j) out[i * wo + j] = in[i * s_outer + j * s_inner];     \
    }                                                                                           \
                                                                                                \
    /* simd_loop<MTM,MR,NR>::helper / helper_double_last_order (simd_loop.hpp:106-158) plus     \
     * copy_from_buff (simd_loop.hpp:160-190): a zero-initialised register tile receives one    \
     * rank-1 update per k (sequential in k, one fused multiply-add per element — the           \
     * reference is built with -ffast-math -march=native, which contracts), then is ADDED into  \
     * C.  Panels hold their actual widths mr/nr (edge panels are packed narrow,                \
     * mtm.hpp:170-199); the reference's lanes >= mr/nr are computed and discarded, so they     \
     * are simply not computed here. */                                                         \
    static void micro_##SFX(T* c, size_t wc0, size_t wc1, const T* ap, const T* bp, size_t kb,  \
                            size_t mr, size_t nr, T* buff, size_t ldbuff) {                     \
        for (size_t j = 0; j < nr; ++j)                                                         \
            for (size_t i = 0; i < mr; ++i) buff[j * ldbuff + i] = (T)0;                        \
        for (size_t k = 0; k < kb; ++k) {                                                       \
            const T* ak = ap + k * mr;                                                          \
            const T* bk = bp + k * nr;                                                          \
            for (size_t j = 0; j < nr; ++j)                                                     \
                for (size_t i = 0; i < mr; ++i)                                                 \
                    buff[j * ldbuff + i] = FMA(bk[j], ak[i], buff[j * ldbuff + i]);             \
        }                                                                                       \
        for (size_t j = 0; j < nr; ++j)                                                         \
            for (size_t i = 0; i < mr; ++i) c[i * wc0 + j * wc1] += buff[j * ldbuff + i];       \
    }                                                                                           \
                                                                                                \
    /* impl::mtm_kernel, mtm.hpp:83-110: NR-panels outer, MR-panels inner, over packed panels   \
     * (panel p of A starts at a + i*K, of B at b + j*K). */                                    \
    static void macro_##SFX(T* c, size_t wc0, size_t wc1, const T* a, const T* b, size_t M,     \
                            size_t N, size_t K, const oracle_blocks* blk, T* buff) {            \
        for (size_t j = 0; j < N; j += blk->nr) {                                               \
            size_t const jb = min_sz(blk->nr, N - j);                                           \
            for (size_t i = 0; i < M; i += blk->mr) {                                           \
                size_t const ib = min_sz(blk->mr, M - i);                                       \
                micro_##SFX(c + wc0 * i + wc1 * j, wc0, wc1, a + i * K, b + j * K, K, ib, jb,   \
                            buff, blk->mr);                                                     \
            }                                                                                   \
        }                                                                                       \
    }                                                                                           \
                                                                                                \
    /* amt::mtm_helper, mtm.hpp:116-206: for j (NB) / for k (KB) { pack B panels; for i (MB)    \
     * { pack A panels; macro-kernel } }, C += A*B.  The reference work-shares the pack-B loop  \
     * and the i loop across an OpenMP team (mtm.hpp:156,169,182); so does this.  Returns 0,    \
     * or 1 if C is not unit-stride in one dimension (the reference silently assumes it:        \
     * ldc = max(wc0,wc1), mtm.hpp:95), or 2 on allocation failure. */                          \
    int oracle_mtm_##SFX(T* c, const size_t* nc, const size_t* wc, const T* a,                  \
                         const size_t* na, const size_t* wa, const T* b, const size_t* nb,      \
                         const size_t* wb, const oracle_blocks* blk_in) {                       \
        oracle_blocks blk;                                                                      \
        size_t const M = na[0], K = na[1], N = nb[1];                                           \
        (void)nc;                                                                               \
        if (wc[0] != 1 && wc[1] != 1) return 1;                                                 \
        if (blk_in) blk = *blk_in;                                                              \
        else oracle_default_blocks(sizeof(T) == 8, wc[0] != 1, &blk);                           \
        if (M == 0 || N == 0 || K == 0) return 0;                                               \
        size_t const NB = min_sz(blk.nb, N + blk.nr), KB = min_sz(blk.kb, K);                   \
        size_t const MB = min_sz(blk.mb, M + blk.mr); /* clamps only shrink scratch buffers */  \
        int const nthreads = omp_get_max_threads();                                             \
        T* pB = (T*)malloc(sizeof(T) * KB * (NB + 1));                                          \
        T* pA = (T*)malloc(sizeof(T) * KB * (MB + 1) * (size_t)nthreads);                       \
        T* bf = (T*)malloc(sizeof(T) * blk.mr * blk.nr * (size_t)nthreads);                     \
        if (!pA || !pB || !bf) { free(pA); free(pB); free(bf); return 2; }                      \
        for (size_t j = 0; j < N; j += NB) {                                                    \
            size_t const jb = min_sz(NB, N - j);                                                \
            for (size_t k = 0; k < K; k += KB) {                                                \
                size_t const kb = min_sz(KB, K - k);                                            \
                const T* bi = b + wb[1] * j + wb[0] * k;                                        \
                const T* ai = a + wa[1] * k;                                                    \
                long const npanels = (long)((jb + blk.nr - 1) / blk.nr);                        \
                _Pragma("omp parallel for schedule(dynamic)")                                   \
                for (long p = 0; p < npanels; ++p) {                                            \
                    size_t const jj = (size_t)p * blk.nr;                                       \
                    size_t const jjb = min_sz(jb - jj, blk.nr);                                 \
                    oracle_pack_##SFX(pB + jj * kb, jjb, bi + jj * wb[1], wb, jjb, kb, 1);      \
                }                                                                               \
                long const nmblk = (long)((M + MB - 1) / MB);                                   \
                _Pragma("omp parallel for schedule(dynamic)")                                   \
                for (long q = 0; q < nmblk; ++q) {                                              \
                    size_t const i = (size_t)q * MB;                                            \
                    size_t const ib = min_sz(MB, M - i);                                        \
                    size_t const tid = (size_t)omp_get_thread_num();                            \
                    T* aptr = pA + tid * kb * MB;                                               \
                    for (size_t ii = 0; ii < ib; ii += blk.mr) {                                \
                        size_t const iib = min_sz(ib - ii, blk.mr);                             \
                        oracle_pack_##SFX(aptr + ii * kb, iib, ai + (i + ii) * wa[0], wa, iib,  \
                                          kb, 0);                                               \
                    }                                                                           \
                    macro_##SFX(c + wc[1] * j + wc[0] * i, wc[0], wc[1], aptr, pB, ib, jb, kb,  \
                                &blk, bf + tid * blk.mr * blk.nr);                              \
                }                                                                               \
            }                                                                                   \
        }                                                                                       \
        free(pA); free(pB); free(bf);                                                           \
        return 0;                                                                               \
    }

#ifndef _OPENMP
#error "build the oracle with -fopenmp (see oracle/Makefile)"
#endif

ORACLE_DEFINE(float, f32, fmaf)
ORACLE_DEFINE(double, f64, fma)

/* Exact comparator for integer-valued inputs (replaces the BLIS call of test/test.mtm.cpp:60-68,
 * alpha = beta = 1): C += A*B in 64-bit integers.  Inputs must be integer-valued and small
 * enough that nothing overflows (the tests use [0,99] like test/test_utils.hpp:4-9). */
void oracle_exact_i64(int64_t* c, const size_t* wc, const int64_t* a, const size_t* na,
                      const size_t* wa, const int64_t* b, const size_t* nb, const size_t* wb) {
    size_t const M = na[0], K = na[1], N = nb[1];
    for (size_t i = 0; i < M; ++i)
        for (size_t j = 0; j < N; ++j) {
            int64_t s = 0;
            for (size_t k = 0; k < K; ++k)
                s += a[i * wa[0] + k * wa[1]] * b[k * wb[0] + j * wb[1]];
            c[i * wc[0] + j * wc[1]] += s;
        }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
