/* b200_mtm.h — C ABI of libb200mtm.so: the B200 (sm_100a) matrix-times-matrix path.
 *
 * This is the drop-in boundary for the mtm hot path of amitsingh19975/OpenMP-BLAS.  Each entry
 * point names the reference interface it replaces (paths relative to the reference root).
 * Signatures carry only plain pointers, sizes and ints: no C++ types, no torch types.
 *
 * Semantics (identical to the reference, include/mtm.hpp:116-206, include/simd_loop.hpp:160-190):
 *   C += A * B            accumulate; no alpha/beta, no transpose flags;
 *   n?[2] = extents {rows, cols}; w?[2] = element strides {row stride, col stride}
 *           (uBLAS first_order / column-major = {1, rows}; last_order / row-major = {cols, 1});
 *   A and B may have any two strides (sub-views included); C must be unit-stride in one
 *   dimension (the reference assumes it: ldc = max(wc[0], wc[1]), mtm.hpp:95);
 *   "transposition" is expressed purely through strides (test/test.mtm.cpp tag order is C,A,B).
 *
 * Every function returns B200_OK (0) or a B200_ERR_* code; b200_last_error() returns a
 * thread-local message for the last failure.  There is NO CPU fallback: without a CUDA device
 * every compute entry fails with B200_ERR_CUDA.
 *
 * Threading: one caller at a time per device (the reference is not re-entrant either: its pack
 * buffers are function-local statics, mtm.hpp:147-151).  Device workspaces are cached inside
 * the library (grow-only) and released by b200_shutdown().
 */
#ifndef B200_MTM_H
#define B200_MTM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ---------------------------------------------------------------------- */
enum {
    B200_OK = 0,
    B200_ERR_INVALID = 1, /* null pointer / bad flag / extent >= 2^31                          */
    B200_ERR_DIM = 2,     /* na[0]!=nc[0] || na[1]!=nb[0] || nc[1]!=nb[1]  (mtm.hpp:243-250)   */
    B200_ERR_LAYOUT = 3,  /* C has no unit stride                                               */
    B200_ERR_CUDA = 4,    /* CUDA runtime/driver failure, or no sm_100 device                   */
    B200_ERR_NOMEM = 5    /* device or pinned allocation failed                                 */
};

/* ---- `flags` argument: kernel family (low byte) | tile config + 1 (second byte, 0 = auto) - */
enum {
    B200_MTM_AUTO = 0,   /* library picks: fp32 -> 3xTF32 for large problems else SIMT; fp64 -> best of DFMA/DMMA */
    B200_MTM_SIMT = 1,   /* CUDA-core FFMA (fp32) / DFMA (fp64)                                 */
    B200_MTM_3XTF32 = 2, /* fp32 only: tcgen05 kind::tf32 x3 (hi*hi + hi*lo + lo*hi), TMEM accumulators */
    B200_MTM_DFMA = 3,   /* fp64 only: alias of SIMT                                            */
    B200_MTM_DMMA = 4    /* fp64 only: mma.sync m8n8k4 tensor-core path                         */
};
#define B200_MTM_FLAGS(variant, config_plus_1) ((int)(variant) | ((int)(config_plus_1) << 8))
/* Third byte: number of SMs the persistent (3xTF32) kernel leaves free, e.g. for a concurrent NCCL
 * broadcast in the row-block sharded driver; 0 = use every SM.                                  */
#define B200_MTM_RESERVE_SMS(n) (((int)(n) & 0xff) << 16)
/* Fourth byte (3xTF32 family): number of K splits per output tile, 0 = automatic (only problems with too few
 * tiles for the 148 SMs are split; the splits of a tile add into C in a fixed order).                 */
#define B200_MTM_SPLIT_K(n) (((int)(n) & 0x7f) << 24)

/* ---- the hot path ------------------------------------------------------------------------
 * Replaces amt::mtm_helper(c,nc,wc,a,na,wa,b,nb,wb,OutLayout) — include/mtm.hpp:116-122 — which
 * is what the callable returned by amt::mtm (mtm.hpp:262-266) invokes.  C's layout tag does not
 * cross the ABI: it is implied by wc (wc[0]==1 -> first_order).
 *
 * Host-pointer form: synchronous.  Stages the operands into device memory, runs the kernel,
 * copies C back, returns when C is valid on the host.  This is the form the reference's tests
 * and harness reach through include/mtm.hpp.  Pinned host memory (b200_host_alloc) is copied
 * asynchronously and overlapped; pageable memory works but copies slower.                     */
int b200_mtm_f32(float* c, const size_t nc[2], const size_t wc[2],
                 const float* a, const size_t na[2], const size_t wa[2],
                 const float* b, const size_t nb[2], const size_t wb[2], int flags);
int b200_mtm_f64(double* c, const size_t nc[2], const size_t wc[2],
                 const double* a, const size_t na[2], const size_t wa[2],
                 const double* b, const size_t nb[2], const size_t wb[2], int flags);

/* Device-pointer form: asynchronous on `stream` (a cudaStream_t / CUstream; NULL = legacy
 * default stream) of the current device.  Same semantics; operands already resident in HBM.
 * Used by the timed harness and the multi-GPU driver.                                          */
int b200_mtm_f32_dev(float* c, const size_t nc[2], const size_t wc[2],
                     const float* a, const size_t na[2], const size_t wa[2],
                     const float* b, const size_t nb[2], const size_t wb[2], int flags, void* stream);
int b200_mtm_f64_dev(double* c, const size_t nc[2], const size_t wc[2],
                     const double* a, const size_t na[2], const size_t wa[2],
                     const double* b, const size_t nb[2], const size_t wb[2], int flags, void* stream);

/* Multi-GPU host-pointer form: ONE call spread over the GPUs of the box, as the reference spreads one call
 * over all cores (OpenMP team over the M-blocks, include/mtm.hpp:156-201; thread_utils.hpp).  Same semantics
 * and host operands as b200_mtm_f32; C's slow dimension is cut into one shard per device (whole multiples of
 * 256 rows: the row-block partition), a host thread per device drives that device's copy / compute / copy-back
 * pipeline, and the operand every shard needs whole is uploaded ONCE: device d fetches the d-th 1/P of it over
 * its own PCIe link and forwards that slice to its peers over NVLink (peer copies ordered by cross-device
 * events).  K is never split, so every element of C comes out of the single-GPU kernel and summation order.
 * `devices`: n_devices CUDA device indices, or NULL for the first n_devices visible ones (n_devices <= 0: all).
 * Problems too small to shard (or operands that cannot be sliced: C and the sliced operand need a unit stride)
 * run on devices[0] through b200_mtm_f32.  Synchronous; no torch, no NCCL, one process.                     */
int b200_mtm_f32_mgpu(float* c, const size_t nc[2], const size_t wc[2],
                      const float* a, const size_t na[2], const size_t wa[2],
                      const float* b, const size_t nb[2], const size_t wb[2], int flags,
                      const int* devices, int n_devices);
int b200_mtm_f64_mgpu(double* c, const size_t nc[2], const size_t wc[2],
                      const double* a, const size_t na[2], const size_t wa[2],
                      const double* b, const size_t nb[2], const size_t wb[2], int flags,
                      const int* devices, int n_devices);

/* Which kernel the last *_dev / host call on this thread resolved to.                          */
typedef struct b200_mtm_choice {
    int variant;          /* B200_MTM_SIMT / _3XTF32 / _DMMA                                   */
    int config;           /* tile-config index within the variant                              */
    int launches;         /* kernels launched by that call                                     */
    int a_mode, b_mode;   /* 0 = 16-byte loads along m/n, 1 = along k, 2 = scalar (strided)     */
    char name[64];        /* e.g. "ffma_128x128x8_t8x8"                                        */
} b200_mtm_choice;
int b200_mtm_last_choice(b200_mtm_choice* out);

/* What AUTO resolves an M x N x K fp32 problem (row-major, aligned operands) to on a device with `sm_count` SMs
 * (<= 0: the current device): kernel family and tile config, in out->variant / config / name (the K splits a launch
 * may add are decided at launch and show in b200_mtm_last_choice).  The reference's counterpart is the block-size
 * choice of matrix_partition (include/mtm.hpp:19-81).  Needs no GPU when sm_count is given.                       */
int b200_mtm_plan_f32(size_t M, size_t N, size_t K, int sm_count, b200_mtm_choice* out);

/* Number of tile configs of a variant for a dtype (is_f64 = 0/1), and their names.            */
int b200_mtm_num_configs(int variant, int is_f64);
const char* b200_mtm_config_name(int variant, int is_f64, int config);

/* Total kernels launched by this process through the library (monotonic).                      */
uint64_t b200_launch_count(void);

/* ---- machine model -----------------------------------------------------------------------
 * Replaces cache_manager / cpu_info (include/cache_manager.hpp:183-223, include/cpuinfo.hpp)
 * and the hard-coded peak of metric.hpp:30: the numbers the harness needs for % of peak.      */
typedef struct b200_device_info {
    char name[128];
    int cc_major, cc_minor;
    int sm_count;
    int sm_clock_khz;           /* cudaDevAttrClockRate (max SM clock)                         */
    int mem_clock_khz;
    int mem_bus_bits;
    size_t smem_per_sm, smem_per_block_optin, l2_bytes, hbm_bytes;
    double peak_fp32_tflops;    /* sm_count * 128 FFMA lanes * 2 * clock                       */
    double peak_fp64_tflops;    /* sm_count *  64 DFMA lanes * 2 * clock                       */
} b200_device_info;
int b200_device_count(int* count);
int b200_set_device(int device);
int b200_get_device_info(int device, b200_device_info* out);

/* ---- resources ---------------------------------------------------------------------------
 * Replace aligned_buff (include/aligned_buff.hpp:12-61) and threads (include/thread_utils.hpp):
 * device / pinned allocations and the stream layer.                                            */
int b200_malloc(void** dptr, size_t bytes);
int b200_free(void* dptr);
int b200_host_alloc(void** hptr, size_t bytes);   /* pinned */
int b200_host_free(void* hptr);
int b200_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream);
int b200_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream);
int b200_memset(void* dptr, int value, size_t bytes, void* stream);
int b200_stream_create(void** stream);
int b200_stream_destroy(void* stream);
int b200_stream_synchronize(void* stream);
int b200_device_synchronize(void);

/* ---- device timing -----------------------------------------------------------------------
 * Replaces amt::benchmark<MaxIter> (include/benchmark.hpp:34-52) for device-resident operands:
 * `warmup` untimed calls, then `iters` back-to-back calls bracketed by CUDA events on `stream`.
 * *mean_ms receives the mean per-call time.  C keeps accumulating, as in src/mtm.cpp:207-208. */
int b200_mtm_bench_f32_dev(float* c, const size_t nc[2], const size_t wc[2],
                           const float* a, const size_t na[2], const size_t wa[2],
                           const float* b, const size_t nb[2], const size_t wb[2], int flags,
                           void* stream, int warmup, int iters, double* mean_ms);
int b200_mtm_bench_f64_dev(double* c, const size_t nc[2], const size_t wc[2],
                           const double* a, const size_t na[2], const size_t wa[2],
                           const double* b, const size_t nb[2], const size_t wb[2], int flags,
                           void* stream, int warmup, int iters, double* mean_ms);

/* Host microseconds the last b200_mtm_bench_*_dev call on this thread spent ENQUEUEING one call (validation, kernel
 * choice, tensor maps, launches): a loop whose device time per call is close to this figure is launch-bound. */
double b200_last_bench_enqueue_us(void);

const char* b200_last_error(void);
int b200_shutdown(void);   /* frees cached workspaces and internal streams                      */

#ifdef __cplusplus
}
#endif
#endif /* B200_MTM_H */
