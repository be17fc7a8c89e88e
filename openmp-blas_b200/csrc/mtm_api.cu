// C-ABI layer of libb200mtm.so (include/b200_mtm.h): validation, canonicalisation of the
// reference's (extents, strides) triples, kernel selection, the CUDA stream / memory layer and
// the host-pointer staging path.  Replaces the reference's amt::mtm_helper entry
// (include/mtm.hpp:116-206) and its resource layer (cache_manager / threads / aligned_buff).
//
// No CPU fallback exists: every compute entry needs a CUDA device and fails loudly without one.
#include "../../include/b200_mtm.h"
#include "../../include/b200_mtv.h"
#include "../../include/b200_trans.h"
#include "../../include/b200_replicate.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: named ranges around the entry points (tracing)

#include "mtm_kernels.h"

namespace b200 {
namespace {

// RAII NVTX range: shows the C-ABI calls (and their shards / slabs) on an Nsight timeline; free when no tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

thread_local std::string g_err;
thread_local b200_mtm_choice g_choice{};
thread_local double g_last_enqueue_us = 0.0;   // b200_mtm_bench_*: host microseconds per enqueued call
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            int code__ = (e__ == cudaErrorMemoryAllocation) ? B200_ERR_NOMEM : B200_ERR_CUDA; \
            return fail(code__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),     \
                        __FILE__, __LINE__);                                                 \
        }                                                                                    \
    } while (0)

// ---- per-device context ---------------------------------------------------------------------
constexpr int kMaxDevices = 32;
constexpr int64_t kGridYLimit = 65535;   // CUDA's gridDim.y / .z maximum
constexpr int kMaxSlabs = 8;   // host-pointer entry: slabs along C's slow dimension for copy/compute overlap

struct Buffer {
    void* ptr = nullptr;
    size_t bytes = 0;
    // Workspaces of the asynchronous *_dev entries are shared by every stream of the device: each use is
    // ordered after the previous one (an event recorded behind the last user, waited on by a user on another
    // stream), and they grow stream-ordered (cudaMallocAsync / cudaFreeAsync) — no device-wide synchronisation.
    cudaEvent_t last_use = nullptr;
    cudaStream_t last_stream = nullptr;
    bool used = false;
    bool pooled = false;                 // allocated with cudaMallocAsync
};

struct DeviceCtx {
    bool ready = false;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    cudaStream_t host_stream = nullptr;  // stream of the synchronous host-pointer entry
    cudaStream_t copy_stream = nullptr;  // second stream for copy/compute overlap
    cudaStream_t out_stream = nullptr;   // third stream: device-to-host copies of finished slabs
    cudaEvent_t ev_in[kMaxSlabs + 2] = {};   // slab i has landed on the device (+2: the tapered tail)
    cudaEvent_t ev_done[kMaxSlabs + 2] = {}; // slab i has been multiplied
    Buffer stage[3];                     // device images of A, B, C for host-pointer calls
    // Pageable host operands: pinned bounce buffers the host threads fill / drain with memcpy while the copy
    // engines move the previous ones (stage_copy_pageable)
    static constexpr int kBounce = 4;
    static constexpr size_t kBounceBytes = (size_t)16 << 20;
    struct Bounce {
        void* ptr = nullptr;
        cudaEvent_t ev = nullptr;
        bool busy = false;                   // ev guards the last DMA that used ptr
        // a device-to-host chunk waiting to be copied on to the caller's memory once ev has fired
        char* out_dst = nullptr;
        size_t out_pitch = 0, out_width = 0, out_rows = 0;
    };
    Bounce up[kBounce], down[kBounce];
    int up_next = 0, down_next = 0;
    cudaEvent_t ev_sent[kMaxDevices] = {};   // multi-GPU host entry: my slice of the shared operand has reached device e
    cudaEvent_t ev_slice = nullptr;          // ... my slice has been uploaded
    bool peer_on[kMaxDevices] = {};          // peer access to device e enabled from this device
    Buffer tf32_ws;                      // hi/lo operand planes of the 3xTF32 path
    Buffer pack_ws;                      // mn-contiguous operand planes of the TMA-fed FFMA path
    Buffer mtv_ws;                       // chunk partials of the matrix-times-vector kernels
};

DeviceCtx g_ctx[kMaxDevices];
std::mutex g_mu;

int current_ctx(DeviceCtx** out) {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return fail(B200_ERR_CUDA, "no usable CUDA device (%s); libb200mtm has no CPU fallback",
                    cudaGetErrorString(e));
    if (dev < 0 || dev >= kMaxDevices) return fail(B200_ERR_CUDA, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceCtx& c = g_ctx[dev];
    if (!c.ready) {
        cudaDeviceProp p;
        CUDA_TRY(cudaGetDeviceProperties(&p, dev));
        if (p.major != 10)
            return fail(B200_ERR_CUDA,
                        "device %d (%s) is sm_%d%d; libb200mtm is built for sm_100a only", dev, p.name,
                        p.major, p.minor);
        c.sm_count = p.multiProcessorCount;
        c.cc_major = p.major;
        c.cc_minor = p.minor;
        CUDA_TRY(cudaStreamCreateWithFlags(&c.host_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c.out_stream, cudaStreamNonBlocking));
        for (auto& ev : c.ev_in) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto& ev : c.ev_done) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&c.ev_slice, cudaEventDisableTiming));
        // Kernels that may be launched while a flag-wait kernel spins (multi-GPU drivers) must never trigger a
        // lazy module load then: load them all now.
        {   // keep freed workspace memory in the stream-ordered pool instead of returning it at every sync
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t keep = ~0ull;
                (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            (void)cudaGetLastError();
        }
        CUDA_TRY(tf32_preload_kernels());
        CUDA_TRY(replicate_preload_kernels());
        c.ready = true;
    }
    *out = &c;
    return B200_OK;
}

// Synchronous grow-only buffer (host-pointer entries: the previous call has completed when this runs).
int ensure(Buffer& b, size_t bytes) {
    if (b.bytes >= bytes) return B200_OK;
    if (b.ptr) {
        CUDA_TRY(cudaFree(b.ptr));       // synchronises the device
        b.ptr = nullptr;
        b.bytes = 0;
    }
    size_t const want = bytes + bytes / 8;  // grow-only with a little slack
    cudaError_t e = cudaMalloc(&b.ptr, want);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        e = cudaMalloc(&b.ptr, bytes);
        if (e != cudaSuccess)
            return fail(B200_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        b.bytes = bytes;
        return B200_OK;
    }
    b.bytes = want;
    return B200_OK;
}

// Stream-ordered workspace of an asynchronous entry: make `st` wait for the previous user on another
// stream, grow without stalling the host or the device.  Pair with release() after the kernels are enqueued.
int acquire(Buffer& b, size_t bytes, cudaStream_t st) {
    if (!b.last_use) CUDA_TRY(cudaEventCreateWithFlags(&b.last_use, cudaEventDisableTiming));
    if (b.used && b.last_stream != st) CUDA_TRY(cudaStreamWaitEvent(st, b.last_use, 0));
    if (b.bytes >= bytes) return B200_OK;
    if (b.ptr) {
        if (b.pooled) CUDA_TRY(cudaFreeAsync(b.ptr, st));     // after everything enqueued on st (incl. the wait above)
        else CUDA_TRY(cudaFree(b.ptr));
        b.ptr = nullptr;
        b.bytes = 0;
    }
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMallocAsync(&b.ptr, want, st);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        want = bytes;
        e = cudaMallocAsync(&b.ptr, want, st);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            b.ptr = nullptr;
            return fail(B200_ERR_NOMEM, "cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        }
    }
    b.bytes = want;
    b.pooled = true;
    return B200_OK;
}
int release(Buffer& b, cudaStream_t st) {
    CUDA_TRY(cudaEventRecord(b.last_use, st));
    b.last_stream = st;
    b.used = true;
    return B200_OK;
}

// ---- problem canonicalisation -----------------------------------------------------------------
template <typename T>
struct Canon {
    MtmShape s;
    T* c;
    const T* a;
    const T* b;
};

int validate(const void* c, const size_t* nc, const size_t* wc, const void* a, const size_t* na,
             const size_t* wa, const void* b, const size_t* nb, const size_t* wb) {
    if (!nc || !wc || !na || !wa || !nb || !wb)
        return fail(B200_ERR_INVALID, "b200_mtm: null extents/strides pointer");
    // Same check, same order, as the reference front-end (include/mtm.hpp:243-250).
    if (!(na[0] == nc[0] && na[1] == nb[0] && nc[1] == nb[1]))
        return fail(B200_ERR_DIM,
                    "b200_mtm: dimension mismatch: C %zux%zu, A %zux%zu, B %zux%zu", nc[0], nc[1],
                    na[0], na[1], nb[0], nb[1]);
    size_t const lim = (size_t)1 << 31;
    if (nc[0] >= lim || nc[1] >= lim || na[1] >= lim)
        return fail(B200_ERR_INVALID, "b200_mtm: extents must be < 2^31");
    if (nc[0] == 0 || nc[1] == 0) return B200_OK;
    if (!c || (na[1] != 0 && (!a || !b))) return fail(B200_ERR_INVALID, "b200_mtm: null data pointer");
    if (wc[0] != 1 && wc[1] != 1 && nc[0] > 1 && nc[1] > 1)
        return fail(B200_ERR_LAYOUT,
                    "b200_mtm: C must be unit-stride in one dimension (strides {%zu,%zu}); "
                    "the reference assumes ldc = max(wc0,wc1) (mtm.hpp:95)", wc[0], wc[1]);
    return B200_OK;
}

template <typename T>
Canon<T> canonicalise(T* c, const size_t* nc, const size_t* wc, const T* a, const size_t* na,
                      const size_t* wa, const T* b, const size_t* nb, const size_t* wb) {
    Canon<T> r;
    (void)nb;
    // Row-contiguous C (last_order), or a single column: use the problem as given.
    bool const row_contig = (wc[1] == 1) || nc[1] == 1;
    bool const col_contig = (wc[0] == 1) || nc[0] == 1;
    if (row_contig || !col_contig) {
        r.s.M = (int64_t)nc[0];
        r.s.N = (int64_t)nc[1];
        r.s.K = (int64_t)na[1];
        r.s.a_sm = (int64_t)wa[0];
        r.s.a_sk = (int64_t)wa[1];
        r.s.b_sk = (int64_t)wb[0];
        r.s.b_sn = (int64_t)wb[1];
        r.s.ldc = (int64_t)wc[0];
        r.c = c;
        r.a = a;
        r.b = b;
    } else {
        // Column-contiguous C (first_order): C^T += B^T * A^T, i.e. swap the operands.
        r.s.M = (int64_t)nc[1];
        r.s.N = (int64_t)nc[0];
        r.s.K = (int64_t)na[1];
        r.s.a_sm = (int64_t)wb[1];  // A'(m', k) = B(k, m')
        r.s.a_sk = (int64_t)wb[0];
        r.s.b_sk = (int64_t)wa[1];  // B'(k, n') = A(n', k)
        r.s.b_sn = (int64_t)wa[0];
        r.s.ldc = (int64_t)wc[1];
        r.c = c;
        r.a = b;
        r.b = a;
    }
    return r;
}

template <typename T>
int load_mode(const T* p, int64_t stride_mn, int64_t stride_k) {
    constexpr int V = 16 / (int)sizeof(T);
    bool const aligned = (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
    if (aligned && stride_mn == 1 && stride_k % V == 0) return LOAD_MN_VEC;
    if (aligned && stride_k == 1 && stride_mn % V == 0) return LOAD_K_VEC;
    return LOAD_GENERIC;
}

// ---- kernel selection ---------------------------------------------------------------------------
// Relative per-SM speed of each tile config when the grid is full (measured on B200, see
// profiles/ and DESIGN.md); combined with wave quantisation and tile fill to pick a config.
// fp32 8192^3 LLL: 39.6 / 32.3 / 49.3 / 46.0 / 48.1 TFLOP/s; fp64: DFMA 21.9 / 17.3 / 20.7 / 22.8,
// DMMA 27.3 / 32.4 / 31.6 / 31.9 (profiles/r01_tune_8192.json).
const double kSimtF32Speed[] = {0.80, 0.66, 1.00, 0.93, 0.975};
const double kSimtF64Speed[] = {0.96, 0.76, 0.91, 1.00};
const double kDmmaF64Speed[] = {0.84, 1.00, 0.975, 0.985, 1.00};

int pick_config(const MtmShape& s, int sm_count, int ncfg, const TileConfig& (*get)(int),
                const double* speed) {
    int best = 0;
    double best_score = -1.0;
    for (int i = 0; i < ncfg; ++i) {
        const TileConfig& t = get(i);
        double const tm = (double)((s.M + t.bm - 1) / t.bm), tn = (double)((s.N + t.bn - 1) / t.bn);
        double const tiles = tm * tn;
        double const slots = (double)sm_count * t.min_blocks;
        double const waves = (double)(int64_t)((tiles + slots - 1) / slots);
        double const wave_eff = tiles / (waves * slots);
        double const fill = ((double)s.M * (double)s.N) / (tiles * t.bm * t.bn);
        double const score = speed[i] * wave_eff * fill;
        if (score > best_score * 1.02) {  // prefer earlier (default) configs on near-ties
            best_score = score;
            best = i;
        }
    }
    return best;
}

void record_choice(int variant, int cfg, const char* name, int launches, int amode, int bmode) {
    g_choice.variant = variant;
    g_choice.config = cfg;
    g_choice.launches = launches;
    g_choice.a_mode = amode;
    g_choice.b_mode = bmode;
    std::snprintf(g_choice.name, sizeof g_choice.name, "%s", name);
    g_launches.fetch_add((uint64_t)launches, std::memory_order_relaxed);
}

// 3xTF32 tile choice: configs 0 (pair, 256x256), 4 (pair, 256x128), 1 (128x128), 5 (128x64); score = relative
// per-SM speed of the config x how well its tile count fills whole waves of the machine x tile fill.
// Static tile assignment: the dynamic scheduler (configs 2, 3) is 0-5% slower on a GPU the kernel has to
// itself (profiles/r01m_*); it exists for runs that share SMs with a concurrent NCCL kernel.
int pick_tf32_config(const MtmShape& s, int sm_count) {
    struct Cand { int cfg, bm, bn, ncta; double speed; };
    // Relative per-SM speeds on full grids, re-measured after the pair kernels lost the peer producer's per-k-block remote arrive
    // (profiles/r03n_ab_peer_arrive.jsonl, r03o_tune.json): 8192^3 256x256 260 = 256x512 259 TFLOP/s, 256x128 236, 128x128 200, 128x64 122;
    // double tiles (config 9) move a quarter less data through L2 -> SM and DRAM, which pays as the problem grows (16384^3 246 against
    // 222) and costs an exposed accumulator drain per unit when K is short (8192^2 x 1024 224 against 245, 65536 x 1024^2 204 against 218).
    double const dbl_speed = s.K >= 16384 ? 1.08 : (s.K >= 8192 ? 1.03 : 0.95);
    Cand const cands[] = {{0, 256, 256, 2, 1.00}, {9, 256, 512, 2, dbl_speed}, {4, 256, 128, 2, 0.90}, {1, 128, 128, 1, 0.78}, {5, 128, 64, 1, 0.47}};
    static bool const no_dbl = std::getenv("B200_TF32_NO_DOUBLE_TILES") != nullptr;     // (measurement aid)
    int best = 0;
    double best_score = -1.0;
    for (const Cand& c : cands) {
        if (c.cfg >= tf32_num_configs() || (c.cfg == 9 && no_dbl)) continue;
        double const tiles = (double)((s.M + c.bm - 1) / c.bm) * (double)((s.N + c.bn - 1) / c.bn);
        double const slots = (double)(sm_count / c.ncta);
        // the K splits AUTO applies (few tiles: all of them; ragged last wave: its tiles) count towards filling the machine
        double const fill = ((double)s.M * (double)s.N) / (tiles * c.bm * c.bn);
        double const score = c.speed * tf32_wave_efficiency((int64_t)tiles, (int)((s.K + 31) / 32), (int)slots) * fill;
        if (score > best_score * 1.02) {
            best_score = score;
            best = c.cfg;
        }
    }
    return best;
}

// AUTO, fp32: which kernel family a problem of this shape runs on.
// The tensor-core path pays a fixed operand-split pass.  Since the MMA kernel is a programmatic dependent launch behind
// it (and the split behind the previous call) that cost is small: measured (profiles/r02z_auto_crossover.jsonl) the path
// is ahead of the CUDA-core kernels from 96^3 on, thin shapes included (4096x64x512 13.9 vs 28.7 us, 33x4096x4096 58 vs
// 182, 4096^2 x 32 29 vs 53); only around 64^3 and 256x256x64 the CUDA cores are level or ahead (7.2 vs 9.1 / 7.3 us).
int auto_variant_f32(const MtmShape& s) {
    bool const big = tf32_num_configs() > 0 && s.M >= 32 && s.N >= 32 && s.K >= 32 &&
                     (double)s.M * (double)s.N * (double)s.K >= 8.0e5;
    return big ? B200_MTM_3XTF32 : B200_MTM_SIMT;
}

// AUTO tile config of the fp32 CUDA-core family.  Large problems: TMA-fed kernel (operands re-laid mn-contiguous when
// needed).  Small or thin ones: the register-staged kernels, whose smaller tiles fill the machine better.
// (thresholds from profiles/r02k_tune_simt_small.json and r02w_tune_simt_small2.json: at 1536^3 the TMA-fed 128x128
// kernel gives 48-51 TFLOP/s against 41 for the best register-staged config; at 1024^3 its 64 tiles leave most SMs
// idle (21) and the 64x128 form, 128 tiles at 3 CTAs/SM, gives 33.6 against 25; at 768^3 the register-staged 64x64
// kernel is still ahead, 22 against 18)
int auto_config_simt_f32(const MtmShape& s, int sm_count) {
    int const n_classic = simt_f32_num_configs();
    bool const big = s.M >= 256 && s.N >= 256 && s.K >= 128 && s.K <= kGridYLimit * 32 &&
                     (double)s.M * (double)s.N >= 1024.0 * 1024.0;
    if (!big) return pick_config(s, sm_count, n_classic, simt_f32_config, kSimtF32Speed);
    int64_t const tiles128 = ((s.M + 127) / 128) * ((s.N + 127) / 128);
    bool const narrow = ffma_tma_num_configs() > 3 && tiles128 * 5 < (int64_t)sm_count * 3;
    return n_classic + (narrow ? 3 : 0);
}

int run_f32(DeviceCtx& ctx, const Canon<float>& p, int flags, cudaStream_t st, int reuse_b) {
    int variant = flags & 0xff;
    int cfg = ((flags >> 8) & 0xff) - 1;
    if (variant == B200_MTM_DFMA || variant == B200_MTM_DMMA || variant > B200_MTM_DMMA)
        return fail(B200_ERR_INVALID, "b200_mtm_f32: variant %d is not an fp32 kernel family", variant);
    if (p.s.K == 0) {
        record_choice(B200_MTM_SIMT, 0, "noop_k0", 0, 0, 0);
        return B200_OK;
    }
    int const amode = load_mode(p.a, p.s.a_sm, p.s.a_sk);
    int const bmode = load_mode(p.b, p.s.b_sn, p.s.b_sk);
    int const vec_c = ((reinterpret_cast<uintptr_t>(p.c) & 15u) == 0 && p.s.ldc % 4 == 0) ? 1 : 0;
    if (variant == B200_MTM_AUTO) variant = auto_variant_f32(p.s);
    if (variant == B200_MTM_3XTF32) {
        if (tf32_num_configs() == 0)
            return fail(B200_ERR_INVALID, "b200_mtm_f32: 3xTF32 path not built into this library");
        if (cfg < 0) cfg = pick_tf32_config(p.s, ctx.sm_count);
        if (cfg >= tf32_num_configs()) return fail(B200_ERR_INVALID, "b200_mtm_f32: bad 3xTF32 config %d", cfg);
        size_t const need = tf32_workspace_bytes(p.s, p.a, p.b);
        int rc = acquire(ctx.tf32_ws, need, st);
        if (rc) return rc;
        int launches = 0, ta = 0, tb = 0, sk = 1;
        CUDA_TRY(launch_3xtf32_f32(p.c, p.a, p.b, p.s, ctx.tf32_ws.ptr, ctx.tf32_ws.bytes, cfg, reuse_b, (flags >> 16) & 0xff,
                                   (flags >> 24) & 0x7f, st, &launches, &ta, &tb, &sk, &cfg));
        if ((rc = release(ctx.tf32_ws, st))) return rc;
        (void)amode;
        (void)bmode;
        char name[64];
        if (sk > 100) std::snprintf(name, sizeof name, "%s_tailsplit%d", tf32_config(cfg).name, sk - 100);
        else if (sk > 1) std::snprintf(name, sizeof name, "%s_splitk%d", tf32_config(cfg).name, sk);
        else if (sk < 0) std::snprintf(name, sizeof name, "%s_streamk", tf32_config(cfg).name);
        else std::snprintf(name, sizeof name, "%s", tf32_config(cfg).name);
        record_choice(B200_MTM_3XTF32, cfg, name, launches, ta, tb);
        return B200_OK;
    }
    int const n_classic = simt_f32_num_configs();
    if (cfg < 0) cfg = auto_config_simt_f32(p.s, ctx.sm_count);
    if (cfg >= n_classic + ffma_tma_num_configs()) return fail(B200_ERR_INVALID, "b200_mtm_f32: bad SIMT config %d", cfg);
    if (cfg >= n_classic && p.s.K > kGridYLimit * 32)
        return fail(B200_ERR_INVALID, "b200_mtm_f32: the TMA-fed FFMA configs take K <= %lld (their pack pass puts K/32 in gridDim.y); "
                    "use AUTO or a register-staged config", (long long)(kGridYLimit * 32));
    if (cfg >= n_classic) {
        int const tcfg = cfg - n_classic;
        bool const a_direct = amode == LOAD_MN_VEC && p.s.a_sk >= p.s.M;
        bool const b_direct = bmode == LOAD_MN_VEC && p.s.b_sk >= p.s.N;
        bool const packs = !a_direct || !b_direct;
        if (packs) {
            int rc = acquire(ctx.pack_ws, ffma_tma_workspace_bytes(p.s), st);
            if (rc) return rc;
        }
        int launches = 0;
        CUDA_TRY(launch_ffma_tma_f32(tcfg, p.c, p.a, p.b, p.s, ctx.pack_ws.ptr, ctx.pack_ws.bytes, vec_c, reuse_b, st, &launches));
        if (packs) {
            int rc = release(ctx.pack_ws, st);
            if (rc) return rc;
        }
        record_choice(B200_MTM_SIMT, cfg, ffma_tma_config(tcfg).name, launches, amode, bmode);
        return B200_OK;
    }
    CUDA_TRY(launch_simt_f32(cfg, p.c, p.a, p.b, p.s, amode, bmode, vec_c, st));
    bool const generic = amode == LOAD_GENERIC || bmode == LOAD_GENERIC;
    record_choice(B200_MTM_SIMT, cfg, simt_f32_config(cfg).name, 1, generic ? 2 : amode, generic ? 2 : bmode);
    return B200_OK;
}

int run_f64(DeviceCtx& ctx, const Canon<double>& p, int flags, cudaStream_t st, int reuse_b) {
    int variant = flags & 0xff;
    int cfg = ((flags >> 8) & 0xff) - 1;
    if (variant == B200_MTM_3XTF32 || variant > B200_MTM_DMMA)
        return fail(B200_ERR_INVALID, "b200_mtm_f64: variant %d is not an fp64 kernel family", variant);
    if (p.s.K == 0) {
        record_choice(B200_MTM_SIMT, 0, "noop_k0", 0, 0, 0);
        return B200_OK;
    }
    int const amode = load_mode(p.a, p.s.a_sm, p.s.a_sk);
    int const bmode = load_mode(p.b, p.s.b_sn, p.s.b_sk);
    int const vec_c = ((reinterpret_cast<uintptr_t>(p.c) & 15u) == 0 && p.s.ldc % 2 == 0) ? 1 : 0;
    bool const generic = amode == LOAD_GENERIC || bmode == LOAD_GENERIC;
    // DMMA reaches 87% of the FP64 peak at 8192^3 against 61% for DFMA (profiles/r01_*): the
    // tensor-core path is the default; DFMA stays selectable.
    if (variant == B200_MTM_AUTO) variant = B200_MTM_DMMA;
    if (variant == B200_MTM_DMMA) {
        int const n_classic = dmma_f64_num_configs();
        if (cfg < 0) {
            // Large problems: TMA-fed DMMA kernel (operands re-laid K-contiguous when needed); small or
            // thin ones: the register-staged kernels.
            bool const big = p.s.M >= 256 && p.s.N >= 256 && p.s.K >= 64 && p.s.M <= kGridYLimit * 32 && p.s.N <= kGridYLimit * 32 &&
                             (double)p.s.M * (double)p.s.N >= 148.0 * 3 * 64 * 64 * 0.75;
            cfg = big ? n_classic : pick_config(p.s, ctx.sm_count, n_classic, dmma_f64_config, kDmmaF64Speed);
        }
        if (cfg >= n_classic + dmma_tma_num_configs()) return fail(B200_ERR_INVALID, "b200_mtm_f64: bad DMMA config %d", cfg);
        if (cfg >= n_classic && (p.s.M > kGridYLimit * 32 || p.s.N > kGridYLimit * 32))
            return fail(B200_ERR_INVALID, "b200_mtm_f64: the TMA-fed DMMA configs take M, N <= %lld (their pack pass puts rows/32 in "
                        "gridDim.y); use AUTO or a register-staged config", (long long)(kGridYLimit * 32));
        if (cfg >= n_classic) {
            int const tcfg = cfg - n_classic;
            int rc = acquire(ctx.pack_ws, dmma_tma_workspace_bytes(p.s), st);
            if (rc) return rc;
            int launches = 0;
            CUDA_TRY(launch_dmma_tma_f64(tcfg, p.c, p.a, p.b, p.s, ctx.pack_ws.ptr, ctx.pack_ws.bytes, vec_c, reuse_b, st, &launches));
            if ((rc = release(ctx.pack_ws, st))) return rc;
            record_choice(B200_MTM_DMMA, cfg, dmma_tma_config(tcfg).name, launches, generic ? 2 : amode, generic ? 2 : bmode);
            return B200_OK;
        }
        CUDA_TRY(launch_dmma_f64(cfg, p.c, p.a, p.b, p.s, amode, bmode, vec_c, st));
        record_choice(B200_MTM_DMMA, cfg, dmma_f64_config(cfg).name, 1, generic ? 2 : amode, generic ? 2 : bmode);
        return B200_OK;
    }
    if (cfg < 0) cfg = pick_config(p.s, ctx.sm_count, simt_f64_num_configs(), simt_f64_config, kSimtF64Speed);
    if (cfg >= simt_f64_num_configs()) return fail(B200_ERR_INVALID, "b200_mtm_f64: bad DFMA config %d", cfg);
    CUDA_TRY(launch_simt_f64(cfg, p.c, p.a, p.b, p.s, amode, bmode, vec_c, st));
    record_choice(B200_MTM_SIMT, cfg, simt_f64_config(cfg).name, 1, generic ? 2 : amode, generic ? 2 : bmode);
    return B200_OK;
}

int run(DeviceCtx& ctx, const Canon<float>& p, int flags, cudaStream_t st, int reuse_b) { return run_f32(ctx, p, flags, st, reuse_b); }
int run(DeviceCtx& ctx, const Canon<double>& p, int flags, cudaStream_t st, int reuse_b) { return run_f64(ctx, p, flags, st, reuse_b); }

template <typename T>
int mtm_dev(T* c, const size_t* nc, const size_t* wc, const T* a, const size_t* na, const size_t* wa,
            const T* b, const size_t* nb, const size_t* wb, int flags, void* stream) {
    NvtxRange const range(sizeof(T) == 4 ? "b200_mtm_f32_dev" : "b200_mtm_f64_dev");
    int rc = validate(c, nc, wc, a, na, wa, b, nb, wb);
    if (rc) return rc;
    DeviceCtx* ctx;
    rc = current_ctx(&ctx);
    if (rc) return rc;
    if (nc[0] == 0 || nc[1] == 0) {
        record_choice(B200_MTM_SIMT, 0, "noop_empty", 0, 0, 0);
        return B200_OK;
    }
    Canon<T> p = canonicalise(c, nc, wc, a, na, wa, b, nb, wb);
    return run(*ctx, p, flags, static_cast<cudaStream_t>(stream), 0);
}

// ---- host-pointer path ----------------------------------------------------------------------------
// A host operand is staged as a "pitched" 2-D copy when one of its strides is 1 (only the matrix
// elements move, and the device image gets a 16-byte-aligned pitch so the vector loaders apply),
// otherwise as one contiguous span with the host strides kept.
struct StagePlan {
    bool pitched;
    int run_dim;             // pitched: dimension with unit stride (runs are contiguous along it)
    size_t n[2];             // extents
    size_t host_w[2];        // host strides
    size_t dev_w[2];         // device-image strides
    size_t span;             // !pitched: elements from the first to the last matrix element
    size_t dev_elems;
};

StagePlan plan_stage(const size_t* n, const size_t* w, size_t vec) {
    StagePlan s{};
    s.n[0] = n[0];
    s.n[1] = n[1];
    s.host_w[0] = w[0];
    s.host_w[1] = w[1];
    auto round_up = [&](size_t x) { return (x + vec - 1) / vec * vec; };
    if (w[1] == 1 && (w[0] >= n[1] || n[0] == 1)) {  // rows are contiguous runs
        s.pitched = true;
        s.run_dim = 1;
        s.dev_w[0] = round_up(n[1]);
        s.dev_w[1] = 1;
        s.dev_elems = n[0] * s.dev_w[0];
    } else if (w[0] == 1 && (w[1] >= n[0] || n[1] == 1)) {  // columns are contiguous runs
        s.pitched = true;
        s.run_dim = 0;
        s.dev_w[0] = 1;
        s.dev_w[1] = round_up(n[0]);
        s.dev_elems = n[1] * s.dev_w[1];
    } else {
        s.pitched = false;
        s.run_dim = -1;
        s.span = (n[0] - 1) * w[0] + (n[1] - 1) * w[1] + 1;
        s.dev_w[0] = w[0];
        s.dev_w[1] = w[1];
        s.dev_elems = s.span;
    }
    return s;
}

// Copy the sub-block [lo0,hi0) x [lo1,hi1) between the host matrix and its device image.
template <typename T>
cudaError_t stage_copy_pageable(DeviceCtx& ctx, const StagePlan& s, T* dev, T* host, const size_t lo[2], const size_t hi[2],
                                bool to_device, cudaStream_t st);
bool is_pageable(const void* p);

// `ctx` given: a large copy from / to PAGEABLE host memory goes through the library's own bounce pipeline
// (the caller must drain_staged(ctx) before it returns to the user).
template <typename T>
cudaError_t stage_copy(const StagePlan& s, T* dev, T* host, const size_t lo[2], const size_t hi[2],
                       bool to_device, cudaStream_t st, DeviceCtx* ctx = nullptr) {
    if (hi[0] <= lo[0] || hi[1] <= lo[1]) return cudaSuccess;
    if (ctx != nullptr && s.pitched) {
        int const r = s.run_dim;
        size_t const width = (hi[r] - lo[r]) * sizeof(T), bytes = width * (hi[1 - r] - lo[1 - r]);
        static bool const off = std::getenv("B200_NO_PAGEABLE_STAGING") != nullptr;
        if (!off && bytes >= ((size_t)4 << 20) && width <= DeviceCtx::kBounceBytes && is_pageable(host))
            return stage_copy_pageable(*ctx, s, dev, host, lo, hi, to_device, st);
    }
    if (s.pitched) {
        int const r = s.run_dim, o = 1 - r;
        size_t const rows = hi[o] - lo[o], width = hi[r] - lo[r];
        T* hp = host + lo[o] * s.host_w[o] + lo[r];
        T* dp = dev + lo[o] * s.dev_w[o] + lo[r];
        size_t const hpitch = rows == 1 ? width : s.host_w[o], dpitch = rows == 1 ? width : s.dev_w[o];
        if (to_device)
            return cudaMemcpy2DAsync(dp, dpitch * sizeof(T), hp, hpitch * sizeof(T), width * sizeof(T), rows,
                                     cudaMemcpyHostToDevice, st);
        return cudaMemcpy2DAsync(hp, hpitch * sizeof(T), dp, dpitch * sizeof(T), width * sizeof(T), rows,
                                 cudaMemcpyDeviceToHost, st);
    }
    // span copy: whole matrix only
    if (to_device) return cudaMemcpyAsync(dev, host, s.span * sizeof(T), cudaMemcpyHostToDevice, st);
    return cudaMemcpyAsync(host, dev, s.span * sizeof(T), cudaMemcpyDeviceToHost, st);
}

// ---- pageable host memory ---------------------------------------------------------------------------------
// The reference's tensors are ordinary heap memory (make_tensor, include/utils.hpp:21-31; src/mtm.cpp:204-208).
// cudaMemcpy from pageable memory is staged by the driver through one internal buffer by one thread (measured:
// 11.5 GB/s, 93 ms for the 8192^3 call against 16.5 ms from pinned memory).  Here the library stages such copies
// itself: a few host threads memcpy 16 MiB chunks into pinned bounce buffers while the copy engine moves the
// chunks before them (and the reverse for C on the way back).
class CopyPool {
    struct Task { char* dst; const char* src; size_t dpitch, spitch, width, rows; std::atomic<int>* left; };
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::vector<Task> q_;
    bool stop_ = false;
    static void do_copy(const Task& t) {
        if (t.dpitch == t.width && t.spitch == t.width) {
            std::memcpy(t.dst, t.src, t.width * t.rows);
            return;
        }
        for (size_t r = 0; r < t.rows; ++r) std::memcpy(t.dst + r * t.dpitch, t.src + r * t.spitch, t.width);
    }
    void loop() {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                t = q_.back();
                q_.pop_back();
            }
            do_copy(t);
            if (t.left->fetch_sub(1) == 1) {          // last piece of its call: wake the caller
                std::lock_guard<std::mutex> lk(m_);
                done_.notify_all();
            }
        }
    }
public:
    CopyPool() {
        unsigned n = std::thread::hardware_concurrency();
        if (const char* v = std::getenv("B200_COPY_THREADS")) n = (unsigned)std::atoi(v) + 1;
        n = n < 2 ? 2 : (n > 13 ? 13 : n);
        for (unsigned i = 0; i + 1 < n; ++i) th_.emplace_back([this] { loop(); });   // the caller is the n-th worker
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    // 2-D copy split by rows (or, for a single long row, by bytes) over the pool; returns when all of it is done.
    // Several callers (the per-device threads of the multi-GPU entry) may use the pool at once: each call waits
    // for its own pieces only.
    void copy2d(char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, size_t rows) {
        size_t const parts = th_.size() + 1;
        std::atomic<int> left{0};
        std::vector<Task> mine;
        if (rows == 1 || (dpitch == width && spitch == width)) {
            size_t const total = width * rows, per = ((total + parts - 1) / parts + 4095) & ~(size_t)4095;
            for (size_t o = 0; o < total; o += per) {
                size_t const w = std::min(per, total - o);
                mine.push_back({dst + o, src + o, w, w, w, 1, &left});
            }
        } else {
            size_t const per = (rows + parts - 1) / parts;
            for (size_t r = 0; r < rows; r += per) mine.push_back({dst + r * dpitch, src + r * spitch, dpitch, spitch, width, std::min(per, rows - r), &left});
        }
        if (mine.size() <= 1) {
            for (auto& t : mine) do_copy(t);
            return;
        }
        Task const own = mine.back();
        mine.pop_back();
        left.store((int)mine.size());
        {
            std::lock_guard<std::mutex> lk(m_);
            for (auto& t : mine) q_.push_back(t);
        }
        cv_.notify_all();
        do_copy(own);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return left.load() == 0; });
    }
};
CopyPool& copy_pool() {
    static CopyPool pool;
    return pool;
}

bool is_pageable(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

// Copy whatever device-to-host chunks are parked in the bounce buffers on to the caller's memory.
cudaError_t drain_bounce(DeviceCtx& ctx, DeviceCtx::Bounce& b) {
    if (b.busy) {
        cudaError_t e = cudaEventSynchronize(b.ev);
        if (e != cudaSuccess) return e;
        b.busy = false;
    }
    if (b.out_dst != nullptr) {
        copy_pool().copy2d(b.out_dst, b.out_pitch, static_cast<const char*>(b.ptr), b.out_width, b.out_width, b.out_rows);
        b.out_dst = nullptr;
    }
    (void)ctx;
    return cudaSuccess;
}
// Entry of a host-pointer call: forget chunks a FAILED earlier call left parked (their destination is gone).
void reset_staged(DeviceCtx& ctx) {
    for (auto* ring : {ctx.up, ctx.down})
        for (int i = 0; i < DeviceCtx::kBounce; ++i) {
            if (ring[i].busy) {
                (void)cudaEventSynchronize(ring[i].ev);
                (void)cudaGetLastError();
                ring[i].busy = false;
            }
            ring[i].out_dst = nullptr;
        }
}
cudaError_t drain_staged(DeviceCtx& ctx) {
    for (int i = 0; i < DeviceCtx::kBounce; ++i) {   // oldest first
        cudaError_t e = drain_bounce(ctx, ctx.down[(ctx.down_next + i) % DeviceCtx::kBounce]);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// stage_copy for a PAGEABLE host matrix (pitched plans): chunks of whole runs through the bounce rings.
template <typename T>
cudaError_t stage_copy_pageable(DeviceCtx& ctx, const StagePlan& s, T* dev, T* host, const size_t lo[2], const size_t hi[2],
                                bool to_device, cudaStream_t st) {
    int const r = s.run_dim, o = 1 - r;
    size_t const rows = hi[o] - lo[o], width = (hi[r] - lo[r]) * sizeof(T);
    char* hp = reinterpret_cast<char*>(host + lo[o] * s.host_w[o] + lo[r]);
    char* dp = reinterpret_cast<char*>(dev + lo[o] * s.dev_w[o] + lo[r]);
    size_t const hpitch = rows == 1 ? width : s.host_w[o] * sizeof(T), dpitch = rows == 1 ? width : s.dev_w[o] * sizeof(T);
    if (width > DeviceCtx::kBounceBytes) return cudaErrorInvalidValue;     // (callers check: a run fits a bounce buffer)
    size_t const rows_per = std::max<size_t>(1, DeviceCtx::kBounceBytes / width);
    for (size_t r0 = 0; r0 < rows; r0 += rows_per) {
        size_t const nr = std::min(rows_per, rows - r0);
        DeviceCtx::Bounce& b = to_device ? ctx.up[ctx.up_next] : ctx.down[ctx.down_next];
        (to_device ? ctx.up_next : ctx.down_next) = ((to_device ? ctx.up_next : ctx.down_next) + 1) % DeviceCtx::kBounce;
        cudaError_t e;
        if (!b.ptr) {
            if ((e = cudaHostAlloc(&b.ptr, DeviceCtx::kBounceBytes, cudaHostAllocDefault)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        if ((e = drain_bounce(ctx, b)) != cudaSuccess) return e;            // its previous chunk has left (and landed)
        if (to_device) {
            copy_pool().copy2d(static_cast<char*>(b.ptr), width, hp + r0 * hpitch, hpitch, width, nr);
            e = cudaMemcpy2DAsync(dp + r0 * dpitch, dpitch, b.ptr, width, width, nr, cudaMemcpyHostToDevice, st);
        } else {
            e = cudaMemcpy2DAsync(b.ptr, width, dp + r0 * dpitch, dpitch, width, nr, cudaMemcpyDeviceToHost, st);
            b.out_dst = hp + r0 * hpitch;
            b.out_pitch = hpitch;
            b.out_width = width;
            b.out_rows = nr;
        }
        if (e != cudaSuccess) return e;
        if ((e = cudaEventRecord(b.ev, st)) != cudaSuccess) return e;
        b.busy = true;
    }
    return cudaSuccess;
}

// Slab boundaries of [x0, x1): up to kMaxSlabs equal slabs (multiples of 256 rows); the LAST one is cut again
// into 1/2, 1/4, 1/4 — what is exposed at the end of the copy/compute/copy-back pipeline is the product and the
// copy-back of the final slab, so that one is made small.
std::vector<size_t> slab_cuts(size_t x0, size_t x1, bool taper) {
    std::vector<size_t> cut{x0};
    size_t slab = (x1 - x0 + kMaxSlabs - 1) / kMaxSlabs;
    slab = (slab + 255) / 256 * 256;
    if (slab < 512) slab = 512;          // thinner slabs fall below what the tensor-core kernels need to be worth launching
    for (size_t r = x0 + slab; r < x1; r += slab) cut.push_back(r);
    size_t const last0 = cut.back(), len = x1 - last0;
    if (taper && cut.size() > 1 && len >= 1024) {
        size_t const h = (len / 2 + 255) / 256 * 256, q = (len / 4 + 255) / 256 * 256;
        if (last0 + h < x1) cut.push_back(last0 + h);
        if (last0 + h + q < x1) cut.push_back(last0 + h + q);
    }
    cut.push_back(x1);
    return cut;
}

// Synchronous host-pointer entry.  Large problems are cut into slabs along C's slow dimension and
// pipelined over three streams: while slab i is multiplied, slab i+1 (its rows of A and C) is on
// its way to the device and the finished slab i-1 of C is on its way back; the operand every
// slab needs in full (B for a row-major C) goes first.  Every slab is an ordinary mtm call on
// device pointers, so the result is bit-identical to the unsliced call.
template <typename T>
int mtm_host(T* c, const size_t* nc, const size_t* wc, const T* a, const size_t* na, const size_t* wa,
             const T* b, const size_t* nb, const size_t* wb, int flags) {
    NvtxRange const range(sizeof(T) == 4 ? "b200_mtm_f32 (host)" : "b200_mtm_f64 (host)");
    int rc = validate(c, nc, wc, a, na, wa, b, nb, wb);
    if (rc) return rc;
    DeviceCtx* ctxp;
    rc = current_ctx(&ctxp);
    if (rc) return rc;
    DeviceCtx& ctx = *ctxp;
    if (nc[0] == 0 || nc[1] == 0 || na[1] == 0) {
        record_choice(B200_MTM_SIMT, 0, "noop_empty", 0, 0, 0);
        return B200_OK;
    }
    constexpr size_t V = 16 / sizeof(T);
    StagePlan const pa = plan_stage(na, wa, V), pb = plan_stage(nb, wb, V), pc = plan_stage(nc, wc, V);
    reset_staged(ctx);
    if ((rc = ensure(ctx.stage[0], pa.dev_elems * sizeof(T)))) return rc;
    if ((rc = ensure(ctx.stage[1], pb.dev_elems * sizeof(T)))) return rc;
    if ((rc = ensure(ctx.stage[2], pc.dev_elems * sizeof(T)))) return rc;
    T* da = static_cast<T*>(ctx.stage[0].ptr);
    T* db = static_cast<T*>(ctx.stage[1].ptr);
    T* dc = static_cast<T*>(ctx.stage[2].ptr);
    T* ha = const_cast<T*>(a);
    T* hb = const_cast<T*>(b);
    cudaStream_t const s_comp = ctx.host_stream, s_in = ctx.copy_stream, s_out = ctx.out_stream;
    size_t const zero[2] = {0, 0};

    // Slice along C's slow dimension: rows of C and A for a row-contiguous C, columns of C and B
    // for a column-contiguous C (the transposed problem of canonicalise()).
    bool const row_contig = (wc[1] == 1) || nc[1] == 1;
    bool const col_contig = (wc[0] == 1) || nc[0] == 1;
    int const slice_dim = (row_contig || !col_contig) ? 0 : 1;
    size_t const extent = nc[slice_dim];
    double const bytes_total = (double)sizeof(T) * ((double)na[0] * na[1] + (double)nb[0] * nb[1] + (double)nc[0] * nc[1]);
    bool const can_slice = pc.pitched && (slice_dim == 0 ? pa.pitched : pb.pitched) && extent >= 512 &&
                           bytes_total >= 48.0 * 1024 * 1024;
    size_t slab = extent;
    if (can_slice) {
        slab = (extent + kMaxSlabs - 1) / kMaxSlabs;
        slab = (slab + 255) / 256 * 256;
    }
    int launches = 0;
    if (slab >= extent) {
        // single shot: B and C on the copy-in stream, A on the compute stream
        CUDA_TRY(stage_copy(pb, db, hb, zero, pb.n, true, s_in, &ctx));
        CUDA_TRY(stage_copy(pc, dc, c, zero, pc.n, true, s_in, &ctx));
        CUDA_TRY(cudaEventRecord(ctx.ev_in[0], s_in));
        CUDA_TRY(stage_copy(pa, da, ha, zero, pa.n, true, s_comp, &ctx));
        CUDA_TRY(cudaStreamWaitEvent(s_comp, ctx.ev_in[0], 0));
        Canon<T> p = canonicalise(dc, nc, pc.dev_w, static_cast<const T*>(da), na, pa.dev_w,
                                  static_cast<const T*>(db), nb, pb.dev_w);
        if ((rc = run(ctx, p, flags, s_comp, 0))) return rc;
        CUDA_TRY(stage_copy(pc, dc, c, zero, pc.n, false, s_comp, &ctx));
        CUDA_TRY(drain_staged(ctx));
        CUDA_TRY(cudaStreamSynchronize(s_comp));
        return B200_OK;
    }

    // The operand shared by all slabs goes first.
    if (slice_dim == 0) CUDA_TRY(stage_copy(pb, db, hb, zero, pb.n, true, s_in, &ctx));
    else CUDA_TRY(stage_copy(pa, da, ha, zero, pa.n, true, s_in, &ctx));
    int i = 0;
    // Slabs never split K on their own (each would decide from its own, smaller tile count): the sliced call
    // then stays bit-identical to the unsliced one, which for problems this large does not split either.
    if (((flags >> 24) & 0x7f) == 0) flags |= B200_MTM_SPLIT_K(1);
    int slab_flags = flags;
    static bool const no_taper = std::getenv("B200_NO_TAPER") != nullptr;    // (measurement aid)
    std::vector<size_t> const cut = slab_cuts(0, extent, !no_taper);
    for (; i + 1 < (int)cut.size(); ++i) {
        size_t const r0 = cut[i], r1 = cut[i + 1];
        size_t lo[2] = {0, 0}, hi_c[2] = {nc[0], nc[1]}, hi_x[2];
        lo[slice_dim] = r0;
        hi_c[slice_dim] = r1;
        // the sliced operand: rows [r0,r1) of A, or columns [r0,r1) of B
        const StagePlan& px = slice_dim == 0 ? pa : pb;
        hi_x[0] = px.n[0];
        hi_x[1] = px.n[1];
        hi_x[slice_dim] = r1;
        CUDA_TRY(stage_copy(px, slice_dim == 0 ? da : db, slice_dim == 0 ? ha : hb, lo, hi_x, true, s_in, &ctx));
        CUDA_TRY(stage_copy(pc, dc, c, lo, hi_c, true, s_in, &ctx));
        CUDA_TRY(cudaEventRecord(ctx.ev_in[i], s_in));
        CUDA_TRY(cudaStreamWaitEvent(s_comp, ctx.ev_in[i], 0));
        size_t ncs[2] = {nc[0], nc[1]}, nas[2] = {na[0], na[1]}, nbs[2] = {nb[0], nb[1]};
        ncs[slice_dim] = r1 - r0;
        T* dcs = dc + r0 * pc.dev_w[slice_dim];
        const T* das = da;
        const T* dbs = db;
        if (slice_dim == 0) {
            nas[0] = r1 - r0;
            das = da + r0 * pa.dev_w[0];
        } else {
            nbs[1] = r1 - r0;
            dbs = db + r0 * pb.dev_w[1];
        }
        Canon<T> p = canonicalise(dcs, ncs, pc.dev_w, das, nas, pa.dev_w, dbs, nbs, pb.dev_w);
        if ((rc = run(ctx, p, slab_flags, s_comp, i > 0 ? 1 : 0))) return rc;
        // The kernel family and tile config are resolved ONCE, by the first slab: a shorter tail slab must not
        // re-resolve AUTO to another family (the re-laid B it reuses lives in that family's workspace, and the
        // result has to be bit-identical to the unsliced call's arithmetic).
        if (i == 0) slab_flags = B200_MTM_FLAGS(g_choice.variant, g_choice.config + 1) | (flags & 0x7fff0000);
        launches += g_choice.launches;
        CUDA_TRY(cudaEventRecord(ctx.ev_done[i], s_comp));
        CUDA_TRY(cudaStreamWaitEvent(s_out, ctx.ev_done[i], 0));
        CUDA_TRY(stage_copy(pc, dc, c, lo, hi_c, false, s_out, &ctx));
    }
    CUDA_TRY(drain_staged(ctx));
    CUDA_TRY(cudaStreamSynchronize(s_out));
    CUDA_TRY(cudaStreamSynchronize(s_comp));
    g_choice.launches = launches;
    return B200_OK;
}

// ---- multi-GPU host-pointer entry --------------------------------------------------------------------
// The reference spreads ONE call over all cores (OpenMP team over M-blocks, include/mtm.hpp:156-201).  Its
// analogue here spreads one call over the GPUs of the box, in ONE process (no torch, no NCCL): C's slow
// dimension is cut into one shard per device (the row-block partition of SURVEY 8e; columns of C and B for a
// column-major C); a host thread per device drives that device's copy / compute / copy-back pipeline; the
// operand every shard needs whole is uploaded ONCE — device d fetches the d-th 1/P of it over its own PCIe
// link and sends that slice to every peer over NVLink (peer copies; cross-device events order them before the
// first product) — instead of P times.  K is never split: every element of C is produced by exactly the
// kernel and summation order of the single-GPU call.
class HostBarrier {
    std::mutex m_;
    std::condition_variable cv_;
    int n_, count_ = 0, gen_ = 0;
public:
    explicit HostBarrier(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        int const g = gen_;
        if (++count_ == n_) {
            count_ = 0;
            ++gen_;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&] { return g != gen_; });
        }
    }
};

struct MgpuShared {
    int P = 0;
    std::vector<int> devs;
    std::vector<DeviceCtx*> ctx;
    std::vector<void*> replica;          // device image of the shared operand on every device
    std::vector<int> rc;
    std::vector<std::string> err;
    std::atomic<bool> failed{false};
    bool gather = true;                  // slices travel peer to peer (else every device uploads all of it)
    // kernel family / tile config: resolved by shard 0's first slab, used by every shard
    std::mutex fm;
    std::condition_variable fcv;
    bool flags_ready = false;
    int flags = 0;
    b200_mtm_choice choice{};
    std::atomic<int> launches{0};
};

template <typename T>
void mgpu_worker(MgpuShared& sh, HostBarrier& bar, int i, T* c, const size_t* nc, const size_t* wc, const T* a,
                 const size_t* na, const size_t* wa, const T* b, const size_t* nb, const size_t* wb, int flags,
                 int slice_dim, const std::vector<size_t>& cut, const StagePlan& pa, const StagePlan& pb,
                 const StagePlan& pc) {
    auto bail = [&](int rc) {
        sh.rc[i] = rc;
        sh.err[i] = g_err;
        sh.failed.store(true);
    };
    auto publish_flags = [&](int f) {
        std::lock_guard<std::mutex> lk(sh.fm);
        if (!sh.flags_ready) {
            sh.flags = f;
            sh.flags_ready = true;
            sh.choice = g_choice;
        }
        sh.fcv.notify_all();
    };
    NvtxRange const range("b200_mtm_mgpu shard");
    int const P = sh.P, dev = sh.devs[i];
    size_t const x0 = cut[i], x1 = cut[i + 1];
    const StagePlan& px = slice_dim == 0 ? pa : pb;          // the sliced operand (rows of A / columns of B)
    const StagePlan& ps = slice_dim == 0 ? pb : pa;          // the shared operand
    int const xi = slice_dim == 0 ? 0 : 1, si = 1 - xi;      // their stage-buffer indices
    T* hx = const_cast<T*>(slice_dim == 0 ? a : b);
    T* hs = const_cast<T*>(slice_dim == 0 ? b : a);
    // My shard as matrices of their own: rows (columns) [x0, x1) of the sliced operand and of C, same host strides, a
    // device image with its own pitch — whatever the layout (a column-major A cut by rows is NOT a contiguous part of
    // the full image), the shard's image is dense.
    constexpr size_t V = 16 / sizeof(T);
    const size_t* const wx = slice_dim == 0 ? wa : wb;
    size_t nxl[2] = {px.n[0], px.n[1]}, ncl[2] = {nc[0], nc[1]};
    nxl[slice_dim] = x1 - x0;
    ncl[slice_dim] = x1 - x0;
    StagePlan const pxl = plan_stage(nxl, wx, V), pcl = plan_stage(ncl, wc, V);
    T* const hxl = hx + x0 * wx[slice_dim];
    T* const hcl = c + x0 * wc[slice_dim];
    DeviceCtx* ctx = nullptr;
    int rc = B200_OK;
    // ---- phase 1: context, buffers, peer access -----------------------------------------------------------
    do {
        cudaError_t e = cudaSetDevice(dev);
        if (e != cudaSuccess) { rc = fail(B200_ERR_CUDA, "cudaSetDevice(%d) failed: %s", dev, cudaGetErrorString(e)); break; }
        if ((rc = current_ctx(&ctx))) break;
        sh.ctx[i] = ctx;
        reset_staged(*ctx);
        if ((rc = ensure(ctx->stage[si], ps.dev_elems * sizeof(T)))) break;
        if ((rc = ensure(ctx->stage[xi], (pxl.dev_elems + 16) * sizeof(T)))) break;
        if ((rc = ensure(ctx->stage[2], (pcl.dev_elems + 16) * sizeof(T)))) break;
        sh.replica[i] = ctx->stage[si].ptr;
        for (int j = 0; j < P && rc == B200_OK; ++j) {
            int const pd = sh.devs[j];
            if (pd == dev) continue;
            if (!ctx->ev_sent[pd]) {
                e = cudaEventCreateWithFlags(&ctx->ev_sent[pd], cudaEventDisableTiming);
                if (e != cudaSuccess) { rc = fail(B200_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(e)); break; }
            }
            if (sh.gather && !ctx->peer_on[pd]) {
                e = cudaDeviceEnablePeerAccess(pd, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); e = cudaSuccess; }
                if (e != cudaSuccess) { rc = fail(B200_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", dev, pd, cudaGetErrorString(e)); break; }
                ctx->peer_on[pd] = true;
            }
        }
    } while (0);
    if (rc) bail(rc);
    bar.wait();                                               // every replica exists
    // ---- phase 2: my slice of the shared operand: host -> me -> every peer ------------------------------------
    int const o = ps.pitched ? 1 - ps.run_dim : 0;            // the shared operand is cut along its line dimension
    size_t const lines = ps.pitched ? ps.n[o] : 0;
    if (!sh.failed.load()) {
        do {
            T* ds = static_cast<T*>(ctx->stage[si].ptr);
            size_t const zero[2] = {0, 0};
            if (!sh.gather || !ps.pitched || P == 1) {        // whole operand over my own link
                cudaError_t e = stage_copy(ps, ds, hs, zero, ps.n, true, ctx->copy_stream, ctx);
                if (e != cudaSuccess) { rc = fail(B200_ERR_CUDA, "staging the shared operand failed: %s", cudaGetErrorString(e)); break; }
                e = cudaEventRecord(ctx->ev_slice, ctx->copy_stream);
                if (e != cudaSuccess) { rc = fail(B200_ERR_CUDA, "cudaEventRecord failed: %s", cudaGetErrorString(e)); break; }
                break;
            }
            size_t const per = (lines + P - 1) / P;
            size_t const l0 = std::min(lines, (size_t)i * per), l1 = std::min(lines, (size_t)(i + 1) * per);
            size_t lo[2] = {0, 0}, hi[2] = {ps.n[0], ps.n[1]};
            lo[o] = l0;
            hi[o] = l1;
            cudaError_t e = stage_copy(ps, ds, hs, lo, hi, true, ctx->copy_stream, ctx);
            if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_slice, ctx->copy_stream);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->out_stream, ctx->ev_slice, 0);
            if (e != cudaSuccess) { rc = fail(B200_ERR_CUDA, "staging my slice of the shared operand failed: %s", cudaGetErrorString(e)); break; }
            size_t const off = l0 * ps.dev_w[o], bytes = (l1 - l0) * ps.dev_w[o] * sizeof(T);   // whole device lines: contiguous
            for (int j = 0; j < P; ++j) {
                int const pd = sh.devs[j];
                if (pd == dev) continue;
                if (bytes) e = cudaMemcpyPeerAsync(static_cast<T*>(sh.replica[j]) + off, pd, ds + off, dev, bytes, ctx->out_stream);
                if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_sent[pd], ctx->out_stream);
                if (e != cudaSuccess) { rc = fail(B200_ERR_CUDA, "peer copy %d -> %d failed: %s", dev, pd, cudaGetErrorString(e)); break; }
            }
        } while (0);
        if (rc) bail(rc);
    }
    bar.wait();                                               // every ev_sent has been recorded
    // ---- phase 3: my shard, slabs pipelined over three streams -------------------------------------------------
    if (!sh.failed.load() && x1 > x0) {
        do {
            cudaStream_t const s_comp = ctx->host_stream, s_in = ctx->copy_stream, s_out = ctx->out_stream;
            cudaError_t e = cudaStreamWaitEvent(s_comp, ctx->ev_slice, 0);
            if (sh.gather && ps.pitched && P > 1)
                for (int j = 0; j < P && e == cudaSuccess; ++j)
                    if (sh.devs[j] != dev) e = cudaStreamWaitEvent(s_comp, sh.ctx[j]->ev_sent[dev], 0);
            if (e != cudaSuccess) { rc = fail(B200_ERR_CUDA, "cudaStreamWaitEvent failed: %s", cudaGetErrorString(e)); break; }
            T* dx = static_cast<T*>(ctx->stage[xi].ptr);
            T* dc = static_cast<T*>(ctx->stage[2].ptr);
            T* ds = static_cast<T*>(ctx->stage[si].ptr);
            std::vector<size_t> const scut = slab_cuts(0, x1 - x0, true);      // slabs in shard-local indices
            int my_flags = ((flags >> 24) & 0x7f) == 0 ? (flags | B200_MTM_SPLIT_K(1)) : flags;   // as in mtm_host
            bool have_flags = false;
            int k = 0;
            for (; k + 1 < (int)scut.size() && rc == B200_OK; ++k) {
                size_t const r0 = scut[k], r1 = scut[k + 1];
                size_t lo[2] = {0, 0}, hi_c[2] = {ncl[0], ncl[1]}, hi_x[2] = {nxl[0], nxl[1]};
                lo[slice_dim] = r0;
                hi_c[slice_dim] = r1;
                hi_x[slice_dim] = r1;
                if ((e = stage_copy(pxl, dx, hxl, lo, hi_x, true, s_in, ctx)) != cudaSuccess ||
                    (e = stage_copy(pcl, dc, hcl, lo, hi_c, true, s_in, ctx)) != cudaSuccess ||
                    (e = cudaEventRecord(ctx->ev_in[k], s_in)) != cudaSuccess ||
                    (e = cudaStreamWaitEvent(s_comp, ctx->ev_in[k], 0)) != cudaSuccess) {
                    rc = fail(B200_ERR_CUDA, "staging slab %d failed: %s", k, cudaGetErrorString(e));
                    break;
                }
                size_t ncs[2] = {ncl[0], ncl[1]}, nas[2] = {na[0], na[1]}, nbs[2] = {nb[0], nb[1]};
                ncs[slice_dim] = r1 - r0;
                T* dcs = dc + r0 * pcl.dev_w[slice_dim];
                const T* dxs = dx + r0 * pxl.dev_w[slice_dim];
                const T* das = slice_dim == 0 ? dxs : ds;
                const T* dbs = slice_dim == 0 ? ds : dxs;
                const size_t* const dwa = slice_dim == 0 ? pxl.dev_w : pa.dev_w;
                const size_t* const dwb = slice_dim == 0 ? pb.dev_w : pxl.dev_w;
                if (slice_dim == 0) nas[0] = r1 - r0; else nbs[1] = r1 - r0;
                if (!have_flags && i != 0) {                  // shard 0 resolves the kernel for everybody
                    std::unique_lock<std::mutex> lk(sh.fm);
                    sh.fcv.wait(lk, [&] { return sh.flags_ready; });
                    my_flags = sh.flags;
                    have_flags = true;
                    if (sh.failed.load()) break;
                }
                Canon<T> p = canonicalise(dcs, ncs, pcl.dev_w, das, nas, dwa, dbs, nbs, dwb);
                if ((rc = run(*ctx, p, my_flags, s_comp, k > 0 ? 1 : 0))) break;
                sh.launches.fetch_add(g_choice.launches);
                if (!have_flags) {
                    my_flags = B200_MTM_FLAGS(g_choice.variant, g_choice.config + 1) | (my_flags & 0x7fff0000);
                    have_flags = true;
                    publish_flags(my_flags);
                }
                if ((e = cudaEventRecord(ctx->ev_done[k], s_comp)) != cudaSuccess ||
                    (e = cudaStreamWaitEvent(s_out, ctx->ev_done[k], 0)) != cudaSuccess ||
                    (e = stage_copy(pcl, dc, hcl, lo, hi_c, false, s_out, ctx)) != cudaSuccess) {
                    rc = fail(B200_ERR_CUDA, "copy-back of slab %d failed: %s", k, cudaGetErrorString(e));
                    break;
                }
            }
            if (rc) break;
            if ((e = drain_staged(*ctx)) != cudaSuccess || (e = cudaStreamSynchronize(s_out)) != cudaSuccess ||
                (e = cudaStreamSynchronize(s_comp)) != cudaSuccess) {
                rc = fail(B200_ERR_CUDA, "shard %d on device %d failed: %s", i, dev, cudaGetErrorString(e));
                break;
            }
        } while (0);
        if (rc) bail(rc);
    }
    if (i == 0) publish_flags(flags);                         // never leave the other shards waiting
    // my sends must have left before the caller may reuse / free anything
    if (ctx) (void)cudaStreamSynchronize(ctx->out_stream);
    bar.wait();
}

template <typename T>
int mtm_host_mgpu(T* c, const size_t* nc, const size_t* wc, const T* a, const size_t* na, const size_t* wa,
                  const T* b, const size_t* nb, const size_t* wb, int flags, const int* devices, int n_devices) {
    NvtxRange const range(sizeof(T) == 4 ? "b200_mtm_f32_mgpu" : "b200_mtm_f64_mgpu");
    int rc = validate(c, nc, wc, a, na, wa, b, nb, wb);
    if (rc) return rc;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        (void)cudaGetLastError();
        return fail(B200_ERR_CUDA, "no usable CUDA device (%s); libb200mtm has no CPU fallback", cudaGetErrorString(e));
    }
    std::vector<int> devs;
    if (devices != nullptr && n_devices > 0) {
        for (int i = 0; i < n_devices; ++i) {
            if (devices[i] < 0 || devices[i] >= count || devices[i] >= kMaxDevices)
                return fail(B200_ERR_INVALID, "b200_mtm_mgpu: device %d out of range (%d visible)", devices[i], count);
            for (int d : devs)
                if (d == devices[i]) return fail(B200_ERR_INVALID, "b200_mtm_mgpu: device %d listed twice", d);
            devs.push_back(devices[i]);
        }
    } else {
        int const want = n_devices > 0 ? std::min(n_devices, count) : count;
        for (int d = 0; d < want && d < kMaxDevices; ++d) devs.push_back(d);
    }
    int cur = 0;
    (void)cudaGetDevice(&cur);
    auto single = [&]() {
        (void)cudaSetDevice(devs[0]);
        int const r = mtm_host<T>(c, nc, wc, a, na, wa, b, nb, wb, flags);
        (void)cudaSetDevice(cur);
        return r;
    };
    if (nc[0] == 0 || nc[1] == 0 || na[1] == 0) return single();
    constexpr size_t V = 16 / sizeof(T);
    StagePlan const pa = plan_stage(na, wa, V), pb = plan_stage(nb, wb, V), pc = plan_stage(nc, wc, V);
    bool const row_contig = (wc[1] == 1) || nc[1] == 1;
    bool const col_contig = (wc[0] == 1) || nc[0] == 1;
    int const slice_dim = (row_contig || !col_contig) ? 0 : 1;
    size_t const extent = nc[slice_dim];
    // shards: whole multiples of the largest tile height, as many devices as there are such blocks
    int P = (int)std::min<size_t>(devs.size(), (extent + 255) / 256);
    bool const sliceable = pc.pitched && (slice_dim == 0 ? pa.pitched : pb.pitched);
    if (P <= 1 || !sliceable) return single();
    size_t per = (extent + P - 1) / P;
    per = (per + 255) / 256 * 256;
    P = (int)((extent + per - 1) / per);
    if (P <= 1) return single();
    devs.resize(P);
    std::vector<size_t> cut(P + 1);
    for (int i = 0; i <= P; ++i) cut[i] = std::min(extent, (size_t)i * per);

    MgpuShared sh;
    sh.P = P;
    sh.devs = devs;
    sh.ctx.assign(P, nullptr);
    sh.replica.assign(P, nullptr);
    sh.rc.assign(P, B200_OK);
    sh.err.assign(P, std::string());
    for (int i = 0; i < P && sh.gather; ++i)
        for (int j = 0; j < P; ++j) {
            if (i == j) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devs[i], devs[j]) != cudaSuccess || !can) {
                (void)cudaGetLastError();
                sh.gather = false;
                break;
            }
        }
    HostBarrier bar(P);
    std::vector<std::thread> th;
    th.reserve(P);
    for (int i = 0; i < P; ++i)
        th.emplace_back([&, i] { mgpu_worker<T>(sh, bar, i, c, nc, wc, a, na, wa, b, nb, wb, flags, slice_dim, cut, pa, pb, pc); });
    for (auto& t : th) t.join();
    (void)cudaSetDevice(cur);
    for (int i = 0; i < P; ++i)
        if (sh.rc[i] != B200_OK) {
            g_err = sh.err[i];
            return sh.rc[i];
        }
    g_choice = sh.choice;
    g_choice.launches = sh.launches.load();
    return B200_OK;
}

template <typename T>
int mtm_bench(T* c, const size_t* nc, const size_t* wc, const T* a, const size_t* na, const size_t* wa,
              const T* b, const size_t* nb, const size_t* wb, int flags, void* stream, int warmup,
              int iters, double* mean_ms) {
    if (!mean_ms || iters <= 0 || warmup < 0) return fail(B200_ERR_INVALID, "b200_mtm_bench: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int i = 0; i < warmup; ++i) {
        int rc = mtm_dev(c, nc, wc, a, na, wa, b, nb, wb, flags, stream);
        if (rc) return rc;
    }
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaEventRecord(e0, st));
    auto const h0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; ++i) {
        int rc = mtm_dev(c, nc, wc, a, na, wa, b, nb, wb, flags, stream);
        if (rc) {
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
            return rc;
        }
    }
    auto const h1 = std::chrono::steady_clock::now();
    CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *mean_ms = (double)ms / iters;
    // host time spent ENQUEUEING one call: when it is close to the device time the loop is launch-bound
    g_last_enqueue_us = std::chrono::duration<double, std::micro>(h1 - h0).count() / iters;
    return B200_OK;
}

// ---- matrix-times-vector ------------------------------------------------------------------------------
int launch_mtv(float* c, const float* a, int64_t M, int64_t K, int64_t si, int64_t sk, const float* b, int acc,
               DeviceCtx& ctx, cudaStream_t st, int* launches, const char** name) {
    CUDA_TRY(launch_mtv_f32(c, a, M, K, si, sk, b, acc, ctx.mtv_ws.ptr, ctx.mtv_ws.bytes, ctx.sm_count, st, launches, name));
    return B200_OK;
}
int launch_mtv(double* c, const double* a, int64_t M, int64_t K, int64_t si, int64_t sk, const double* b, int acc,
               DeviceCtx& ctx, cudaStream_t st, int* launches, const char** name) {
    CUDA_TRY(launch_mtv_f64(c, a, M, K, si, sk, b, acc, ctx.mtv_ws.ptr, ctx.mtv_ws.bytes, ctx.sm_count, st, launches, name));
    return B200_OK;
}

int validate_mtv(const void* c, const void* a, const size_t* na, const size_t* wa, const void* b) {
    if (!na || !wa) return fail(B200_ERR_INVALID, "b200_mtv: null extents/strides pointer");
    size_t const lim = (size_t)1 << 31;
    if (na[0] >= lim || na[1] >= lim) return fail(B200_ERR_INVALID, "b200_mtv: extents must be < 2^31");
    if (na[0] == 0) return B200_OK;
    if (!c || (na[1] != 0 && (!a || !b))) return fail(B200_ERR_INVALID, "b200_mtv: null data pointer");
    return B200_OK;
}

template <typename T>
int mtv_dev(T* c, const T* a, const size_t* na, const size_t* wa, const T* b, int a_last_order, int /*flags*/,
            void* stream) {
    int rc = validate_mtv(c, a, na, wa, b);
    if (rc) return rc;
    DeviceCtx* ctx;
    if ((rc = current_ctx(&ctx))) return rc;
    if (na[0] == 0) {
        record_choice(B200_MTM_SIMT, 0, "noop_empty", 0, 0, 0);
        return B200_OK;
    }
    cudaStream_t const st = static_cast<cudaStream_t>(stream);
    if ((rc = acquire(ctx->mtv_ws, mtv_workspace_bytes((int64_t)na[0], (int)sizeof(T), ctx->sm_count), st))) return rc;
    int launches = 0;
    const char* name = "mtv";
    // first_order path accumulates, last_order path assigns (mtv.hpp:15-100)
    if ((rc = launch_mtv(c, a, (int64_t)na[0], (int64_t)na[1], (int64_t)wa[0], (int64_t)wa[1], b, a_last_order ? 0 : 1,
                         *ctx, st, &launches, &name)))
        return rc;
    if ((rc = release(ctx->mtv_ws, st))) return rc;
    record_choice(B200_MTM_SIMT, 0, name, launches, wa[0] == 1 ? 0 : (wa[1] == 1 ? 1 : 2), 0);
    return B200_OK;
}

template <typename T>
int mtv_host(T* c, const T* a, const size_t* na, const size_t* wa, const T* b, int a_last_order, int flags) {
    int rc = validate_mtv(c, a, na, wa, b);
    if (rc) return rc;
    DeviceCtx* ctxp;
    if ((rc = current_ctx(&ctxp))) return rc;
    DeviceCtx& ctx = *ctxp;
    if (na[0] == 0) return B200_OK;
    constexpr size_t V = 16 / sizeof(T);
    size_t const M = na[0], K = na[1];
    size_t const n1[2] = {M, K ? K : 1};
    StagePlan const pa = plan_stage(n1, wa, V);
    size_t const vb = (K + V - 1) / V * V + V, vc = (M + V - 1) / V * V + V;
    if ((rc = ensure(ctx.stage[0], (K ? pa.dev_elems : 1) * sizeof(T)))) return rc;
    if ((rc = ensure(ctx.stage[1], vb * sizeof(T)))) return rc;
    if ((rc = ensure(ctx.stage[2], vc * sizeof(T)))) return rc;
    T* da = static_cast<T*>(ctx.stage[0].ptr);
    T* db = static_cast<T*>(ctx.stage[1].ptr);
    T* dc = static_cast<T*>(ctx.stage[2].ptr);
    cudaStream_t const st = ctx.host_stream;
    size_t const zero[2] = {0, 0};
    if (K) {
        CUDA_TRY(stage_copy(pa, da, const_cast<T*>(a), zero, pa.n, true, st));
        CUDA_TRY(cudaMemcpyAsync(db, b, K * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    if (!a_last_order) CUDA_TRY(cudaMemcpyAsync(dc, c, M * sizeof(T), cudaMemcpyHostToDevice, st));  // accumulating path reads c
    if ((rc = mtv_dev<T>(dc, da, na, pa.dev_w, db, a_last_order, flags, st))) return rc;
    CUDA_TRY(cudaMemcpyAsync(c, dc, M * sizeof(T), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return B200_OK;
}

template <typename T>
int mtv_bench(T* c, const T* a, const size_t* na, const size_t* wa, const T* b, int a_last_order, int flags,
              void* stream, int warmup, int iters, double* mean_ms) {
    if (!mean_ms || iters <= 0 || warmup < 0) return fail(B200_ERR_INVALID, "b200_mtv_bench: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int i = 0; i < warmup; ++i) {
        int rc = mtv_dev<T>(c, a, na, wa, b, a_last_order, flags, stream);
        if (rc) return rc;
    }
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) {
        int rc = mtv_dev<T>(c, a, na, wa, b, a_last_order, flags, stream);
        if (rc) {
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
            return rc;
        }
    }
    CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *mean_ms = (double)ms / iters;
    return B200_OK;
}

// ---- transpose --------------------------------------------------------------------------------------------
cudaError_t launch_transpose(float* c, const float* a, int64_t M, int64_t N, int64_t sai, int64_t saj, int64_t scj,
                             int64_t sci, cudaStream_t st) {
    return launch_transpose_f32(c, a, M, N, sai, saj, scj, sci, st);
}
cudaError_t launch_transpose(double* c, const double* a, int64_t M, int64_t N, int64_t sai, int64_t saj, int64_t scj,
                             int64_t sci, cudaStream_t st) {
    return launch_transpose_f64(c, a, M, N, sai, saj, scj, sci, st);
}
cudaError_t launch_transpose_inplace(float* a, int64_t n, cudaStream_t st) { return launch_transpose_inplace_f32(a, n, st); }
cudaError_t launch_transpose_inplace(double* a, int64_t n, cudaStream_t st) { return launch_transpose_inplace_f64(a, n, st); }

int validate_transpose(const void* c, const size_t* nc, const size_t* wc, const void* a, const size_t* na,
                       const size_t* wa) {
    if (!nc || !wc || !na || !wa) return fail(B200_ERR_INVALID, "b200_transpose: null extents/strides pointer");
    if (!((na[0] == nc[1]) && (na[1] == nc[0])))      // trans.hpp:121-126
        return fail(B200_ERR_DIM, "b200_transpose: dimension mismatch: c %zux%zu, a %zux%zu", nc[0], nc[1], na[0], na[1]);
    size_t const lim = (size_t)1 << 31;
    if (na[0] >= lim || na[1] >= lim) return fail(B200_ERR_INVALID, "b200_transpose: extents must be < 2^31");
    if (na[0] == 0 || na[1] == 0) return B200_OK;
    if (!c || !a) return fail(B200_ERR_INVALID, "b200_transpose: null data pointer");
    return B200_OK;
}

template <typename T>
int transpose_dev(T* c, const size_t* nc, const size_t* wc, const T* a, const size_t* na, const size_t* wa,
                  void* stream) {
    int rc = validate_transpose(c, nc, wc, a, na, wa);
    if (rc) return rc;
    DeviceCtx* ctx;
    if ((rc = current_ctx(&ctx))) return rc;
    if (na[0] == 0 || na[1] == 0) {
        record_choice(B200_MTM_SIMT, 0, "noop_empty", 0, 0, 0);
        return B200_OK;
    }
    CUDA_TRY(launch_transpose(c, a, (int64_t)na[0], (int64_t)na[1], (int64_t)wa[0], (int64_t)wa[1], (int64_t)wc[0],
                              (int64_t)wc[1], static_cast<cudaStream_t>(stream)));
    record_choice(B200_MTM_SIMT, 0, "transpose_tile64", 1, wa[1] <= wa[0] ? 1 : 0, wc[1] <= wc[0] ? 1 : 0);
    return B200_OK;
}

template <typename T>
int transpose_host(T* c, const size_t* nc, const size_t* wc, const T* a, const size_t* na, const size_t* wa) {
    int rc = validate_transpose(c, nc, wc, a, na, wa);
    if (rc) return rc;
    DeviceCtx* ctxp;
    if ((rc = current_ctx(&ctxp))) return rc;
    DeviceCtx& ctx = *ctxp;
    if (na[0] == 0 || na[1] == 0) return B200_OK;
    constexpr size_t V = 16 / sizeof(T);
    StagePlan const pa = plan_stage(na, wa, V), pc = plan_stage(nc, wc, V);
    if ((rc = ensure(ctx.stage[0], pa.dev_elems * sizeof(T)))) return rc;
    if ((rc = ensure(ctx.stage[2], pc.dev_elems * sizeof(T)))) return rc;
    T* da = static_cast<T*>(ctx.stage[0].ptr);
    T* dc = static_cast<T*>(ctx.stage[2].ptr);
    cudaStream_t const st = ctx.host_stream;
    size_t const zero[2] = {0, 0};
    CUDA_TRY(stage_copy(pa, da, const_cast<T*>(a), zero, pa.n, true, st));
    if (!pc.pitched) CUDA_TRY(stage_copy(pc, dc, c, zero, pc.n, true, st));   // span copy-back must preserve the gaps
    if ((rc = transpose_dev<T>(dc, nc, pc.dev_w, da, na, pa.dev_w, st))) return rc;
    CUDA_TRY(stage_copy(pc, dc, c, zero, pc.n, false, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return B200_OK;
}

template <typename T>
int transpose_inplace_dev(T* a, const size_t* na, void* stream) {
    if (!na) return fail(B200_ERR_INVALID, "b200_transpose_inplace: null extents pointer");
    if (na[0] != na[1])
        return fail(B200_ERR_DIM, "b200_transpose_inplace: %zux%zu is not square (the reference's in-place swap, "
                                  "trans.hpp:63-92, is only meaningful for square matrices)", na[0], na[1]);
    if (na[0] >= ((size_t)1 << 31)) return fail(B200_ERR_INVALID, "b200_transpose_inplace: extent must be < 2^31");
    if (na[0] == 0) return B200_OK;
    if (!a) return fail(B200_ERR_INVALID, "b200_transpose_inplace: null data pointer");
    DeviceCtx* ctx;
    int rc = current_ctx(&ctx);
    if (rc) return rc;
    CUDA_TRY(launch_transpose_inplace(a, (int64_t)na[0], static_cast<cudaStream_t>(stream)));
    record_choice(B200_MTM_SIMT, 0, "transpose_inplace_tile32", na[0] > 1 ? 1 : 0, 0, 0);
    return B200_OK;
}

template <typename T>
int transpose_inplace_host(T* a, const size_t* na) {
    if (!na) return fail(B200_ERR_INVALID, "b200_transpose_inplace: null extents pointer");
    if (na[0] != na[1]) return transpose_inplace_dev<T>(a, na, nullptr);   // reports the error
    if (na[0] == 0) return B200_OK;
    if (!a) return fail(B200_ERR_INVALID, "b200_transpose_inplace: null data pointer");
    DeviceCtx* ctxp;
    int rc = current_ctx(&ctxp);
    if (rc) return rc;
    DeviceCtx& ctx = *ctxp;
    size_t const bytes = na[0] * na[1] * sizeof(T);
    if ((rc = ensure(ctx.stage[0], bytes))) return rc;
    T* da = static_cast<T*>(ctx.stage[0].ptr);
    cudaStream_t const st = ctx.host_stream;
    CUDA_TRY(cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, st));
    if ((rc = transpose_inplace_dev<T>(da, na, st))) return rc;
    CUDA_TRY(cudaMemcpyAsync(a, da, bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return B200_OK;
}

template <typename T>
int transpose_bench(T* c, const size_t* nc, const size_t* wc, const T* a, const size_t* na, const size_t* wa,
                    void* stream, int warmup, int iters, double* mean_ms) {
    if (!mean_ms || iters <= 0 || warmup < 0) return fail(B200_ERR_INVALID, "b200_transpose_bench: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int i = 0; i < warmup; ++i) {
        int rc = transpose_dev<T>(c, nc, wc, a, na, wa, stream);
        if (rc) return rc;
    }
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) {
        int rc = transpose_dev<T>(c, nc, wc, a, na, wa, stream);
        if (rc) {
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
            return rc;
        }
    }
    CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *mean_ms = (double)ms / iters;
    return B200_OK;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int b200_mtm_f32(float* c, const size_t nc[2], const size_t wc[2], const float* a, const size_t na[2],
                 const size_t wa[2], const float* b, const size_t nb[2], const size_t wb[2], int flags) {
    return mtm_host<float>(c, nc, wc, a, na, wa, b, nb, wb, flags);
}
int b200_mtm_f64(double* c, const size_t nc[2], const size_t wc[2], const double* a, const size_t na[2],
                 const size_t wa[2], const double* b, const size_t nb[2], const size_t wb[2], int flags) {
    return mtm_host<double>(c, nc, wc, a, na, wa, b, nb, wb, flags);
}
int b200_mtm_f32_mgpu(float* c, const size_t nc[2], const size_t wc[2], const float* a, const size_t na[2],
                      const size_t wa[2], const float* b, const size_t nb[2], const size_t wb[2], int flags,
                      const int* devices, int n_devices) {
    return mtm_host_mgpu<float>(c, nc, wc, a, na, wa, b, nb, wb, flags, devices, n_devices);
}
int b200_mtm_f64_mgpu(double* c, const size_t nc[2], const size_t wc[2], const double* a, const size_t na[2],
                      const size_t wa[2], const double* b, const size_t nb[2], const size_t wb[2], int flags,
                      const int* devices, int n_devices) {
    return mtm_host_mgpu<double>(c, nc, wc, a, na, wa, b, nb, wb, flags, devices, n_devices);
}
int b200_mtm_f32_dev(float* c, const size_t nc[2], const size_t wc[2], const float* a,
                     const size_t na[2], const size_t wa[2], const float* b, const size_t nb[2],
                     const size_t wb[2], int flags, void* stream) {
    return mtm_dev<float>(c, nc, wc, a, na, wa, b, nb, wb, flags, stream);
}
int b200_mtm_f64_dev(double* c, const size_t nc[2], const size_t wc[2], const double* a,
                     const size_t na[2], const size_t wa[2], const double* b, const size_t nb[2],
                     const size_t wb[2], int flags, void* stream) {
    return mtm_dev<double>(c, nc, wc, a, na, wa, b, nb, wb, flags, stream);
}
int b200_mtm_bench_f32_dev(float* c, const size_t nc[2], const size_t wc[2], const float* a,
                           const size_t na[2], const size_t wa[2], const float* b, const size_t nb[2],
                           const size_t wb[2], int flags, void* stream, int warmup, int iters,
                           double* mean_ms) {
    return mtm_bench<float>(c, nc, wc, a, na, wa, b, nb, wb, flags, stream, warmup, iters, mean_ms);
}
int b200_mtm_bench_f64_dev(double* c, const size_t nc[2], const size_t wc[2], const double* a,
                           const size_t na[2], const size_t wa[2], const double* b, const size_t nb[2],
                           const size_t wb[2], int flags, void* stream, int warmup, int iters,
                           double* mean_ms) {
    return mtm_bench<double>(c, nc, wc, a, na, wa, b, nb, wb, flags, stream, warmup, iters, mean_ms);
}

int b200_mtm_last_choice(b200_mtm_choice* out) {
    if (!out) return fail(B200_ERR_INVALID, "b200_mtm_last_choice: null output");
    *out = g_choice;
    return B200_OK;
}

int b200_mtm_plan_f32(size_t M, size_t N, size_t K, int sm_count, b200_mtm_choice* out) {
    if (!out || M == 0 || N == 0 || K == 0) return fail(B200_ERR_INVALID, "b200_mtm_plan_f32: bad arguments");
    if (sm_count <= 0) {
        DeviceCtx* ctx;
        int const rc = current_ctx(&ctx);
        if (rc) return rc;
        sm_count = ctx->sm_count;
    }
    MtmShape s{};
    s.M = (int64_t)M;
    s.N = (int64_t)N;
    s.K = (int64_t)K;
    s.ldc = (int64_t)N;
    s.a_sm = (int64_t)K;
    s.a_sk = 1;
    s.b_sk = (int64_t)N;
    s.b_sn = 1;
    std::memset(out, 0, sizeof *out);
    out->variant = auto_variant_f32(s);
    out->config = out->variant == B200_MTM_3XTF32 ? pick_tf32_config(s, sm_count) : auto_config_simt_f32(s, sm_count);
    std::snprintf(out->name, sizeof out->name, "%s", b200_mtm_config_name(out->variant, 0, out->config));
    return B200_OK;
}

int b200_mtm_num_configs(int variant, int is_f64) {
    switch (variant) {
        case B200_MTM_SIMT: return is_f64 ? simt_f64_num_configs() : simt_f32_num_configs() + ffma_tma_num_configs();
        case B200_MTM_DFMA: return is_f64 ? simt_f64_num_configs() : 0;
        case B200_MTM_DMMA: return is_f64 ? dmma_f64_num_configs() + dmma_tma_num_configs() : 0;
        case B200_MTM_3XTF32: return is_f64 ? 0 : tf32_num_configs();
        default: return 0;
    }
}

const char* b200_mtm_config_name(int variant, int is_f64, int config) {
    if (config < 0 || config >= b200_mtm_num_configs(variant, is_f64)) return "";
    switch (variant) {
        case B200_MTM_SIMT:
            if (is_f64) return simt_f64_config(config).name;
            return config < simt_f32_num_configs() ? simt_f32_config(config).name
                                                   : ffma_tma_config(config - simt_f32_num_configs()).name;
        case B200_MTM_DFMA: return simt_f64_config(config).name;
        case B200_MTM_DMMA:
            return config < dmma_f64_num_configs() ? dmma_f64_config(config).name
                                                   : dmma_tma_config(config - dmma_f64_num_configs()).name;
        case B200_MTM_3XTF32: return tf32_config(config).name;
        default: return "";
    }
}

uint64_t b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

double b200_last_bench_enqueue_us(void) { return g_last_enqueue_us; }

int b200_device_count(int* count) {
    if (!count) return fail(B200_ERR_INVALID, "b200_device_count: null output");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        *count = 0;
        return fail(B200_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    return B200_OK;
}

int b200_set_device(int device) {
    CUDA_TRY(cudaSetDevice(device));
    return B200_OK;
}

int b200_get_device_info(int device, b200_device_info* out) {
    if (!out) return fail(B200_ERR_INVALID, "b200_get_device_info: null output");
    cudaDeviceProp p;
    CUDA_TRY(cudaGetDeviceProperties(&p, device));
    std::memset(out, 0, sizeof *out);
    std::snprintf(out->name, sizeof out->name, "%s", p.name);
    out->cc_major = p.major;
    out->cc_minor = p.minor;
    out->sm_count = p.multiProcessorCount;
    int clk = 0, mclk = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, device));
    CUDA_TRY(cudaDeviceGetAttribute(&mclk, cudaDevAttrMemoryClockRate, device));
    out->sm_clock_khz = clk;
    out->mem_clock_khz = mclk;
    out->mem_bus_bits = p.memoryBusWidth;
    out->smem_per_sm = p.sharedMemPerMultiprocessor;
    out->smem_per_block_optin = p.sharedMemPerBlockOptin;
    out->l2_bytes = (size_t)p.l2CacheSize;
    out->hbm_bytes = p.totalGlobalMem;
    // sm_100: 128 FP32 lanes and 64 FP64 lanes per SM, one FMA (2 flop) per lane per clock.
    out->peak_fp32_tflops = (double)p.multiProcessorCount * 128.0 * 2.0 * (double)clk * 1e3 / 1e12;
    out->peak_fp64_tflops = (double)p.multiProcessorCount * 64.0 * 2.0 * (double)clk * 1e3 / 1e12;
    return B200_OK;
}

int b200_malloc(void** dptr, size_t bytes) {
    if (!dptr) return fail(B200_ERR_INVALID, "b200_malloc: null output");
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? B200_ERR_NOMEM : B200_ERR_CUDA,
                    "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    return B200_OK;
}
int b200_free(void* dptr) {
    if (dptr) CUDA_TRY(cudaFree(dptr));
    return B200_OK;
}
int b200_host_alloc(void** hptr, size_t bytes) {
    if (!hptr) return fail(B200_ERR_INVALID, "b200_host_alloc: null output");
    cudaError_t e = cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? B200_ERR_NOMEM : B200_ERR_CUDA,
                    "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    return B200_OK;
}
int b200_host_free(void* hptr) {
    if (hptr) CUDA_TRY(cudaFreeHost(hptr));
    return B200_OK;
}
int b200_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    return B200_OK;
}
int b200_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    return B200_OK;
}
int b200_memset(void* dptr, int value, size_t bytes, void* stream) {
    CUDA_TRY(cudaMemsetAsync(dptr, value, bytes, static_cast<cudaStream_t>(stream)));
    return B200_OK;
}
int b200_stream_create(void** stream) {
    if (!stream) return fail(B200_ERR_INVALID, "b200_stream_create: null output");
    cudaStream_t s;
    CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return B200_OK;
}
int b200_stream_destroy(void* stream) {
    if (stream) CUDA_TRY(cudaStreamDestroy(static_cast<cudaStream_t>(stream)));
    return B200_OK;
}
int b200_stream_synchronize(void* stream) {
    CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return B200_OK;
}
int b200_device_synchronize(void) {
    CUDA_TRY(cudaDeviceSynchronize());
    return B200_OK;
}

int b200_mtv_f32(float* c, const float* a, const size_t na[2], const size_t wa[2], const float* b, int a_last_order,
                 int flags) {
    return mtv_host<float>(c, a, na, wa, b, a_last_order, flags);
}
int b200_mtv_f64(double* c, const double* a, const size_t na[2], const size_t wa[2], const double* b, int a_last_order,
                 int flags) {
    return mtv_host<double>(c, a, na, wa, b, a_last_order, flags);
}
int b200_mtv_f32_dev(float* c, const float* a, const size_t na[2], const size_t wa[2], const float* b,
                     int a_last_order, int flags, void* stream) {
    return mtv_dev<float>(c, a, na, wa, b, a_last_order, flags, stream);
}
int b200_mtv_f64_dev(double* c, const double* a, const size_t na[2], const size_t wa[2], const double* b,
                     int a_last_order, int flags, void* stream) {
    return mtv_dev<double>(c, a, na, wa, b, a_last_order, flags, stream);
}
int b200_mtv_bench_f32_dev(float* c, const float* a, const size_t na[2], const size_t wa[2], const float* b,
                           int a_last_order, int flags, void* stream, int warmup, int iters, double* mean_ms) {
    return mtv_bench<float>(c, a, na, wa, b, a_last_order, flags, stream, warmup, iters, mean_ms);
}
int b200_mtv_bench_f64_dev(double* c, const double* a, const size_t na[2], const size_t wa[2], const double* b,
                           int a_last_order, int flags, void* stream, int warmup, int iters, double* mean_ms) {
    return mtv_bench<double>(c, a, na, wa, b, a_last_order, flags, stream, warmup, iters, mean_ms);
}

int b200_transpose_f32(float* c, const size_t nc[2], const size_t wc[2], const float* a, const size_t na[2],
                       const size_t wa[2], int) {
    return transpose_host<float>(c, nc, wc, a, na, wa);
}
int b200_transpose_f64(double* c, const size_t nc[2], const size_t wc[2], const double* a, const size_t na[2],
                       const size_t wa[2], int) {
    return transpose_host<double>(c, nc, wc, a, na, wa);
}
int b200_transpose_inplace_f32(float* a, const size_t na[2], int) { return transpose_inplace_host<float>(a, na); }
int b200_transpose_inplace_f64(double* a, const size_t na[2], int) { return transpose_inplace_host<double>(a, na); }
int b200_transpose_f32_dev(float* c, const size_t nc[2], const size_t wc[2], const float* a, const size_t na[2],
                           const size_t wa[2], int, void* stream) {
    return transpose_dev<float>(c, nc, wc, a, na, wa, stream);
}
int b200_transpose_f64_dev(double* c, const size_t nc[2], const size_t wc[2], const double* a, const size_t na[2],
                           const size_t wa[2], int, void* stream) {
    return transpose_dev<double>(c, nc, wc, a, na, wa, stream);
}
int b200_transpose_inplace_f32_dev(float* a, const size_t na[2], int, void* stream) {
    return transpose_inplace_dev<float>(a, na, stream);
}
int b200_transpose_inplace_f64_dev(double* a, const size_t na[2], int, void* stream) {
    return transpose_inplace_dev<double>(a, na, stream);
}
int b200_transpose_bench_f32_dev(float* c, const size_t nc[2], const size_t wc[2], const float* a, const size_t na[2],
                                 const size_t wa[2], int, void* stream, int warmup, int iters, double* mean_ms) {
    return transpose_bench<float>(c, nc, wc, a, na, wa, stream, warmup, iters, mean_ms);
}
int b200_transpose_bench_f64_dev(double* c, const size_t nc[2], const size_t wc[2], const double* a, const size_t na[2],
                                 const size_t wa[2], int, void* stream, int warmup, int iters, double* mean_ms) {
    return transpose_bench<double>(c, nc, wc, a, na, wa, stream, warmup, iters, mean_ms);
}

// ---- operand replication (include/b200_replicate.h) ---------------------------------------------
int b200_replicate_push_2d(void* const* dst, int n_dst, int multicast, const void* src, size_t rows, size_t row_bytes,
                           size_t src_pitch, size_t dst_pitch, uint32_t* const* flag_dst, int n_flag_dst,
                           int flag_multicast, uint32_t flag_value, int ctas, void* stream) {
    if (!dst || !src || (n_flag_dst > 0 && !flag_dst)) return fail(B200_ERR_INVALID, "b200_replicate_push_2d: null pointer");
    CUDA_TRY(launch_replicate_push_2d(dst, n_dst, multicast, src, rows, row_bytes, src_pitch, dst_pitch, flag_dst,
                                      n_flag_dst, flag_multicast, flag_value, ctas, static_cast<cudaStream_t>(stream)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return B200_OK;
}
int b200_replicate_push(void* const* dst, int n_dst, int multicast, const void* src, size_t bytes,
                        uint32_t* const* flag_dst, int n_flag_dst, int flag_multicast, uint32_t flag_value,
                        int ctas, void* stream) {
    if (!dst || !src || (n_flag_dst > 0 && !flag_dst)) return fail(B200_ERR_INVALID, "b200_replicate_push: null pointer");
    if (bytes == 0 && n_flag_dst == 0) return B200_OK;
    CUDA_TRY(launch_replicate_push(dst, n_dst, multicast, src, bytes, flag_dst, n_flag_dst, flag_multicast, flag_value,
                                   ctas, static_cast<cudaStream_t>(stream)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return B200_OK;
}
int b200_flag_wait(const uint32_t* flag, uint32_t value, int count, int stride, int skip, void* stream) {
    if (!flag) return fail(B200_ERR_INVALID, "b200_flag_wait: null pointer");
    CUDA_TRY(launch_flag_wait(flag, value, count, stride, skip, static_cast<cudaStream_t>(stream)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return B200_OK;
}
int b200_flag_signal(uint32_t* flag, uint32_t value, void* stream) {
    if (!flag) return fail(B200_ERR_INVALID, "b200_flag_signal: null pointer");
    CUDA_TRY(launch_flag_signal(flag, value, static_cast<cudaStream_t>(stream)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return B200_OK;
}

const char* b200_last_error(void) { return g_err.c_str(); }

int b200_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) {
        (void)cudaGetLastError();
        return B200_OK;
    }
    for (int d = 0; d < kMaxDevices; ++d) {
        DeviceCtx& c = g_ctx[d];
        if (!c.ready) continue;
        cudaSetDevice(d);
        cudaDeviceSynchronize();
        for (Buffer* b : {&c.stage[0], &c.stage[1], &c.stage[2], &c.tf32_ws, &c.pack_ws, &c.mtv_ws}) {
            if (b->ptr) cudaFree(b->ptr);            // also valid for stream-ordered allocations (synchronises)
            if (b->last_use) cudaEventDestroy(b->last_use);
            *b = Buffer{};
        }
        for (auto& ev : c.ev_in)
            if (ev) cudaEventDestroy(ev);
        for (auto& ev : c.ev_done)
            if (ev) cudaEventDestroy(ev);
        for (auto* ring : {c.up, c.down})
            for (int i = 0; i < DeviceCtx::kBounce; ++i) {
                if (ring[i].ptr) cudaFreeHost(ring[i].ptr);
                if (ring[i].ev) cudaEventDestroy(ring[i].ev);
            }
        for (auto& ev : c.ev_sent)
            if (ev) cudaEventDestroy(ev);
        if (c.ev_slice) cudaEventDestroy(c.ev_slice);
        if (c.out_stream) cudaStreamDestroy(c.out_stream);
        if (c.host_stream) cudaStreamDestroy(c.host_stream);
        if (c.copy_stream) cudaStreamDestroy(c.copy_stream);
        c = DeviceCtx{};
    }
    cudaSetDevice(cur);
    return B200_OK;
}

}  // extern "C"
