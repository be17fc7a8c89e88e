// fp32 mtm on the 5th-generation tensor cores: 3xTF32 with tcgen05.mma / TMEM / TMA (sm_100a).
//
//   C += A*B  with every fp32 operand split as x = hi + lo and
//   A*B ~= Alo*Bhi + Ahi*Blo + Ahi*Bhi   (the lo*lo term is dropped), all three products accumulated in
//   fp32 in tensor memory.
//
// The split.  tcgen05.mma kind::tf32 reads fp32 bit patterns from shared memory and ignores the low 13
// mantissa bits, so the RAW operand already is the `hi` factor (hi = trunc_tf32(x)) and only
// lo = rn_tf32(x - trunc_tf32(x)) has to be materialised.  An operand the TMA can fetch as it lies in HBM
// (16-byte aligned, unit stride in one dimension, the other stride a multiple of 4 elements) is therefore
// used IN PLACE as hi — K-contiguous operands through K-major shared-memory descriptors, m/n-contiguous
// ones (a row-major B, a column-major A) through MN-major descriptors (SWIZZLE_128B_BASE32B, fed by TMA boxes
// with 32-byte swizzle atoms), no transpose anywhere — and the
// pre-pass writes one plane (lo, same layout as the operand) instead of two.  Operands with arbitrary
// strides / alignment are gathered into K-contiguous hi + lo planes as before ("packed").  This replaces
// the reference's pack step (include/utils.hpp:99-141, called from mtm.hpp:169-199).
//
// Two launches per call (plane-fed configs, the default), chained by programmatic dependent launch: the MMA kernel's
// CTAs set themselves up while the split still runs and wait for the planes in the kernel; the split of the NEXT call is
// scheduled while this call's MMA kernel runs and waits for it the same way (griddepcontrol.wait / .launch_dependents).
//   1. split_kernel — ONE launch covering both operands (A's blocks, then B's): the lo planes.
//   2. mtm_tf32x3_kernel — persistent, warp-specialised GEMM: warp 0 = TMA producer (Ahi, Alo, Bhi, Blo
//      tiles, 128B-swizzled, 3 stages — 4 with the narrow B tiles, 2 with double tiles), warp 1 = tcgen05.mma issuer
//      (warp-uniform loop, one elected lane issues; 3 MMAs per 8-wide k step, fp32 accumulators in TMEM, two
//      accumulator buffers so the epilogue of tile i overlaps the main loop of tile i+1), warp 2 = TMEM allocator,
//      warps 4-7 = epilogue: tcgen05.ld -> swizzled
//      shared memory -> TMA reduce-add into C (cp.reduce.async.bulk.tensor ... .add, SASS UTMAREDG: the L2 does
//      C += acc, C never enters the SM, edges are clipped by the tensor map) — the reference's
//      copy_from_buff (simd_loop.hpp:160-190).  A C that is not 16-byte aligned / ldc % 4 != 0 is
//      read-modify-written through registers, one 128-byte line per warp instruction.  NCTA = 2 pairs two SMs
//      on one tile (cta_group::2): each CTA stages its own 128 rows of A and its half of B's columns, halving
//      shared-memory reads per SM; the copies of both CTAs complete on the leader's barrier and only the leader arrives
//      on it (a per-k-block remote arrive of the peer bounded every pair config at ~1 us per k-block).  Double tiles
//      (256 x 512 per pair): a work unit is two neighbouring 256 x 256 tiles that share their A tiles, one accumulator
//      buffer each — a quarter less L2 -> SM traffic per flop, for the largest problems.
// Work units: whole tiles handed out statically (or by a dynamic counter for runs that share SMs with a
// collective); problems with few tiles split every tile along K, problems with a ragged last wave split only
// that wave's tiles (tail split); the units of a tile add into C in a fixed order (turnstile).  Opt-in and
// measured slower on B200: stream-K ranges; the FUSED configs, whose extra converter warps
// compute the lo tiles in shared memory instead of reading planes (one launch, shared-memory-bandwidth bound).
//
// Exactness: integers of magnitude < 2^11 are exact in TF32 (lo == 0) and the fp32 accumulation of exact
// products is exact while partial sums stay below 2^24, so the reference's integer test cases are
// reproduced bit for bit.
#include <cstdlib>

#include "mtm_kernels.h"
#include "sm100_ptx.cuh"

namespace b200 {
namespace {

using namespace ptx;

constexpr int TILE_R = 128;                    // rows of A each CTA stages per k-block (and at most as many of B^T)
constexpr int BK = 32;                         // 32 fp32 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 8;                      // kind::tf32 consumes 32 bytes of K per MMA
constexpr int STAGES = 3;                      // with full-width (128-row) B tiles: 3 stages of 64 KiB
constexpr int MAX_STAGES = 4;                  // narrow B tiles (64 rows) make a stage 48 KiB: 4 of them fit the same ring
constexpr int TILE_BYTES = TILE_R * BK * 4;    // 16 KiB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;    // Ahi, Alo, Bhi, Blo
constexpr int MN_CHUNK = 32;                   // MN-major staging: one TMA box = 32 (m/n) x BK (k), 128-byte rows
constexpr int MN_CHUNK_BYTES = MN_CHUNK * BK * 4;   // 4 KiB = BK rows (k) of 128 bytes
constexpr int MN_ATOM_BYTES = 512;             // MN-major swizzle atom (SWIZZLE_128B_BASE32B): 4 k-rows of 128 bytes
constexpr int MN_KSTEP_BYTES = UMMA_K * 128;   // one MMA consumes 8 k-rows = 2 atoms
constexpr int NUM_THREADS = 256;
// FUSED kernels split the same 192 KiB differently: a deeper ring of RAW stages (A tile + B tile as fetched) and a
// short ring of LO stages the converter warps fill — the TMA then runs as far ahead of the MMA as in the plane-fed
// kernel although a conversion step sits in between.
constexpr int RAW_STAGES = 4, LO_STAGES = 2;
constexpr int RAW_STAGE_BYTES = 2 * TILE_BYTES;   // A raw, B raw
constexpr int LO_STAGE_BYTES = 2 * TILE_BYTES;    // A lo, B lo
static_assert(RAW_STAGES * RAW_STAGE_BYTES + LO_STAGES * LO_STAGE_BYTES == STAGES * STAGE_BYTES, "same shared-memory footprint");
constexpr int ACC_STAGES = 2;
constexpr int SCHED_STAGES = 4;                // ring of tile indices handed out by the dynamic scheduler
constexpr int EPI_WARPS = 4;
constexpr int EPI_BOX = 32;                    // epilogue box: 32 rows x 32 columns of C per TMA reduce
constexpr int EPI_BUF_BYTES = EPI_BOX * EPI_BOX * 4;
constexpr int EPI_BUFS = 2;                    // per warp: fill one while the TMA reads the other
constexpr int EPI_BYTES = EPI_WARPS * EPI_BUFS * EPI_BUF_BYTES;   // 32 KiB
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 512 /*barriers, scheduler ring*/;

struct Tf32Params {
    float* C;
    int64_t ldc;
    int M, N;
    int num_k_blocks;
    int tiles_m, tiles_n;
    int* tile_counter;   // DYNAMIC: next unclaimed tile (initialised to the number of CTA groups)
    int group;           // tile-rows walked together before moving to the next tile-column
    int bn_cta;          // rows of B^T (columns of C) each CTA stages: 128 or 64; the tile is 128*NCTA x bn_cta*NCTA
    int a_mn, b_mn;      // operand tiles are MN-major in shared memory (else K-major)
    int c_tma;           // epilogue: TMA reduce-add (else register read-modify-write)
    int peer_arrive;     // (measurement aid) pairs: the peer CTA's producer also arrives on the leader's full barrier
    int dbl;             // double tiles (pairs only): a work unit is two neighbouring 256 x 256 tiles, 256 x 512, that share
                         // their A tiles — one accumulator buffer each, B staged for both: 3/4 of the L2 -> SM bytes per flop
    uint32_t mn_lbo, mn_sbo;   // MN-major descriptor strides in bytes (between atoms along m/n, along k)
    uint32_t mn_layout;        // MN-major descriptor layout type (1 = SWIZZLE_128B_BASE32B)
    // Split-K (small problems: too few tiles for the machine): a work unit is (tile, split s), split s covers
    // k-blocks [s * kb_per_split, ...).  The splits of a tile add into C in the order s = 0, 1, ... — a turnstile
    // word per (tile, 32-row block) — so the result does not depend on which split finishes first.
    int split_k, kb_per_split;
    int64_t split_from;  // units below this index are whole tiles; from here on every tile is cut into split_k units
                         // (0: the whole problem is split; tiles - tail: only the ragged LAST WAVE is — its tiles then
                         // take 1/split_k of a tile-time while all groups still walk K in lock-step)
    uint32_t* turn;
    // Stream-K (static scheduling, when whole tiles would leave a ragged last wave): the T * nkb k-block
    // iterations are cut into equal contiguous ranges of sk_width, one per CTA group; a tile cut by a range
    // boundary is finished by two (or more) groups, which add their parts into C in a fixed order (highest
    // group first: that is also the order in which they get there) through the same turnstile words.
    int stream_k;
    int64_t sk_width, sk_total;
};

// Tile hand-out.  STATIC: CTA group g takes tiles g, g + G, g + 2G, ...  DYNAMIC: the first tile is
// g, every further one comes from a global counter — one scheduler thread per CTA group claims it and
// publishes it through a small shared-memory ring (to both CTAs of a pair), so a group that starts
// late or runs slowly (e.g. SMs shared with a concurrent NCCL broadcast) simply takes fewer tiles.
template <int NCTA, bool DYNAMIC>
struct TileSource {
    int it = 0;
    int64_t cur = -1, end = 0;      // stream-K: this group's range of global k-block indices
    // Stream-K: hand out the START index of the next item of this group's range (an item never crosses a tile).
    __device__ __forceinline__ int64_t next_stream_k(const Tf32Params& p, int group_id) {
        if (cur < 0) {
            cur = (int64_t)group_id * p.sk_width;
            end = cur + p.sk_width;
            if (cur > p.sk_total) cur = p.sk_total;
            if (end > p.sk_total) end = p.sk_total;
        }
        if (cur >= end) return -1;
        int64_t const at = cur;
        int64_t len = p.num_k_blocks - at % p.num_k_blocks;
        if (len > end - at) len = end - at;
        cur += len;
        return at;
    }
    // full_warp: all 32 lanes call next() together (epilogue warps) and lane 0 releases the slot once
    // every lane has read it; otherwise the caller is a single elected thread.
    __device__ __forceinline__ int64_t next(int group_id, int num_groups, int64_t total_tiles, uint64_t* sched_full,
                                            uint64_t* sched_empty, const volatile int* sched_tile, bool do_arrive,
                                            bool full_warp, const Tf32Params* sk = nullptr) {
        if constexpr (!DYNAMIC) {
            if (sk != nullptr) return next_stream_k(*sk, group_id);
            int64_t const t = (int64_t)group_id + (int64_t)it * num_groups;
            ++it;
            return t < total_tiles ? t : -1;
        } else {
            int const s = it % SCHED_STAGES;
            uint32_t const ph = (uint32_t)(it / SCHED_STAGES) & 1u;
            mbar_wait_cluster(&sched_full[s], ph);
            int const t = sched_tile[s];
            if (full_warp) __syncwarp();                    // every lane of the warp has read the slot
            if (do_arrive) mbar_arrive_cluster(&sched_empty[s], 0);   // slot may be refilled (leader's barrier)
            ++it;
            return (int64_t)t;
        }
    }
};

// ---- the GEMM kernel ----------------------------------------------------------------------------------
// ---- the split itself -----------------------------------------------------------------------------------------
// Default split: hi = x itself (the tensor core drops the low 13 mantissa bits), lo = rn_tf32(x - trunc_tf32(x)).
// Truncation never overflows (|hi| <= |x|), x - trunc(x) is exact, and non-finite x gets lo = 0 so that an
// Inf stays an Inf instead of turning into Inf - Inf.  ROUND_HI (measurement / fallback aid): the classic
// split hi = rn_tf32(x), lo = rn_tf32(x - hi) with both planes materialised.
template <bool ROUND_HI>
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    uint32_t const xb = __float_as_uint(x);
    if constexpr (ROUND_HI) {
        uint32_t h, l;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
        if ((h & 0x7f800000u) == 0x7f800000u && (xb & 0x7f800000u) != 0x7f800000u) h = xb & 0xffffe000u;  // rounding overflowed
        hi = __uint_as_float(h);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
        lo = ((xb & 0x7f800000u) == 0x7f800000u) ? 0.f : __uint_as_float(l);
    } else {
        hi = x;
        float const r = x - __uint_as_float(xb & 0xffffe000u);
        uint32_t l;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
        lo = ((xb & 0x7f800000u) == 0x7f800000u) ? 0.f : __uint_as_float(l);
    }
}

// One unit of work of a CTA group: k-blocks [kb0, kb1) of output tile `tile`; turn_word >= 0: several units add
// into this tile, in turn order (word index of the tile's turnstile).
struct WorkItem {
    int64_t tile;
    int kb0, kb1;
    int turn;
    int64_t turn_word;
};
__device__ __forceinline__ WorkItem decode_unit(const Tf32Params& p, int64_t unit, int group_id) {
    WorkItem w;
    int const nkb = p.num_k_blocks;
    if (p.stream_k) {
        w.tile = unit / nkb;
        w.kb0 = (int)(unit - w.tile * nkb);
        int64_t end = (int64_t)(group_id + 1) * p.sk_width;
        if (end > p.sk_total) end = p.sk_total;
        int64_t len = end - unit;
        if (len > nkb - w.kb0) len = nkb - w.kb0;
        w.kb1 = w.kb0 + (int)len;
        int64_t const g_first = (w.tile * nkb) / p.sk_width, g_last = ((w.tile + 1) * nkb - 1) / p.sk_width;
        w.turn = (int)(g_last - group_id);
        w.turn_word = g_last > g_first ? g_first : -1;      // among split tiles the first contributing group is unique
    } else {
        if (unit < p.split_from) {
            w.tile = unit;
            w.kb0 = 0;
            w.kb1 = nkb;
            w.turn = 0;
            w.turn_word = -1;
        } else {
            int64_t const u = unit - p.split_from, t = u / p.split_k;
            int const sidx = (int)(u - t * p.split_k);
            w.tile = p.split_from + t;
            w.kb0 = sidx * p.kb_per_split;
            w.kb1 = w.kb0 + p.kb_per_split < nkb ? w.kb0 + p.kb_per_split : nkb;
            w.turn = sidx;
            w.turn_word = p.split_k > 1 ? t : -1;
        }
    }
    return w;
}

// FUSED: the lo tiles are not fetched from planes in HBM — four extra converter warps compute them in shared
// memory from the raw tiles the TMA just delivered (lo = rn_tf32(x - trunc_tf32(x)), element for element at the
// same tile offset, so any swizzle / major carries over).  No pre-pass launch, half the HBM and L2->SM traffic;
// the price is shared-memory bandwidth (32 KiB read + 32 KiB written per k-block next to the MMA's own reads).
// Needs both operands fetchable in place; static tile assignment only.
template <int NCTA, bool DYNAMIC, bool FUSED>
__global__ void __launch_bounds__(FUSED ? NUM_THREADS + 128 : NUM_THREADS, 1)
mtm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                  const __grid_constant__ CUtensorMap map_c, Tf32Params p) {
    constexpr int UMMA_M = 128 * NCTA;
    constexpr int TMEM_COLS = ACC_STAGES * 128 * NCTA;   // 256 or 512 (power of two); sized for the widest tile
    constexpr int EPI_THREADS = 128;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint32_t const align_off = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;  // SWIZZLE_128B needs 1 KiB alignment
    uint8_t* smem = smem_raw + align_off;
    uint8_t* epi_smem = smem + STAGES * STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + EPI_BYTES);
    uint64_t* full_bar = bars;                          // [MAX_STAGES]   TMA -> MMA
    uint64_t* empty_bar = bars + MAX_STAGES;            // [MAX_STAGES]   MMA -> TMA
    uint64_t* tmem_full_bar = bars + 2 * MAX_STAGES;    // [ACC_STAGES] MMA -> epilogue
    uint64_t* tmem_empty_bar = bars + 2 * MAX_STAGES + ACC_STAGES;  // [ACC_STAGES] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 2 * ACC_STAGES);
    uint64_t* sched_full = bars + 2 * MAX_STAGES + 2 * ACC_STAGES + 1;      // [SCHED_STAGES] scheduler -> roles (per CTA)
    uint64_t* sched_empty = sched_full + SCHED_STAGES;                  // [SCHED_STAGES] roles -> scheduler (leader's)
    volatile int* sched_tile = reinterpret_cast<volatile int*>(sched_empty + SCHED_STAGES);   // [SCHED_STAGES]
    uint64_t* raw_bar = sched_empty + SCHED_STAGES + 2;    // FUSED [RAW_STAGES]: TMA -> converters (per CTA, local)
    uint64_t* raw_empty = raw_bar + RAW_STAGES;            // FUSED [RAW_STAGES]: MMA commit -> producer (every CTA)
    uint64_t* conv_bar = raw_empty + RAW_STAGES;           // FUSED [LO_STAGES]: converters of both CTAs -> MMA (leader's)
    uint64_t* lo_empty = conv_bar + LO_STAGES;             // FUSED [LO_STAGES]: MMA commit -> converters (every CTA)
    uint8_t* const lo_smem = smem + RAW_STAGES * RAW_STAGE_BYTES;   // FUSED: the lo ring behind the raw ring

    // If the next kernel on the stream is a programmatic dependent launch (the split pass of the next mtm call), its
    // CTAs may be scheduled from now on; they wait for this grid to complete before they touch memory.
    griddep_launch_dependents();

    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t const cta_rank = NCTA == 1 ? 0u : cluster_ctarank();
    bool const is_leader = cta_rank == 0;
    int const umma_n = p.bn_cta * NCTA;
    // plane-fed stage: [A hi | A lo | B hi | B lo]; a narrow B tile shrinks the stage, and one more stage fits
    int const b_tile_bytes = p.bn_cta * BK * 4;
    int const tiles_w = p.dbl ? 2 : 1;                  // 256-wide tiles per work unit
    int const stage_bytes = 2 * TILE_BYTES + 2 * tiles_w * b_tile_bytes;   // double tiles: [A hi | A lo | B0 hi | B0 lo | B1 hi | B1 lo]
    int const num_stages = (STAGES * STAGE_BYTES) / stage_bytes < MAX_STAGES ? (STAGES * STAGE_BYTES) / stage_bytes : MAX_STAGES;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_hi)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_lo)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b_hi)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b_lo)) : "memory");
        if (p.c_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_c)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < MAX_STAGES; ++i) {
            mbar_init(&full_bar[i], 1 + (NCTA == 2 && !FUSED ? p.peer_arrive : 0));      // the leader's arrive + expect_tx for the bytes of BOTH CTAs (the peer's loads
                                             // complete on this barrier too; it does not arrive itself — see the producer)
            mbar_init(&empty_bar[i], 1);     // one tcgen05.commit
        }
        for (int i = 0; i < ACC_STAGES; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], NCTA * EPI_THREADS);
        }
        if constexpr (FUSED) {
            for (int i = 0; i < RAW_STAGES; ++i) {
                mbar_init(&raw_bar[i], 1);                   // this CTA's producer (arrive + tx)
                mbar_init(&raw_empty[i], 1);                 // one tcgen05.commit
            }
            for (int i = 0; i < LO_STAGES; ++i) {
                mbar_init(&conv_bar[i], NCTA * 4);           // one arrive per converter warp of every CTA
                mbar_init(&lo_empty[i], 1);                  // one tcgen05.commit
            }
        }
        for (int i = 0; i < SCHED_STAGES; ++i) {
            mbar_init(&sched_full[i], 1);                    // the scheduler's arrive
            mbar_init(&sched_empty[i], NCTA * 5);            // leader: MMA thread + 4 epilogue warps; peer: producer + 4 epilogue warps
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<NCTA>(tmem_slot, TMEM_COLS);
    tcgen05_fence_before();
    if constexpr (NCTA == 1) __syncthreads(); else cluster_sync_all();
    tcgen05_fence_after();
    uint32_t const tmem_base = *tmem_slot;
    // Everything above (barriers, tensor memory, descriptor prefetch) may have run while the split pass before this
    // kernel was still working; the lo planes, the tile counter and the turnstiles are only valid from here on.
    griddep_wait();

    int const num_groups = gridDim.x / NCTA;
    int const group_id = blockIdx.x / NCTA;
    int64_t const total_tiles = p.split_from + ((int64_t)p.tiles_m * p.tiles_n - p.split_from) * p.split_k;   // work units
    const Tf32Params* const skp = (!DYNAMIC && p.stream_k) ? &p : nullptr;     // stream-K ranges instead of whole units

    if (warp == 0) {
        // ===== TMA producer (one elected lane) =====
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            TileSource<NCTA, DYNAMIC> src;
            // DYNAMIC: the leader's producer is the first role to need the next tile, so it claims it
            // (just in time, after issuing the current tile's loads: CTA groups then take consecutive
            // tiles in completion order and the concurrently running tiles stay neighbours in L2) and
            // publishes it through the ring; every other role — including the peer CTA's producer —
            // reads the ring.
            bool const claims = DYNAMIC && is_leader;
            int const b_chunks = p.bn_cta / MN_CHUNK;
            uint32_t const stage_tx = (uint32_t)stage_bytes * NCTA;
            int64_t tile = claims ? (int64_t)group_id
                                  : src.next(group_id, num_groups, total_tiles, sched_full, sched_empty, sched_tile, true, false, skp);
            for (int it = 0;; ++it) {
                if (claims) {
                    int const s = it % SCHED_STAGES;
                    uint32_t const ph = (uint32_t)(it / SCHED_STAGES) & 1u;
                    mbar_wait_cluster(&sched_empty[s], ph ^ 1);               // every reader has consumed this slot
                    for (uint32_t r = 0; r < (uint32_t)NCTA; ++r) {
                        st_shared_cluster_u32(const_cast<const int*>(&sched_tile[s]), r, (uint32_t)(int)tile);
                        mbar_arrive_cluster(&sched_full[s], r);               // release: publishes the store above
                    }
                }
                if (tile < 0) break;
                int64_t pm, pn;
                WorkItem const w = decode_unit(p, tile, group_id);
                tile_coords_rt(w.tile, p.tiles_m, p.tiles_n, p.group, pm, pn);
                int const row_a = (int)(pm * UMMA_M) + (int)cta_rank * TILE_R;
                int const row_b = (int)(pn * umma_n * tiles_w) + (int)cta_rank * p.bn_cta;   // (+ umma_n for the second tile of a double)
                int const kb0 = w.kb0, kb1 = w.kb1;
                if constexpr (FUSED) {
                    // raw tiles only, completing on THIS CTA's barrier: its own converter warps pick them up
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&raw_empty[stage], phase ^ 1);
                        uint8_t* s = smem + stage * RAW_STAGE_BYTES;
                        int const k0 = kb * BK;
                        if (!p.a_mn) {
                            tma_load_2d<1>(&map_a_hi, &raw_bar[stage], s, k0, row_a);
                        } else {
#pragma unroll
                            for (int j = 0; j < TILE_R / MN_CHUNK; ++j)
                                tma_load_2d<1>(&map_a_hi, &raw_bar[stage], s + j * MN_CHUNK_BYTES, row_a + j * MN_CHUNK, k0);
                        }
                        if (!p.b_mn) {
                            tma_load_2d<1>(&map_b_hi, &raw_bar[stage], s + TILE_BYTES, k0, row_b);
                        } else {
                            for (int j = 0; j < b_chunks; ++j)
                                tma_load_2d<1>(&map_b_hi, &raw_bar[stage], s + TILE_BYTES + j * MN_CHUNK_BYTES, row_b + j * MN_CHUNK, k0);
                        }
                        mbar_arrive_expect_tx(&raw_bar[stage], (uint32_t)(TILE_BYTES + p.bn_cta * BK * 4));
                        if (++stage == RAW_STAGES) { stage = 0; phase ^= 1; }
                    }
                } else
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* s = smem + stage * stage_bytes;
                    int const k0 = kb * BK;
                    if (!p.a_mn) {
                        tma_load_2d<NCTA>(&map_a_hi, &full_bar[stage], s + 0 * TILE_BYTES, k0, row_a);
                        tma_load_2d<NCTA>(&map_a_lo, &full_bar[stage], s + 1 * TILE_BYTES, k0, row_a);
                    } else {
#pragma unroll
                        for (int j = 0; j < TILE_R / MN_CHUNK; ++j) {
                            tma_load_2d<NCTA>(&map_a_hi, &full_bar[stage], s + 0 * TILE_BYTES + j * MN_CHUNK_BYTES, row_a + j * MN_CHUNK, k0);
                            tma_load_2d<NCTA>(&map_a_lo, &full_bar[stage], s + 1 * TILE_BYTES + j * MN_CHUNK_BYTES, row_a + j * MN_CHUNK, k0);
                        }
                    }
                    for (int t = 0; t < tiles_w; ++t) {
                        uint8_t* const sb = s + 2 * TILE_BYTES + t * 2 * b_tile_bytes;
                        int const rb = row_b + t * umma_n;
                        if (!p.b_mn) {
                            tma_load_2d<NCTA>(&map_b_hi, &full_bar[stage], sb, k0, rb);
                            tma_load_2d<NCTA>(&map_b_lo, &full_bar[stage], sb + b_tile_bytes, k0, rb);
                        } else {
                            for (int j = 0; j < b_chunks; ++j) {
                                tma_load_2d<NCTA>(&map_b_hi, &full_bar[stage], sb + j * MN_CHUNK_BYTES, rb + j * MN_CHUNK, k0);
                                tma_load_2d<NCTA>(&map_b_lo, &full_bar[stage], sb + b_tile_bytes + j * MN_CHUNK_BYTES, rb + j * MN_CHUNK, k0);
                            }
                        }
                    }
                    // Only the leader arrives (and announces the bytes of both CTAs); the peer's copies are accounted by their
                    // complete_tx alone — if they land before the leader's expect_tx the transaction count just goes negative
                    // for a moment, the phase cannot complete without the leader's arrival.  (A remote arrive with cluster-scope
                    // release per k-block in the peer's producer loop was measured to bound EVERY pair config at the same
                    // ~1870 clocks per k-block, whatever its stage count, bytes or MMA time: profiles/r03j_*, r03m_*.)
                    if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
                    else if (p.peer_arrive) mbar_arrive_cluster(&full_bar[stage], 0);
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
                if (claims) {
                    int64_t const claimed = (int64_t)atomicAdd(p.tile_counter, 1);
                    tile = claimed < total_tiles ? claimed : -1;
                } else {
                    tile = src.next(group_id, num_groups, total_tiles, sched_full, sched_empty, sched_tile, true, false, skp);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        // The whole warp runs the (warp-uniform) loop and ONE elected lane issues: with a single thread in the loop the
        // compiler keeps descriptors and addresses in vector registers and moves them to uniform registers in front of every
        // MMA — about 130 clocks of dependent instructions per MMA, measured as a fixed ~1870 clocks per k-block that capped
        // the single-tile pair configs at 82 % (256 x 256) and 41 % (256 x 128) tensor-pipe activity whatever their stage
        // count or byte volume (profiles/r03f_*, r03j_*).  Descriptors are two 32-bit words; only the low one moves.
        if (is_leader) {
            uint32_t const issue = elect_one() ? 1u : 0u;
            uint32_t const idesc = make_idesc_tf32(UMMA_M, (uint32_t)umma_n, (uint32_t)p.a_mn, (uint32_t)p.b_mn);
            uint32_t const a_hiw = p.a_mn ? desc_hi_mnmajor(p.mn_sbo, p.mn_layout) : desc_hi_kmajor_sw128();
            uint32_t const b_hiw = p.b_mn ? desc_hi_mnmajor(p.mn_sbo, p.mn_layout) : desc_hi_kmajor_sw128();
            uint32_t const a_low0 = p.a_mn ? desc_lo_lbo_mnmajor(p.mn_lbo) : 0u;      // + (address >> 4)
            uint32_t const b_low0 = p.b_mn ? desc_lo_lbo_mnmajor(p.mn_lbo) : 0u;
            // per k step (8 of K): K-major +32 B inside the 128-byte swizzle row; MN-major the next pair of 4-row atoms
            uint32_t const a_kstep = (uint32_t)(p.a_mn ? MN_KSTEP_BYTES : UMMA_K * 4) >> 4;
            uint32_t const b_kstep = (uint32_t)(p.b_mn ? MN_KSTEP_BYTES : UMMA_K * 4) >> 4;
            uint32_t const smem0 = smem_u32(smem), lo0 = smem_u32(lo_smem);
            int stage = 0;
            uint32_t phase = 0;
            int lstage = 0;                 // FUSED: position in the lo ring
            uint32_t lphase = 0;
            int it = 0;
            TileSource<NCTA, DYNAMIC> src;
            for (int64_t unit; (unit = src.next(group_id, num_groups, total_tiles, sched_full, sched_empty, sched_tile, lane == 0, true, skp)) >= 0; ++it) {
                WorkItem const w = decode_unit(p, unit, group_id);
                int const kb0 = w.kb0, kb1 = w.kb1;
                // single tiles alternate between the two accumulator buffers; a double tile takes both (tile t -> buffer t)
                int const acc = p.dbl ? 0 : it % ACC_STAGES;
                uint32_t const acc_phase = p.dbl ? (uint32_t)it & 1u : (uint32_t)(it / ACC_STAGES) & 1u;
                if (!p.dbl) {
                    mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
                    tcgen05_fence_after();
                }
                for (int kb = kb0; kb < kb1; ++kb) {
                    uint32_t t_ahi, t_alo, t_bhi, t_blo;        // shared-memory addresses (>> 4) of the four operand tiles
                    if constexpr (FUSED) {
                        // the lo tiles were written by ordinary stores, in the peer CTA too: acquire at cluster scope
                        // (the converters waited for the raw tiles, so those have landed as well)
                        if constexpr (NCTA == 2) mbar_wait_cluster(&conv_bar[lstage], lphase);
                        else mbar_wait(&conv_bar[lstage], lphase);
                        uint32_t const rs = (smem0 + (uint32_t)(stage * RAW_STAGE_BYTES)) >> 4, ls = (lo0 + (uint32_t)(lstage * LO_STAGE_BYTES)) >> 4;
                        t_ahi = rs;
                        t_bhi = rs + (TILE_BYTES >> 4);
                        t_alo = ls;
                        t_blo = ls + (TILE_BYTES >> 4);
                    } else {
                        mbar_wait(&full_bar[stage], phase);
                        uint32_t const sb = (smem0 + (uint32_t)(stage * stage_bytes)) >> 4;
                        t_ahi = sb;
                        t_alo = sb + (TILE_BYTES >> 4);
                        t_bhi = sb + (2 * TILE_BYTES >> 4);
                        t_blo = t_bhi + (uint32_t)(b_tile_bytes >> 4);
                    }
                    tcgen05_fence_after();
                    for (int t = 0; t < tiles_w; ++t) {
                        if (p.dbl && kb == kb0) {
                            mbar_wait(&tmem_empty_bar[t], acc_phase ^ 1);   // the epilogue has drained buffer t (the other one may still be draining)
                            tcgen05_fence_after();
                        }
                        uint32_t const tmem_d = tmem_base + (uint32_t)((acc + t) * umma_n);
                        uint32_t a_hi = a_low0 + t_ahi, a_lo = a_low0 + t_alo, b_hi = b_low0 + t_bhi, b_lo = b_low0 + t_blo;
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            // small terms first, then the dominant hi*hi product
                            umma_tf32_w<NCTA>(tmem_d, a_lo, a_hiw, b_hi, b_hiw, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u, issue);
                            umma_tf32_w<NCTA>(tmem_d, a_hi, a_hiw, b_lo, b_hiw, idesc, 1u, issue);
                            umma_tf32_w<NCTA>(tmem_d, a_hi, a_hiw, b_hi, b_hiw, idesc, 1u, issue);
                            a_hi += a_kstep;
                            a_lo += a_kstep;
                            b_hi += b_kstep;
                            b_lo += b_kstep;
                        }
                        t_bhi += (uint32_t)(2 * b_tile_bytes >> 4);      // the second tile's B
                        t_blo += (uint32_t)(2 * b_tile_bytes >> 4);
                    }
                    if constexpr (FUSED) {
                        umma_commit_w<NCTA>(&raw_empty[stage], issue);          // raw stage -> the producers (both CTAs)
                        umma_commit_w<NCTA>(&lo_empty[lstage], issue);          // lo stage -> the converters (both CTAs)
                        if (kb == kb1 - 1) umma_commit_w<NCTA>(&tmem_full_bar[acc], issue);
                        if (++stage == RAW_STAGES) { stage = 0; phase ^= 1; }
                        if (++lstage == LO_STAGES) { lstage = 0; lphase ^= 1; }
                        continue;
                    }
                    umma_commit_w<NCTA>(&empty_bar[stage], issue);              // frees the smem stage (both CTAs)
                    if (kb == kb1 - 1) {
                        umma_commit_w<NCTA>(&tmem_full_bar[acc], issue);        // accumulator ready
                        if (p.dbl) umma_commit_w<NCTA>(&tmem_full_bar[1], issue);
                    }
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (FUSED && warp >= 8) {
        // ===== converters: lo tiles from the raw tiles, in shared memory =====
        int const tid_c = (warp - 8) * 32 + lane;                 // 0 .. 127
        int const b_vec = p.bn_cta * (BK * 4 / 16);               // 16-byte units of this CTA's B tile
        int stage = 0, lstage = 0;
        uint32_t phase = 0, lphase = 0;
        TileSource<NCTA, DYNAMIC> src;
        constexpr int NV = TILE_BYTES / 16 / 128;                 // 16-byte units per thread and tile
        for (int64_t unit; (unit = src.next(group_id, num_groups, total_tiles, sched_full, sched_empty, sched_tile, false, true, skp)) >= 0;) {
            WorkItem const w = decode_unit(p, unit, group_id);
            int const kb0 = w.kb0, kb1 = w.kb1;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&raw_bar[stage], phase);                // this CTA's raw tiles have landed
                uint32_t const rs = smem_u32(smem + stage * RAW_STAGE_BYTES), ls = smem_u32(lo_smem + lstage * LO_STAGE_BYTES);
                float4 xa[NV], xb[NV];
#pragma unroll
                for (int i = 0; i < NV; ++i) xa[i] = ld_shared_v4(rs + (uint32_t)(tid_c + i * 128) * 16u);
#pragma unroll
                for (int i = 0; i < NV; ++i)
                    if (tid_c + i * 128 < b_vec) xb[i] = ld_shared_v4(rs + TILE_BYTES + (uint32_t)(tid_c + i * 128) * 16u);
                mbar_wait(&lo_empty[lstage], lphase ^ 1);         // the MMAs that read this lo stage have completed
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    float h;
                    float4 l;
                    split_tf32<false>(xa[i].x, h, l.x);
                    split_tf32<false>(xa[i].y, h, l.y);
                    split_tf32<false>(xa[i].z, h, l.z);
                    split_tf32<false>(xa[i].w, h, l.w);
                    st_shared_v4(ls + (uint32_t)(tid_c + i * 128) * 16u, __float_as_uint(l.x), __float_as_uint(l.y),
                                 __float_as_uint(l.z), __float_as_uint(l.w));
                }
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    if (tid_c + i * 128 < b_vec) {
                        float h;
                        float4 l;
                        split_tf32<false>(xb[i].x, h, l.x);
                        split_tf32<false>(xb[i].y, h, l.y);
                        split_tf32<false>(xb[i].z, h, l.z);
                        split_tf32<false>(xb[i].w, h, l.w);
                        st_shared_v4(ls + TILE_BYTES + (uint32_t)(tid_c + i * 128) * 16u, __float_as_uint(l.x), __float_as_uint(l.y),
                                     __float_as_uint(l.z), __float_as_uint(l.w));
                    }
                }
                fence_proxy_async_shared();                       // my stores -> visible to the tensor core's reads
                __syncwarp();
                // plain arrive (measured: the release.cluster form costs a MEMBAR per k-block, 12 us of a 37 us call at
                // 1024^3, profiles/r02h_*); the proxy fence above is what orders the tile for the tensor core
                if (lane == 0) mbar_arrive_remote(&conv_bar[lstage], 0);
                if (++stage == RAW_STAGES) { stage = 0; phase ^= 1; }
                if (++lstage == LO_STAGES) { lstage = 0; lphase ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== epilogue: TMEM -> registers -> (shared memory -> TMA reduce-add | C += acc) =====
        int const ew = warp & 3;                        // TMEM lane quarter this warp may access
        uint8_t* const my_epi = epi_smem + ew * (EPI_BUFS * EPI_BUF_BYTES);
        uint32_t q = 0;                                 // boxes sent so far by this warp (selects the staging buffer)
        int it = 0;
        TileSource<NCTA, DYNAMIC> src;
        for (int64_t tile; (tile = src.next(group_id, num_groups, total_tiles, sched_full, sched_empty, sched_tile, lane == 0, true, skp)) >= 0; ++it) {
            int64_t pm, pn;
            WorkItem const w = decode_unit(p, tile, group_id);
            uint32_t const split = (uint32_t)w.turn;
            tile_coords_rt(w.tile, p.tiles_m, p.tiles_n, p.group, pm, pn);
            int64_t const row0 = pm * UMMA_M + (int64_t)cta_rank * TILE_R + ew * 32;   // first row of this warp
            volatile uint32_t* my_turn = nullptr;
            for (int t = 0; t < tiles_w; ++t) {      // a double tile: buffer 0, then buffer 1
            int const acc = p.dbl ? t : it % ACC_STAGES;
            uint32_t const acc_phase = p.dbl ? (uint32_t)it & 1u : (uint32_t)(it / ACC_STAGES) & 1u;
            int64_t const col0 = (pn * tiles_w + t) * umma_n;
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tcgen05_fence_after();
            if (w.turn_word >= 0 && t == 0) {
                // my 32 rows of this unit: wait until the units before mine (split-K: lower k; stream-K: higher group)
                // have added theirs
                my_turn = p.turn + (w.turn_word * (4 * NCTA) + (int64_t)cta_rank * 4 + ew);
                if (lane == 0) {
                    long long const t0 = clock64();
                    while (*my_turn != split) {
                        __nanosleep(64);
                        if (clock64() - t0 > 8000000000LL) __trap();
                    }
                    __threadfence();
                }
                __syncwarp();
            }
#pragma unroll 1
            for (int c = 0; c < umma_n / 32; ++c) {
                uint32_t v[32];
                uint32_t const taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * umma_n + c * 32);
                tmem_ld_32x32b_x32(taddr, v);
                tmem_ld_wait();
                int64_t const n0 = col0 + c * 32;
                if (p.c_tma) {
                    if (row0 < p.M && n0 < p.N) {       // warp-uniform; partial boxes are clipped by the tensor map
                        uint8_t* const buf = my_epi + (q & 1u) * EPI_BUF_BYTES;
                        if (lane == 0) bulk_wait_group_read<EPI_BUFS - 1>();   // the box that used this buffer has been read
                        __syncwarp();
                        uint32_t const rbase = smem_u32(buf) + (uint32_t)lane * 128u;
#pragma unroll
                        for (int j = 0; j < 8; ++j)     // row `lane`, 16-byte chunk j at its SWIZZLE_128B position
                            st_shared_v4(rbase + (uint32_t)((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        fence_proxy_async_shared();
                        __syncwarp();
                        if (lane == 0) {
                            tma_reduce_add_2d(&map_c, buf, (int)n0, (int)row0);
                            bulk_commit_group();
                        }
                        ++q;
                    }
                } else if (row0 < p.M && n0 < p.N) {
                    // C cannot be a TMA target (odd leading dimension / unaligned base): transpose the box through
                    // the warp's staging buffer so that every read-modify-write instruction covers 32 CONSECUTIVE
                    // columns of one row (one 128-byte line) instead of one column of 32 rows
                    float* const fb = reinterpret_cast<float*>(my_epi);
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 32; ++j) fb[lane * 33 + j] = __uint_as_float(v[j]);     // padded: conflict-free both ways
                    __syncwarp();
                    bool const col_ok = n0 + lane < p.N;
                    float* const cbase = p.C + row0 * p.ldc + n0 + lane;
                    int const rows_here = p.M - row0 < 32 ? (int)(p.M - row0) : 32;
                    for (int r = 0; r < rows_here; ++r)
                        if (col_ok) cbase[(int64_t)r * p.ldc] += fb[r * 33 + lane];
                }
            }
            tcgen05_fence_before();
            mbar_arrive_cluster(&tmem_empty_bar[acc], 0);   // accumulator may be overwritten
            if (my_turn != nullptr && t == tiles_w - 1) {
                // hand the rows to the next unit once my additions have been performed
                __syncwarp();
                if (lane == 0) {
                    if (p.c_tma) bulk_wait_group<0>();
                    __threadfence();
                    *my_turn = split + 1;
                }
            }
            }
        }
        // the staging buffers must have been READ before the CTA retires; the reduces themselves complete with the grid
        // (kernel completion covers the CTA's outstanding bulk operations), which saves their write latency on the way out
        if (p.c_tma && lane == 0) bulk_wait_group_read<0>();
        __syncwarp();
    }

    tcgen05_fence_before();
    if constexpr (NCTA == 1) __syncthreads(); else cluster_sync_all();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc<NCTA>(tmem_base, TMEM_COLS);
    }
}

// ---- operand split pre-pass (kernels) ----------------------------------------------------------------------
enum SplitMode : int { SPLIT_NONE = 0, SPLIT_ELEMENTWISE = 1, SPLIT_GATHER = 2 };

struct SplitJob {
    const float* in;
    float* hi;            // GATHER only
    float* lo;
    int64_t s_line, s_k;  // ELEMENTWISE: input pitch between lines (elements); GATHER: in(r, k) = in[r * s_line + k * s_k]
    int64_t pitch;        // output pitch (elements)
    int lines, len;       // ELEMENTWISE: lines x len (len contiguous); GATHER: rows x K
    int mode;
    int bx;               // blocks along len / K
    int64_t blocks;       // CTAs of this job
};

constexpr int EW_LINES = 4, EW_LEN = 1024;     // elementwise CTA tile: 4 lines x 1024 contiguous floats (one float4 per thread and line)

__device__ __forceinline__ void split_elementwise(const SplitJob& j, int64_t blk) {
    int64_t const by = blk / j.bx;
    int const bx = (int)(blk - by * j.bx);
    int const col = bx * EW_LEN + (int)threadIdx.x * 4;
    if (col >= j.len) return;
    bool const vec = col + 4 <= j.len;
    float4 x[EW_LINES];
#pragma unroll
    for (int i = 0; i < EW_LINES; ++i) {
        int64_t const line = by * EW_LINES + i;
        x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (line < j.lines) {
            const float* src = j.in + line * j.s_line + col;
            if (vec) {
                x[i] = *reinterpret_cast<const float4*>(src);
            } else {                                        // ragged end of a line: never read past the matrix
                x[i].x = src[0];
                if (col + 1 < j.len) x[i].y = src[1];
                if (col + 2 < j.len) x[i].z = src[2];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < EW_LINES; ++i) {
        int64_t const line = by * EW_LINES + i;
        if (line < j.lines) {
            float h;
            float4 l;
            split_tf32<false>(x[i].x, h, l.x);
            split_tf32<false>(x[i].y, h, l.y);
            split_tf32<false>(x[i].z, h, l.z);
            split_tf32<false>(x[i].w, h, l.w);
            *reinterpret_cast<float4*>(j.lo + line * j.pitch + col) = l;     // pitch % 4 == 0: in bounds, padding gets 0
        }
    }
}

template <bool ROUND_HI>
__device__ __forceinline__ void split_gather(const SplitJob& j, int64_t blk, float (*tile)[33]) {
    int64_t const by = blk / j.bx;
    int const bx = (int)(blk - by * j.bx);
    int const tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    int64_t const r0 = by * 32;
    int const k0 = bx * 32;
    int const rows = j.lines, K = j.len;
    if (j.s_k == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int64_t const r = r0 + ty + 8 * i;
            int const k = k0 + tx;
            if (r < rows && k < K) {
                float hi, lo;
                split_tf32<ROUND_HI>(j.in[r * j.s_line + k], hi, lo);
                j.hi[r * j.pitch + k] = hi;
                j.lo[r * j.pitch + k] = lo;
            }
        }
    } else {
        // read with the warp running along r (coalesced when s_line == 1), transpose through smem
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int64_t const r = r0 + tx;
            int const k = k0 + ty + 8 * i;
            tile[ty + 8 * i][tx] = (r < rows && k < K) ? j.in[r * j.s_line + (int64_t)k * j.s_k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int const rl = ty + 8 * i;
            int64_t const r = r0 + rl;
            int const k = k0 + tx;
            if (r < rows && k < K) {
                float hi, lo;
                split_tf32<ROUND_HI>(tile[tx][rl], hi, lo);
                j.hi[r * j.pitch + k] = hi;
                j.lo[r * j.pitch + k] = lo;
            }
        }
    }
}

// One launch for both operands: CTAs [0, a.blocks) work on A, the rest on B.
// Shared memory (the 32 x 33 transpose tile) is dynamic and only requested when a job gathers: an elementwise-only
// launch takes none, so the MMA CTAs of the dependent launch fit next to its last CTAs (an MMA CTA leaves 1.5 KiB of
// the SM's shared memory).
constexpr int SPLIT_GATHER_SMEM = 32 * 33 * 4;
template <bool ROUND_HI>
__global__ void __launch_bounds__(256)
split_kernel(SplitJob a, SplitJob b, int* tile_counter, int counter_init, uint32_t* turn, int n_turn) {
    extern __shared__ __align__(16) float split_smem[];
    float (*tile)[33] = reinterpret_cast<float (*)[33]>(split_smem);
    // Programmatic dependent launch on both sides.  This launch may have become resident while the kernel before it on
    // the stream — the MMA kernel of the previous call, which still reads the planes this one overwrites — was running:
    // wait for it to complete before touching memory.  The MMA kernel behind this launch may set itself up right away;
    // it waits for the planes in turn (griddep_wait), so the chain of calls stays fully ordered.
    griddep_wait();
    griddep_launch_dependents();
    // The split always precedes the MMA kernel on the stream: it also re-arms the dynamic tile counter and the
    // split-K turnstiles.
    if (blockIdx.x == 0) {
        if (tile_counter != nullptr && threadIdx.x == 0) *tile_counter = counter_init;
        for (int i = threadIdx.x; i < n_turn; i += blockDim.x) turn[i] = 0u;
    }
    int64_t blk = (int64_t)blockIdx.x;
    const SplitJob& j = blk < a.blocks ? a : b;
    if (blk >= a.blocks) blk -= a.blocks;
    if (j.mode == SPLIT_ELEMENTWISE) split_elementwise(j, blk);
    else if (j.mode == SPLIT_GATHER) split_gather<ROUND_HI>(j, blk, tile);
}

// ---- host side ---------------------------------------------------------------------------------------------
inline int64_t round_up64(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// How one operand reaches the tensor cores.  `mn` = extent along m (A) or n (B); in(mn, k) = p[mn * s_mn + k * s_k].
enum OperandMode : int { OP_K_DIRECT = 0, OP_MN_DIRECT = 1, OP_PACKED = 2 };
struct OperandPlan {
    int mode;
    int64_t pitch;       // plane pitch in elements (K-major planes: along k; MN-major plane: along m/n)
    size_t bytes;        // workspace bytes of this operand's planes (1 KiB aligned)
};

int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}
bool force_packed() {
    static bool const f = env_int("B200_TF32_FORCE_PACKED", 0) != 0 || env_int("B200_TF32_ROUND_HI", 0) != 0;
    return f;
}
bool round_hi() {
    static bool const f = env_int("B200_TF32_ROUND_HI", 0) != 0;
    return f;
}

OperandPlan plan_operand(const float* p, int64_t mn, int64_t K, int64_t s_mn, int64_t s_k) {
    OperandPlan r{};
    bool const aligned = (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
    if (!force_packed() && aligned && s_k == 1 && s_mn % 4 == 0 && s_mn >= K) {
        r.mode = OP_K_DIRECT;
        r.pitch = round_up64(K, 4);
        r.bytes = (size_t)round_up64(mn * r.pitch * 4, 1024);
    } else if (!force_packed() && aligned && s_mn == 1 && s_k % 4 == 0 && s_k >= mn) {
        r.mode = OP_MN_DIRECT;
        r.pitch = round_up64(mn, 4);
        r.bytes = (size_t)round_up64(K * r.pitch * 4, 1024);
    } else {
        r.mode = OP_PACKED;
        r.pitch = round_up64(K, 4);
        r.bytes = 2 * (size_t)round_up64(mn * r.pitch * 4, 1024);
    }
    return r;
}

SplitJob make_job(const OperandPlan& pl, const float* in, int64_t mn, int64_t K, int64_t s_mn, int64_t s_k, float* planes) {
    SplitJob j{};
    j.in = in;
    j.pitch = pl.pitch;
    if (pl.mode == OP_PACKED) {
        j.mode = SPLIT_GATHER;
        j.hi = planes;
        j.lo = planes + pl.bytes / 8;          // second half of the operand's region
        j.s_line = s_mn;
        j.s_k = s_k;
        j.lines = (int)mn;
        j.len = (int)K;
        j.bx = (int)((K + 31) / 32);
        j.blocks = (int64_t)j.bx * ((mn + 31) / 32);
    } else {
        j.mode = SPLIT_ELEMENTWISE;
        j.hi = nullptr;
        j.lo = planes;
        bool const kd = pl.mode == OP_K_DIRECT;
        j.s_line = kd ? s_mn : s_k;
        j.s_k = 1;
        j.lines = (int)(kd ? mn : K);
        j.len = (int)(kd ? K : mn);
        j.bx = (int)((j.len + EW_LEN - 1) / EW_LEN);
        j.blocks = (int64_t)j.bx * ((j.lines + EW_LINES - 1) / EW_LINES);
    }
    return j;
}

// Tensor maps of one operand (hi, lo).  K-major: dims {K, mn}, box {BK, box_rows}.  MN-major: dims {mn, K}, box {32, BK}.
bool make_operand_maps(CUtensorMap* hi, CUtensorMap* lo, const OperandPlan& pl, const float* in, const float* planes,
                       int64_t mn, int64_t K, int64_t s_mn, int64_t s_k, int box_rows) {
    if (pl.mode == OP_PACKED) {
        const float* hp = planes;
        const float* lp = planes + pl.bytes / 8;
        return make_map_2d_f32(hi, hp, (uint64_t)K, (uint64_t)mn, (uint64_t)pl.pitch, BK, box_rows, CU_TENSOR_MAP_SWIZZLE_128B) &&
               make_map_2d_f32(lo, lp, (uint64_t)K, (uint64_t)mn, (uint64_t)pl.pitch, BK, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
    }
    if (pl.mode == OP_K_DIRECT)
        return make_map_2d_f32(hi, in, (uint64_t)K, (uint64_t)mn, (uint64_t)s_mn, BK, box_rows, CU_TENSOR_MAP_SWIZZLE_128B) &&
               make_map_2d_f32(lo, planes, (uint64_t)K, (uint64_t)mn, (uint64_t)pl.pitch, BK, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
    // MN-major tf32 tiles: 32-byte swizzle atoms (the only MN-major form kind::tf32 reads, see sm100_ptx.cuh)
    static int const env_sw = env_int("B200_TF32_MN_TMA_SWIZZLE", (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);   // (bring-up aid)
    CUtensorMapSwizzle const sw = (CUtensorMapSwizzle)env_sw;
    return make_map_2d_f32(hi, in, (uint64_t)mn, (uint64_t)K, (uint64_t)s_k, MN_CHUNK, BK, sw) &&
           make_map_2d_f32(lo, planes, (uint64_t)mn, (uint64_t)K, (uint64_t)pl.pitch, MN_CHUNK, BK, sw);
}

struct Tf32Tile {
    TileConfig cfg;
    int ncta, bn_cta;
    bool dynamic;
    bool fused;      // lo tiles computed in shared memory by converter warps (no pre-pass); sibling = same tile without
    int sibling;
    bool dbl = false;   // double tiles: 256 x 512 per pair, two accumulator buffers sharing the A tiles
};
const Tf32Tile kCfg[] = {
    {{"tf32x3_2cta_256x256x32", 256, 256, 32, NUM_THREADS, 1}, 2, 128, false, false, 0},        // static tile assignment
    {{"tf32x3_1cta_128x128x32", 128, 128, 32, NUM_THREADS, 1}, 1, 128, false, false, 1},
    {{"tf32x3_2cta_256x256x32_dyn", 256, 256, 32, NUM_THREADS, 1}, 2, 128, true, false, 2},     // dynamic tile scheduler
    {{"tf32x3_1cta_128x128x32_dyn", 128, 128, 32, NUM_THREADS, 1}, 1, 128, true, false, 3},
    {{"tf32x3_2cta_256x128x32", 256, 128, 32, NUM_THREADS, 1}, 2, 64, false, false, 4},         // narrower tiles: more of them
    {{"tf32x3_1cta_128x64x32", 128, 64, 32, NUM_THREADS, 1}, 1, 64, false, false, 5},
    {{"tf32x3_2cta_256x256x32_fused", 256, 256, 32, NUM_THREADS + 128, 1}, 2, 128, false, true, 0},   // in-kernel lo conversion, ONE launch
    {{"tf32x3_1cta_128x128x32_fused", 128, 128, 32, NUM_THREADS + 128, 1}, 1, 128, false, true, 1},
    {{"tf32x3_1cta_128x64x32_fused", 128, 64, 32, NUM_THREADS + 128, 1}, 1, 64, false, true, 5},
    {{"tf32x3_2cta_256x512x32", 256, 512, 32, NUM_THREADS, 1}, 2, 128, false, false, 9, true},     // double tiles: 3/4 of the L2 -> SM bytes per flop
    {{"tf32x3_2cta_256x512x32_dyn", 256, 512, 32, NUM_THREADS, 1}, 2, 128, true, false, 10, true}, // the same with the dynamic tile scheduler
};

// pdl: launch with programmatic stream serialization — the kernel may become resident while the kernel before it on
// the stream (this call's split pass) is still running; it orders itself behind that kernel with griddep_wait().
template <int NCTA, bool DYNAMIC, bool FUSED = false>
cudaError_t launch_gemm(const CUtensorMap* maps, const Tf32Params& p, int groups, int dev, cudaStream_t stream, bool pdl = false) {
    static bool attr_done[64] = {};     // per instantiation and device; racing callers set the same value
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        cudaError_t ea = cudaFuncSetAttribute(mtm_tf32x3_kernel<NCTA, DYNAMIC, FUSED>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (ea != cudaSuccess) return ea;
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(groups * NCTA));
    cfg.blockDim = dim3(FUSED ? NUM_THREADS + 128 : NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NCTA;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, mtm_tf32x3_kernel<NCTA, DYNAMIC, FUSED>, maps[0], maps[1], maps[2], maps[3], maps[4], p);
}

// The split pass: one launch, A's blocks then B's.  pdl: a programmatic dependent launch as well — behind the MMA kernel
// of a previous call its CTAs are scheduled early and wait in the kernel (hides the launch latency between calls).
cudaError_t launch_split(const SplitJob& ja, const SplitJob& jb, int* tile_counter, int counter_init, uint32_t* turn, int n_turn,
                         bool round_hi_planes, bool pdl, cudaStream_t stream) {
    int64_t const nblk = ja.blocks + jb.blocks;
    if (nblk <= 0 || nblk > 0x7fffffffLL) return cudaErrorInvalidValue;
    bool const gathers = (ja.blocks > 0 && ja.mode == SPLIT_GATHER) || (jb.blocks > 0 && jb.mode == SPLIT_GATHER);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)nblk);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = gathers ? SPLIT_GATHER_SMEM : 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return round_hi_planes ? cudaLaunchKernelEx(&cfg, split_kernel<true>, ja, jb, tile_counter, counter_init, turn, n_turn)
                           : cudaLaunchKernelEx(&cfg, split_kernel<false>, ja, jb, tile_counter, counter_init, turn, n_turn);
}

}  // namespace

// CUDA loads kernels lazily, at their first launch, and that load can need the context to be idle.  The
// multi-GPU drivers launch kernels that wait in-kernel for other GPUs, so every kernel of this path is
// made resident at context creation (CUDA programming guide, lazy loading: concurrent execution).
cudaError_t tf32_preload_kernels() {
    cudaFuncAttributes fa;
    cudaError_t e;
    if ((e = cudaFuncGetAttributes(&fa, split_kernel<false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, split_kernel<true>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, mtm_tf32x3_kernel<1, false, false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, mtm_tf32x3_kernel<1, true, false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, mtm_tf32x3_kernel<2, false, false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, mtm_tf32x3_kernel<2, true, false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, mtm_tf32x3_kernel<1, false, true>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, mtm_tf32x3_kernel<2, false, true>)) != cudaSuccess) return e;
    return cudaSuccess;
}

int tf32_num_configs() { return (int)(sizeof(kCfg) / sizeof(kCfg[0])); }
const TileConfig& tf32_config(int cfg) { return kCfg[cfg].cfg; }

constexpr size_t WS_SLACK = 16384;           // alignment slack, tile counter (word 0), split-K turnstiles (words 64 ...)
constexpr int TURN_WORD0 = 64, TURN_WORDS = 3072;

size_t tf32_workspace_bytes(const MtmShape& s, const float* A, const float* B) {
    OperandPlan const pa = plan_operand(A, s.M, s.K, s.a_sm, s.a_sk);
    OperandPlan const pb = plan_operand(B, s.N, s.K, s.b_sn, s.b_sk);
    return pa.bytes + pb.bytes + WS_SLACK;
}

int tf32_auto_split(int64_t tiles, int nkb, int slots);

// Tail split: with more tiles than CTA groups, only the tiles of a ragged LAST wave are cut along K (2..4 ways, >= 32
// k-blocks each) when they leave at least half of the groups idle: that wave then takes 1/S of a tile-time.  Returns
// the factor (1 = none).
int tf32_tail_split(int64_t tiles, int nkb, int slots) {
    if (tiles <= slots || slots <= 0) return 1;
    if (tiles / slots > 16) return 1;            // beyond ~16 waves half a tile-time is under 3 % of the call
    int64_t const tail = tiles % slots;
    if (tail == 0 || tail * 2 > slots) return 1;
    int64_t sk = slots / tail;
    if (sk > nkb / 32) sk = nkb / 32;
    if (sk > 4) sk = 4;
    return sk < 2 ? 1 : (int)sk;
}
// Fraction of the machine-time of ceil-ed waves that does work, with the K splits AUTO would apply.
double tf32_wave_efficiency(int64_t tiles, int nkb, int slots) {
    if (tiles <= 0 || slots <= 0) return 0.0;
    int const whole = tf32_auto_split(tiles, nkb, slots);
    if (whole > 1) {
        // the splits of a tile add into C one after the other and each pays an epilogue: measured, a split problem runs
        // at about 0.85 of what its unit count promises (2048^3: 32 double tiles x 2 splits 198 TFLOP/s against 212 for
        // 64 whole 256 x 256 tiles, profiles/r03f_bench.json / r03b_ab_pdl_final.jsonl)
        double const units = (double)tiles * whole, waves = (double)((tiles * whole + slots - 1) / slots);
        return 0.85 * units / (waves * slots);
    }
    int const ts = tf32_tail_split(tiles, nkb, slots);
    double const full = (double)(tiles / slots), tail = (double)(tiles % slots);
    double const waves = full + (tail > 0 ? 1.0 / ts : 0.0);
    return ((double)tiles / slots) / waves;
}

// Split-K factor for `tiles` output tiles of `nkb` k-blocks on `slots` CTA groups.  Measured
// (profiles/r02d_tune_small_splitk.json): the splits of a tile take turns adding into C, so many short splits
// serialise on their epilogues (S = 16 is 3-10x SLOWER than S = 1 at K <= 1024), while long-K problems with few
// tiles gain (512x512x8192: 35 -> 78 TFLOP/s at S = 4; 256x4096x4096: 85 -> 116 at S = 2).  Hence: only when the
// tiles leave at least half of the machine idle, at least 16 k-blocks (K = 512) per split, at most 4 splits.
int tf32_auto_split(int64_t tiles, int nkb, int slots) {
    if (tiles <= 0 || tiles * 2 > slots || nkb < 32) return 1;
    int64_t sk = slots / tiles;
    if (sk > nkb / 16) sk = nkb / 16;      // (1024^3 on 32 pair tiles: 2 splits of 16 k-blocks 23.9 against 25.9 us, profiles/r03o_tune.json)
    if (sk > 4) sk = 4;
    return sk < 1 ? 1 : (int)sk;
}

cudaError_t launch_3xtf32_f32(float* C, const float* A, const float* B, const MtmShape& s, void* ws,
                              size_t ws_bytes, int cfg, int reuse_b, int reserve_sms, int split_k, cudaStream_t stream,
                              int* launches, int* a_mode, int* b_mode, int* split_used, int* cfg_used) {
    if (launches) *launches = 0;
    if (cfg < 0 || cfg >= tf32_num_configs()) return cudaErrorInvalidValue;
    if (ws_bytes < tf32_workspace_bytes(s, A, B)) return cudaErrorInvalidValue;
    OperandPlan const pa = plan_operand(A, s.M, s.K, s.a_sm, s.a_sk);
    OperandPlan const pb = plan_operand(B, s.N, s.K, s.b_sn, s.b_sk);
    if (a_mode) *a_mode = pa.mode;
    if (b_mode) *b_mode = pb.mode;
    // the fused form needs both operands in place; otherwise its plane-fed sibling does the call
    if (kCfg[cfg].fused && (pa.mode == OP_PACKED || pb.mode == OP_PACKED)) cfg = kCfg[cfg].sibling;
    if (cfg_used) *cfg_used = cfg;
    float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
    // B's planes first: their position does not depend on M, so a caller slicing M can keep them (reuse_b).
    float* b_planes = base;
    float* a_planes = b_planes + pb.bytes / 4;
    int* tile_counter = reinterpret_cast<int*>(a_planes + pa.bytes / 4);   // inside the slack
    int n_launch = 0;

    int dev = 0, sm_count = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if (reserve_sms > 0 && reserve_sms < sm_count - 2) sm_count -= reserve_sms;
    const Tf32Tile& tc = kCfg[cfg];
    int const ncta = tc.ncta;
    Tf32Params p{};
    p.C = C;
    p.ldc = s.ldc;
    p.M = (int)s.M;
    p.N = (int)s.N;
    p.num_k_blocks = (int)((s.K + BK - 1) / BK);
    p.tiles_m = (int)((s.M + tc.cfg.bm - 1) / tc.cfg.bm);
    p.tiles_n = (int)((s.N + tc.cfg.bn - 1) / tc.cfg.bn);
    p.tile_counter = tile_counter;
    p.bn_cta = tc.bn_cta;
    p.dbl = tc.dbl ? 1 : 0;
    p.peer_arrive = env_int("B200_TF32_PEER_ARRIVE", 0) != 0 ? 1 : 0;     // (measurement aid, read per call)
    p.a_mn = pa.mode == OP_MN_DIRECT;
    p.b_mn = pb.mode == OP_MN_DIRECT;
    static int const env_lbo = env_int("B200_TF32_MN_LBO", MN_CHUNK_BYTES), env_sbo = env_int("B200_TF32_MN_SBO", MN_ATOM_BYTES);
    p.mn_lbo = (uint32_t)env_lbo;
    p.mn_sbo = (uint32_t)env_sbo;
    static int const env_layout = env_int("B200_TF32_MN_LAYOUT", 1);   // (bring-up aid)
    p.mn_layout = (uint32_t)env_layout;
    static bool const no_tma_epi = env_int("B200_TF32_NO_TMA_EPI", 0) != 0;
    p.c_tma = (!no_tma_epi && (reinterpret_cast<uintptr_t>(C) & 15u) == 0 && s.ldc % 4 == 0 && s.ldc >= s.N) ? 1 : 0;
    // Groups of 8 tile-rows: A row-panels and B column-panels of the running wave stay in L2.
    int const env_group = env_int("B200_TF32_GROUP", 0);   // (measurement aid, read per call: A/B the L2 locality of the walk)
    p.group = env_group > 0 ? env_group : (tc.dbl ? 16 : 8);
    int64_t const total_tiles = (int64_t)p.tiles_m * p.tiles_n;
    int groups = sm_count / ncta;
    // split-K: requested (flags) or automatic; no empty splits; the turnstile words must fit the slack
    static int const env_split = env_int("B200_TF32_SPLIT_K", 0);      // (measurement aid)
    int sk = split_k > 0 ? split_k : (env_split > 0 ? env_split : tf32_auto_split(total_tiles, p.num_k_blocks, groups));
    if (sk > p.num_k_blocks) sk = p.num_k_blocks;
    if (sk < 1 || total_tiles * 4 * ncta > TURN_WORDS) sk = 1;
    p.split_from = 0;
    static int const env_tail = env_int("B200_TF32_TAIL_SPLIT", 1);       // (measurement aid: 0 = off)
    if (sk == 1 && split_k == 0 && env_split == 0 && env_tail != 0) {
        // not a few-tiles problem: maybe its last wave is ragged
        int const ts = tf32_tail_split(total_tiles, p.num_k_blocks, groups);
        if (ts > 1) {
            sk = ts;
            p.split_from = total_tiles - total_tiles % groups;
        }
    }
    p.kb_per_split = (p.num_k_blocks + sk - 1) / sk;
    p.split_k = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;
    p.turn = reinterpret_cast<uint32_t*>(tile_counter) + TURN_WORD0;
    int const n_turn = p.split_k > 1 ? (int)((total_tiles - p.split_from) * 4 * ncta) : 0;
    int64_t const total_units = p.split_from + (total_tiles - p.split_from) * p.split_k;
    // Stream-K instead of whole tiles (static configs, no split-K): every group gets the same number of k-block
    // iterations, so a ragged last wave costs nothing.  MEASURED (profiles/r02n_stream_k_ab.txt) it does not pay here
    // and stays OPT-IN (B200_TF32_STREAM_K=1): 2048^3 (0.86 waves) 198.8 vs 198.1 TFLOP/s, 1024^3 76.6 vs 81.8,
    // 2560^3 162.6 vs 198.9 — the groups then sit at different k offsets of the tiles they share A and B panels with,
    // which gives up the lock-step L2 reuse of whole-tile waves (and a split tile pays a second epilogue + a turn).
    static int const env_sk = env_int("B200_TF32_STREAM_K", 0);
    p.stream_k = 0;
    p.sk_total = total_tiles * (int64_t)p.num_k_blocks;
    p.sk_width = (p.sk_total + groups - 1) / groups;
    if (env_sk == 1 && !tc.dynamic && !tc.dbl && split_k == 0 && env_split == 0 && p.split_k == 1 && p.sk_width >= 8) {   // (an explicit split factor, 1 included, means: no K splitting of any kind)
        int64_t const waves = (total_tiles + groups - 1) / groups;
        double const eff = (double)total_tiles / (double)(waves * groups);
        if (eff < 0.95) p.stream_k = 1;
    }
    if (split_used) *split_used = p.stream_k ? -1 : (p.split_from > 0 && p.split_k > 1 ? 100 + p.split_k : p.split_k);   // -1: stream-K, 100 + S: tail split
    int n_turn_words = n_turn;
    if (p.stream_k) n_turn_words = groups * 4 * ncta;                   // one turnstile per group that starts a split tile
    else if (total_units < groups) groups = (int)total_units;

    // 1. tensor maps (before the split: an operand whose direct map cannot be encoded would have to be packed)
    CUtensorMap maps[5];
    if (!make_operand_maps(&maps[0], &maps[1], pa, A, a_planes, s.M, s.K, s.a_sm, s.a_sk, TILE_R) ||
        !make_operand_maps(&maps[2], &maps[3], pb, B, b_planes, s.N, s.K, s.b_sn, s.b_sk, tc.bn_cta))
        return cudaErrorInvalidValue;
    if (p.c_tma) {
        if (!make_map_2d_f32(&maps[4], C, (uint64_t)s.N, (uint64_t)s.M, (uint64_t)s.ldc, EPI_BOX, EPI_BOX, CU_TENSOR_MAP_SWIZZLE_128B))
            p.c_tma = 0;
    }
    if (!p.c_tma) maps[4] = maps[0];   // unused by the kernel

    // 2. split pre-pass: one launch, A's blocks then B's
    // (measurement aid, read on every call: tools/ab_env.py)  1 = no programmatic launches, 2 = only the MMA kernel's
    int const pdl_off = env_int("B200_TF32_NO_PDL", 0);
    bool const no_pdl = pdl_off == 1;
    if (!tc.fused) {
        SplitJob ja = make_job(pa, A, s.M, s.K, s.a_sm, s.a_sk, a_planes);
        SplitJob jb = make_job(pb, B, s.N, s.K, s.b_sn, s.b_sk, b_planes);
        if (reuse_b) {
            jb.mode = SPLIT_NONE;
            jb.blocks = 0;
        }
        if ((e = launch_split(ja, jb, tile_counter, groups, p.turn, n_turn_words, round_hi(), pdl_off == 0, stream)) != cudaSuccess) return e;
        ++n_launch;
    } else if (n_turn_words > 0) {
        // no pre-pass at all; only a split-K call has words to reset (the fused tiles use static assignment)
        if ((e = cudaMemsetAsync(p.turn, 0, sizeof(uint32_t) * (size_t)n_turn_words, stream)) != cudaSuccess) return e;
    }

    // 3. the MMA kernel — behind a split pass it is a programmatic dependent launch: its CTAs set themselves up
    // (barriers, tensor memory, descriptor fetches) while the split is still running and wait for the planes in the
    // kernel (griddep_wait).  Measured (profiles/r02x_ab_pdl.jsonl): 128^3 12.3 -> 9.3 us per call, 512^3 16.4 -> 14.4,
    // 1024^3 25.2 -> 22.5, 2048^3 86.1 -> 83.3; no difference from 8192^3 on.
    bool const pdl = !tc.fused && !no_pdl;
    if (tc.fused) e = ncta == 2 ? launch_gemm<2, false, true>(maps, p, groups, dev, stream) : launch_gemm<1, false, true>(maps, p, groups, dev, stream);
    else if (ncta == 2) e = tc.dynamic ? launch_gemm<2, true>(maps, p, groups, dev, stream, pdl) : launch_gemm<2, false>(maps, p, groups, dev, stream, pdl);
    else e = tc.dynamic ? launch_gemm<1, true>(maps, p, groups, dev, stream, pdl) : launch_gemm<1, false>(maps, p, groups, dev, stream, pdl);
    if (e != cudaSuccess) return e;
    ++n_launch;
    if (launches) *launches = n_launch;
    return cudaSuccess;
}

}  // namespace b200
