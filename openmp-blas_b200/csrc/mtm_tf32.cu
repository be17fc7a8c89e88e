// fp32 mtm on the 5th-generation tensor cores: 3xTF32 with tcgen05.mma / TMEM / TMA (sm_100a).
//
//   C += A*B  with every fp32 operand split as x = hi + lo (both representable in TF32) and
//   A*B ~= Ahi*Bhi + Ahi*Blo + Alo*Bhi   (the lo*lo term, ~2^-22 relative, is dropped),
//   all three products accumulated in fp32 in tensor memory.
//
// Two kernels per call:
//   1. split_planes_kernel — replaces the reference's pack step (include/utils.hpp:99-141,
//      called from mtm.hpp:169-199): reads an operand through its (row, col) strides ONCE and
//      writes two K-contiguous planes hi/lo ([rows_p][K_p], zero padded to tile multiples) into
//      the workspace.  Any layout / stride / alignment of A and B is absorbed here, so the MMA
//      kernel sees a single canonical form: A planes [M_p][K_p], B^T planes [N_p][K_p].
//   2. mtm_tf32x3_kernel — persistent, warp-specialised GEMM: warp 0 = TMA producer (4 tiles per
//      stage: Ahi, Alo, Bhi, Blo, 128B-swizzled, 3 stages), warp 1 = single-thread tcgen05.mma
//      issuer (3 MMAs per 8-wide k step, kind::tf32, fp32 accumulators in TMEM, two accumulator
//      buffers so the epilogue of tile i overlaps the main loop of tile i+1), warp 2 = TMEM
//      allocator, warps 4-7 = epilogue (tcgen05.ld -> C += acc, the reference's copy_from_buff,
//      simd_loop.hpp:160-190).  NCTA = 2 pairs two SMs on one 256 x 256 tile (cta_group::2): each
//      CTA stages its own 128 rows of A and 128 rows of B^T, halving shared-memory reads per SM.
//
// Exactness: integers of magnitude < 2^11 are exact in TF32 (lo == 0) and the fp32 accumulation
// of exact products is exact while partial sums stay below 2^24, so the reference's integer test
// cases are reproduced bit for bit.
#include <cstdlib>

#include "mtm_kernels.h"
#include "sm100_ptx.cuh"

namespace b200 {
namespace {

using namespace ptx;

constexpr int TILE_R = 128;                    // rows of A / of B^T each CTA stages per k-block
constexpr int BK = 32;                         // 32 fp32 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 8;                      // kind::tf32 consumes 32 bytes of K per MMA
constexpr int STAGES = 3;
constexpr int TILE_BYTES = TILE_R * BK * 4;    // 16 KiB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;    // Ahi, Alo, Bhi, Blo
constexpr int NUM_THREADS = 256;
constexpr int ACC_STAGES = 2;
constexpr int SCHED_STAGES = 4;                // ring of tile indices handed out by the dynamic scheduler
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers, scheduler ring*/;
constexpr int PLANE_ROW_ALIGN = 256;           // planes are padded to the largest pair tile

struct Tf32Params {
    float* C;
    int64_t ldc;
    int M, N;
    int num_k_blocks;
    int tiles_m, tiles_n;
    int* tile_counter;   // DYNAMIC: next unclaimed tile (initialised to the number of CTA groups)
    // Gated form (multi-GPU receiver): B's planes are produced WHILE this kernel runs, one 256-column
    // panel of B at a time, in the order the tile schedule first touches them; panel_ready[j] != 0 once
    // rows [256 j, 256 j + 256) of the B^T planes are complete.  nullptr: everything is there already.
    const uint32_t* panel_ready;
    unsigned int* started;    // gated form: CTA groups that are resident and through their set-up (see the host side)
    int group;                // tile-rows walked together before moving to the next tile-column (8 when everything is resident)
};

constexpr int GATE_PANEL = 256;                // columns of B per gating panel (= the pair tile's N)
constexpr int GATE_MAX_PANELS = 512;
constexpr long long GATE_TIMEOUT_CLK = 60000000000LL;   // ~30 s: a lost sender fails loudly instead of hanging the GPU

// Tile hand-out.  STATIC: CTA group g takes tiles g, g + G, g + 2G, ...  DYNAMIC: the first tile is
// g, every further one comes from a global counter — one scheduler thread per CTA group claims it and
// publishes it through a small shared-memory ring (to both CTAs of a pair), so a group that starts
// late or runs slowly (e.g. SMs shared with a concurrent NCCL broadcast) simply takes fewer tiles.
template <int NCTA, bool DYNAMIC>
struct TileSource {
    int it = 0;
    // full_warp: all 32 lanes call next() together (epilogue warps) and lane 0 releases the slot once
    // every lane has read it; otherwise the caller is a single elected thread.
    __device__ __forceinline__ int64_t next(int group_id, int num_groups, int64_t total_tiles, uint64_t* sched_full,
                                            uint64_t* sched_empty, const volatile int* sched_tile, bool do_arrive,
                                            bool full_warp) {
        if constexpr (!DYNAMIC) {
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
template <int NCTA, bool DYNAMIC>
__global__ void __launch_bounds__(NUM_THREADS, 1)
mtm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                  Tf32Params p) {
    constexpr int UMMA_M = 128 * NCTA;
    constexpr int UMMA_N = 128 * NCTA;            // each CTA stages 128 of the N rows of B^T
    constexpr int TMEM_COLS = ACC_STAGES * UMMA_N;  // 256 or 512 (power of two)
    constexpr int EPI_THREADS = 128;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint32_t const align_off = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;  // SWIZZLE_128B needs 1 KiB alignment
    uint8_t* smem = smem_raw + align_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full_bar = bars;                          // [STAGES]   TMA -> MMA
    uint64_t* empty_bar = bars + STAGES;                // [STAGES]   MMA -> TMA
    uint64_t* tmem_full_bar = bars + 2 * STAGES;        // [ACC_STAGES] MMA -> epilogue
    uint64_t* tmem_empty_bar = bars + 2 * STAGES + ACC_STAGES;  // [ACC_STAGES] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * ACC_STAGES);
    uint64_t* sched_full = bars + 2 * STAGES + 2 * ACC_STAGES + 1;      // [SCHED_STAGES] scheduler -> roles (per CTA)
    uint64_t* sched_empty = sched_full + SCHED_STAGES;                  // [SCHED_STAGES] roles -> scheduler (leader's)
    volatile int* sched_tile = reinterpret_cast<volatile int*>(sched_empty + SCHED_STAGES);   // [SCHED_STAGES]

    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t const cta_rank = NCTA == 1 ? 0u : cluster_ctarank();
    bool const is_leader = cta_rank == 0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_hi)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_lo)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b_hi)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b_lo)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], NCTA);   // one arrive(+tx) per CTA of the pair, on the leader's barrier
            mbar_init(&empty_bar[i], 1);     // one tcgen05.commit
        }
        for (int i = 0; i < ACC_STAGES; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], NCTA * EPI_THREADS);
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
    if (p.started != nullptr && is_leader && threadIdx.x == 0) atomicAdd(p.started, 1u);

    int const num_groups = gridDim.x / NCTA;
    int const group_id = blockIdx.x / NCTA;
    int64_t const total_tiles = (int64_t)p.tiles_m * p.tiles_n;

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
            int ready_panel = -1;
            int64_t tile = claims ? (int64_t)group_id
                                  : src.next(group_id, num_groups, total_tiles, sched_full, sched_empty, sched_tile, true, false);
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
                tile_coords_rt(tile, p.tiles_m, p.tiles_n, p.group, pm, pn);
                int const row_a = (int)(pm * UMMA_M) + (int)cta_rank * TILE_R;
                int const row_b = (int)(pn * UMMA_N) + (int)cta_rank * TILE_R;
                if (p.panel_ready != nullptr) {
                    int const panel = row_b / GATE_PANEL;
                    if (panel != ready_panel) {
                        const volatile uint32_t* f = p.panel_ready + panel;
                        long long const t0 = clock64();
                        while (*f == 0u) {
                            __nanosleep(100);
                            if (clock64() - t0 > GATE_TIMEOUT_CLK) __trap();
                        }
                        __threadfence();                                         // acquire the planes' stores ...
                        asm volatile("fence.proxy.async.global;" ::: "memory");  // ... for the TMA (async proxy) reads
                        ready_panel = panel;
                    }
                }
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* s = smem + stage * STAGE_BYTES;
                    int const k0 = kb * BK;
                    tma_load_2d<NCTA>(&map_a_hi, &full_bar[stage], s + 0 * TILE_BYTES, k0, row_a);
                    tma_load_2d<NCTA>(&map_a_lo, &full_bar[stage], s + 1 * TILE_BYTES, k0, row_a);
                    tma_load_2d<NCTA>(&map_b_hi, &full_bar[stage], s + 2 * TILE_BYTES, k0, row_b);
                    tma_load_2d<NCTA>(&map_b_lo, &full_bar[stage], s + 3 * TILE_BYTES, k0, row_b);
                    if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES * NCTA);
                    else mbar_arrive_cluster(&full_bar[stage], 0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (claims) {
                    int64_t const claimed = (int64_t)atomicAdd(p.tile_counter, 1);
                    tile = claimed < total_tiles ? claimed : -1;
                } else {
                    tile = src.next(group_id, num_groups, total_tiles, sched_full, sched_empty, sched_tile, true, false);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only, one elected lane) =====
        if (is_leader && elect_one()) {
            constexpr uint32_t idesc = make_idesc_tf32(UMMA_M, UMMA_N);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            TileSource<NCTA, DYNAMIC> src;
            for (; src.next(group_id, num_groups, total_tiles, sched_full, sched_empty, sched_tile, true, false) >= 0; ++it) {
                int const acc = it % ACC_STAGES;
                uint32_t const acc_phase = (uint32_t)(it / ACC_STAGES) & 1u;
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
                tcgen05_fence_after();
                uint32_t const tmem_d = tmem_base + (uint32_t)(acc * UMMA_N);
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    uint32_t const s = smem_u32(smem + stage * STAGE_BYTES);
                    uint64_t const a_hi = make_kmajor_sw128_desc(s + 0 * TILE_BYTES);
                    uint64_t const a_lo = make_kmajor_sw128_desc(s + 1 * TILE_BYTES);
                    uint64_t const b_hi = make_kmajor_sw128_desc(s + 2 * TILE_BYTES);
                    uint64_t const b_lo = make_kmajor_sw128_desc(s + 3 * TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        uint64_t const adv = (uint64_t)((k * UMMA_K * 4) >> 4);  // +32 B per k step, in 16-B units
                        // small terms first, then the dominant hi*hi product
                        umma_tf32<NCTA>(tmem_d, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_tf32<NCTA>(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
                        umma_tf32<NCTA>(tmem_d, a_hi + adv, b_hi + adv, idesc, 1u);
                    }
                    umma_commit<NCTA>(&empty_bar[stage]);                       // frees the smem stage (both CTAs)
                    if (kb == p.num_k_blocks - 1) umma_commit<NCTA>(&tmem_full_bar[acc]);  // accumulator ready
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> C += acc =====
        int const ew = warp & 3;                        // TMEM lane quarter this warp may access
        int it = 0;
        TileSource<NCTA, DYNAMIC> src;
        for (int64_t tile; (tile = src.next(group_id, num_groups, total_tiles, sched_full, sched_empty, sched_tile, lane == 0, true)) >= 0; ++it) {
            int const acc = it % ACC_STAGES;
            uint32_t const acc_phase = (uint32_t)(it / ACC_STAGES) & 1u;
            int64_t pm, pn;
            tile_coords_rt(tile, p.tiles_m, p.tiles_n, p.group, pm, pn);
            int64_t const row = pm * UMMA_M + (int64_t)cta_rank * TILE_R + ew * 32 + lane;
            int64_t const col0 = pn * UMMA_N;
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tcgen05_fence_after();
            float* crow = p.C + row * p.ldc;
            bool const row_ok = row < p.M;
            bool const vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15u) == 0);
#pragma unroll 1
            for (int c = 0; c < UMMA_N / 32; ++c) {
                uint32_t v[32];
                uint32_t const taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * UMMA_N + c * 32);
                tmem_ld_32x32b_x32(taddr, v);
                tmem_ld_wait();
                int64_t const n0 = col0 + c * 32;
                if (row_ok && n0 < p.N) {
                    if (vec_ok && n0 + 32 <= p.N) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 cv = *reinterpret_cast<float4*>(crow + n0 + j);
                            cv.x += __uint_as_float(v[j]);
                            cv.y += __uint_as_float(v[j + 1]);
                            cv.z += __uint_as_float(v[j + 2]);
                            cv.w += __uint_as_float(v[j + 3]);
                            *reinterpret_cast<float4*>(crow + n0 + j) = cv;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n0 + j < p.N) crow[n0 + j] += __uint_as_float(v[j]);
                    }
                }
            }
            tcgen05_fence_before();
            mbar_arrive_cluster(&tmem_empty_bar[acc], 0);   // accumulator may be overwritten
        }
    }

    tcgen05_fence_before();
    if constexpr (NCTA == 1) __syncthreads(); else cluster_sync_all();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc<NCTA>(tmem_base, TMEM_COLS);
    }
}

// ---- operand split pre-pass ------------------------------------------------------------------------------
// out_hi/out_lo: [rows_p][kp] K-contiguous planes.  in(r, k) = in[r * s_r + k * s_k] for r < rows, k < K,
// zero outside.  hi = rn_tf32(x), lo = rn_tf32(x - hi): both exactly representable in TF32, so the
// tensor core's handling of the low 13 mantissa bits of its inputs is irrelevant.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
    lo = __uint_as_float(l);
}

struct SplitGate {
    int rblk0;                               // first 32-row block of the operand this launch covers
    unsigned int* done_counter;              // CTAs of this launch that have finished (zeroed by the caller)
    uint32_t* ready_flag;                    // nullptr: nobody waits for this launch in-kernel
};

template <bool K_CONTIG>
__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ in, int64_t s_r, int64_t s_k, int rows, int K,
                    float* __restrict__ out_hi, float* __restrict__ out_lo, int kp, int* tile_counter,
                    int counter_init, SplitGate gate) {
    __shared__ float tile[32][33];
    // The A split always precedes the MMA kernel on the stream: it also re-arms the dynamic tile counter.
    if (tile_counter != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *tile_counter = counter_init;
    int const tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    int const r0 = (blockIdx.y + gate.rblk0) * 32, k0 = blockIdx.x * 32;
    if constexpr (K_CONTIG) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int const r = r0 + ty + 8 * i, k = k0 + tx;
            float x = (r < rows && k < K) ? in[(int64_t)r * s_r + (int64_t)k * s_k] : 0.f;
            float hi, lo;
            split_tf32(x, hi, lo);
            out_hi[(int64_t)r * kp + k] = hi;
            out_lo[(int64_t)r * kp + k] = lo;
        }
    } else {
        // read with the warp running along r (coalesced when s_r == 1), transpose through smem
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int const r = r0 + tx, k = k0 + ty + 8 * i;
            tile[ty + 8 * i][tx] = (r < rows && k < K) ? in[(int64_t)r * s_r + (int64_t)k * s_k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int const rl = ty + 8 * i;
            float hi, lo;
            split_tf32(tile[tx][rl], hi, lo);
            out_hi[(int64_t)(r0 + rl) * kp + k0 + tx] = hi;
            out_lo[(int64_t)(r0 + rl) * kp + k0 + tx] = lo;
        }
    }
    if (gate.ready_flag != nullptr) {
        // Publish the panel: every CTA fences its plane stores, the last one to finish raises the flag the
        // running MMA kernel's producers poll.
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int const prev = atomicAdd(gate.done_counter, 1u);
            if (prev + 1 == gridDim.x * gridDim.y) {
                __threadfence();
                *reinterpret_cast<volatile uint32_t*>(gate.ready_flag) = 1u;
            }
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
bool make_plane_map(CUtensorMap* map, float* plane, int rows_p, int kp) {
    return ptx::make_map_2d_f32(map, plane, (uint64_t)kp, (uint64_t)rows_p, (uint64_t)kp, BK, TILE_R,
                                CU_TENSOR_MAP_SWIZZLE_128B);
}

inline int round_up(int64_t x, int a) { return (int)((x + a - 1) / a * a); }

const TileConfig kCfg[] = {
    {"tf32x3_2cta_256x256x32", 256, 256, 32, NUM_THREADS, 1},        // static tile assignment
    {"tf32x3_1cta_128x128x32", 128, 128, 32, NUM_THREADS, 1},
    {"tf32x3_2cta_256x256x32_dyn", 256, 256, 32, NUM_THREADS, 1},    // dynamic tile scheduler
    {"tf32x3_1cta_128x128x32_dyn", 128, 128, 32, NUM_THREADS, 1},
};

template <int NCTA, bool DYNAMIC>
cudaError_t launch_gemm(const CUtensorMap* maps, const Tf32Params& p, int groups, cudaStream_t stream, bool gated) {
    cudaError_t ea = cudaFuncSetAttribute(mtm_tf32x3_kernel<NCTA, DYNAMIC>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (ea != cudaSuccess) return ea;
    // The kernel's 193.5 KiB fit the 196 KiB shared-memory configuration of an SM, which leaves nothing for
    // another CTA.  The gated form needs the panel-split CTAs (5 KiB each) to run NEXT TO the resident
    // persistent CTAs, so it asks for the full 228 KiB carve-out (measured: without it the splits never get
    // an SM and the product deadlocks on its own flags).
    static bool const force_max = std::getenv("B200_TF32_MAX_CARVEOUT") != nullptr;   // measurement aid: A/B the carve-out alone
    ea = cudaFuncSetAttribute(mtm_tf32x3_kernel<NCTA, DYNAMIC>, cudaFuncAttributePreferredSharedMemoryCarveout,
                              (gated || force_max) ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault);
    if (ea != cudaSuccess) return ea;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(groups * NCTA));
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NCTA;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, mtm_tf32x3_kernel<NCTA, DYNAMIC>, maps[0], maps[1], maps[2], maps[3], p);
}

}  // namespace

// CUDA loads kernels lazily, at their first launch, and that load can need the context to be idle.  The
// gated form launches kernels that spin until OTHER kernels have run, so every kernel involved has to be
// resident before the first spinning one starts (CUDA programming guide, lazy loading: concurrent execution).
cudaError_t tf32_preload_kernels() {
    cudaFuncAttributes fa;
    cudaError_t e;
    if ((e = cudaFuncGetAttributes(&fa, split_planes_kernel<true>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, split_planes_kernel<false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, mtm_tf32x3_kernel<1, false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, mtm_tf32x3_kernel<1, true>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, mtm_tf32x3_kernel<2, false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, mtm_tf32x3_kernel<2, true>)) != cudaSuccess) return e;
    return cudaSuccess;
}

int tf32_num_configs() { return (int)(sizeof(kCfg) / sizeof(kCfg[0])); }
const TileConfig& tf32_config(int cfg) { return kCfg[cfg]; }

size_t tf32_workspace_bytes(const MtmShape& s) {
    size_t const kp = (size_t)round_up(s.K, BK);
    size_t const mp = (size_t)round_up(s.M, PLANE_ROW_ALIGN), np = (size_t)round_up(s.N, PLANE_ROW_ALIGN);
    return 2 * sizeof(float) * kp * (mp + np) + 8192;   // + alignment slack, tile counter, gating flags
}

cudaError_t launch_3xtf32_f32(float* C, const float* A, const float* B, const MtmShape& s, void* ws,
                              size_t ws_bytes, int cfg, int reuse_b, int reserve_sms, cudaStream_t stream,
                              int* launches, const Tf32Gate* gate) {
    if (launches) *launches = 0;
    if (gate != nullptr && (s.b_sn != 1 || reuse_b || (s.N + GATE_PANEL - 1) / GATE_PANEL > GATE_MAX_PANELS))
        return cudaErrorInvalidValue;   // gating is by column panels of a row-major B
    if (ws_bytes < tf32_workspace_bytes(s)) return cudaErrorInvalidValue;
    int const kp = round_up(s.K, BK);
    int const mp = round_up(s.M, PLANE_ROW_ALIGN), np = round_up(s.N, PLANE_ROW_ALIGN);
    float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
    // B planes first: their position does not depend on M, so a caller slicing M can keep them.
    float* b_hi = base;
    float* b_lo = b_hi + (size_t)np * kp;
    float* a_hi = b_lo + (size_t)np * kp;
    float* a_lo = a_hi + (size_t)mp * kp;
    int n_launch = 0;
    int* tile_counter = reinterpret_cast<int*>(a_lo + (size_t)mp * kp);   // inside the 4 KiB of slack

    int dev = 0, sm_count = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if (reserve_sms > 0 && reserve_sms < sm_count - 2) sm_count -= reserve_sms;
    int const ncta = (cfg == 0 || cfg == 2) ? 2 : 1;
    bool const dynamic = cfg >= 2;
    Tf32Params p;
    p.C = C;
    p.ldc = s.ldc;
    p.M = (int)s.M;
    p.N = (int)s.N;
    p.num_k_blocks = kp / BK;
    p.tiles_m = (int)((s.M + 128 * ncta - 1) / (128 * ncta));
    p.tiles_n = (int)((s.N + 128 * ncta - 1) / (128 * ncta));
    p.tile_counter = tile_counter;
    // gating words live in the workspace slack behind the tile counter: [64, 64+512) ready, [640, 640+512) done
    uint32_t* panel_ready = reinterpret_cast<uint32_t*>(tile_counter) + 64;
    unsigned int* panel_done = reinterpret_cast<unsigned int*>(tile_counter) + 64 + GATE_MAX_PANELS + 64;
    p.panel_ready = gate != nullptr ? panel_ready : nullptr;
    unsigned int* started = reinterpret_cast<unsigned int*>(tile_counter) + 32;   // zeroed with the flags below
    p.started = gate != nullptr ? started : nullptr;
    // Ungated: groups of 8 tile-rows (A row-panels and B column-panels of the running wave stay in L2).
    // Gated: B arrives one column panel at a time, so a wave should touch FEW panels — with 8 rows per group
    // the first wave of 74 pair tiles needs panels 0..9 at once, with 16 rows 5, with 32 rows 3 — but taller
    // groups cost L2 locality on A.  Measured at 8192^3 with every panel already there (profiles/r01x_*):
    // 8 rows 4.37-4.46 ms, 16 rows 4.22-4.45, 32 rows 4.56-4.62 (ungated 4.08): 16 is the default.
    static int const env_ungated_group = [] { const char* v = std::getenv("B200_TF32_GROUP"); return v ? std::atoi(v) : 0; }();
    p.group = env_ungated_group > 0 ? env_ungated_group : 8;   // (measurement aid: A/B the L2 locality of the walk)
    if (gate != nullptr) {
        static int const env_group = [] { const char* v = std::getenv("B200_GATE_GROUP"); return v ? std::atoi(v) : 0; }();
        int const want = env_group > 0 ? env_group : 16;
        p.group = p.tiles_m < want ? (p.tiles_m > 0 ? p.tiles_m : 1) : want;
    }
    int64_t const total_tiles = (int64_t)p.tiles_m * p.tiles_n;
    int groups = sm_count / ncta;
    if (total_tiles < groups) groups = (int)total_tiles;

    // 1. split pre-pass: A as rows = m; B as rows = n (i.e. B^T), both K-contiguous planes.
    dim3 const blk(256);
    dim3 const ga((unsigned)(kp / 32), (unsigned)(mp / 32)), gb((unsigned)(kp / 32), (unsigned)(np / 32));
    SplitGate const no_gate{0, nullptr, nullptr};
    auto launch_split_a = [&]() {
        if (s.a_sk == 1)
            split_planes_kernel<true><<<ga, blk, 0, stream>>>(A, s.a_sm, s.a_sk, (int)s.M, (int)s.K, a_hi, a_lo, kp, tile_counter, groups, no_gate);
        else
            split_planes_kernel<false><<<ga, blk, 0, stream>>>(A, s.a_sm, s.a_sk, (int)s.M, (int)s.K, a_hi, a_lo, kp, tile_counter, groups, no_gate);
        ++n_launch;
    };
    if (gate != nullptr) {
        // Gated form.  Launch ORDER matters: streams may share a hardware work queue, and a queued launch
        // that waits on its stream predecessor blocks everything behind it in that queue.  So the MMA kernel
        // goes in first (it only spins on flags), the panel splits — each of which has to wait for the one
        // before — go in after it.
        if ((e = cudaMemsetAsync(started, 0, sizeof(uint32_t) * (32 + 2 * GATE_MAX_PANELS + 64), stream)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(gate->fork, stream)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(gate->side, gate->fork, 0)) != cudaSuccess) return e;
    }
    launch_split_a();
    if (!reuse_b && gate == nullptr) {
        if (s.b_sk == 1)
            split_planes_kernel<true><<<gb, blk, 0, stream>>>(B, s.b_sn, s.b_sk, (int)s.N, (int)s.K, b_hi, b_lo, kp, nullptr, 0, no_gate);
        else
            split_planes_kernel<false><<<gb, blk, 0, stream>>>(B, s.b_sn, s.b_sk, (int)s.N, (int)s.K, b_hi, b_lo, kp, nullptr, 0, no_gate);
        ++n_launch;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (launches) *launches = n_launch;

    // 2. tensor maps over the planes + the MMA kernel
    CUtensorMap maps[4];
    if (!make_plane_map(&maps[0], a_hi, mp, kp) || !make_plane_map(&maps[1], a_lo, mp, kp) ||
        !make_plane_map(&maps[2], b_hi, np, kp) || !make_plane_map(&maps[3], b_lo, np, kp))
        return cudaErrorInvalidValue;
    bool const gated = gate != nullptr;
    if (ncta == 2) e = dynamic ? launch_gemm<2, true>(maps, p, groups, stream, gated) : launch_gemm<2, false>(maps, p, groups, stream, gated);
    else e = dynamic ? launch_gemm<1, true>(maps, p, groups, stream, gated) : launch_gemm<1, false>(maps, p, groups, stream, gated);
    if (e != cudaSuccess) return e;
    ++n_launch;
    if (gate != nullptr) {
        // B's planes are built panel by panel on the side stream while the MMA kernel already runs on
        // `stream`: wait for the panel's arrival, split it, raise panel_ready (last CTA of the split).
        int const n_panels = (int)((s.N + GATE_PANEL - 1) / GATE_PANEL);
        // Waits are stream memory operations where the driver offers them (no SM, no CTA slot); the first one
        // holds the side stream until every CTA group of the MMA kernel is resident: small kernels that get to an
        // SM first keep its shared-memory configuration small, the persistent CTAs then trickle in over the whole
        // panel chain, and a statically scheduled pair that starts late finishes late (measured: +0.5-0.7 ms at
        // 8192^3 with every panel already there, profiles/r01v_*).
        StreamWaitValue32Fn const wait32 = get_stream_wait_value32_fn();
        if (wait32 != nullptr) {
            if (wait32(gate->side, (CUdeviceptr)(uintptr_t)started, (cuuint32_t)groups, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                return cudaErrorUnknown;
        }
        for (int j = 0; j < n_panels; ++j) {
            int const rblk0 = j * (GATE_PANEL / 32);
            int const rblks = (np / 32 - rblk0) < GATE_PANEL / 32 ? (np / 32 - rblk0) : GATE_PANEL / 32;
            SplitGate const g{rblk0, panel_done + j, panel_ready + j};
            // One warp waits for the panel's arrival; the split itself never spins (spinning CTAs all over the
            // machine would keep the SMs from being re-configured for the MMA kernel and, being many, would
            // sit in front of anything else that has to run).
            if (wait32 != nullptr) {
                if (wait32(gate->side, (CUdeviceptr)(uintptr_t)gate->arrival_flag, (cuuint32_t)(gate->first_seq + (uint32_t)j),
                           CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                    return cudaErrorUnknown;
            } else {
                if ((e = launch_flag_wait(gate->arrival_flag, gate->first_seq + (uint32_t)j, 1, 1, -1, gate->side)) != cudaSuccess) return e;
                ++n_launch;
            }
            split_planes_kernel<false><<<dim3((unsigned)(kp / 32), (unsigned)rblks), blk, 0, gate->side>>>(
                B, s.b_sn, s.b_sk, (int)s.N, (int)s.K, b_hi, b_lo, kp, nullptr, 0, g);
            ++n_launch;
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if ((e = cudaEventRecord(gate->join, gate->side)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(stream, gate->join, 0)) != cudaSuccess) return e;   // formal join of the side stream
    }
    if (launches) *launches = n_launch;
    return cudaSuccess;
}

}  // namespace b200
