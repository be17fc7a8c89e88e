// fp32 FFMA mtm kernel fed by TMA (sm_100a): C += A*B on the CUDA cores with shared-memory tiles
// staged by cp.async.bulk.tensor into a multi-stage mbarrier ring.
//
// Compared with the register-staged kernel (mtm_simt.cuh) the inner loop carries no global loads,
// no shared-memory stores and no address arithmetic: each thread executes only 16-byte LDS
// fragment reads and FFMAs (8x8 accumulator block per thread, 128x128 tile per CTA, 2 CTAs per
// SM), one elected thread re-arms a stage with two TMA copies per BK-slice, and out-of-range
// rows/columns/k are zero-filled by the TMA unit (no predicated loads).
//
// Operand form: TMA cannot transpose, so both tiles are fetched "mn-contiguous" — A(m,k) with unit
// stride along m, B(k,n) with unit stride along n, landing in shared memory as [k][m] / [k][n].
// An operand that is k-contiguous, arbitrarily strided or mis-aligned is first re-laid into that
// form by pack_mn_kernel (one pass over the operand, the GPU counterpart of the reference's
// amt::pack, include/utils.hpp:99-141, which the reference runs for every block anyway); an
// operand that already has the form is read in place.
#include "mtm_kernels.h"
#include "sm100_ptx.cuh"

namespace b200 {
namespace {

using namespace ptx;

constexpr int BN = 128, TN = 8;   // every config: 16 threads across n, 8 columns per thread
constexpr int NT = 256;           // 16 x 16 threads; rows per thread TM = BM / 16

__device__ __forceinline__ void tma_load_tile(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// Packed 2-wide FP32 FMA (PTX fma.rn.f32x2, SASS FFMA2 on sm_100): d.lo = a.lo*b.lo + d.lo,
// d.hi = a.hi*b.hi + d.hi, each IEEE round-to-nearest — bit-identical to two fmaf().  ptxas folds a
// {x, x} pair into the scalar-broadcast operand form (FFMA2 Rd, Ra.F32, Rb.F32x2.HI_LO, Rd...), so
// an 8x8 outer-product step is 32 issue slots instead of 64: the FMA pipe (2 cycles per FFMA2)
// stays busy while the LDS fragment reads take the free slots.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void fma2(uint64_t& d, uint64_t a, uint64_t b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ float2 unpack2(uint64_t v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

struct FfmaTmaParams {
    float* C;
    int64_t ldc;
    int M, N;
    int num_k_blocks;
    int64_t tiles_m, tiles_n;
    int vec_c;
};

template <int BM, int BK, int STAGES, bool PACKED>
__global__ void __launch_bounds__(NT, (BM <= 64 ? 3 : (BM <= 128 ? 2 : 1)))
mtm_ffma_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    FfmaTmaParams p) {
    constexpr int TM = BM / 16;
    constexpr int TX = BN / TN, TY = BM / TM;
    static_assert(TX * TY == NT && TM % 4 == 0, "thread grid");
    constexpr int A_ELEMS = BK * BM, B_ELEMS = BK * BN;
    constexpr uint32_t STAGE_BYTES = (A_ELEMS + B_ELEMS) * 4;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // Offset computed in the shared window so the compiler keeps LDS (not generic LD) addressing.
    uint32_t const align_off = (128u - (smem_u32(smem_raw) & 127u)) & 127u;
    float* tiles = reinterpret_cast<float*>(smem_raw + align_off);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + STAGES * (A_ELEMS + B_ELEMS));

    int const tid = threadIdx.x;
    int const lane = tid & 31, warp = tid >> 5;
    constexpr int WX = TX / 8;
    int const tx = (warp % WX) * 8 + (lane & 7);
    int const ty = (warp / WX) * 4 + (lane >> 3);

    int64_t pid_m, pid_n;
    tile_coords<8>(blockIdx.x, p.tiles_m, p.tiles_n, pid_m, pid_n);
    int const m0 = (int)(pid_m * BM), n0 = (int)(pid_n * BN);
    int const nkb = p.num_k_blocks;

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
        for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        int const pre = nkb < STAGES ? nkb : STAGES;
        for (int s = 0; s < pre; ++s) {
            float* sa = tiles + s * (A_ELEMS + B_ELEMS);
            mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
            tma_load_tile(&map_a, &full_bar[s], sa, m0, s * BK);
            tma_load_tile(&map_b, &full_bar[s], sa + A_ELEMS, n0, s * BK);
        }
    }

    // Accumulators: scalar acc[i][j], or packed pairs acc2[i][j/2] = {acc[i][j], acc[i][j+1]}.
    float acc[TM][TN];
    uint64_t acc2[TM][TN / 2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            acc[i][j] = 0.f;
            acc2[i][j / 2] = 0ull;
        }

    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait_lean(&full_bar[stage], phase);
        const float* as = tiles + stage * (A_ELEMS + B_ELEMS) + ty * 4;
        const float* bs = tiles + stage * (A_ELEMS + B_ELEMS) + A_ELEMS + tx * 4;
#pragma unroll 8
        for (int k = 0; k < BK; ++k) {
            float af[TM];
#pragma unroll
            for (int i = 0; i < TM / 4; ++i)
                *reinterpret_cast<float4*>(&af[i * 4]) = *reinterpret_cast<const float4*>(as + k * BM + i * TY * 4);
            float4 const b0 = *reinterpret_cast<const float4*>(bs + k * BN);
            float4 const b1 = *reinterpret_cast<const float4*>(bs + k * BN + TX * 4);
            if constexpr (PACKED) {
                uint64_t const bp[TN / 2] = {pack2(b0.x, b0.y), pack2(b0.z, b0.w), pack2(b1.x, b1.y), pack2(b1.z, b1.w)};
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    uint64_t const ai = pack2(af[i], af[i]);
#pragma unroll
                    for (int j = 0; j < TN / 2; ++j) fma2(acc2[i][j], ai, bp[j]);
                }
            } else {
                float const bf[TN] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
            }
        }
        __syncthreads();  // every thread is done reading this stage
        if (tid == 0 && kb + STAGES < nkb) {
            float* sa = tiles + stage * (A_ELEMS + B_ELEMS);
            mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            tma_load_tile(&map_a, &full_bar[stage], sa, m0, (kb + STAGES) * BK);
            tma_load_tile(&map_b, &full_bar[stage], sa + A_ELEMS, n0, (kb + STAGES) * BK);
        }
        if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
        }
    }

    if constexpr (PACKED) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN / 2; ++j) {
                float2 const v = unpack2(acc2[i][j]);
                acc[i][2 * j] = v.x;
                acc[i][2 * j + 1] = v.y;
            }
    }

    // Epilogue: C += acc   (reference: copy_from_buff, simd_loop.hpp:160-190).
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int64_t const m = (int64_t)m0 + ty * 4 + (i / 4) * (TY * 4) + (i % 4);
        if (m >= p.M) continue;
        float* crow = p.C + m * p.ldc;
#pragma unroll
        for (int j = 0; j < TN / 4; ++j) {
            int64_t const n = (int64_t)n0 + tx * 4 + j * (TX * 4);
            if (p.vec_c && n + 4 <= p.N) {
                float4 c = *reinterpret_cast<float4*>(crow + n);
                c.x += acc[i][j * 4 + 0];
                c.y += acc[i][j * 4 + 1];
                c.z += acc[i][j * 4 + 2];
                c.w += acc[i][j * 4 + 3];
                *reinterpret_cast<float4*>(crow + n) = c;
            } else {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    if (n + jj < p.N) crow[n + jj] += acc[i][j * 4 + jj];
            }
        }
    }
}

// out[k * ldp + mn] = in(mn, k) = in[mn * s_mn + k * s_k]  for mn < MN, k < K.
// K_CONTIG: the warp reads along k (coalesced when s_k == 1) and transposes through shared memory;
// otherwise the warp reads along mn.  Writes are always coalesced along mn.
template <bool K_CONTIG>
__global__ void __launch_bounds__(256)
pack_mn_kernel(const float* __restrict__ in, int64_t s_mn, int64_t s_k, int MN, int K, float* __restrict__ out,
               int64_t ldp) {
    __shared__ float tile[32][33];
    int const tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    int const mn0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    if constexpr (K_CONTIG) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int const mn = mn0 + ty + 8 * i, k = k0 + tx;
            tile[ty + 8 * i][tx] = (mn < MN && k < K) ? in[(int64_t)mn * s_mn + (int64_t)k * s_k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int const k = k0 + ty + 8 * i, mn = mn0 + tx;
            if (k < K && mn < MN) out[(int64_t)k * ldp + mn] = tile[tx][ty + 8 * i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int const k = k0 + ty + 8 * i, mn = mn0 + tx;
            if (k < K && mn < MN) out[(int64_t)k * ldp + mn] = in[(int64_t)mn * s_mn + (int64_t)k * s_k];
        }
    }
}

inline int64_t round_up64(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

bool tma_direct_ok(const float* p, int64_t s_mn, int64_t s_k, int64_t extent_mn) {
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && s_mn == 1 && s_k % 4 == 0 && s_k >= extent_mn;
}

const TileConfig kCfg[] = {
    {"ffma_tma_128x128x32_s3", 128, 128, 32, NT, 2},     // scalar FFMA, 8x8 per thread (default)
    {"ffma2_tma_128x128x32_s3", 128, 128, 32, NT, 2},    // packed FFMA2, same pipeline
    {"ffma2_tma_256x128x32_s3", 256, 128, 32, NT, 1},    // 16x8 per thread: 25% less smem->RF traffic per FMA,
                                                         // but 1 CTA/SM (178 regs): measured slower (57 vs 60 TFLOP/s)
    {"ffma_tma_64x128x32_s3", 64, 128, 32, NT, 3},       // 4x8 per thread, 3 CTAs/SM: twice the tiles for mid-size problems
};

template <int BM, int BK, int STAGES, bool PACKED>
cudaError_t launch_cfg(const CUtensorMap& ma, const CUtensorMap& mb, FfmaTmaParams p, int K, cudaStream_t stream) {
    constexpr size_t smem = (size_t)STAGES * BK * (BM + BN) * 4 + STAGES * 8 + 256;
    cudaError_t const ea = cudaFuncSetAttribute(mtm_ffma_tma_kernel<BM, BK, STAGES, PACKED>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ea != cudaSuccess) return ea;
    p.num_k_blocks = (K + BK - 1) / BK;
    int64_t const grid = p.tiles_m * p.tiles_n;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    mtm_ffma_tma_kernel<BM, BK, STAGES, PACKED><<<dim3((unsigned)grid), dim3(NT), smem, stream>>>(ma, mb, p);
    return cudaGetLastError();
}

}  // namespace

int ffma_tma_num_configs() { return (int)(sizeof(kCfg) / sizeof(kCfg[0])); }
const TileConfig& ffma_tma_config(int cfg) { return kCfg[cfg]; }

size_t ffma_tma_workspace_bytes(const MtmShape& s) {
    return 4 * (size_t)s.K * (size_t)(round_up64(s.M, 4) + round_up64(s.N, 4)) + 1024;
}

cudaError_t launch_ffma_tma_f32(int cfg, float* C, const float* A, const float* B, const MtmShape& s, void* ws,
                                size_t ws_bytes, int vec_c, int reuse_b, cudaStream_t stream, int* launches) {
    if (launches) *launches = 0;
    int n_launch = 0;
    float* wsf = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
    const float* a_src = A;
    int64_t a_ld = s.a_sk;
    const float* b_src = B;
    int64_t b_ld = s.b_sk;
    dim3 const blk(256);
    if (!tma_direct_ok(A, s.a_sm, s.a_sk, s.M)) {
        if (ws_bytes < ffma_tma_workspace_bytes(s)) return cudaErrorInvalidValue;
        int64_t const ldp = round_up64(s.M, 4);
        // B's plane comes first in the workspace (its position must not depend on M, see reuse_b)
        float* dst = wsf + (size_t)s.K * (size_t)round_up64(s.N, 4);
        dim3 const g((unsigned)((s.M + 31) / 32), (unsigned)((s.K + 31) / 32));
        if (s.a_sk == 1 || s.a_sk < s.a_sm)
            pack_mn_kernel<true><<<g, blk, 0, stream>>>(A, s.a_sm, s.a_sk, (int)s.M, (int)s.K, dst, ldp);
        else
            pack_mn_kernel<false><<<g, blk, 0, stream>>>(A, s.a_sm, s.a_sk, (int)s.M, (int)s.K, dst, ldp);
        a_src = dst;
        a_ld = ldp;
        ++n_launch;
    }
    if (!tma_direct_ok(B, s.b_sn, s.b_sk, s.N)) {
        if (ws_bytes < ffma_tma_workspace_bytes(s)) return cudaErrorInvalidValue;
        float* dst = wsf;
        int64_t const ldp = round_up64(s.N, 4);
        dim3 const g((unsigned)((s.N + 31) / 32), (unsigned)((s.K + 31) / 32));
        if (!reuse_b) {
            if (s.b_sk == 1 || s.b_sk < s.b_sn)
                pack_mn_kernel<true><<<g, blk, 0, stream>>>(B, s.b_sn, s.b_sk, (int)s.N, (int)s.K, dst, ldp);
            else
                pack_mn_kernel<false><<<g, blk, 0, stream>>>(B, s.b_sn, s.b_sk, (int)s.N, (int)s.K, dst, ldp);
            ++n_launch;
        }
        b_src = dst;
        b_ld = ldp;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;

    int const bk = kCfg[cfg].bk;
    int const BM = kCfg[cfg].bm;
    CUtensorMap ma, mb;
    if (!make_map_2d_f32(&ma, a_src, (uint64_t)s.M, (uint64_t)s.K, (uint64_t)a_ld, (uint32_t)BM, (uint32_t)bk,
                         CU_TENSOR_MAP_SWIZZLE_NONE) ||
        !make_map_2d_f32(&mb, b_src, (uint64_t)s.N, (uint64_t)s.K, (uint64_t)b_ld, BN, (uint32_t)bk,
                         CU_TENSOR_MAP_SWIZZLE_NONE))
        return cudaErrorInvalidValue;
    FfmaTmaParams p;
    p.C = C;
    p.ldc = s.ldc;
    p.M = (int)s.M;
    p.N = (int)s.N;
    p.num_k_blocks = 0;
    p.tiles_m = (s.M + BM - 1) / BM;
    p.tiles_n = (s.N + BN - 1) / BN;
    p.vec_c = vec_c;
    switch (cfg) {
        case 0: e = launch_cfg<128, 32, 3, false>(ma, mb, p, (int)s.K, stream); break;
        case 1: e = launch_cfg<128, 32, 3, true>(ma, mb, p, (int)s.K, stream); break;
        case 2: e = launch_cfg<256, 32, 3, true>(ma, mb, p, (int)s.K, stream); break;
        case 3: e = launch_cfg<64, 32, 3, false>(ma, mb, p, (int)s.K, stream); break;
        default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    if (launches) *launches = n_launch + 1;
    return cudaSuccess;
}

}  // namespace b200
