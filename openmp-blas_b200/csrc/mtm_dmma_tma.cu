// fp64 DMMA mtm kernel fed by TMA (sm_100a): C += A*B with mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) and
// shared-memory tiles staged by cp.async.bulk.tensor into a multi-stage mbarrier ring.
//
// Both operands are fetched K-contiguous — A(m, k) with unit stride along k, B as B^T(n, k) — into
// 128-byte-swizzled tiles [row][16 k].  That layout makes the DMMA fragment reads conflict-free: a
// fragment is (8 rows) x (4 k) doubles, lane l reading row l/4, k l%4; with SWIZZLE_128B the 16-byte
// chunk index is (k/2) ^ (row % 8), so the 16 chunks a warp touches spread over all 8 chunk columns
// exactly twice — the two-wavefront minimum for 256 bytes.  (A dense [k][m] tile, as the FFMA kernel
// uses, would put the 4 k-rows of a fragment on the same banks: a 4-way conflict.)
// An operand that is not K-contiguous / aligned is first re-laid by pack_k_kernel (one pass, the GPU
// counterpart of the reference's amt::pack, include/utils.hpp:99-141).
//
// 64 x 64 tile per CTA, 4 warps (2 x 2, 32 x 32 per warp = 16 DMMA tiles, 32 accumulator doubles per
// thread), BK = 16, 3 stages of 16 KiB with 4 CTAs per SM (or 4 stages / 3 CTAs).  Inner loop: LDS.64
// fragment reads + DMMA only.  Measured 8192^3: 36.2 TFLOP/s = 0.97 of the FP64 pipe (register-staged
// kernel: 32.4).
#include "mtm_kernels.h"
#include "sm100_ptx.cuh"

namespace b200 {
namespace {

using namespace ptx;

constexpr int DT = 64;          // tile edge (rows of A / rows of B^T per CTA)
constexpr int DBK = 16;         // doubles per swizzled row (128 bytes)
constexpr int DTHREADS = 128;
constexpr int TILE_BYTES_D = DT * DBK * 8;          // 8 KiB
constexpr int STAGE_BYTES_D = 2 * TILE_BYTES_D;     // A + B^T
constexpr size_t smem_bytes_d(int stages) { return (size_t)stages * STAGE_BYTES_D + 1024 + 128; }

__device__ __forceinline__ void tma_load_tile_d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

struct DmmaTmaParams {
    double* C;
    int64_t ldc;
    int M, N;
    int num_k_blocks;
    int64_t tiles_m, tiles_n;
    int vec_c;
};

template <int DSTAGES, int MINB>
__global__ void __launch_bounds__(DTHREADS, MINB)
mtm_dmma_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    DmmaTmaParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint32_t const align_off = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;   // SWIZZLE_128B: 1 KiB alignment
    uint8_t* tiles = smem_raw + align_off;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + DSTAGES * STAGE_BYTES_D);

    int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int const wm0 = (warp >> 1) * 32, wn0 = (warp & 1) * 32;
    int const lq = lane >> 2, lr = lane & 3;

    int64_t pid_m, pid_n;
    tile_coords<8>(blockIdx.x, p.tiles_m, p.tiles_n, pid_m, pid_n);
    int const m0 = (int)(pid_m * DT), n0 = (int)(pid_n * DT);
    int const nkb = p.num_k_blocks;

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
        for (int s = 0; s < DSTAGES; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        int const pre = nkb < DSTAGES ? nkb : DSTAGES;
        for (int s = 0; s < pre; ++s) {
            uint8_t* sa = tiles + s * STAGE_BYTES_D;
            mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES_D);
            tma_load_tile_d(&map_a, &full_bar[s], sa, s * DBK, m0);
            tma_load_tile_d(&map_b, &full_bar[s], sa + TILE_BYTES_D, s * DBK, n0);
        }
    }

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // Byte offset of this lane's element inside a tile, for k4 = 0..3: row r = base + i*8 + lq (r % 8 == lq),
    // k = k4*4 + lr  ->  chunk = ((k >> 1) ^ lq), byte = r*128 + chunk*16 + (k & 1)*8.
    int koff[4];
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) koff[k4] = (((k4 * 2 + (lr >> 1)) ^ lq) << 4) + ((lr & 1) << 3);
    int const a_row = (wm0 + lq) * 128, b_row = (wn0 + lq) * 128;

    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait_lean(&full_bar[stage], phase);
        const uint8_t* as = tiles + stage * STAGE_BYTES_D + a_row;
        const uint8_t* bs = tiles + stage * STAGE_BYTES_D + TILE_BYTES_D + b_row;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = *reinterpret_cast<const double*>(as + i * 8 * 128 + koff[k4]);
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = *reinterpret_cast<const double*>(bs + j * 8 * 128 + koff[k4]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                        : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                        : "d"(af[i]), "d"(bf[j]));
        }
        __syncthreads();  // every warp is done reading this stage
        if (tid == 0 && kb + DSTAGES < nkb) {
            uint8_t* sa = tiles + stage * STAGE_BYTES_D;
            mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES_D);
            tma_load_tile_d(&map_a, &full_bar[stage], sa, (kb + DSTAGES) * DBK, m0);
            tma_load_tile_d(&map_b, &full_bar[stage], sa + TILE_BYTES_D, (kb + DSTAGES) * DBK, n0);
        }
        if (++stage == DSTAGES) {
            stage = 0;
            phase ^= 1;
        }
    }

    // Epilogue: C += acc.  C fragment: lane holds C[m = lq][n = 2*lr + {0,1}] of each 8x8 tile.
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t const m = (int64_t)m0 + wm0 + i * 8 + lq;
        if (m >= p.M) continue;
        double* crow = p.C + m * p.ldc;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t const n = (int64_t)n0 + wn0 + j * 8 + 2 * lr;
            if (p.vec_c && n + 2 <= p.N) {
                double2 c = *reinterpret_cast<double2*>(crow + n);
                c.x += acc[i][j][0];
                c.y += acc[i][j][1];
                *reinterpret_cast<double2*>(crow + n) = c;
            } else {
                if (n < p.N) crow[n] += acc[i][j][0];
                if (n + 1 < p.N) crow[n + 1] += acc[i][j][1];
            }
        }
    }
}

// out[r * ldp + k] = in(r, k) = in[r * s_r + k * s_k]  for r < R, k < K (K-contiguous plane).
// R_CONTIG: the warp reads along r (coalesced when s_r == 1) and transposes through shared memory.
template <bool R_CONTIG>
__global__ void __launch_bounds__(256)
pack_k_kernel(const double* __restrict__ in, int64_t s_r, int64_t s_k, int R, int K, double* __restrict__ out,
              int64_t ldp) {
    __shared__ double tile[32][33];
    int const tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    int const r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    if constexpr (R_CONTIG) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int const r = r0 + tx, k = k0 + ty + 8 * i;
            tile[ty + 8 * i][tx] = (r < R && k < K) ? in[(int64_t)r * s_r + (int64_t)k * s_k] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int const r = r0 + ty + 8 * i, k = k0 + tx;
            if (r < R && k < K) out[(int64_t)r * ldp + k] = tile[tx][ty + 8 * i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int const r = r0 + ty + 8 * i, k = k0 + tx;
            if (r < R && k < K) out[(int64_t)r * ldp + k] = in[(int64_t)r * s_r + (int64_t)k * s_k];
        }
    }
}

inline int64_t round_up64(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

bool make_map_kmajor_f64(CUtensorMap* map, const double* base, uint64_t K, uint64_t rows, uint64_t ld) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {K, rows};
    cuuint64_t strides[1] = {ld * sizeof(double)};
    cuuint32_t box[2] = {(cuuint32_t)DBK, (cuuint32_t)DT};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool direct_ok(const double* p, int64_t s_r, int64_t s_k, int64_t K) {
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && s_k == 1 && s_r % 2 == 0 && s_r >= K;
}

const TileConfig kCfg[] = {
    {"dmma_tma_64x64x16_s3", 64, 64, 16, DTHREADS, 4},   // 3 stages, 4 CTAs / SM: 36.2 TFLOP/s at 8192^3 (default)
    {"dmma_tma_64x64x16_s4", 64, 64, 16, DTHREADS, 3},   // 4 stages, 3 CTAs / SM: 35.9
};

template <int DSTAGES, int MINB>
cudaError_t launch_dmma_cfg(const CUtensorMap& ma, const CUtensorMap& mb, const DmmaTmaParams& p, int64_t grid,
                            cudaStream_t stream) {
    constexpr size_t smem = smem_bytes_d(DSTAGES);
    cudaError_t e = cudaFuncSetAttribute(mtm_dmma_tma_kernel<DSTAGES, MINB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mtm_dmma_tma_kernel<DSTAGES, MINB><<<dim3((unsigned)grid), dim3(DTHREADS), smem, stream>>>(ma, mb, p);
    return cudaGetLastError();
}

}  // namespace

int dmma_tma_num_configs() { return (int)(sizeof(kCfg) / sizeof(kCfg[0])); }
const TileConfig& dmma_tma_config(int cfg) { return kCfg[cfg]; }

size_t dmma_tma_workspace_bytes(const MtmShape& s) {
    return 8 * (size_t)round_up64(s.K, 2) * (size_t)(s.M + s.N) + 1024;
}

cudaError_t launch_dmma_tma_f64(int cfg, double* C, const double* A, const double* B, const MtmShape& s, void* ws,
                                size_t ws_bytes, int vec_c, int reuse_b, cudaStream_t stream, int* launches) {
    if (launches) *launches = 0;
    if (cfg < 0 || cfg > 1) return cudaErrorInvalidValue;
    int n_launch = 0;
    double* wsd = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
    int64_t const kp = round_up64(s.K, 2);
    const double* a_src = A;
    int64_t a_ld = s.a_sm;
    const double* b_src = B;      // as B^T: rows n, k-stride b_sk
    int64_t b_ld = s.b_sn;
    dim3 const blk(256);
    // B^T's plane comes first in the workspace so that its position does not depend on M (reuse_b)
    if (!direct_ok(B, s.b_sn, s.b_sk, s.K)) {
        if (ws_bytes < dmma_tma_workspace_bytes(s)) return cudaErrorInvalidValue;
        dim3 const g((unsigned)((s.K + 31) / 32), (unsigned)((s.N + 31) / 32));
        if (!reuse_b) {
            if (s.b_sn == 1 || s.b_sn < s.b_sk)
                pack_k_kernel<true><<<g, blk, 0, stream>>>(B, s.b_sn, s.b_sk, (int)s.N, (int)s.K, wsd, kp);
            else
                pack_k_kernel<false><<<g, blk, 0, stream>>>(B, s.b_sn, s.b_sk, (int)s.N, (int)s.K, wsd, kp);
            ++n_launch;
        }
        b_src = wsd;
        b_ld = kp;
    }
    if (!direct_ok(A, s.a_sm, s.a_sk, s.K)) {
        if (ws_bytes < dmma_tma_workspace_bytes(s)) return cudaErrorInvalidValue;
        double* dst = wsd + (size_t)kp * (size_t)s.N;
        dim3 const g((unsigned)((s.K + 31) / 32), (unsigned)((s.M + 31) / 32));
        if (s.a_sm == 1 || s.a_sm < s.a_sk)
            pack_k_kernel<true><<<g, blk, 0, stream>>>(A, s.a_sm, s.a_sk, (int)s.M, (int)s.K, dst, kp);
        else
            pack_k_kernel<false><<<g, blk, 0, stream>>>(A, s.a_sm, s.a_sk, (int)s.M, (int)s.K, dst, kp);
        ++n_launch;
        a_src = dst;
        a_ld = kp;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;

    CUtensorMap ma, mb;
    if (!make_map_kmajor_f64(&ma, a_src, (uint64_t)s.K, (uint64_t)s.M, (uint64_t)a_ld) ||
        !make_map_kmajor_f64(&mb, b_src, (uint64_t)s.K, (uint64_t)s.N, (uint64_t)b_ld))
        return cudaErrorInvalidValue;
    DmmaTmaParams p;
    p.C = C;
    p.ldc = s.ldc;
    p.M = (int)s.M;
    p.N = (int)s.N;
    p.num_k_blocks = (int)((s.K + DBK - 1) / DBK);
    p.tiles_m = (s.M + DT - 1) / DT;
    p.tiles_n = (s.N + DT - 1) / DT;
    p.vec_c = vec_c;
    int64_t const grid = p.tiles_m * p.tiles_n;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    e = cfg == 0 ? launch_dmma_cfg<3, 4>(ma, mb, p, grid, stream) : launch_dmma_cfg<4, 3>(ma, mb, p, grid, stream);
    if (e != cudaSuccess) return e;
    if (launches) *launches = n_launch + 1;
    return cudaSuccess;
}

}  // namespace b200
