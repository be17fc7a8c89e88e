// Inline-PTX wrappers for the sm_100a features the mtm kernels use: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05.mma / TMEM, clusters.  No CUTLASS: these are the raw instructions.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

namespace b200 {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Spin on the barrier's phase parity.  A watchdog turns a protocol bug into a trapped kernel
// (an error the host sees) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t const addr = smem_u32(bar);
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (spins == 1024) t0 = clock64();
        if (spins > 1024 && (spins & 1023) == 0 && clock64() - t0 > 4000000000LL) {
            printf("mtm_tf32x3: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}
// Wait with acquire semantics at cluster scope: the data guarded by the barrier was written by
// another CTA of the cluster (st.shared::cluster + mbarrier.arrive.release.cluster).
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t const addr = smem_u32(bar);
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (spins == 1024) t0 = clock64();
        if (spins > 1024 && (spins & 1023) == 0 && clock64() - t0 > 4000000000LL) {
            printf("mtm: cluster mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}
// Cluster-scope wait for a thread with slack (the tile scheduler runs several tiles ahead): sleeps
// between probes so that it does not burn issue slots / power next to the MMA pipeline.
__device__ __forceinline__ void mbar_wait_cluster_relaxed(uint64_t* bar, uint32_t parity) {
    uint32_t const addr = smem_u32(bar);
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        __nanosleep(spins < 8 ? 200 : 4000);
        if (spins == 64) t0 = clock64();
        if (spins > 64 && (spins & 255) == 0 && clock64() - t0 > 8000000000LL) __trap();
    }
}
// Store a 32-bit value into the shared memory of CTA `cta` of this cluster, at the same offset.
__device__ __forceinline__ void st_shared_cluster_u32(const void* local_addr, uint32_t cta, uint32_t value) {
    asm volatile(
        "{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\nst.shared::cluster.u32 [ra], %2;\n}"
        ::"r"(smem_u32(local_addr)), "r"(cta), "r"(value) : "memory");
}
// Same, without the diagnostic printf (keeps hot kernels free of a stack frame).
__device__ __forceinline__ void mbar_wait_lean(uint64_t* bar, uint32_t parity) {
    uint32_t const addr = smem_u32(bar);
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (spins == 4096) t0 = clock64();
        if (spins > 4096 && (spins & 4095) == 0 && clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Arrive on the barrier at the same shared-memory offset in CTA `cta` of this cluster.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\nmbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n}"
        ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
// Same without the cluster-scope release (default semantics): for arrivals whose payload was already made visible
// by other means (e.g. fence.proxy.async for tiles the tensor core reads) — no MEMBAR in the SASS.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\nmbarrier.arrive.shared::cluster.b64 _, [ra];\n}"
        ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
template <int NCTA>
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    if constexpr (NCTA == 1) {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
            " [%0], [%1, {%3, %4}], [%2], %5;"
            ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
              "l"(kEvictNormal)
            : "memory");
    } else {
        // Both CTAs of the pair load into their own shared memory; the bytes are accounted on the
        // leader CTA's barrier (peer bit cleared in the shared::cluster address).
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
            " [%0], [%1, {%3, %4}], [%2], %5;"
            ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0),
              "r"(c1), "l"(kEvictNormal)
            : "memory");
    }
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int NCTA>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    if constexpr (NCTA == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int NCTA>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    if constexpr (NCTA == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <int NCTA>
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (NCTA == 1)
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// The same instruction for a WARP-UNIFORM issue loop: every lane runs the loop (so the compiler can keep descriptors,
// addresses and counters on the uniform datapath instead of moving them over from vector registers before every MMA),
// the lane with `issue` != 0 issues.  Descriptors come as two 32-bit words: only the low word (start address, leading
// byte offset) changes from MMA to MMA.
template <int NCTA>
__device__ __forceinline__ void umma_tf32_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate, uint32_t issue) {
    if constexpr (NCTA == 1)
        asm volatile(
            "{\n.reg .pred p, q;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nsetp.ne.b32 q, %7, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
            "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n}"
            ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(issue) : "memory");
    else
        asm volatile(
            "{\n.reg .pred p, q;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nsetp.ne.b32 q, %7, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
            "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n}"
            ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(issue) : "memory");
}
template <int NCTA>
__device__ __forceinline__ void umma_commit_w(uint64_t* bar, uint32_t issue) {
    if constexpr (NCTA == 1)
        asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %1, 0;\n@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}"
                     ::"r"(smem_u32(bar)), "r"(issue) : "memory");
    else
        asm volatile(
            "{\n.reg .pred q;\nsetp.ne.b32 q, %1, 0;\n@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %2;\n}"
            ::"r"(smem_u32(bar)), "r"(issue), "h"((uint16_t)3) : "memory");
}
// The two words of a shared-memory matrix descriptor (see make_kmajor_sw128_desc / make_mnmajor_sw128_32b_desc below): the
// high word is constant per operand; the low word is (address >> 4) + the leading-byte-offset field.
__device__ __forceinline__ uint32_t desc_hi_kmajor_sw128() { return (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ uint32_t desc_hi_mnmajor(uint32_t sbo_bytes, uint32_t layout_type) {
    return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | ((layout_type & 7u) << 29);
}
__device__ __forceinline__ uint32_t desc_lo_lbo_mnmajor(uint32_t lbo_bytes) { return ((lbo_bytes >> 4) & 0x3FFFu) << 16; }
// tcgen05.commit: the barrier is arrived on once every MMA issued so far by this thread has
// completed (implies tcgen05.fence::before_thread_sync).  NCTA == 2: arrive in both CTAs.
template <int NCTA>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    if constexpr (NCTA == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
            ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major, 128B-swizzled tile (rows of 128 bytes, 8-row
// swizzle atoms 1024 bytes apart): start address, SBO = 1024 B, descriptor version 1 (sm_100),
// layout type 2 = SWIZZLE_128B.  The leading-dimension offset is unused for this layout.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Shared-memory matrix descriptor of an MN-major tile of 32-bit elements.  For tf32 the only MN-major layout
// the tensor core accepts is SWIZZLE_128B_BASE32B (layout type 1): rows of 128 bytes (32 fp32 along m/n), one
// row per k, 32-byte chunks XOR-swizzled with the row index inside atoms of 4 rows (byte-address bits [5,7) ^=
// bits [7,9)) — what a TMA box {32 (m/n), k} with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B produces.  Canonical form
// in 16-byte units: ((8,n),(4,k)):((1,LBO),(8,SBO)): LBO = byte distance between atoms along m/n, SBO = between
// the 4-row atoms along k.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_32b_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                               uint32_t layout_type = 1) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout_type & 7u) << 61;
    return d;
}
// TMA reduce-add of a shared-memory box into global memory (SASS UTMAREDG): the L2 performs
// C += box element-wise in the tensor map's data type; no C data enters the SM.  Bulk-group completion.
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// At most N of this thread's bulk groups may still be READING their shared-memory source.
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// Generic-proxy shared-memory writes -> visible to the async proxy (TMA / tcgen05 reads).
__device__ __forceinline__ void fence_proxy_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Programmatic dependent launch.  A kernel launched with the programmatic-stream-serialization attribute may become
// resident while the kernel before it on the stream is still running; griddep_wait() blocks until that kernel has
// completed and its memory operations are visible (a no-op for an ordinary launch).  griddep_launch_dependents():
// once every CTA of this grid has executed it (or exited), the next kernel's CTAs may start to be scheduled.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Instruction descriptor, kind::tf32: D = fp32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2),
// operand majors at bits 15 (A) and 16 (B): 0 = K-major, 1 = MN-major, N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t m, uint32_t n, uint32_t a_mn_major = 0, uint32_t b_mn_major = 0) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}


// Host: cuTensorMapEncodeTiled through the runtime's driver-entry-point lookup (no libcuda link).
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            sym = nullptr;
        return reinterpret_cast<EncodeTiledFn>(sym);
    }();
    return fn;
}

// cuStreamWaitValue32: the stream itself waits until (int32)(*addr - value) >= 0 — no SM, no CTA slot.
using StreamWaitValue32Fn = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
inline StreamWaitValue32Fn get_stream_wait_value32_fn() {
    static StreamWaitValue32Fn fn = [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &sym, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            sym = nullptr;
        return reinterpret_cast<StreamWaitValue32Fn>(sym);
    }();
    return fn;
}

// 2-D fp32 tensor map: element (i0, i1) at base[i1 * stride1 + i0]; box {box0, box1}; OOB reads give 0.
inline bool make_map_2d_f32(CUtensorMap* map, const float* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_elems,
                            uint32_t box0, uint32_t box1, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {stride1_elems * sizeof(float)};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace ptx
}  // namespace b200
