// Replicating an operand across the GPUs of one NVSwitch box — the exchange step of the row-block
// sharded mtm ("B is broadcast"), done with this library's own kernels over NVLink
// instead of a collective library's copy kernels:
//
//   * the root streams its copy of B into a MULTICAST address (multimem.st): the NVSwitch
//     replicates every 16-byte store into the symmetric buffer of every GPU, so the root sends B
//     once whatever the number of GPUs, and the receiving GPUs run no communication kernel at all
//     (their SMs stay with the product);  without multicast support the same kernel stores to each
//     peer's mapped buffer in turn;
//   * arrival is published per K-chunk by a flag word written after the data (system-scope fence,
//     last-CTA election), and consumed by a one-warp wait kernel in front of the chunk's product;
//   * a receiver tells the root it has finished reading the buffer with a one-thread signal kernel.
//
// The buffers and flag words live in symmetric (peer-mapped) allocations that the Python driver
// obtains from torch.distributed's symmetric memory; this file only sees raw addresses.
#include <cstdint>

#include <atomic>

#include "mtm_kernels.h"

namespace b200 {
namespace {

constexpr int kPushThreads = 512;
constexpr int kMaxDst = 8;

struct PushDst {
    void* dst[kMaxDst];
    uint32_t* flag[kMaxDst];
    int n_dst, n_flag;
};

// CTAs that have finished their stores, one counter PER LAUNCH: the host hands every push the next slot of
// this ring, so pushes that overlap on different streams (two replicators, a probe next to a step) never
// share a counter; the last CTA of a launch leaves its slot at zero for the launch that reuses it 64 later.
constexpr int kPushSlots = 64;
__device__ unsigned int g_push_done[kPushSlots];

__device__ __forceinline__ void multimem_st_v4(void* mc, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void multimem_st_u32(uint32_t* mc, uint32_t v) {
    asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(mc), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_stream_v4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

// 16-byte units, grid-stride, four independent loads in flight per thread before the stores.
// TWO_D: the region is `rows` runs of `upr` units with separate source / destination pitches (a column
// panel of a row-major matrix); unit i sits in run i / upr.
struct Pitch {
    size_t upr, src_pitch16, dst_pitch16;
};

template <bool MC, bool TWO_D>
__global__ void __launch_bounds__(kPushThreads) replicate_push_kernel(PushDst d, const float4* __restrict__ src, size_t n16,
                                                                      Pitch pt, int flag_mc, uint32_t flag_value, int slot) {
    size_t const stride = (size_t)gridDim.x * kPushThreads;
    size_t i = (size_t)blockIdx.x * kPushThreads + threadIdx.x;
    auto src_of = [&](size_t u) -> size_t {
        if (!TWO_D) return u;
        size_t const r = u / pt.upr;
        return r * pt.src_pitch16 + (u - r * pt.upr);
    };
    auto dst_of = [&](size_t u) -> size_t {
        if (!TWO_D) return u;
        size_t const r = u / pt.upr;
        return r * pt.dst_pitch16 + (u - r * pt.upr);
    };
    auto store = [&](size_t o, const float4& v) {
        if (MC) {
            multimem_st_v4(reinterpret_cast<float4*>(d.dst[0]) + o, v);
        } else {
#pragma unroll 1
            for (int p = 0; p < d.n_dst; ++p) reinterpret_cast<float4*>(d.dst[p])[o] = v;
        }
    };
    for (; i + 3 * stride < n16; i += 4 * stride) {
        float4 const v0 = ld_stream_v4(src + src_of(i)), v1 = ld_stream_v4(src + src_of(i + stride)),
                     v2 = ld_stream_v4(src + src_of(i + 2 * stride)), v3 = ld_stream_v4(src + src_of(i + 3 * stride));
        store(dst_of(i), v0);
        store(dst_of(i + stride), v1);
        store(dst_of(i + 2 * stride), v2);
        store(dst_of(i + 3 * stride), v3);
    }
    for (; i < n16; i += stride) store(dst_of(i), ld_stream_v4(src + src_of(i)));
    // Publish: every CTA fences its stores system-wide, the last one to finish writes the flag(s).
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int const prev = atomicAdd(&g_push_done[slot], 1u);
        if (prev + 1 == gridDim.x) {
            atomicExch(&g_push_done[slot], 0u);
            __threadfence_system();
            for (int p = 0; p < d.n_flag; ++p) {
                if (flag_mc)
                    multimem_st_u32(d.flag[p], flag_value);
                else
                    *reinterpret_cast<volatile uint32_t*>(d.flag[p]) = flag_value;
            }
            __threadfence_system();
        }
    }
}

// One warp; lane i watches flag[i * stride] (i < count, i != skip) until it has reached `value`
// (wrap-safe signed comparison).  Traps after ~30 s so that a lost peer fails the run loudly
// instead of hanging the GPU.
__global__ void flag_wait_kernel(const uint32_t* flag, uint32_t value, int count, int stride, int skip) {
    int const lane = threadIdx.x;
    if (lane < count && lane != skip) {
        const volatile uint32_t* p = flag + (size_t)lane * stride;
        long long const t0 = clock64();
        while ((int32_t)(*p - value) < 0) {
            __nanosleep(200);
            if (clock64() - t0 > 60000000000LL) __trap();
        }
    }
    __syncwarp();
    __threadfence_system();
}

__global__ void flag_signal_kernel(uint32_t* flag, uint32_t value) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(flag) = value;
    __threadfence_system();
}

}  // namespace

cudaError_t launch_replicate_push_2d(void* const* dst, int n_dst, int multicast, const void* src, size_t rows,
                                     size_t row_bytes, size_t src_pitch, size_t dst_pitch, uint32_t* const* flag_dst,
                                     int n_flag_dst, int flag_multicast, uint32_t flag_value, int ctas, cudaStream_t stream) {
    if (n_dst < 1 || n_dst > kMaxDst || n_flag_dst < 0 || n_flag_dst > kMaxDst || (multicast && n_dst != 1) ||
        (flag_multicast && n_flag_dst != 1))
        return cudaErrorInvalidValue;
    if ((row_bytes & 15u) || (reinterpret_cast<uintptr_t>(src) & 15u)) return cudaErrorMisalignedAddress;
    bool const two_d = rows > 1 && (src_pitch != row_bytes || dst_pitch != row_bytes);
    if (two_d && ((src_pitch & 15u) || (dst_pitch & 15u) || src_pitch < row_bytes || dst_pitch < row_bytes))
        return cudaErrorMisalignedAddress;
    PushDst d{};
    d.n_dst = n_dst;
    d.n_flag = n_flag_dst;
    for (int i = 0; i < n_dst; ++i) {
        if (reinterpret_cast<uintptr_t>(dst[i]) & 15u) return cudaErrorMisalignedAddress;
        d.dst[i] = dst[i];
    }
    for (int i = 0; i < n_flag_dst; ++i) d.flag[i] = flag_dst[i];
    size_t n16 = rows * (row_bytes / 16);
    Pitch const pt{row_bytes / 16, src_pitch / 16, dst_pitch / 16};
    if (ctas < 0) {
        // Copy-engine form: the DMA engines move the data (no SM touches it), then a one-CTA launch of
        // the same kernel with nothing left to copy publishes the flag(s) behind them in stream order.
        for (int i = 0; i < n_dst && n16 > 0; ++i) {
            cudaError_t const e = two_d ? cudaMemcpy2DAsync(dst[i], dst_pitch, src, src_pitch, row_bytes, rows,
                                                            cudaMemcpyDeviceToDevice, stream)
                                        : cudaMemcpyAsync(dst[i], src, rows * row_bytes, cudaMemcpyDeviceToDevice, stream);
            if (e != cudaSuccess) return e;
        }
        if (n_flag_dst == 0) return cudaSuccess;
        n16 = 0;
        ctas = 1;
    }
    if (ctas < 1) ctas = 32;
    size_t const need = (n16 + kPushThreads - 1) / kPushThreads;
    if ((size_t)ctas > need) ctas = (int)(need ? need : 1);
    const float4* s4 = static_cast<const float4*>(src);
    static std::atomic<unsigned> next_slot{0};
    int const slot = (int)(next_slot.fetch_add(1, std::memory_order_relaxed) % kPushSlots);
    if (multicast) {
        if (two_d) replicate_push_kernel<true, true><<<ctas, kPushThreads, 0, stream>>>(d, s4, n16, pt, flag_multicast, flag_value, slot);
        else replicate_push_kernel<true, false><<<ctas, kPushThreads, 0, stream>>>(d, s4, n16, pt, flag_multicast, flag_value, slot);
    } else {
        if (two_d) replicate_push_kernel<false, true><<<ctas, kPushThreads, 0, stream>>>(d, s4, n16, pt, flag_multicast, flag_value, slot);
        else replicate_push_kernel<false, false><<<ctas, kPushThreads, 0, stream>>>(d, s4, n16, pt, flag_multicast, flag_value, slot);
    }
    return cudaGetLastError();
}

cudaError_t launch_replicate_push(void* const* dst, int n_dst, int multicast, const void* src, size_t bytes,
                                  uint32_t* const* flag_dst, int n_flag_dst, int flag_multicast, uint32_t flag_value,
                                  int ctas, cudaStream_t stream) {
    return launch_replicate_push_2d(dst, n_dst, multicast, src, 1, bytes, bytes, bytes, flag_dst, n_flag_dst, flag_multicast,
                                    flag_value, ctas, stream);
}

cudaError_t replicate_preload_kernels() {
    cudaFuncAttributes fa;
    cudaError_t e;
    if ((e = cudaFuncGetAttributes(&fa, replicate_push_kernel<true, true>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, replicate_push_kernel<true, false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, replicate_push_kernel<false, true>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, replicate_push_kernel<false, false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, flag_wait_kernel)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, flag_signal_kernel)) != cudaSuccess) return e;
    return cudaSuccess;
}

cudaError_t launch_flag_wait(const uint32_t* flag, uint32_t value, int count, int stride, int skip, cudaStream_t stream) {
    if (count < 1 || count > 32) return cudaErrorInvalidValue;
    flag_wait_kernel<<<1, 32, 0, stream>>>(flag, value, count, stride, skip);
    return cudaGetLastError();
}

cudaError_t launch_flag_signal(uint32_t* flag, uint32_t value, cudaStream_t stream) {
    flag_signal_kernel<<<1, 1, 0, stream>>>(flag, value);
    return cudaGetLastError();
}

}  // namespace b200
