// amt::benchmark — mean time of MaxIter back-to-back invocations.
//
// benchmark<MaxIter>(fn, args...) keeps the reference's host protocol (include/benchmark.hpp:34-52:
// steady_clock, no warm-up, mean in nanoseconds) so `amt::benchmark<4>(amt::mtm(res, A, B, nullopt))`
// from src/mtm.cpp:207-208 compiles and means the same thing; through the host-tensor front-end it
// therefore includes the host<->device copies.  device_benchmark() is the device-resident
// counterpart: warm-up calls, then CUDA events around the timed launches on one stream.
#ifndef B200_AMT_BENCHMARK_HPP
#define B200_AMT_BENCHMARK_HPP

#include <cstddef>
#include <functional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>

#include "b200_mtm.h"
#include "device_matrix.hpp"
#include "timer.hpp"

namespace amt {

template <typename T>
inline void no_opt(T const& val) noexcept {
    asm volatile("" : : "r,m"(val) : "memory");
}
template <typename T>
inline void no_opt(T& val) noexcept {
    asm volatile("" : "+m,r"(val) : : "memory");
}
inline void clobber_mem() noexcept { asm volatile("" : : : "memory"); }

template <std::size_t MaxIter = 100u, typename Fn, typename... Args>
double benchmark(Fn&& fn, Args&&... args) {
    double total = 0.0;
    for (std::size_t i = 0; i < MaxIter; ++i) {
        timer t;
        if constexpr (std::is_void_v<std::invoke_result_t<Fn, Args...>>) {
            std::invoke(fn, args...);
        } else {
            auto r = std::invoke(fn, args...);
            no_opt(r);
        }
        total += t.stop();
    }
    return total / static_cast<double>(MaxIter);
}

// Mean nanoseconds per C += A*B with device-resident operands (CUDA events on `stream`).
template <std::size_t MaxIter = 10u, std::size_t WarmUp = 3u, typename T, typename LC, typename LA, typename LB>
double device_benchmark(device_matrix<T, LC>& c, device_matrix<T, LA> const& a, device_matrix<T, LB> const& b,
                        int flags = B200_MTM_AUTO, void* stream = nullptr) {
    double ms = 0.0;
    int rc;
    if constexpr (std::is_same_v<T, float>)
        rc = b200_mtm_bench_f32_dev(c.data(), c.extents(), c.strides(), a.data(), a.extents(), a.strides(), b.data(),
                                    b.extents(), b.strides(), flags, stream, (int)WarmUp, (int)MaxIter, &ms);
    else
        rc = b200_mtm_bench_f64_dev(c.data(), c.extents(), c.strides(), a.data(), a.extents(), a.strides(), b.data(),
                                    b.extents(), b.strides(), flags, stream, (int)WarmUp, (int)MaxIter, &ms);
    if (rc != B200_OK) throw std::runtime_error(std::string("amt::device_benchmark [B200]: ") + b200_last_error());
    return ms * 1e6;
}

}  // namespace amt

#endif  // B200_AMT_BENCHMARK_HPP
