// amt::mtm — matrix-times-matrix, C += A * B, executed on an NVIDIA B200.
//
// Drop-in for the reference's include/mtm.hpp:208-267: same template signature, same validation
// (same exception type and messages), same "returns a nullary callable that accumulates into c"
// contract, same layout/stride conventions.  The body is different: instead of the OpenMP
// 5-loop blocked algorithm (mtm.hpp:116-206) the callable hands the raw pointers, extents and
// strides to the C ABI of libb200mtm.so (include/b200_mtm.h), whose sm_100a kernels do the work.
// There is no CPU path.
//
//   auto fn = amt::mtm(c, a, b, std::nullopt);   // validates, throws std::runtime_error
//   fn();                                        // c += a * b   (host tensors: synchronous)
//
// Link with -lb200mtm (openmp-blas_b200/libb200mtm.so).
#ifndef B200_AMT_MTM_HPP
#define B200_AMT_MTM_HPP

#include <boost/numeric/ublas/tensor.hpp>

#include <cstddef>
#include <cstdlib>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>

#include "b200_mtm.h"
#include "device_matrix.hpp"
#include "utils.hpp"

namespace amt {

namespace b200 {

// Kernel family / tile config used by callables created afterwards (process-wide default:
// B200_MTM_AUTO).  The reference has no such knob; it exists so the harness can time the
// FFMA, 3xTF32, DFMA and DMMA variants through the unchanged amt::mtm signature.
inline int& default_flags() noexcept {
    static int flags = B200_MTM_AUTO;
    return flags;
}
inline void set_variant(int variant, int config = -1) noexcept {
    default_flags() = B200_MTM_FLAGS(variant, config + 1);
}

// How many GPUs a host-tensor call may use.  The reference's num_threads argument only ever raises the team to
// the maximum (thread_utils.hpp:47-56); the analogue here is "all visible GPUs" for problems large enough to
// shard (>= 2 * 4096^3 flop per call), overridable with set_devices(n) or the environment variable
// B200_MTM_DEVICES (1 = single GPU).
inline int& max_devices() noexcept {
    static int n = [] {
        char const* v = std::getenv("B200_MTM_DEVICES");
        return v ? std::atoi(v) : 0;       // 0 = all visible
    }();
    return n;
}
inline void set_devices(int n) noexcept { max_devices() = n; }
inline int visible_devices() noexcept {
    static int n = [] {
        int c = 0;
        return b200_device_count(&c) == B200_OK ? c : 0;
    }();
    return n;
}

template <typename T>
inline int call_host(T* c, std::size_t const* nc, std::size_t const* wc, T const* a,
                     std::size_t const* na, std::size_t const* wa, T const* b,
                     std::size_t const* nb, std::size_t const* wb, int flags) {
    int const want = max_devices() > 0 ? max_devices() : visible_devices();
    double const flop = 2.0 * (double)nc[0] * (double)nc[1] * (double)na[1];
    if (want > 1 && visible_devices() > 1 && flop >= 2.0 * 4096.0 * 4096.0 * 4096.0) {
        if constexpr (std::is_same_v<T, float>)
            return b200_mtm_f32_mgpu(c, nc, wc, a, na, wa, b, nb, wb, flags, nullptr, want);
        else
            return b200_mtm_f64_mgpu(c, nc, wc, a, na, wa, b, nb, wb, flags, nullptr, want);
    }
    if constexpr (std::is_same_v<T, float>)
        return b200_mtm_f32(c, nc, wc, a, na, wa, b, nb, wb, flags);
    else
        return b200_mtm_f64(c, nc, wc, a, na, wa, b, nb, wb, flags);
}

}  // namespace b200

template <typename Out, typename E1, typename E2>
constexpr auto mtm(boost::numeric::ublas::tensor_core<Out>& c,
                   boost::numeric::ublas::tensor_core<E1> const& a,
                   boost::numeric::ublas::tensor_core<E2> const& b,
                   [[maybe_unused]] std::optional<std::size_t> num_threads) {
    namespace ub = boost::numeric::ublas;
    using value_type = typename ub::tensor_core<Out>::value_type;
    static_assert(std::is_same_v<typename ub::tensor_core<E1>::value_type,
                                 typename ub::tensor_core<E2>::value_type> &&
                      std::is_same_v<value_type, typename ub::tensor_core<E2>::value_type>,
                  "both tensor type and result type must be of same value_type");
    static_assert(std::is_same_v<value_type, float> || std::is_same_v<value_type, double>,
                  "the B200 mtm path supports float and double");

    auto const& na = a.extents();
    auto const& nb = b.extents();
    auto const& nc = c.extents();

    // Same two checks, order and messages as the reference front-end (mtm.hpp:234-250; the
    // text says "amt::mtv" there too).
    if (!(ub::is_matrix(na) && ub::is_matrix(nb) && ub::is_matrix(nc))) {
        throw std::runtime_error(
            "amt::mtv(boost::numeric::ublas::tensor_core<Out>& c, boost::numeric::ublas::tensor_core<E1> const& a, "
            "boost::numeric::ublas::tensor_core<E2> const& b) : "
            "a, b, and c must be the matrices");
    }
    // num_threads: the reference only ever raises the OpenMP team to the maximum
    // (thread_utils.hpp:47-56); on the GPU there is nothing to clip, so it is ignored.
    if (!((na[0] == nc[0]) && (na[1] == nb[0]) && (nc[1] == nb[1]))) {
        throw std::runtime_error(
            "amt::mtv(boost::numeric::ublas::tensor_core<Out>&, boost::numeric::ublas::tensor_core<E1> const&, "
            "boost::numeric::ublas::tensor_core<E2> const&) : "
            "dimension mismatch");
    }

    // The callable borrows the tensors' storage and their extents/strides arrays, exactly like
    // the reference's lambda (mtm.hpp:252-266): the tensors must outlive it.
    value_type* c_ptr = c.data();
    value_type const* a_ptr = a.data();
    value_type const* b_ptr = b.data();
    std::size_t const* wc_ptr = c.strides().data();
    std::size_t const* wa_ptr = a.strides().data();
    std::size_t const* wb_ptr = b.strides().data();
    std::size_t const* nc_ptr = nc.data();
    std::size_t const* na_ptr = na.data();
    std::size_t const* nb_ptr = nb.data();
    int const flags = b200::default_flags();

    return [=] {
        int const rc = b200::call_host<value_type>(c_ptr, nc_ptr, wc_ptr, a_ptr, na_ptr, wa_ptr, b_ptr,
                                                   nb_ptr, wb_ptr, flags);
        if (rc != B200_OK)
            throw std::runtime_error(std::string("amt::mtm [B200]: ") + b200_last_error());
    };
}

// Device-resident operands: same contract (validate now, accumulate on every call), but the
// callable only enqueues the kernel on `stream` (nullptr = default stream) and returns.
template <typename T, typename LC, typename LA, typename LB>
auto mtm(device_matrix<T, LC>& c, device_matrix<T, LA> const& a, device_matrix<T, LB> const& b,
         void* stream = nullptr) {
    std::size_t const* na = a.extents();
    std::size_t const* nb = b.extents();
    std::size_t const* nc = c.extents();
    if (!((na[0] == nc[0]) && (na[1] == nb[0]) && (nc[1] == nb[1]))) {
        throw std::runtime_error(
            "amt::mtv(boost::numeric::ublas::tensor_core<Out>&, boost::numeric::ublas::tensor_core<E1> const&, "
            "boost::numeric::ublas::tensor_core<E2> const&) : "
            "dimension mismatch");
    }
    T* c_ptr = c.data();
    T const* a_ptr = a.data();
    T const* b_ptr = b.data();
    std::size_t const* wa = a.strides();
    std::size_t const* wb = b.strides();
    std::size_t const* wc = c.strides();
    int const flags = b200::default_flags();
    return [=] {
        int rc;
        if constexpr (std::is_same_v<T, float>)
            rc = b200_mtm_f32_dev(c_ptr, nc, wc, a_ptr, na, wa, b_ptr, nb, wb, flags, stream);
        else
            rc = b200_mtm_f64_dev(c_ptr, nc, wc, a_ptr, na, wa, b_ptr, nb, wb, flags, stream);
        if (rc != B200_OK) throw std::runtime_error(std::string("amt::mtm [B200]: ") + b200_last_error());
    };
}

}  // namespace amt

#endif  // B200_AMT_MTM_HPP
