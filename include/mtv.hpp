// amt::mtv / amt::vtm — matrix-times-vector and vector-times-matrix on an NVIDIA B200.
//
// Drop-in for the reference's include/mtv.hpp:102-236: same template signatures, same validation
// (exception type and messages), same "returns a nullary callable" contract.  The callable hands
// the raw pointers to the C ABI of libb200mtm.so (include/b200_mtv.h) instead of the OpenMP
// routine mtv_helper (mtv.hpp:15-100).
//
// Semantics are the reference's, including its per-layout asymmetry:
//   mtv, A first_order :  c += A v      (accumulates, simd_loop.hpp:58-75)
//   mtv, A last_order  :  c  = A v      (assigns, mtv.hpp:94-99)
//   vtm(c, A, v) = mtv on A^T with the other layout's path (mtv.hpp:206-236):
//        A first_order :  c  = v A ;    A last_order : c += v A
// One deliberate deviation: for a last_order A the reference's vtm passes the UN-flipped strides to
// the first_order helper (mtv.hpp:224-226 leaves wa = {na[0], 1}) and so computes
// sum_k a[i + k] * v[k]; its own test (test/test.vtm.cpp:77-124, BLIS_TRANSPOSE) expects v*A.  This
// header computes v*A.  Likewise the reference's vtm strides the transposed view by new_na[0]
// (mtv.hpp:222-226), which is only right for square matrices (all its tests are square); here any
// M x N works.  (tests/test_oracle_mtv.py pins these facts.)
#ifndef B200_AMT_MTV_HPP
#define B200_AMT_MTV_HPP

#include <boost/numeric/ublas/tensor.hpp>

#include <array>
#include <cstddef>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>

#include "b200_mtv.h"
#include "utils.hpp"

namespace amt {

namespace b200 {
template <typename T>
inline int call_mtv_host(T* c, T const* a, std::size_t const* na, std::size_t const* wa, T const* b, int a_last_order) {
    if constexpr (std::is_same_v<T, float>) return b200_mtv_f32(c, a, na, wa, b, a_last_order, B200_MTM_AUTO);
    else return b200_mtv_f64(c, a, na, wa, b, a_last_order, B200_MTM_AUTO);
}
}  // namespace b200

namespace detail {
template <bool IsVtm, typename Out, typename E1, typename E2>
auto mtv_impl(boost::numeric::ublas::tensor_core<Out>& c, boost::numeric::ublas::tensor_core<E1> const& a,
              boost::numeric::ublas::tensor_core<E2> const& b) {
    namespace ub = boost::numeric::ublas;
    using value_type = typename ub::tensor_core<Out>::value_type;
    using layout_type = typename ub::tensor_core<E1>::layout_type;
    static_assert(std::is_same_v<typename ub::tensor_core<E1>::value_type, typename ub::tensor_core<E2>::value_type> &&
                      std::is_same_v<value_type, typename ub::tensor_core<E2>::value_type>,
                  "both tensor type and result type must be of same value_type");
    static_assert(std::is_same_v<value_type, float> || std::is_same_v<value_type, double>,
                  "the B200 mtv path supports float and double");
    auto const& na = a.extents();
    auto const& nb = b.extents();
    auto const& nc = c.extents();
    if (!(ub::is_matrix(na) && ub::is_vector(nb) && ub::is_vector(nc))) {
        throw std::runtime_error(
            "amt::mtv(boost::numeric::ublas::tensor_core<Out>& c, boost::numeric::ublas::tensor_core<E1> const& a, "
            "boost::numeric::ublas::tensor_core<E2> const& b) : "
            "c and b must be vector, and a must be a matrix");
    }
    std::size_t const NB = ub::product(nb), NC = ub::product(nc);
    bool const mismatch = IsVtm ? ((na[1] != NC) || (na[0] != NB)) : ((na[1] != NB) || (na[0] != NC));
    if (mismatch) {
        throw std::runtime_error(
            "amt::mtv(boost::numeric::ublas::tensor_core<Out>&, boost::numeric::ublas::tensor_core<E1> const&, "
            "boost::numeric::ublas::tensor_core<E2> const&) : "
            "dimension mismatch");
    }
    constexpr bool first = std::is_same_v<layout_type, ub::layout::first_order>;
    // Strides follow from the layout, as in the reference (mtv.hpp:157-159), not from the tensor.
    std::array<std::size_t, 2> ext = {na[0], na[1]};
    std::array<std::size_t, 2> wa = first ? std::array<std::size_t, 2>{1, na[0]} : std::array<std::size_t, 2>{na[1], 1};
    int a_last_order = first ? 0 : 1;
    if constexpr (IsVtm) {   // c = v A  ==  A^T v : flip extents and strides, take the other layout's path
        ext = {na[1], na[0]};
        wa = {wa[1], wa[0]};
        a_last_order = first ? 1 : 0;
    }
    value_type* c_ptr = c.data();
    value_type const* a_ptr = a.data();
    value_type const* b_ptr = b.data();
    return [=] {
        int const rc = b200::call_mtv_host<value_type>(c_ptr, a_ptr, ext.data(), wa.data(), b_ptr, a_last_order);
        if (rc != B200_OK) throw std::runtime_error(std::string("amt::mtv [B200]: ") + b200_last_error());
    };
}
}  // namespace detail

template <typename Out, typename E1, typename E2>
constexpr auto mtv(boost::numeric::ublas::tensor_core<Out>& c, boost::numeric::ublas::tensor_core<E1> const& a,
                   boost::numeric::ublas::tensor_core<E2> const& b,
                   [[maybe_unused]] std::optional<std::size_t> num_threads) {
    return detail::mtv_impl<false>(c, a, b);
}

template <typename Out, typename E1, typename E2>
constexpr auto vtm(boost::numeric::ublas::tensor_core<Out>& c, boost::numeric::ublas::tensor_core<E1> const& a,
                   boost::numeric::ublas::tensor_core<E2> const& b,
                   [[maybe_unused]] std::optional<std::size_t> num_threads) {
    return detail::mtv_impl<true>(c, a, b);
}

}  // namespace amt

#endif  // B200_AMT_MTV_HPP
