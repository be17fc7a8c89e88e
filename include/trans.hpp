// amt::transpose — out-of-place c = a^T and in-place (square) transpose on an NVIDIA B200.
//
// Drop-in for the reference's include/trans.hpp:94-168: same two overloads, same validation and
// messages, same "returns a nullary callable" contract.  The callable calls the C ABI of
// libb200mtm.so (include/b200_trans.h) instead of the OpenMP transpose_helper (trans.hpp:33-92).
#ifndef B200_AMT_TRANS_HPP
#define B200_AMT_TRANS_HPP

#include <boost/numeric/ublas/tensor.hpp>

#include <cstddef>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>

#include "b200_trans.h"
#include "utils.hpp"

namespace amt {

template <typename Out, typename E>
constexpr auto transpose(boost::numeric::ublas::tensor_core<Out>& c, boost::numeric::ublas::tensor_core<E> const& a,
                         [[maybe_unused]] std::optional<std::size_t> num_threads) {
    namespace ub = boost::numeric::ublas;
    using value_type = typename ub::tensor_core<E>::value_type;
    static_assert(std::is_same_v<typename ub::tensor_core<Out>::value_type, value_type>,
                  "input value type and result value type must be of same value type");
    static_assert(std::is_same_v<value_type, float> || std::is_same_v<value_type, double>,
                  "the B200 transpose path supports float and double");
    auto const& na = a.extents();
    auto const& nc = c.extents();
    if (!(ub::is_matrix(na) && ub::is_matrix(nc))) {
        throw std::runtime_error(
            "amt::transpose(boost::numeric::ublas::tensor_core<Out>& c, boost::numeric::ublas::tensor_core<E> const& a) : "
            "a and c must be the matrices");
    }
    if (!((na[0] == nc[1]) && (na[1] == nc[0]))) {
        throw std::runtime_error(
            "amt::transpose(boost::numeric::ublas::tensor_core<Out>& c, boost::numeric::ublas::tensor_core<E> const& a) : "
            "dimension mismatch");
    }
    value_type* c_ptr = c.data();
    value_type const* a_ptr = a.data();
    std::size_t const* wc = c.strides().data();
    std::size_t const* wa = a.strides().data();
    std::size_t const* na_ptr = na.data();
    std::size_t const* nc_ptr = nc.data();
    return [=] {
        int rc;
        if constexpr (std::is_same_v<value_type, float>) rc = b200_transpose_f32(c_ptr, nc_ptr, wc, a_ptr, na_ptr, wa, 0);
        else rc = b200_transpose_f64(c_ptr, nc_ptr, wc, a_ptr, na_ptr, wa, 0);
        if (rc != B200_OK) throw std::runtime_error(std::string("amt::transpose [B200]: ") + b200_last_error());
    };
}

// In place.  Like the reference (trans.hpp:143-168) this treats the storage as an n x n first_order
// block; it is only meaningful for square matrices (the reference's own test is square), and the
// B200 path reports an error for anything else instead of scrambling the storage.
template <typename E>
constexpr auto transpose(boost::numeric::ublas::tensor_core<E>& a,
                         [[maybe_unused]] std::optional<std::size_t> num_threads) {
    namespace ub = boost::numeric::ublas;
    using value_type = typename ub::tensor_core<E>::value_type;
    auto const& na = a.extents();
    if (!ub::is_matrix(na)) {
        throw std::runtime_error(
            "amt::transpose(boost::numeric::ublas::tensor_core<E> const& a) : "
            "a must be a matrix");
    }
    value_type* a_ptr = a.data();
    std::size_t const* na_ptr = na.data();
    return [=] {
        int rc;
        if constexpr (std::is_same_v<value_type, float>) rc = b200_transpose_inplace_f32(a_ptr, na_ptr, 0);
        else rc = b200_transpose_inplace_f64(a_ptr, na_ptr, 0);
        if (rc != B200_OK) throw std::runtime_error(std::string("amt::transpose [B200]: ") + b200_last_error());
    };
}

}  // namespace amt

#endif  // B200_AMT_TRANS_HPP
