// amt utility surface used by callers of amt::mtm: tensor aliases, make_tensor, layout traits
// and tags.  Counterpart of the reference's include/utils.hpp:13-31, 75-97 — only the parts of
// that header that belong to the mtm API surface; its CPU packing routines (utils.hpp:99-141)
// have no GPU counterpart (strides feed the kernel's tile loaders directly).
#ifndef B200_AMT_UTILS_HPP
#define B200_AMT_UTILS_HPP

#include <boost/numeric/ublas/tensor.hpp>

#include <algorithm>
#include <cstddef>
#include <type_traits>

namespace amt {

namespace ub = boost::numeric::ublas;
using shape_t = ub::extents<2u>;

template <typename T, typename L>
using tensor_t = ub::tensor_static_rank<T, 2, L>;

// Zero-initialised M x N matrix; first_order (column-major) unless L says otherwise.
template <typename T, typename L = ub::layout::first_order>
auto make_tensor(ub::integral auto M, ub::integral auto N) {
    return tensor_t<T, L>(static_cast<std::size_t>(M), static_cast<std::size_t>(N));
}

template <typename T, typename L = ub::layout::first_order>
auto make_tensor(ub::integral auto M, ub::integral auto N, T val) {
    auto t = make_tensor<T, L>(M, N);
    std::fill(t.begin(), t.end(), val);
    return t;
}

template <typename L>
struct is_first_order : std::is_same<L, ub::layout::first_order> {};
template <typename L>
inline constexpr bool is_first_order_v = is_first_order<L>::value;

template <typename L>
struct is_last_order : std::is_same<L, ub::layout::last_order> {};
template <typename L>
inline constexpr bool is_last_order_v = is_last_order<L>::value;

namespace tag {
struct trans {};
struct inplace {};
struct outplace {};
}  // namespace tag

}  // namespace amt

#endif  // B200_AMT_UTILS_HPP
