// Minimal stand-in for the Boost.uBLAS *develop-branch* tensor API that amt::mtm is
// written against (tensor_core / tensor_static_rank / extents<2> / layout tags).
//
// Real Boost is not installed in this image, and the API in question was never part of a
// Boost release, so the B200 build ships this header so that callers written against the
// reference (test/test.mtm.cpp, src/mtm.cpp) compile unchanged.  If a real uBLAS develop
// tree is on the include path *before* include/compat, it wins and this file is unused.
//
// What amt::mtm needs from a tensor (reference include/mtm.hpp:216-260, include/utils.hpp:15-31):
//   value_type, layout_type, extents() / strides() with operator[] and data() -> size_t const*,
//   data(), plus is_matrix(extents).  make_tensor additionally needs the (M, N) constructor
//   that value-initialises storage to zero, begin()/end() and size().
#ifndef B200_COMPAT_BOOST_UBLAS_TENSOR_HPP
#define B200_COMPAT_BOOST_UBLAS_TENSOR_HPP

#include <algorithm>
#include <array>
#include <climits>
#include <cmath>
#include <concepts>
#include <cstddef>
#include <initializer_list>
#include <iostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace boost::numeric::ublas {

namespace layout {
struct first_order {};   // column-major: strides {1, rows}
struct last_order {};    // row-major:    strides {cols, 1}
}  // namespace layout

template <class T>
concept integral = std::is_integral_v<std::remove_cvref_t<T>>;

template <unsigned N>
struct extents {
    using value_type = std::size_t;
    using size_type = std::size_t;
    using const_pointer = std::size_t const*;

    constexpr extents() noexcept = default;
    constexpr extents(std::initializer_list<std::size_t> il) {
        if (il.size() != N) throw std::length_error("ublas::extents<N>: wrong number of extents");
        std::copy(il.begin(), il.end(), m_data.begin());
    }
    constexpr std::size_t operator[](std::size_t i) const noexcept { return m_data[i]; }
    constexpr std::size_t& operator[](std::size_t i) noexcept { return m_data[i]; }
    constexpr std::size_t at(std::size_t i) const { return m_data.at(i); }
    constexpr const_pointer data() const noexcept { return m_data.data(); }
    constexpr std::size_t size() const noexcept { return N; }
    constexpr auto begin() const noexcept { return m_data.begin(); }
    constexpr auto end() const noexcept { return m_data.end(); }
    constexpr bool operator==(extents const& o) const noexcept { return m_data == o.m_data; }

private:
    std::array<std::size_t, N> m_data{};
};

template <unsigned N>
constexpr std::size_t product(extents<N> const& e) noexcept {
    std::size_t p = 1;
    for (auto v : e) p *= v;
    return p;
}
// Rank-2 with both extents >= 1.  Real uBLAS additionally distinguishes vector/scalar for
// 1xN / Nx1 / 1x1; no reference test pins that, so degenerate matrices are accepted here.
template <unsigned N>
constexpr bool is_matrix(extents<N> const& e) noexcept {
    return N == 2 && e[0] >= 1 && e[1] >= 1;
}
template <unsigned N>
constexpr bool is_vector(extents<N> const& e) noexcept {
    return N == 2 && ((e[0] == 1) != (e[1] == 1));
}
template <unsigned N>
constexpr bool is_scalar(extents<N> const& e) noexcept {
    return product(e) == 1;
}

template <class T, std::size_t Rank, class Layout>
struct static_rank_engine {
    using value_type = T;
    using layout_type = Layout;
    static constexpr std::size_t rank = Rank;
};

template <class Engine>
class tensor_core {
public:
    using engine_type = Engine;
    using value_type = typename Engine::value_type;
    using layout_type = typename Engine::layout_type;
    using extents_type = ::boost::numeric::ublas::extents<static_cast<unsigned>(Engine::rank)>;
    using strides_type = extents_type;
    using container_type = std::vector<value_type>;
    using size_type = std::size_t;
    using pointer = value_type*;
    using const_pointer = value_type const*;
    using iterator = typename container_type::iterator;
    using const_iterator = typename container_type::const_iterator;

    static_assert(Engine::rank == 2, "compat tensor_core supports rank-2 tensors only");

    tensor_core() = default;
    tensor_core(size_type m, size_type n)
        : m_extents{m, n}, m_strides(make_strides(m, n)), m_data(m * n, value_type{}) {}
    explicit tensor_core(extents_type const& e) : tensor_core(e[0], e[1]) {}

    extents_type const& extents() const noexcept { return m_extents; }
    strides_type const& strides() const noexcept { return m_strides; }
    pointer data() noexcept { return m_data.data(); }
    const_pointer data() const noexcept { return m_data.data(); }
    size_type size() const noexcept { return m_data.size(); }
    size_type size(size_type i) const noexcept { return m_extents[i]; }
    static constexpr size_type rank() noexcept { return Engine::rank; }
    bool empty() const noexcept { return m_data.empty(); }

    iterator begin() noexcept { return m_data.begin(); }
    iterator end() noexcept { return m_data.end(); }
    const_iterator begin() const noexcept { return m_data.begin(); }
    const_iterator end() const noexcept { return m_data.end(); }

    value_type& operator[](size_type i) noexcept { return m_data[i]; }
    value_type const& operator[](size_type i) const noexcept { return m_data[i]; }
    value_type& operator()(size_type i, size_type j) noexcept {
        return m_data[i * m_strides[0] + j * m_strides[1]];
    }
    value_type const& operator()(size_type i, size_type j) const noexcept {
        return m_data[i * m_strides[0] + j * m_strides[1]];
    }
    value_type& at(size_type i, size_type j) {
        if (i >= m_extents[0] || j >= m_extents[1]) throw std::out_of_range("tensor_core::at");
        return (*this)(i, j);
    }
    value_type const& at(size_type i, size_type j) const {
        if (i >= m_extents[0] || j >= m_extents[1]) throw std::out_of_range("tensor_core::at");
        return (*this)(i, j);
    }

private:
    static strides_type make_strides(size_type m, size_type n) {
        if constexpr (std::is_same_v<layout_type, layout::first_order>) {
            return strides_type{1, m};
        } else {
            return strides_type{n, 1};
        }
    }

    extents_type m_extents{};
    strides_type m_strides{};
    container_type m_data{};
};

template <class T, std::size_t Rank, class Layout = layout::first_order>
using tensor_static_rank = tensor_core<static_rank_engine<T, Rank, Layout>>;

}  // namespace boost::numeric::ublas

#endif  // B200_COMPAT_BOOST_UBLAS_TENSOR_HPP
