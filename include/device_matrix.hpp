// amt::device_matrix<T, Layout> — an M x N matrix resident in HBM, the device-side counterpart of
// the host tensors `amt::make_tensor` builds.  Replaces amt::aligned_buff (include/aligned_buff.hpp)
// as the RAII owner of working storage, and gives the harness operands whose timing excludes PCIe.
//
//   auto dA = amt::make_device_matrix<float, L>(M, K, 1.f);       // like make_tensor<T, L>(M, K, val)
//   amt::mtm(dC, dA, dB)();                                        // C += A*B on the current stream
//   dC.copy_to(host_tensor);
#ifndef B200_AMT_DEVICE_MATRIX_HPP
#define B200_AMT_DEVICE_MATRIX_HPP

#include <boost/numeric/ublas/tensor.hpp>

#include <cstddef>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "b200_mtm.h"

namespace amt {

namespace detail {
inline void b200_check(int rc, const char* what) {
    if (rc != B200_OK) throw std::runtime_error(std::string(what) + " [B200]: " + b200_last_error());
}
}  // namespace detail

template <typename T, typename Layout = boost::numeric::ublas::layout::first_order>
class device_matrix {
public:
    using value_type = T;
    using layout_type = Layout;
    static_assert(std::is_same_v<T, float> || std::is_same_v<T, double>, "float or double");

    device_matrix(std::size_t m, std::size_t n) : m_n{m, n} {
        if constexpr (std::is_same_v<Layout, boost::numeric::ublas::layout::first_order>) {
            m_w[0] = 1;
            m_w[1] = m;
        } else {
            m_w[0] = n;
            m_w[1] = 1;
        }
        detail::b200_check(b200_malloc(reinterpret_cast<void**>(&m_ptr), bytes()), "device_matrix");
        detail::b200_check(b200_memset(m_ptr, 0, bytes(), nullptr), "device_matrix");   // tensors start at zero (utils.hpp:23)
    }
    device_matrix(device_matrix const&) = delete;
    device_matrix& operator=(device_matrix const&) = delete;
    device_matrix(device_matrix&& o) noexcept : m_ptr(o.m_ptr) {
        m_n[0] = o.m_n[0]; m_n[1] = o.m_n[1]; m_w[0] = o.m_w[0]; m_w[1] = o.m_w[1];
        o.m_ptr = nullptr;
    }
    ~device_matrix() {
        if (m_ptr) b200_free(m_ptr);
    }

    T* data() noexcept { return m_ptr; }
    T const* data() const noexcept { return m_ptr; }
    std::size_t const* extents() const noexcept { return m_n; }
    std::size_t const* strides() const noexcept { return m_w; }
    std::size_t size() const noexcept { return m_n[0] * m_n[1]; }
    std::size_t size(std::size_t i) const noexcept { return m_n[i]; }
    std::size_t bytes() const noexcept { return size() * sizeof(T); }

    void fill(T val) {
        std::vector<T> h(size(), val);
        detail::b200_check(b200_memcpy_h2d(m_ptr, h.data(), bytes(), nullptr), "device_matrix::fill");
        detail::b200_check(b200_stream_synchronize(nullptr), "device_matrix::fill");
    }
    // Host tensors of the same layout and extents (flat storage copy).
    template <typename Tensor>
    void copy_from(Tensor const& t) {
        if (t.size() != size()) throw std::runtime_error("device_matrix::copy_from: size mismatch");
        detail::b200_check(b200_memcpy_h2d(m_ptr, t.data(), bytes(), nullptr), "device_matrix::copy_from");
        detail::b200_check(b200_stream_synchronize(nullptr), "device_matrix::copy_from");
    }
    template <typename Tensor>
    void copy_to(Tensor& t) const {
        if (t.size() != size()) throw std::runtime_error("device_matrix::copy_to: size mismatch");
        detail::b200_check(b200_memcpy_d2h(t.data(), m_ptr, bytes(), nullptr), "device_matrix::copy_to");
        detail::b200_check(b200_stream_synchronize(nullptr), "device_matrix::copy_to");
    }

private:
    T* m_ptr{nullptr};
    std::size_t m_n[2]{};
    std::size_t m_w[2]{};
};

template <typename T, typename L = boost::numeric::ublas::layout::first_order>
device_matrix<T, L> make_device_matrix(std::size_t M, std::size_t N) {
    return device_matrix<T, L>(M, N);
}
template <typename T, typename L = boost::numeric::ublas::layout::first_order>
device_matrix<T, L> make_device_matrix(std::size_t M, std::size_t N, T val) {
    device_matrix<T, L> d(M, N);
    d.fill(val);
    return d;
}

}  // namespace amt

#endif  // B200_AMT_DEVICE_MATRIX_HPP
