// amt::range — fill a container with an arithmetic / geometric size sweep, as used by the
// reference harness (`amt::range(x, 32., 3072., 32., std::plus<>{})`, src/mtm.cpp:373-376;
// reference implementation include/range.hpp:29-53).
#ifndef B200_AMT_RANGE_HPP
#define B200_AMT_RANGE_HPP

#include <functional>
#include <stdexcept>

namespace amt {

// start, fn(start, stride), fn(fn(start, stride), stride), ... while < end
template <typename Container, typename Fn, typename V = typename Container::value_type>
void range(Container& c, V start, V end, V stride, Fn&& fn) {
    if (start > end) throw std::runtime_error("amt::range(Container&, ValueType, ValueType, ValueType, Fn&&) : start > end");
    c.clear();
    for (V v = start; v < end; v = std::invoke(fn, v, stride)) c.push_back(v);
}

template <typename Container, typename V = typename Container::value_type>
void range(Container& c, V start, V end, V stride) {
    range(c, start, end, stride, std::plus<>{});
}

}  // namespace amt

#endif  // B200_AMT_RANGE_HPP
