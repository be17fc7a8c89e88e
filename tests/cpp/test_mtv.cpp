// C++ front-end tests for amt::mtv / amt::vtm on the B200, re-expressing the reference's
// test/test.mtv.cpp and test/test.vtm.cpp: first/last order x {float,double} x sizes 2..511 with
// rand()%100 inputs.  Catch2 and BLIS are not installed: a CHECK harness and an exact 64-bit integer
// comparator (gemv with alpha = beta = 1 on a zero vector, as the reference's BLIS call) replace them.
#include <boost/numeric/ublas/tensor.hpp>

#include <mtv.hpp>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <optional>
#include <string>

namespace ub = boost::numeric::ublas;
using F = ub::layout::first_order;
using L = ub::layout::last_order;

static int g_failures = 0, g_checks = 0;
#define CHECK(cond)                                                                   \
    do {                                                                              \
        ++g_checks;                                                                   \
        if (!(cond)) {                                                                \
            ++g_failures;                                                             \
            if (g_failures <= 20) std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
        }                                                                             \
    } while (0)

template <typename T, typename Tensor>
void rand_gen(Tensor& t) {
    for (auto& v : t) v = static_cast<T>(std::rand() % 100);
}

template <typename T, typename Layout, bool IsVtm>
void range_case(const char* name) {
    int const before = g_failures;
    for (std::size_t sz = 2; sz < 512; ++sz) {   // MinSize = 2, MaxSize = 512 (test/test.mtv.cpp:31-35)
        auto A = amt::make_tensor<T, Layout>(sz, sz);
        auto v = amt::make_tensor<T>(1, sz);
        rand_gen<T>(A);
        rand_gen<T>(v);
        auto rres = amt::make_tensor<T>(1, sz);
        if constexpr (IsVtm) amt::vtm(rres, A, v, std::nullopt)();
        else amt::mtv(rres, A, v, std::nullopt)();
        bool ok = true;
        for (std::size_t i = 0; i < sz && ok; ++i) {
            int64_t s = 0;
            for (std::size_t k = 0; k < sz; ++k)
                s += IsVtm ? static_cast<int64_t>(A(k, i)) * static_cast<int64_t>(v[k])
                           : static_cast<int64_t>(A(i, k)) * static_cast<int64_t>(v[k]);
            ok = static_cast<double>(rres[i]) == static_cast<double>(s);
        }
        CHECK(ok);
    }
    std::printf("%-24s %-6s sz 2..511 : %s\n", name, sizeof(T) == 4 ? "float" : "double",
                before == g_failures ? "ok" : "FAILED");
}

template <typename T>
void all_cases() {
    range_case<T, F, false>("mtv first_order");
    range_case<T, L, false>("mtv last_order");
    range_case<T, F, true>("vtm first_order");
    range_case<T, L, true>("vtm last_order");
    // validation throws, mtv.hpp:121-146
    auto A = amt::make_tensor<T>(4, 5);
    auto v = amt::make_tensor<T>(1, 6);
    auto r = amt::make_tensor<T>(1, 4);
    bool threw = false;
    try {
        (void)amt::mtv(r, A, v, std::nullopt);
    } catch (std::runtime_error const& e) {
        threw = std::string(e.what()).find("dimension mismatch") != std::string::npos;
    }
    CHECK(threw);
    auto M2 = amt::make_tensor<T>(3, 3);
    threw = false;
    try {
        (void)amt::mtv(r, A, M2, std::nullopt);   // b is a matrix, not a vector
    } catch (std::runtime_error const& e) {
        threw = std::string(e.what()).find("must be vector") != std::string::npos;
    }
    CHECK(threw);
}

int main() {
    int ndev = 0;
    if (b200_device_count(&ndev) != B200_OK || ndev == 0) {
        std::fprintf(stderr, "no CUDA device: %s\n", b200_last_error());
        return 77;
    }
    std::srand(1);
    all_cases<float>();
    all_cases<double>();
    std::printf("%d checks, %d failures\n", g_checks, g_failures);
    b200_shutdown();
    return g_failures ? 1 : 0;
}
