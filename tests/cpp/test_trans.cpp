// C++ front-end tests for amt::transpose on the B200, re-expressing the reference's
// test/test.trans.cpp (out-of-place and in-place, sizes 2..31, rand()%100; Eigen's .transpose()
// replaced by direct element comparison), plus layout pairings, rectangular shapes and the throws.
#include <boost/numeric/ublas/tensor.hpp>

#include <trans.hpp>

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

template <typename T, typename LC, typename LA>
void outplace(std::size_t M, std::size_t N) {
    auto tA = amt::make_tensor<T, LA>(M, N);
    auto tres = amt::make_tensor<T, LC>(N, M);
    rand_gen<T>(tA);
    amt::transpose(tres, tA, std::nullopt)();
    bool ok = true;
    for (std::size_t i = 0; i < M; ++i)
        for (std::size_t j = 0; j < N; ++j) ok = ok && tres(j, i) == tA(i, j);
    CHECK(ok);
}

template <typename T>
void all_cases() {
    for (std::size_t sz = 2; sz < 32; ++sz) {      // test/test.trans.cpp:12-38
        outplace<T, F, F>(sz, sz);
        auto tA = amt::make_tensor<T>(sz, sz);     // :46-72, in place
        rand_gen<T>(tA);
        auto temp = tA;
        amt::transpose(tA, std::nullopt)();
        bool ok = true;
        for (std::size_t i = 0; i < sz; ++i)
            for (std::size_t j = 0; j < sz; ++j) ok = ok && tA(i, j) == temp(j, i);
        CHECK(ok);
    }
    outplace<T, F, L>(37, 129);
    outplace<T, L, F>(129, 37);
    outplace<T, L, L>(300, 65);
    outplace<T, F, F>(1, 50);
    auto A = amt::make_tensor<T>(4, 5);
    auto C = amt::make_tensor<T>(4, 5);
    bool threw = false;
    try {
        (void)amt::transpose(C, A, std::nullopt);
    } catch (std::runtime_error const& e) {
        threw = std::string(e.what()).find("dimension mismatch") != std::string::npos;
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
    std::printf("transpose: %d checks, %d failures\n", g_checks, g_failures);
    b200_shutdown();
    return g_failures ? 1 : 0;
}
