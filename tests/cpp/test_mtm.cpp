// C++ front-end tests for amt::mtm on the B200, re-expressing the reference's test/test.mtm.cpp
// (8 (C,A,B) layout combinations x {float,double} x square sizes 2..31, inputs rand()%100 —
// test/test_utils.hpp:4-9) through the unchanged amt::mtm signature.
//
// Differences from the reference test, all deliberate:
//   * Catch2 is not installed: a 30-line CHECK harness replaces it;
//   * BLIS is not installed: the comparator is an exact 64-bit integer triple loop.  With inputs
//     in [0,99] and K <= 31 every product and partial sum is exactly representable in fp32, so
//     this is a bit-exact known-answer test (stricter than the reference's Approx compare);
//   * added: the two validation throws (mtm.hpp:234-250), repeated-call accumulation
//     (src/mtm.cpp:207-208 relies on it), rectangular and non-zero-initial-C cases.
//
// Build/run: see tests/test_cpp_frontend.py.  Exit code 0 = all passed.
#include <boost/numeric/ublas/tensor.hpp>

#include <benchmark.hpp>
#include <mtm.hpp>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <optional>
#include <string>
#include <vector>

namespace ub = boost::numeric::ublas;
using F = ub::layout::first_order;
using L = ub::layout::last_order;

static int g_failures = 0;
static int g_checks = 0;

#define CHECK(cond)                                                              \
    do {                                                                         \
        ++g_checks;                                                              \
        if (!(cond)) {                                                           \
            ++g_failures;                                                        \
            if (g_failures <= 20)                                                \
                std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
        }                                                                        \
    } while (0)

template <typename T, typename Tensor>
void rand_gen(Tensor& t) {  // test/test_utils.hpp:4-9
    for (auto& v : t) v = static_cast<T>(std::rand() % 100);
}

template <typename Tensor>
int64_t at_i64(Tensor const& t, std::size_t i, std::size_t j) {
    return static_cast<int64_t>(t(i, j));
}

// C_expected(i,j) = C0(i,j) + calls * sum_k A(i,k) B(k,j), exactly.
template <typename TC, typename TA, typename TB>
bool matches_exact(TC const& c, TC const& c0, TA const& a, TB const& b, int calls) {
    std::size_t const M = a.size(0), K = a.size(1), N = b.size(1);
    for (std::size_t i = 0; i < M; ++i)
        for (std::size_t j = 0; j < N; ++j) {
            int64_t s = 0;
            for (std::size_t k = 0; k < K; ++k) s += at_i64(a, i, k) * at_i64(b, k, j);
            int64_t const want = at_i64(c0, i, j) + calls * s;
            if (static_cast<double>(c(i, j)) != static_cast<double>(want)) return false;
        }
    return true;
}

// One reference TEMPLATE_TEST_CASE: "(XYZ) Matrix Matrix Product for Range[Start: 2, End: 32, Step: 1]".
template <typename T, typename LC, typename LA, typename LB>
void range_case(const char* name) {
    constexpr std::size_t MinSize = 2, MaxSize = 32, Step = 1;  // test/test.mtm.cpp:32-36
    int before = g_failures;
    for (std::size_t sz = MinSize; sz < MaxSize; sz += Step) {
        auto A = amt::make_tensor<T, LA>(sz, sz);
        auto B = amt::make_tensor<T, LB>(sz, sz);
        rand_gen<T>(A);
        rand_gen<T>(B);
        auto rres = amt::make_tensor<T, LC>(sz, sz);
        auto zero = amt::make_tensor<T, LC>(sz, sz);
        amt::mtm(rres, A, B, std::nullopt)();  // test/test.mtm.cpp:70
        CHECK(matches_exact(rres, zero, A, B, 1));
    }
    std::printf("%-4s %-6s sz 2..31 : %s\n", name, sizeof(T) == 4 ? "float" : "double",
                before == g_failures ? "ok" : "FAILED");
}

template <typename T>
void all_layouts() {
    range_case<T, F, F, F>("FFF");  // test/test.mtm.cpp:30
    range_case<T, F, F, L>("FFL");  // :84
    range_case<T, F, L, F>("FLF");  // :138
    range_case<T, L, F, F>("LFF");  // :192
    range_case<T, F, L, L>("FLL");  // :246
    range_case<T, L, F, L>("LFL");  // :300
    range_case<T, L, L, F>("LLF");  // :354
    range_case<T, L, L, L>("LLL");  // :408
}

template <typename T>
void extra_cases() {
    // Rectangular, non-zero initial C, two invocations of the same callable (accumulates twice).
    {
        auto A = amt::make_tensor<T, L>(37, 5);
        auto B = amt::make_tensor<T, F>(5, 129);
        auto C = amt::make_tensor<T, L>(37, 129);
        rand_gen<T>(A);
        rand_gen<T>(B);
        rand_gen<T>(C);
        auto C0 = C;
        auto fn = amt::mtm(C, A, B, std::optional<std::size_t>{4});
        fn();
        CHECK(matches_exact(C, C0, A, B, 1));
        fn();
        CHECK(matches_exact(C, C0, A, B, 2));
    }
    {
        auto A = amt::make_tensor<T, F>(200, 67);
        auto B = amt::make_tensor<T, L>(67, 3);
        auto C = amt::make_tensor<T, F>(200, 3);
        rand_gen<T>(A);
        rand_gen<T>(B);
        auto C0 = C;
        amt::mtm(C, A, B, std::nullopt)();
        CHECK(matches_exact(C, C0, A, B, 1));
    }
    // Validation throws at mtm() call time, with the reference's message (mtm.hpp:243-250).
    {
        auto A = amt::make_tensor<T>(4, 5);
        auto B = amt::make_tensor<T>(6, 3);  // K mismatch
        auto C = amt::make_tensor<T>(4, 3);
        bool threw = false;
        try {
            (void)amt::mtm(C, A, B, std::nullopt);
        } catch (std::runtime_error const& e) {
            threw = std::string(e.what()).find("dimension mismatch") != std::string::npos;
        }
        CHECK(threw);
    }
    {
        auto A = amt::make_tensor<T>(4, 5);
        auto B = amt::make_tensor<T>(5, 3);
        auto C = amt::make_tensor<T>(5, 3);  // M mismatch
        bool threw = false;
        try {
            (void)amt::mtm(C, A, B, std::nullopt);
        } catch (std::runtime_error const& e) {
            threw = std::string(e.what()).find("dimension mismatch") != std::string::npos;
        }
        CHECK(threw);
    }
}

// Device-resident operands through amt::device_matrix + the device overload of amt::mtm.
template <typename T>
void device_matrix_cases() {
    auto A = amt::make_tensor<T, L>(70, 33);
    auto B = amt::make_tensor<T, F>(33, 45);
    auto C = amt::make_tensor<T, L>(70, 45);
    rand_gen<T>(A);
    rand_gen<T>(B);
    rand_gen<T>(C);
    auto C0 = C;
    auto dA = amt::make_device_matrix<T, L>(70, 33);
    auto dB = amt::make_device_matrix<T, F>(33, 45);
    auto dC = amt::make_device_matrix<T, L>(70, 45);
    dA.copy_from(A);
    dB.copy_from(B);
    dC.copy_from(C);
    auto fn = amt::mtm(dC, dA, dB);
    fn();
    fn();
    dC.copy_to(C);
    CHECK(matches_exact(C, C0, A, B, 2));
    auto ones = amt::make_device_matrix<T, F>(64, 64, T(1));     // src/mtm.cpp:204-206 inputs
    auto res = amt::make_device_matrix<T, F>(64, 64);
    double const ns = amt::device_benchmark<4, 1>(res, ones, ones);
    CHECK(ns > 0.0);
    auto H = amt::make_tensor<T, F>(64, 64);
    res.copy_to(H);
    CHECK(H(0, 0) == T(5 * 64) && H(63, 63) == T(5 * 64));      // 1 warm-up + 4 timed calls accumulate
    bool threw = false;
    try {
        auto bad = amt::make_device_matrix<T, F>(65, 64);
        (void)amt::mtm(bad, ones, ones);
    } catch (std::runtime_error const& e) {
        threw = std::string(e.what()).find("dimension mismatch") != std::string::npos;
    }
    CHECK(threw);
}

// Medium shapes forced onto the TMA-fed FFMA configs (AUTO only picks them for large problems), so
// that compute-sanitizer (tools/sanitize.sh) also covers the mbarrier / cp.async.bulk.tensor kernels
// and the operand pack pass, including ragged edges.
void tma_ffma_cases() {
    int const n_classic = 5;   // register-staged configs come first (csrc/mtm_simt_f32.cu)
    int const n_all = b200_mtm_num_configs(B200_MTM_SIMT, 0);
    for (int cfg = n_classic; cfg < n_all; ++cfg) {
        amt::b200::set_variant(B200_MTM_SIMT, cfg);
        {
            auto A = amt::make_tensor<float, L>(261, 131);     // k-contiguous, odd sizes: packed + zero-filled edges
            auto B = amt::make_tensor<float, L>(131, 387);
            auto C = amt::make_tensor<float, L>(261, 387);
            rand_gen<float>(A); rand_gen<float>(B); rand_gen<float>(C);
            auto C0 = C;
            amt::mtm(C, A, B, std::nullopt)();
            CHECK(matches_exact(C, C0, A, B, 1));
        }
        {
            auto A = amt::make_tensor<float, F>(256, 160);     // mn-contiguous, aligned: TMA reads the operands in place
            auto B = amt::make_tensor<float, L>(160, 384);
            auto C = amt::make_tensor<float, F>(256, 384);
            rand_gen<float>(A); rand_gen<float>(B);
            auto C0 = C;
            amt::mtm(C, A, B, std::nullopt)();
            CHECK(matches_exact(C, C0, A, B, 1));
        }
    }
    std::printf("tma-fed ffma configs %d..%d : done\n", n_classic, n_all - 1);
}

// Same for the TMA-fed fp64 DMMA configs (K-contiguous swizzled tiles, pack pass for other layouts).
void tma_dmma_cases() {
    int const n_classic = 5;   // register-staged DMMA configs come first (csrc/mtm_dmma_f64.cu)
    int const n_all = b200_mtm_num_configs(B200_MTM_DMMA, 1);
    for (int cfg = n_classic; cfg < n_all; ++cfg) {
        amt::b200::set_variant(B200_MTM_DMMA, cfg);
        {
            auto A = amt::make_tensor<double, F>(133, 77);      // m-contiguous, odd sizes: both operands packed
            auto B = amt::make_tensor<double, L>(77, 195);
            auto C = amt::make_tensor<double, L>(133, 195);
            rand_gen<double>(A); rand_gen<double>(B); rand_gen<double>(C);
            auto C0 = C;
            amt::mtm(C, A, B, std::nullopt)();
            CHECK(matches_exact(C, C0, A, B, 1));
        }
        {
            auto A = amt::make_tensor<double, L>(128, 96);      // k-contiguous, aligned: TMA reads A in place
            auto B = amt::make_tensor<double, F>(96, 192);      // B^T k-contiguous as well
            auto C = amt::make_tensor<double, F>(128, 192);
            rand_gen<double>(A); rand_gen<double>(B);
            auto C0 = C;
            amt::mtm(C, A, B, std::nullopt)();
            CHECK(matches_exact(C, C0, A, B, 1));
        }
    }
    std::printf("tma-fed dmma configs %d..%d : done\n", n_classic, n_all - 1);
}

int main(int argc, char** argv) {
    // Optional argument: kernel family to force (1 = SIMT, 2 = 3xTF32 [float only], 4 = DMMA [double only]).
    int const variant = argc > 1 ? std::atoi(argv[1]) : B200_MTM_AUTO;
    int ndev = 0;
    if (b200_device_count(&ndev) != B200_OK || ndev == 0) {
        std::fprintf(stderr, "no CUDA device: %s\n", b200_last_error());
        return 77;
    }
    std::srand(1);
    if (variant != B200_MTM_DMMA) {
        amt::b200::set_variant(variant == B200_MTM_3XTF32 ? B200_MTM_3XTF32 : variant);
        all_layouts<float>();
        extra_cases<float>();
        device_matrix_cases<float>();
        if (variant == B200_MTM_AUTO || variant == B200_MTM_SIMT) tma_ffma_cases();
    }
    if (variant != B200_MTM_3XTF32) {
        amt::b200::set_variant(variant);
        all_layouts<double>();
        extra_cases<double>();
        device_matrix_cases<double>();
        if (variant == B200_MTM_AUTO || variant == B200_MTM_DMMA) tma_dmma_cases();
    }
    std::printf("%d checks, %d failures\n", g_checks, g_failures);
    b200_shutdown();
    return g_failures ? 1 : 0;
}
