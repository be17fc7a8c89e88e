// One amt::mtm call on HOST tensors spread over the GPUs of the box (include/mtm.hpp -> b200_mtm_f32_mgpu):
// BASELINE.json configs[4] (fp32 32768^3, row-major) through the reference's own entry point, the way
// src/mtm.cpp:204-208 calls it (make_tensor storage = ordinary pageable host memory, amt::benchmark-style
// repeated calls).  Integer data in [0, 9]: every summation order is exact, sampled rows are compared with an
// integer product on the host.  Prints one JSON line.
//
//   mtm_mgpu_check [--size 32768] [--devices N (0 = all)] [--calls 2] [--layout L|F]
//
// Build: g++ -std=c++20 -O2 -fopenmp -Iinclude -Iinclude/compat tools/mtm_mgpu_check.cpp -Lopenmp-blas_b200 -lb200mtm
#include <boost/numeric/ublas/tensor.hpp>

#include <mtm.hpp>
#include <timer.hpp>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace ub = boost::numeric::ublas;

static inline std::uint32_t mix(std::uint64_t x) {       // counter-based: any thread can fill any range
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return static_cast<std::uint32_t>(x);
}

template <typename L>
int run(std::size_t n, int devices, int calls) {
    int visible = 0;
    if (b200_device_count(&visible) != B200_OK || visible == 0) {
        std::fprintf(stderr, "no CUDA device: %s\n", b200_last_error());
        return 77;
    }
    amt::b200::set_devices(devices);
    auto A = amt::make_tensor<float, L>(n, n);
    auto B = amt::make_tensor<float, L>(n, n);
    auto C = amt::make_tensor<float, L>(n, n);
    float* a = A.data();
    float* b = B.data();
    std::size_t const total = n * n;
#pragma omp parallel for schedule(static)
    for (std::size_t i = 0; i < total; ++i) {
        a[i] = static_cast<float>(mix(i * 2 + 1) % 10);
        b[i] = static_cast<float>(mix(i * 2 + 0x9e3779b97f4a7c15ULL) % 10);
    }
    auto fn = amt::mtm(C, A, B, std::nullopt);
    std::vector<double> ms;
    for (int it = 0; it < calls; ++it) {
        auto t0 = std::chrono::steady_clock::now();
        fn();
        auto t1 = std::chrono::steady_clock::now();
        ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    b200_mtm_choice ch{};
    b200_mtm_last_choice(&ch);
    // exactness: C == calls * A*B on sampled rows (all columns), integer arithmetic on the host
    bool const row_major = std::is_same_v<L, ub::layout::last_order>;
    auto at = [&](float const* p, std::size_t i, std::size_t j) { return row_major ? p[i * n + j] : p[i + j * n]; };
    std::size_t const n_rows = 24;
    long long bad = 0;
    float const* c = C.data();
#pragma omp parallel for schedule(dynamic) reduction(+ : bad)
    for (std::size_t s = 0; s < n_rows; ++s) {
        std::size_t const i = (s * (n - 1)) / (n_rows - 1);
        std::vector<long long> acc(n, 0);
        for (std::size_t k = 0; k < n; ++k) {
            long long const av = static_cast<long long>(at(a, i, k));
            if (av == 0) continue;
            for (std::size_t j = 0; j < n; ++j) acc[j] += av * static_cast<long long>(at(b, k, j));
        }
        for (std::size_t j = 0; j < n; ++j)
            if (static_cast<long long>(at(c, i, j)) != acc[j] * calls) ++bad;
    }
    double best = ms[0];
    for (double v : ms) best = v < best ? v : best;
    double const flop = static_cast<double>(n) * n * (2.0 * n - 1.0);
    std::printf("{\"tool\": \"mtm_mgpu_check\", \"size\": %zu, \"layout\": \"%s\", \"devices_visible\": %d, \"devices_requested\": %d, "
                "\"calls\": %d, \"ms_first\": %.3f, \"ms_best\": %.3f, \"tflops_e2e_best\": %.3f, \"kernel\": \"%s\", \"launches\": %d, "
                "\"exact\": %s, \"mismatches\": %lld, \"host_memory\": \"pageable (make_tensor)\"}\n",
                n, row_major ? "L" : "F", visible, devices, calls, ms[0], best, flop / best / 1e9, ch.name, ch.launches,
                bad == 0 ? "true" : "false", bad);
    return bad == 0 ? 0 : 1;
}

int main(int argc, char** argv) {
    std::size_t n = 32768;
    int devices = 0, calls = 2;
    bool first_order = false;
    for (int i = 1; i < argc; ++i) {
        std::string const arg = argv[i];
        if (arg == "--size" && i + 1 < argc) n = std::strtoull(argv[++i], nullptr, 10);
        else if (arg == "--devices" && i + 1 < argc) devices = std::atoi(argv[++i]);
        else if (arg == "--calls" && i + 1 < argc) calls = std::atoi(argv[++i]);
        else if (arg == "--layout" && i + 1 < argc) first_order = argv[++i][0] == 'F';
        else {
            std::fprintf(stderr, "usage: mtm_mgpu_check [--size n] [--devices N] [--calls c] [--layout L|F]\n");
            return 2;
        }
    }
    if (calls < 1) calls = 1;
    return first_order ? run<ub::layout::first_order>(n, devices, calls) : run<ub::layout::last_order>(n, devices, calls);
}
