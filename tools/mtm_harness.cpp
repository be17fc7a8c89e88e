// mtm benchmark harness for the B200 path — the counterpart of the reference's src/mtm.cpp
// (main() at src/mtm.cpp:357-418, openmp_gemm at :189-214): the same size sweep
// (amt::range(x, 32, 3072, 32), src/mtm.cpp:373-376), the same inputs (all-ones A and B, zero C
// that keeps accumulating, :204-206), the same flop count M*N*(2K-1) (:203), the same report
// (metric.str(), tensor.csv).  One series per kernel family, device-resident (CUDA-event timed),
// plus the reference's own call shape — amt::benchmark<4>(amt::mtm(res, A, B, nullopt)) on host
// tensors — which includes the PCIe copies.
//
//   mtm_harness [--type f32|f64] [--layout F|L] [--max 3072] [--step 32] [--iters 4]
//               [--mrect|--nrect|--krect --fixed 1024] [--csv tensor.csv] [--host]
//
// Build: g++ -std=c++20 -O2 -Iinclude -Iinclude/compat tools/mtm_harness.cpp -Lopenmp-blas_b200 -lb200mtm
#include <boost/numeric/ublas/tensor.hpp>

#include <benchmark.hpp>
#include <metric.hpp>
#include <mtm.hpp>
#include <range.hpp>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <optional>
#include <string>
#include <vector>

namespace ub = boost::numeric::ublas;

struct options {
    bool f64 = false, last_order = false, host = false;
    bool mrect = false, nrect = false, krect = false;   // src/mtm.cpp:25-40
    std::size_t fixed = 1024, max = 3072, step = 32, iters = 4;
    std::string csv = "tensor.csv";
};

template <typename T, typename L>
int run(options const& o) {
    std::vector<double> x;
    amt::range(x, 32., static_cast<double>(o.max) + 1., static_cast<double>(o.step), std::plus<>{});
    amt::metric<T> m(x.size());
    b200_device_info info;
    if (b200_get_device_info(0, &info) != B200_OK) {
        std::cerr << "no CUDA device: " << b200_last_error() << '\n';
        return 77;
    }
    std::cout << info.name << ": " << info.sm_count << " SMs @ " << info.sm_clock_khz / 1000 << " MHz, peak "
              << (o.f64 ? info.peak_fp64_tflops : info.peak_fp32_tflops) << " TFLOP/s ("
              << (o.f64 ? "fp64" : "fp32") << " FMA pipe)\n";

    struct family { const char* name; int variant; };
    std::vector<family> fams;
    if (o.f64) fams = {{"B200.dmma", B200_MTM_DMMA}, {"B200.dfma", B200_MTM_DFMA}};
    else fams = {{"B200.3xtf32", B200_MTM_3XTF32}, {"B200.ffma", B200_MTM_SIMT}};

    amt::timer total;
    for (double el : x) {
        auto const sz = static_cast<std::size_t>(el);
        std::size_t const M = o.mrect ? o.fixed : sz, N = o.nrect ? o.fixed : sz, K = o.krect ? o.fixed : sz;
        double const ops = static_cast<double>(M) * static_cast<double>(N) * (2. * static_cast<double>(K) - 1.);
        auto dA = amt::make_device_matrix<T, L>(M, K, T(1));
        auto dB = amt::make_device_matrix<T, L>(K, N, T(1));
        for (auto const& f : fams) {
            if (b200_mtm_num_configs(f.variant, o.f64) == 0) continue;
            auto dC = amt::make_device_matrix<T, L>(M, N);
            double ns = 0;
            // amt::benchmark<4> protocol with a warm-up; iteration count is a run-time option here
            if (o.iters == 4) ns = amt::device_benchmark<4, 2>(dC, dA, dB, f.variant);
            else ns = amt::device_benchmark<16, 3>(dC, dA, dB, f.variant);
            m[f.name].update(ops / ns);   // flop/ns == GFLOP/s (src/mtm.cpp:210)
        }
        if (o.host) {
            auto A = amt::make_tensor<T, L>(M, K, T(1));
            auto B = amt::make_tensor<T, L>(K, N, T(1));
            auto res = amt::make_tensor<T, L>(M, N);
            auto bench_fn = amt::mtm(res, A, B, std::nullopt);         // src/mtm.cpp:207
            double const st = amt::benchmark<4>(std::move(bench_fn));   // src/mtm.cpp:208
            amt::no_opt(res);
            m["B200.host-tensors(e2e)"].update(ops / st);
        }
    }
    std::cerr << "sweep has completed! ( " << total << " )\n";
    std::cout << m.str(o.f64 ? "B200.dmma" : "B200.3xtf32") << '\n';
    m.csv(o.csv);
    return 0;
}

int main(int argc, char** argv) {
    options o;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&](const char* what) -> std::string {
            if (i + 1 >= argc) { std::cerr << "missing value for " << what << '\n'; std::exit(2); }
            return argv[++i];
        };
        if (a == "--type") o.f64 = next("--type") == "f64";
        else if (a == "--layout") o.last_order = next("--layout") == "L";
        else if (a == "--max") o.max = std::stoul(next("--max"));
        else if (a == "--step") o.step = std::stoul(next("--step"));
        else if (a == "--iters") o.iters = std::stoul(next("--iters"));
        else if (a == "--fixed") o.fixed = std::stoul(next("--fixed"));
        else if (a == "--csv") o.csv = next("--csv");
        else if (a == "--mrect") o.mrect = true;
        else if (a == "--nrect") o.nrect = true;
        else if (a == "--krect") o.krect = true;
        else if (a == "--host") o.host = true;
        else { std::cerr << "unknown option " << a << '\n'; return 2; }
    }
    int rc;
    if (o.f64) rc = o.last_order ? run<double, ub::layout::last_order>(o) : run<double, ub::layout::first_order>(o);
    else rc = o.last_order ? run<float, ub::layout::last_order>(o) : run<float, ub::layout::first_order>(o);
    b200_shutdown();
    return rc;
}
