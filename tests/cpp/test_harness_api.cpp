// CPU-only checks of the measurement utilities the mtm harness uses (range / metric / timer /
// benchmark): same call shapes and report layout as the reference's include/{range,metric,timer,
// benchmark}.hpp.  No GPU needed: metric gets an explicit peak, benchmark times a host lambda.
#include <benchmark.hpp>
#include <metric.hpp>
#include <range.hpp>

#include <cstdio>
#include <fstream>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { ++g_fail; std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); } } while (0)

int main(int argc, char** argv) {
    std::string const csv_path = argc > 1 ? argv[1] : "/tmp/b200_metric_test.csv";
    // amt::range(x, 32., 3072., 32., std::plus<>{}) -> 32, 64, ..., 3040 (95 points), src/mtm.cpp:373-376
    std::vector<double> x;
    amt::range(x, 32., 3072., 32., std::plus<>{});
    CHECK(x.size() == 95 && x.front() == 32. && x.back() == 3040.);
    amt::range(x, 2., 16384., 2., std::multiplies<>{});     // the commented geometric sweep, src/mtm.cpp:377-378
    CHECK(x.size() == 13 && x.back() == 8192.);
    bool threw = false;
    try { amt::range(x, 10., 5., 1.); } catch (std::runtime_error const&) { threw = true; }
    CHECK(threw);

    amt::metric<float> m(3, 1000.0);
    for (double g : {100., 300., 200.}) m["tensor"].update(g);
    for (double g : {50., 150., 100.}) m["other"].update(g);
    std::string const s = m.str("tensor");
    CHECK(s.find("Peak Performance: 1000 GFlops") != std::string::npos);
    CHECK(s.find("Name: tensor") != std::string::npos && s.find("Name: other") != std::string::npos);
    CHECK(s.find("Min GFlops: 100") != std::string::npos && s.find("Max GFlops: 300") != std::string::npos);
    CHECK(s.find("Avg GFlops: 200") != std::string::npos);
    CHECK(s.find("Max Peak Utilization in %: 30") != std::string::npos);
    CHECK(s.find("Max SpeedUp with respect to tensor: 2") != std::string::npos);   // 300 / 150
    CHECK(s.find("Avg SpeedUp with respect to tensor: 2") != std::string::npos);   // 200 / 100
    m.csv(csv_path);
    std::ifstream f(csv_path);
    std::string l0, l1, l2, l3;
    std::getline(f, l0); std::getline(f, l1); std::getline(f, l2); std::getline(f, l3);
    CHECK(l0 == "\"tensor\",\"other\"");
    CHECK(l1 == "100,50" && l2 == "300,150" && l3 == "200,100");

    int calls = 0;
    double const ns = amt::benchmark<4>([&] { ++calls; });
    CHECK(calls == 4 && ns >= 0.0);
    double const ns2 = amt::benchmark<3>([&](int a) { return a + calls; }, 2);   // non-void result goes through no_opt
    CHECK(ns2 >= 0.0);
    amt::timer t;
    t.stop();
    CHECK(t.nano() >= 0.0 && t.milli() <= t.micro());
    std::ostringstream os;
    os << t;
    CHECK(!os.str().empty());
    std::printf("harness api: %d failures\n", g_fail);
    return g_fail ? 1 : 0;
}
