// amt::metric<T> — named GFLOP/s series with min / max / average, speed-up against a named
// series and % of peak; `str()` and `csv()` keep the reference's report layout
// (include/metric.hpp:435-465, 479-500) so its tables (doc/matrix_times_matrix.tex:501-724) can be
// regenerated with a B200 column.  Differences: the peak is not the hard-coded CPU constant of
// metric.hpp:30 but is taken from the device (SMs x lanes x 2 x clock, b200_get_device_info), and
// the Matplot++ plotting members are dropped (no Matplot++ here; csv()/raw() feed any plotter).
#ifndef B200_AMT_METRIC_HPP
#define B200_AMT_METRIC_HPP

#include <algorithm>
#include <cstddef>
#include <fstream>
#include <iomanip>
#include <limits>
#include <optional>
#include <ostream>
#include <sstream>
#include <string>
#include <string_view>
#include <type_traits>
#include <utility>
#include <vector>

#include "b200_mtm.h"

namespace amt {

template <typename T>
class metric {
    static_assert(std::is_same_v<T, float> || std::is_same_v<T, double>);

public:
    struct flops_data {
        std::vector<double> plot{};
        double min{std::numeric_limits<double>::max()};
        double max{0.};
        double agg{0.};
        void update(double gflops) {
            agg += gflops;
            min = std::min(min, gflops);
            max = std::max(max, gflops);
            plot.push_back(gflops);
        }
    };
    using size_type = std::size_t;

    // `total` = number of points of the sweep (the divisor of the averages, as in the reference).
    explicit metric(size_type total, std::optional<double> peak_gflops = std::nullopt) : m_total(total) {
        if (peak_gflops) {
            m_peak = *peak_gflops;
        } else {
            b200_device_info info;
            int dev_count = 0;
            if (b200_device_count(&dev_count) == B200_OK && dev_count > 0 && b200_get_device_info(0, &info) == B200_OK)
                m_peak = 1e3 * (std::is_same_v<T, double> ? info.peak_fp64_tflops : info.peak_fp32_tflops);
        }
    }

    // Series are kept in insertion order (the reference iterates an unordered_map).
    flops_data& operator[](std::string_view name) {
        for (auto& kv : m_data)
            if (kv.first == name) return kv.second;
        m_data.emplace_back(std::string(name), flops_data{});
        m_data.back().second.plot.reserve(m_total);
        return m_data.back().second;
    }
    flops_data& insert_or_update(std::string const& name, double gflops) {
        auto& d = (*this)[name];
        d.update(gflops);
        return d;
    }
    double peak() const noexcept { return m_peak; }

    std::string str(std::optional<std::string_view> pattern = std::nullopt) const {
        std::stringstream ss;
        flops_data const* pref = nullptr;
        if (pattern)
            for (auto const& [k, v] : m_data)
                if (k.find(*pattern) != std::string::npos) {
                    pref = &v;
                    break;
                }
        ss << (std::is_same_v<T, double> ? "[Double-Precision]" : "[Single-Precision]") << '\n';
        ss << "Peak Performance: " << m_peak << " GFlops\n";
        for (auto const& [k, v] : m_data) {
            double const avg = v.agg / static_cast<double>(m_total);
            ss << "Name: " << k << '\n';
            ss << '\t' << "Min GFlops: " << v.min << '\n';
            ss << '\t' << "Max GFlops: " << v.max << '\n';
            if (pref) {
                double const pavg = pref->agg / static_cast<double>(m_total);
                ss << '\t' << "Max SpeedUp with respect to " << *pattern << ": " << (pref->max / v.max) << '\n';
                ss << '\t' << "Avg SpeedUp with respect to " << *pattern << ": " << (pavg / avg) << '\n';
            }
            ss << '\t' << "Max Peak Utilization in %: " << (v.max / m_peak) * 100. << '\n';
            ss << '\t' << "Avg GFlops: " << avg << '\n';
            ss << '\t' << "Avg Peak Utilization in %: " << (avg / m_peak) * 100. << '\n' << '\n';
        }
        return ss.str();
    }

    void raw(std::string_view filename = "raw_data.txt") const {
        std::ofstream f(filename.data());
        for (auto const& [k, v] : m_data) {
            f << k << ' ';
            for (double d : v.plot) f << d << ' ';
            f << '\n';
        }
    }

    // One quoted header per series, then one row per sweep point (reference csv(), metric.hpp:479-500).
    void csv(std::string_view filename = "raw_data.csv") const {
        std::ofstream f(filename.data());
        for (std::size_t j = 0; j < m_data.size(); ++j) f << std::quoted(m_data[j].first) << (j + 1 == m_data.size() ? '\n' : ',');
        for (size_type i = 0; i < m_total; ++i)
            for (std::size_t j = 0; j < m_data.size(); ++j) {
                auto const& p = m_data[j].second.plot;
                if (i < p.size()) f << p[i];
                f << (j + 1 == m_data.size() ? '\n' : ',');
            }
    }

    friend std::ostream& operator<<(std::ostream& os, metric const& m) { return os << m.str(); }

private:
    std::vector<std::pair<std::string, flops_data>> m_data{};
    size_type m_total{};
    double m_peak{1.};
};

}  // namespace amt

#endif  // B200_AMT_METRIC_HPP
