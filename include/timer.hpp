// amt::timer — host wall-clock timer (steady_clock), API-compatible with the reference's
// include/timer.hpp:9-110 as far as the mtm harness uses it (start/stop, operator() in ns,
// unit accessors, stream output).  Device work is timed with CUDA events instead
// (amt::device_benchmark in benchmark.hpp).
#ifndef B200_AMT_TIMER_HPP
#define B200_AMT_TIMER_HPP

#include <chrono>
#include <ostream>

namespace amt {

struct timer {
    using clock_type = std::chrono::steady_clock;

    timer() noexcept { start(); }
    void start() noexcept { m_start = clock_type::now(); m_running = true; }
    double stop() noexcept { m_end = clock_type::now(); m_running = false; return nano(); }

    double nano() const noexcept {
        auto const end = m_running ? clock_type::now() : m_end;
        return std::chrono::duration<double, std::nano>(end - m_start).count();
    }
    double micro() const noexcept { return nano() * 1e-3; }
    double milli() const noexcept { return nano() * 1e-6; }
    double sec() const noexcept { return nano() * 1e-9; }
    double min() const noexcept { return sec() / 60.0; }
    double operator()() const noexcept { return nano(); }   // nanoseconds, like the reference
    operator double() const noexcept { return nano(); }

    friend std::ostream& operator<<(std::ostream& os, timer const& t) {
        double const s = t.sec();
        if (s >= 60.0) return os << s / 60.0 << "min";
        if (s >= 1.0) return os << s << "s";
        if (s >= 1e-3) return os << s * 1e3 << "ms";
        return os << s * 1e6 << "us";
    }

private:
    clock_type::time_point m_start{}, m_end{};
    bool m_running{false};
};

}  // namespace amt

#endif  // B200_AMT_TIMER_HPP
