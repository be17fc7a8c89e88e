// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Thin C exports around the *unmodified* reference headers under /root/reference/include
// (amitsingh19975/OpenMP-BLAS).  This translation unit contains no arithmetic of its own:
// every result it returns is produced by the reference's amt::mtm / amt::mtm_helper /
// amt::pack.  It is compiled by oracle/Makefile into oracle/_ref/libref_mtm_<isa>.so and
// used (a) to pin oracle/oracle_mtm.c, (b) as the checker in tests/ and smoke(), and
// (c) as the timed CPU arm of bench.py (`--impl reference`, `cpu_baseline.kind = "reference"`).
// Nothing under openmp-blas_b200/ or include/ may link or load it.
//
// The reference headers define non-inline namespace-scope objects
// (cache_manager.hpp:223, thread_utils.hpp:82-83), so they must be included in exactly
// one TU of the library — this one.
#include <boost/numeric/ublas/tensor.hpp>  // include/compat stand-in (Boost is not installed)

#include <mtm.hpp>  // /root/reference/include/mtm.hpp, unmodified
#include <mtv.hpp>  // /root/reference/include/mtv.hpp, unmodified (amt::mtv, amt::vtm)
#include <trans.hpp>  // /root/reference/include/trans.hpp, unmodified (amt::transpose)

#include <chrono>
#include <cstring>
#include <string>

namespace ub = boost::numeric::ublas;

namespace {

thread_local std::string g_last_error;

template <class T, class LC, class LA, class LB>
int run_frontend(std::size_t M, std::size_t N, std::size_t Ka, std::size_t Kb, std::size_t Mc,
                 std::size_t Nc, T const* a, T const* b, T* c) {
    // Mirrors test/test.mtm.cpp:38-70: build tensors, call amt::mtm(...)().
    auto A = amt::make_tensor<T, LA>(M, Ka);
    auto B = amt::make_tensor<T, LB>(Kb, N);
    auto C = amt::make_tensor<T, LC>(Mc, Nc);
    std::memcpy(A.data(), a, sizeof(T) * A.size());
    std::memcpy(B.data(), b, sizeof(T) * B.size());
    std::memcpy(C.data(), c, sizeof(T) * C.size());
    try {
        amt::mtm(C, A, B, std::nullopt)();
    } catch (std::exception const& e) {
        g_last_error = e.what();
        return 1;
    }
    std::memcpy(c, C.data(), sizeof(T) * C.size());
    return 0;
}

template <class T>
int dispatch_frontend(int lc, int la, int lb, std::size_t M, std::size_t N, std::size_t Ka,
                      std::size_t Kb, std::size_t Mc, std::size_t Nc, T const* a, T const* b,
                      T* c) {
    using F = ub::layout::first_order;
    using L = ub::layout::last_order;
    int const key = (lc ? 4 : 0) | (la ? 2 : 0) | (lb ? 1 : 0);  // 1 = last_order
    switch (key) {
        case 0: return run_frontend<T, F, F, F>(M, N, Ka, Kb, Mc, Nc, a, b, c);
        case 1: return run_frontend<T, F, F, L>(M, N, Ka, Kb, Mc, Nc, a, b, c);
        case 2: return run_frontend<T, F, L, F>(M, N, Ka, Kb, Mc, Nc, a, b, c);
        case 3: return run_frontend<T, F, L, L>(M, N, Ka, Kb, Mc, Nc, a, b, c);
        case 4: return run_frontend<T, L, F, F>(M, N, Ka, Kb, Mc, Nc, a, b, c);
        case 5: return run_frontend<T, L, F, L>(M, N, Ka, Kb, Mc, Nc, a, b, c);
        case 6: return run_frontend<T, L, L, F>(M, N, Ka, Kb, Mc, Nc, a, b, c);
        default: return run_frontend<T, L, L, L>(M, N, Ka, Kb, Mc, Nc, a, b, c);
    }
}

template <class T>
void raw_helper(T* c, std::size_t const* nc, std::size_t const* wc, T const* a,
                std::size_t const* na, std::size_t const* wa, T const* b, std::size_t const* nb,
                std::size_t const* wb, int c_last_order) {
    amt::threads::clip_num_threads(std::optional<std::size_t>{});  // mtm.hpp:241
    if (c_last_order)
        amt::mtm_helper(c, nc, wc, a, na, wa, b, nb, wb, ub::layout::last_order{});
    else
        amt::mtm_helper(c, nc, wc, a, na, wa, b, nb, wb, ub::layout::first_order{});
}

template <class T>
double bench_helper(int iters, T* c, std::size_t const* nc, std::size_t const* wc, T const* a,
                    std::size_t const* na, std::size_t const* wa, T const* b,
                    std::size_t const* nb, std::size_t const* wb, int c_last_order) {
    // amt::benchmark protocol (benchmark.hpp:34-52): mean steady_clock ns over back-to-back calls.
    double total = 0.0;
    for (int i = 0; i < iters; ++i) {
        auto t0 = std::chrono::steady_clock::now();
        raw_helper(c, nc, wc, a, na, wa, b, nb, wb, c_last_order);
        auto t1 = std::chrono::steady_clock::now();
        total += std::chrono::duration<double, std::nano>(t1 - t0).count();
    }
    return total / static_cast<double>(iters);
}

// amt::mtv / amt::vtm through the reference front-end on freshly built tensors, as
// test/test.mtv.cpp:36-65 and test/test.vtm.cpp:36-65 do: A is M x N in layout LA, the vectors are
// 1 x len first_order tensors.  `c` is in/out (the first_order path accumulates).
template <class T, class LA>
int run_mtv(bool is_vtm, std::size_t M, std::size_t N, std::size_t nb_len, std::size_t nc_len, T const* a,
            T const* b, T* c) {
    auto A = amt::make_tensor<T, LA>(M, N);
    auto v = amt::make_tensor<T>(1, nb_len);
    auto r = amt::make_tensor<T>(1, nc_len);
    std::memcpy(A.data(), a, sizeof(T) * A.size());
    std::memcpy(v.data(), b, sizeof(T) * v.size());
    std::memcpy(r.data(), c, sizeof(T) * r.size());
    try {
        if (is_vtm) amt::vtm(r, A, v, std::nullopt)();
        else amt::mtv(r, A, v, std::nullopt)();
    } catch (std::exception const& e) {
        g_last_error = e.what();
        return 1;
    }
    std::memcpy(c, r.data(), sizeof(T) * r.size());
    return 0;
}

// amt::transpose through the reference front-end (test/test.trans.cpp:20-38, :58-72).
template <class T, class LC, class LA>
int run_transpose(std::size_t M, std::size_t N, std::size_t Mc, std::size_t Nc, T const* a, T* c) {
    auto A = amt::make_tensor<T, LA>(M, N);
    auto Cc = amt::make_tensor<T, LC>(Mc, Nc);
    std::memcpy(A.data(), a, sizeof(T) * A.size());
    std::memcpy(Cc.data(), c, sizeof(T) * Cc.size());
    try {
        amt::transpose(Cc, A, std::nullopt)();
    } catch (std::exception const& e) {
        g_last_error = e.what();
        return 1;
    }
    std::memcpy(c, Cc.data(), sizeof(T) * Cc.size());
    return 0;
}
template <class T>
int run_transpose_inplace(std::size_t n, T* a) {
    auto A = amt::make_tensor<T>(n, n);
    std::memcpy(A.data(), a, sizeof(T) * A.size());
    try {
        amt::transpose(A, std::nullopt)();
    } catch (std::exception const& e) {
        g_last_error = e.what();
        return 1;
    }
    std::memcpy(a, A.data(), sizeof(T) * A.size());
    return 0;
}
template <class T>
int dispatch_transpose(int lc, int la, std::size_t M, std::size_t N, std::size_t Mc, std::size_t Nc, T const* a, T* c) {
    using F = ub::layout::first_order;
    using L = ub::layout::last_order;
    if (lc) return la ? run_transpose<T, L, L>(M, N, Mc, Nc, a, c) : run_transpose<T, L, F>(M, N, Mc, Nc, a, c);
    return la ? run_transpose<T, F, L>(M, N, Mc, Nc, a, c) : run_transpose<T, F, F>(M, N, Mc, Nc, a, c);
}

template <class T, class L>
void block_sizes(std::size_t* out) {
    using P = amt::impl::matrix_partition<256ul, T, L>;  // mtm.hpp:131
    out[0] = P::mr();
    out[1] = P::nr();
    out[2] = P::kc();
    out[3] = P::mc();
    out[4] = P::nc();
}

}  // namespace

extern "C" {

char const* ref_last_error() { return g_last_error.c_str(); }

int ref_max_threads() { return amt::threads::get_max_threads(); }

// Full front-end (validation + callable), layouts: 0 = first_order (column-major), 1 = last_order.
// Extents are passed independently (A: M x Ka, B: Kb x N, C: Mc x Nc) so the dimension-mismatch
// throw (mtm.hpp:243-250) can be exercised.  Returns 0, or 1 with ref_last_error() set.
int ref_mtm_tensor_f32(int lc, int la, int lb, std::size_t M, std::size_t N, std::size_t Ka,
                       std::size_t Kb, std::size_t Mc, std::size_t Nc, float const* a,
                       float const* b, float* c) {
    return dispatch_frontend<float>(lc, la, lb, M, N, Ka, Kb, Mc, Nc, a, b, c);
}
int ref_mtm_tensor_f64(int lc, int la, int lb, std::size_t M, std::size_t N, std::size_t Ka,
                       std::size_t Kb, std::size_t Mc, std::size_t Nc, double const* a,
                       double const* b, double* c) {
    return dispatch_frontend<double>(lc, la, lb, M, N, Ka, Kb, Mc, Nc, a, b, c);
}

// Raw pointer/extent/stride interface == amt::mtm_helper (mtm.hpp:116-122); in place, C += A*B.
void ref_mtm_f32(float* c, std::size_t const* nc, std::size_t const* wc, float const* a,
                 std::size_t const* na, std::size_t const* wa, float const* b,
                 std::size_t const* nb, std::size_t const* wb, int c_last_order) {
    raw_helper(c, nc, wc, a, na, wa, b, nb, wb, c_last_order);
}
void ref_mtm_f64(double* c, std::size_t const* nc, std::size_t const* wc, double const* a,
                 std::size_t const* na, std::size_t const* wa, double const* b,
                 std::size_t const* nb, std::size_t const* wb, int c_last_order) {
    raw_helper(c, nc, wc, a, na, wa, b, nb, wb, c_last_order);
}

// Mean ns per call over `iters` back-to-back calls (C keeps accumulating, as in src/mtm.cpp:207-208).
double ref_mtm_bench_f32(int iters, float* c, std::size_t const* nc, std::size_t const* wc,
                         float const* a, std::size_t const* na, std::size_t const* wa,
                         float const* b, std::size_t const* nb, std::size_t const* wb,
                         int c_last_order) {
    return bench_helper(iters, c, nc, wc, a, na, wa, b, nb, wb, c_last_order);
}
double ref_mtm_bench_f64(int iters, double* c, std::size_t const* nc, std::size_t const* wc,
                         double const* a, std::size_t const* na, std::size_t const* wa,
                         double const* b, std::size_t const* nb, std::size_t const* wb,
                         int c_last_order) {
    return bench_helper(iters, c, nc, wc, a, na, wa, b, nb, wb, c_last_order);
}

// amt::pack (utils.hpp:99-118) and amt::pack(..., tag::trans) (utils.hpp:120-141).
void ref_pack_f32(float* out, std::size_t wo, float const* in, std::size_t const* wi,
                  std::size_t m, std::size_t n, int trans) {
    if (trans)
        amt::pack(out, wo, in, wi, m, n, amt::tag::trans{});
    else
        amt::pack(out, wo, in, wi, m, n);
}
void ref_pack_f64(double* out, std::size_t wo, double const* in, std::size_t const* wi,
                  std::size_t m, std::size_t n, int trans) {
    if (trans)
        amt::pack(out, wo, in, wi, m, n, amt::tag::trans{});
    else
        amt::pack(out, wo, in, wi, m, n);
}

// amt::mtv (is_vtm = 0) / amt::vtm (is_vtm = 1); a_last_order selects A's layout; A is M x N.
int ref_mtv_tensor_f32(int is_vtm, int a_last_order, std::size_t M, std::size_t N, std::size_t nb_len,
                       std::size_t nc_len, float const* a, float const* b, float* c) {
    namespace ub = boost::numeric::ublas;
    return a_last_order ? run_mtv<float, ub::layout::last_order>(is_vtm, M, N, nb_len, nc_len, a, b, c)
                        : run_mtv<float, ub::layout::first_order>(is_vtm, M, N, nb_len, nc_len, a, b, c);
}
int ref_mtv_tensor_f64(int is_vtm, int a_last_order, std::size_t M, std::size_t N, std::size_t nb_len,
                       std::size_t nc_len, double const* a, double const* b, double* c) {
    namespace ub = boost::numeric::ublas;
    return a_last_order ? run_mtv<double, ub::layout::last_order>(is_vtm, M, N, nb_len, nc_len, a, b, c)
                        : run_mtv<double, ub::layout::first_order>(is_vtm, M, N, nb_len, nc_len, a, b, c);
}

// amt::transpose(c, a): a is M x N in layout la, c is Mc x Nc in layout lc (0 = first_order).
int ref_transpose_tensor_f32(int lc, int la, std::size_t M, std::size_t N, std::size_t Mc, std::size_t Nc,
                             float const* a, float* c) {
    return dispatch_transpose<float>(lc, la, M, N, Mc, Nc, a, c);
}
int ref_transpose_tensor_f64(int lc, int la, std::size_t M, std::size_t N, std::size_t Mc, std::size_t Nc,
                             double const* a, double* c) {
    return dispatch_transpose<double>(lc, la, M, N, Mc, Nc, a, c);
}
// amt::transpose(a): in place, n x n first_order tensor.
int ref_transpose_inplace_f32(std::size_t n, float* a) { return run_transpose_inplace<float>(n, a); }
int ref_transpose_inplace_f64(std::size_t n, double* a) { return run_transpose_inplace<double>(n, a); }

// out[5] = {MR, NR, KB, MB, NB} (mtm.hpp:19-81) for dtype (0 = f32, 1 = f64) and C layout.
void ref_block_sizes(int is_f64, int c_last_order, std::size_t* out) {
    using F = ub::layout::first_order;
    using L = ub::layout::last_order;
    if (is_f64) {
        if (c_last_order) block_sizes<double, L>(out); else block_sizes<double, F>(out);
    } else {
        if (c_last_order) block_sizes<float, L>(out); else block_sizes<float, F>(out);
    }
}

}  // extern "C"
