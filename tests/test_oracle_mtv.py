"""Pin the CPU oracle of the matrix-times-vector path (oracle/oracle_mtv.c) against the reference's
own cases: test/test.mtv.cpp and test/test.vtm.cpp — first/last order x {f32,f64} x sizes 2..511,
rand()%100 inputs — with an exact integer comparator (the reference uses BLIS gemv with
alpha = beta = 1 on a zero result vector), and against the reference itself (oracle/_ref)."""
import numpy as np
import pytest

DTYPES = [np.float32, np.float64]


def _case(rng, sz, dtype, order, rows=None):
    rows = sz if rows is None else rows
    a = np.asarray(rng.integers(0, 100, (rows, sz)).astype(dtype), order=order)
    return a


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("order", ["F", "C"])
@pytest.mark.parametrize("is_vtm", [False, True])
def test_reference_cases_oracle_exact(is_vtm, order, dtype, oracle_lib):
    """test/test.mtv.cpp:29-126 and test/test.vtm.cpp:29-125 (Range[2, 512)), result vector starts at 0."""
    rng = np.random.default_rng(abs(hash((is_vtm, order, np.dtype(dtype).name))) % 2**32)
    for sz in range(2, 512):
        a = _case(rng, sz, dtype, order)
        v = rng.integers(0, 100, sz).astype(dtype)
        c = np.zeros(sz, dtype)
        (oracle_lib.vtm if is_vtm else oracle_lib.mtv)(c, a, v)
        ai, vi = a.astype(np.int64), v.astype(np.int64)
        want = vi @ ai if is_vtm else ai @ vi
        assert np.array_equal(c.astype(np.int64), want), (is_vtm, order, sz)


@pytest.mark.parametrize("dtype", DTYPES)
def test_accumulate_vs_assign_semantics(dtype, oracle_lib, reference_lib):
    """first_order mtv / last_order vtm accumulate into c, the other two assign — in the reference
    (simd_loop.hpp:58-75 vs mtv.hpp:94-99) and therefore in the oracle."""
    rng = np.random.default_rng(5)
    for sz in (7, 64, 300):
        for order in "FC":
            a = _case(rng, sz, dtype, order)
            v = rng.integers(0, 100, sz).astype(dtype)
            c0 = rng.integers(0, 100, sz).astype(dtype)
            for is_vtm in (False, True):
                accumulates = (order == "F") != is_vtm
                exact = (v.astype(np.int64) @ a.astype(np.int64)) if is_vtm else (a.astype(np.int64) @ v.astype(np.int64))
                want = exact + (c0.astype(np.int64) if accumulates else 0)
                c = c0.copy()
                (oracle_lib.vtm if is_vtm else oracle_lib.mtv)(c, a, v)
                assert np.array_equal(c.astype(np.int64), want)
                if not (is_vtm and order == "C"):          # see test_reference_vtm_last_order_bug
                    r = c0.copy()
                    reference_lib.mtv_tensor(is_vtm, a, v, r)
                    assert np.array_equal(r.astype(np.int64), want), (is_vtm, order, sz)


def test_reference_vtm_last_order_bug(reference_lib, oracle_lib):
    """The reference's amt::vtm for a last_order matrix passes un-flipped strides to the first_order
    helper (mtv.hpp:224-226 leaves wa = {na[0], 1}) and so computes sum_k a[i + k] v[k], not v*A; its own
    test (test/test.vtm.cpp:77-124) expects v*A via BLIS_TRANSPOSE.  The oracle and the B200 path
    compute v*A.  This test documents the deviation and fails if the reference ever changes."""
    a = np.arange(1, 10, dtype=np.float32).reshape(3, 3)     # last_order
    v = np.array([1, 2, 3], np.float32)
    r = np.zeros(3, np.float32)
    reference_lib.mtv_tensor(True, a, v, r)
    flat = a.ravel()
    buggy = np.array([sum(flat[i + k] * v[k] for k in range(3)) for i in range(3)], np.float32)
    assert np.array_equal(r, buggy) and not np.array_equal(r, v @ a)
    c = np.zeros(3, np.float32)
    oracle_lib.vtm(c, a, v)
    assert np.array_equal(c, v @ a)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("order,is_vtm", [("F", False), ("C", False), ("F", True)])
def test_oracle_vs_reference_random(order, is_vtm, dtype, oracle_lib, reference_lib):
    """Non-integer data: the reference vectorises its k loops (`omp simd reduction`), so its summation
    order is the compiler's; agreement is within rounding: |diff| <= 2 (K+1) u (|A||v| + |c0|)."""
    rng = np.random.default_rng(11)
    u = np.finfo(dtype).eps / 2
    # vtm in the reference is only right for SQUARE matrices (it strides the transposed view by
    # new_na[0] instead of the column length, mtv.hpp:222-226; all its tests are square), so the
    # rectangular comparisons are mtv-only.
    shapes = ((33, 33), (260, 260)) if is_vtm else ((33, 77), (500, 260), (1, 40), (129, 1))
    for (rows, cols) in shapes:
        a = np.asarray(rng.uniform(-1, 1, (rows, cols)).astype(dtype), order=order)
        nb, nc = (rows, cols) if is_vtm else (cols, rows)
        if min(nb, nc) < 2:
            continue                                         # uBLAS is_vector needs length >= 2 in the front-end
        v = rng.uniform(-1, 1, nb).astype(dtype)
        c0 = rng.uniform(-1, 1, nc).astype(dtype)
        r, c = c0.copy(), c0.copy()
        reference_lib.mtv_tensor(is_vtm, a, v, r)
        (oracle_lib.vtm if is_vtm else oracle_lib.mtv)(c, a, v)
        absab = (np.abs(v).astype(np.float64) @ np.abs(a).astype(np.float64)) if is_vtm else (np.abs(a).astype(np.float64) @ np.abs(v).astype(np.float64))
        bound = 2 * (nb + 1) * u * (absab + np.abs(c0)) + 1e-300
        assert np.all(np.abs(r.astype(np.float64) - c.astype(np.float64)) <= bound)


def test_reference_mtv_throws(reference_lib):
    a = np.zeros((4, 5), np.float32, order="F")
    with pytest.raises(RuntimeError, match="dimension mismatch"):        # mtv.hpp:141-146
        reference_lib.mtv_tensor(False, a, np.zeros(6, np.float32), np.zeros(4, np.float32))
    with pytest.raises(RuntimeError, match="dimension mismatch"):        # mtv.hpp:209-214
        reference_lib.mtv_tensor(True, a, np.zeros(5, np.float32), np.zeros(5, np.float32))
