"""Pin the CPU oracle of the transpose path (oracle/oracle_trans.c) against the reference's cases
(test/test.trans.cpp: in- and out-of-place, sizes 2..31, rand()%100 — comparator Eigen's
.transpose(), here numpy's .T) and against the reference itself (oracle/_ref), bit for bit."""
import numpy as np
import pytest

DTYPES = [np.float32, np.float64]


@pytest.mark.parametrize("dtype", DTYPES)
def test_reference_cases(dtype, oracle_lib, reference_lib):
    rng = np.random.default_rng(1)
    for sz in range(2, 32):                                   # test/test.trans.cpp:12-14, :46-48
        a = np.asfortranarray(rng.integers(0, 100, (sz, sz)).astype(dtype))
        for lib in (oracle_lib, reference_lib):
            c = np.zeros((sz, sz), dtype, order="F")
            (lib.transpose if lib is oracle_lib else lib.transpose_tensor)(c, a)
            assert np.array_equal(c, a.T)
            b = a.copy(order="K")
            lib.transpose_inplace(b)
            assert np.array_equal(b, a.T)


@pytest.mark.parametrize("dtype", DTYPES)
def test_layout_pairs_rectangular_and_views(dtype, oracle_lib, reference_lib):
    rng = np.random.default_rng(2)
    for (M, N) in ((1, 7), (7, 1), (33, 65), (300, 129), (64, 64)):
        for oa in "FC":
            for oc in "FC":
                a = np.asarray(rng.uniform(-1, 1, (M, N)).astype(dtype), order=oa)
                c1 = np.zeros((N, M), dtype, order=oc)
                c2 = np.zeros((N, M), dtype, order=oc)
                oracle_lib.transpose(c1, a)
                assert np.array_equal(c1, a.T)
                if min(M, N) >= 1:
                    reference_lib.transpose_tensor(c2, a)
                    assert np.array_equal(c2, a.T)
    big = rng.uniform(-1, 1, (90, 120)).astype(dtype)
    cbig = np.zeros((200, 150), dtype)
    ref = cbig.copy()
    oracle_lib.transpose(cbig[5:65:2, 10:50], big[3:43, 7:97:3])       # 30x40 view <- (40x30 view)^T
    ref[5:65:2, 10:50] = big[3:43, 7:97:3].T
    assert np.array_equal(cbig, ref)


def test_reference_transpose_throws(reference_lib):
    a = np.zeros((4, 5), np.float32, order="F")
    with pytest.raises(RuntimeError, match="dimension mismatch"):      # trans.hpp:121-126
        reference_lib.transpose_tensor(np.zeros((4, 5), np.float32, order="F"), a)
