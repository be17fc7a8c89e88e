"""GPU parity tests of the transpose path (b200_transpose_* C ABI via the Python / C++ mirrors):
bit-exact against the CPU oracle on the reference's cases (test/test.trans.cpp) and beyond —
every layout pairing, rectangular, strided views, in place."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
DTYPES = [np.float32, np.float64]


@pytest.mark.parametrize("dtype", DTYPES)
def test_reference_cases(dtype, ob, oracle_lib):
    rng = np.random.default_rng(1)
    for sz in range(2, 32):                                   # test/test.trans.cpp: Range[2, 32)
        a = np.asfortranarray(rng.integers(0, 100, (sz, sz)).astype(dtype))
        want = np.zeros((sz, sz), dtype, order="F")
        oracle_lib.transpose(want, a)
        got = np.zeros((sz, sz), dtype, order="F")
        ob.transpose(got, a)()
        assert np.array_equal(got, want)
        b = a.copy(order="K")
        ob.transpose_inplace(b)()
        assert np.array_equal(b, want)


@pytest.mark.parametrize("dtype", DTYPES)
def test_layout_pairs_shapes_host_and_device(dtype, ob, oracle_lib):
    import torch
    rng = np.random.default_rng(2)

    def dev(x):
        return torch.from_numpy(x).cuda() if x.flags["C_CONTIGUOUS"] else torch.from_numpy(np.ascontiguousarray(x.T)).cuda().t()

    for (M, N) in ((1, 1), (1, 70), (70, 1), (63, 65), (64, 64), (129, 300), (1000, 37), (2049, 2051)):
        for oa in "FC":
            for oc in "FC":
                a = np.asarray(rng.uniform(-1, 1, (M, N)).astype(dtype), order=oa)
                want = np.zeros((N, M), dtype, order=oc)
                oracle_lib.transpose(want, a)
                got = np.full((N, M), 7, dtype, order=oc)
                ob.transpose(got, a)()
                assert np.array_equal(got, want), ("host", M, N, oa, oc)
                tc = dev(np.full((N, M), 7, dtype, order=oc))
                ob.transpose(tc, dev(a))()
                torch.cuda.synchronize()
                assert np.array_equal(tc.cpu().numpy(), want), ("dev", M, N, oa, oc)


@pytest.mark.parametrize("dtype", DTYPES)
def test_strided_views_leave_parents_untouched(dtype, ob, oracle_lib):
    import torch
    rng = np.random.default_rng(3)
    big_a = rng.uniform(-1, 1, (300, 400)).astype(dtype)
    big_c = rng.uniform(-1, 1, (500, 350)).astype(dtype)
    sa, sc = np.s_[3:203, 7:307:3], np.s_[5:205:2, 10:210]        # a view 200x100 -> c view 100x200
    want = big_c.copy()
    oracle_lib.transpose(want[sc], big_a[sa])
    got = big_c.copy()
    ob.transpose(got[sc], big_a[sa])()
    assert np.array_equal(got, want)
    tc, ta = torch.from_numpy(big_c).cuda(), torch.from_numpy(big_a).cuda()
    ob.transpose(tc[5:205:2, 10:210], ta[3:203, 7:307:3])()
    torch.cuda.synchronize()
    assert np.array_equal(tc.cpu().numpy(), want)


@pytest.mark.parametrize("dtype", DTYPES)
def test_inplace_square_and_errors(dtype, ob, oracle_lib):
    import torch
    rng = np.random.default_rng(4)
    for n in (1, 2, 31, 32, 33, 64, 100, 1025):
        a = np.asfortranarray(rng.uniform(-1, 1, (n, n)).astype(dtype))
        want = a.copy(order="K")
        oracle_lib.transpose_inplace(want)
        b = a.copy(order="K")
        ob.transpose_inplace(b)()
        assert np.array_equal(b, want) and np.array_equal(b, a.T)
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        ob.transpose_inplace(t)()
        torch.cuda.synchronize()
        assert np.array_equal(t.cpu().numpy(), a.T)
    with pytest.raises(ob.B200Error) as ei:                       # non-square in place is refused, not scrambled
        ob.transpose_inplace(np.zeros((4, 6), dtype))()
    assert ei.value.code == 2
    with pytest.raises(RuntimeError, match="dimension mismatch"):  # trans.hpp:121-126
        ob.transpose(np.zeros((4, 5), dtype), np.zeros((4, 5), dtype))


def test_full_size_device(ob):
    """16384 x 16384 fp32 (1 GiB in, 1 GiB out), all four layout pairings, exact against torch."""
    import torch
    n = 16384
    a = torch.rand((n, n), device="cuda")
    for a_view in (a, a.t()):
        for c_first in (False, True):
            c = torch.zeros((n, n), device="cuda")
            cv = c.t() if c_first else c
            ob.transpose(cv, a_view)()
            torch.cuda.synchronize()
            assert torch.equal(cv, a_view.t())
    b = a.clone()
    ob.transpose_inplace(b)()
    torch.cuda.synchronize()
    assert torch.equal(b, a.t())


def test_cpp_front_end_trans(ob, tmp_path):
    exe = tmp_path / "test_trans"
    lib = ob.library_path().parent
    cmd = ["/usr/bin/g++", "-std=c++20", "-O2", f"-I{ROOT / 'include' / 'compat'}", f"-I{ROOT / 'include'}",
           str(ROOT / "tests" / "cpp" / "test_trans.cpp"), "-o", str(exe), f"-L{lib}", "-lb200mtm", f"-Wl,-rpath,{lib}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    print(r.stdout[-800:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
