"""GPU parity tests of the matrix-times-vector path (b200_mtv_* C ABI through the Python / C++ host
mirrors) against the CPU oracle: the reference's own test cases (test/test.mtv.cpp, test/test.vtm.cpp:
first/last order x {f32,f64} x sizes 2..511 and the 32Ki case) bit-exact on their integer inputs;
tolerance on floating-point data; strided views; accumulate/assign semantics."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
DTYPES = [np.float32, np.float64]


def _fns(ob, oracle_lib, is_vtm):
    return (ob.vtm, oracle_lib.vtm) if is_vtm else (ob.mtv, oracle_lib.mtv)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("order", ["F", "C"])
@pytest.mark.parametrize("is_vtm", [False, True])
def test_reference_range_cases_host_entry(is_vtm, order, dtype, ob, oracle_lib):
    """Range[2, 512) of test/test.mtv.cpp:29-126 / test/test.vtm.cpp:29-125, non-zero start vector."""
    gpu, cpu = _fns(ob, oracle_lib, is_vtm)
    rng = np.random.default_rng(abs(hash((is_vtm, order, np.dtype(dtype).name, "gpu"))) % 2**32)
    for sz in list(range(2, 70)) + list(range(70, 512, 7)) + [255, 256, 257, 511]:
        a = np.asarray(rng.integers(0, 100, (sz, sz)).astype(dtype), order=order)
        v = rng.integers(0, 100, sz).astype(dtype)
        c0 = rng.integers(0, 100, sz).astype(dtype)
        want, got = c0.copy(), c0.copy()
        cpu(want, a, v)
        gpu(got, a, v)()
        assert np.array_equal(got, want), (is_vtm, order, sz)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("order", ["F", "C"])
@pytest.mark.parametrize("is_vtm", [False, True])
def test_rectangular_and_device_entry(is_vtm, order, dtype, ob, oracle_lib):
    import torch
    gpu, cpu = _fns(ob, oracle_lib, is_vtm)
    rng = np.random.default_rng(3)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    for (rows, cols) in ((1, 1), (1, 300), (300, 1), (37, 1025), (1025, 37), (4099, 263), (130, 20000), (20000, 130),
                         (3, 70001)):
        a = np.asarray(rng.integers(0, 10, (rows, cols)).astype(dtype), order=order)
        nb, nc = (rows, cols) if is_vtm else (cols, rows)
        v = rng.integers(0, 10, nb).astype(dtype)
        c0 = rng.integers(0, 10, nc).astype(dtype)
        want = c0.copy()
        cpu(want, a, v)
        got = c0.copy()
        gpu(got, a, v)()
        assert np.array_equal(got, want), ("host", is_vtm, order, rows, cols)
        ta = torch.from_numpy(a).cuda() if order == "C" else torch.from_numpy(np.ascontiguousarray(a.T)).cuda().t()
        tv, tc = torch.from_numpy(v).cuda(), torch.from_numpy(c0).cuda()
        assert ta.dtype == tdt
        gpu(tc, ta, tv)()
        torch.cuda.synchronize()
        assert np.array_equal(tc.cpu().numpy(), want), ("dev", is_vtm, order, rows, cols, ob.last_choice())


@pytest.mark.parametrize("dtype", DTYPES)
def test_strided_matrix_views(dtype, ob, oracle_lib):
    """Sub-views with a leading dimension larger than the extent, mis-aligned starts, and views with
    no unit stride at all."""
    import torch
    rng = np.random.default_rng(9)
    big = rng.integers(0, 10, (700, 900)).astype(dtype)
    tbig = torch.from_numpy(big).cuda()
    for sl in (np.s_[5:505, 3:603], np.s_[1:400:3, 2:800:2], np.s_[8:520, 16:528], np.s_[0:700:7, 1:900:9]):
        a = big[sl]
        for is_vtm in (False, True):
            gpu, cpu = _fns(ob, oracle_lib, is_vtm)
            nb, nc = (a.shape[0], a.shape[1]) if is_vtm else (a.shape[1], a.shape[0])
            v = rng.integers(0, 10, nb).astype(dtype)
            c0 = rng.integers(0, 10, nc).astype(dtype)
            want = c0.copy()
            cpu(want, a, v, a_last_order=True)           # row-major parent: the last_order tag
            got = c0.copy()
            gpu(got, a, v, layout="L")()
            assert np.array_equal(got, want), ("host", sl, is_vtm)
            tc = torch.from_numpy(c0).cuda()
            gpu(tc, tbig[sl], torch.from_numpy(v).cuda(), layout="L")()
            torch.cuda.synchronize()
            assert np.array_equal(tc.cpu().numpy(), want), ("dev", sl, is_vtm)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("order", ["F", "C"])
@pytest.mark.parametrize("is_vtm", [False, True])
def test_tolerance_uniform(is_vtm, order, dtype, ob):
    gpu = ob.vtm if is_vtm else ob.mtv
    rng = np.random.default_rng(21)
    u = np.finfo(dtype).eps / 2
    for (rows, cols) in ((2000, 3000), (64, 50000), (50000, 64)):
        a = np.asarray(rng.uniform(-1, 1, (rows, cols)).astype(dtype), order=order)
        nb, nc = (rows, cols) if is_vtm else (cols, rows)
        v = rng.uniform(-1, 1, nb).astype(dtype)
        c0 = rng.uniform(-1, 1, nc).astype(dtype)
        accumulates = (order == "F") != is_vtm
        al, vl = a.astype(np.longdouble), v.astype(np.longdouble)
        exact = (vl @ al if is_vtm else al @ vl) + (c0.astype(np.longdouble) if accumulates else 0)
        absab = (np.abs(vl) @ np.abs(al)) if is_vtm else (np.abs(al) @ np.abs(vl))
        got = c0.copy()
        gpu(got, a, v)()
        ratio = float(np.max(np.abs(got.astype(np.longdouble) - exact) / ((nb + 1) * u * (absab + np.abs(c0)) + 1e-300)))
        print(f"\n[mtv tolerance] vtm={is_vtm} {order} {np.dtype(dtype).name} {rows}x{cols}: ratio {ratio:.4f} (limit 2)")
        assert ratio <= 2.0


@pytest.mark.parametrize("order", ["F", "C"])
@pytest.mark.parametrize("is_vtm", [False, True])
def test_reference_32k_case(is_vtm, order, ob):
    """test/test.mtv.cpp:128-210 / test/test.vtm.cpp:127-209: the 32Ki x 32Ki case (fp32, 4 GiB,
    device-resident).  Integers in [0,9] keep every partial sum exact in fp32; checked against fp64."""
    import torch
    sz = 32 * 1024
    g = torch.Generator(device="cuda").manual_seed(4)
    a = torch.randint(0, 10, (sz, sz), device="cuda", generator=g, dtype=torch.int32).float()
    if order == "F":
        a = a.t()                        # same storage viewed column-major
    v = torch.randint(0, 10, (sz,), device="cuda", generator=g).float()
    c0 = torch.randint(0, 10, (sz,), device="cuda", generator=g).float()
    c = c0.clone()
    (ob.vtm if is_vtm else ob.mtv)(c, a, v)()
    torch.cuda.synchronize()
    accumulates = (order == "F") != is_vtm
    want = torch.zeros(sz, device="cuda", dtype=torch.float64)
    for r0 in range(0, sz, 4096):        # fp64 reference in slabs (an fp64 copy of A would be 8 GiB)
        if is_vtm:
            want += v[r0:r0 + 4096].double() @ a[r0:r0 + 4096].double()
        else:
            want[r0:r0 + 4096] = a[r0:r0 + 4096].double() @ v.double()
    if accumulates:
        want += c0.double()
    assert torch.equal(c.double(), want), ob.last_choice()


def test_cpp_front_end_mtv(ob, tmp_path):
    exe = tmp_path / "test_mtv"
    lib = ob.library_path().parent
    cmd = ["/usr/bin/g++", "-std=c++20", "-O2", f"-I{ROOT / 'include' / 'compat'}", f"-I{ROOT / 'include'}",
           str(ROOT / "tests" / "cpp" / "test_mtv.cpp"), "-o", str(exe), f"-L{lib}", "-lb200mtm", f"-Wl,-rpath,{lib}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900)
    print(r.stdout[-1500:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
