"""TEST INFRASTRUCTURE — ctypes loader for the CPU oracle of the mtm path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the timed
CPU baseline.  The product (``openmp-blas_b200/``, ``include/``) must never import it.

Two libraries:

* ``liboracle_mtm.so``  — plain-C restatement (``oracle_mtm.c``), always present after
  ``make -C oracle``;
* ``_ref/libref_mtm_v{3,4}.so`` — the unmodified reference headers compiled behind C exports
  (``ref_mtm.cpp``); built only where ``/root/reference`` exists, shipped prebuilt elsewhere.

Matrices are numpy 2-D arrays; layout and sub-views are carried by the array's strides
(``order="F"`` = uBLAS ``first_order`` = column-major, ``order="C"`` = ``last_order``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_SIZE2 = C.c_size_t * 2


class OracleBlocks(C.Structure):
    _fields_ = [("mr", C.c_size_t), ("nr", C.c_size_t), ("kb", C.c_size_t),
                ("mb", C.c_size_t), ("nb", C.c_size_t)]


def build(quiet: bool = True) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", str(HERE)], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _desc(x: np.ndarray):
    assert x.ndim == 2, "matrices are 2-D"
    it = x.dtype.itemsize
    assert all(s % it == 0 and s >= 0 for s in x.strides)
    return _SIZE2(*x.shape), _SIZE2(*(s // it for s in x.strides))


def _ptr(x: np.ndarray):
    return C.c_void_p(x.ctypes.data)


def _ptr_flat(x: np.ndarray):
    assert x.flags["C_CONTIGUOUS"] or x.flags["F_CONTIGUOUS"], "vectors are contiguous"
    return C.c_void_p(x.ctypes.data)


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"mtm supports float32/float64 only, got {dtype}")


def _cpu_has_avx512() -> bool:
    try:
        flags = Path("/proc/cpuinfo").read_text()
    except OSError:
        return False
    return all(f in flags for f in (" avx512f", " avx512vl", " avx512bw", " avx512dq", " avx512cd"))


class Oracle:
    """The plain-C restatement (oracle_mtm.c)."""

    kind = "port"

    def __init__(self):
        path = HERE / "liboracle_mtm.so"
        if not path.exists():
            build()
        self.path = path
        self.lib = C.CDLL(str(path))
        for sfx in ("f32", "f64"):
            fn = getattr(self.lib, f"oracle_mtm_{sfx}")
            fn.restype = C.c_int
            fn.argtypes = [C.c_void_p, _SIZE2, _SIZE2, C.c_void_p, _SIZE2, _SIZE2,
                           C.c_void_p, _SIZE2, _SIZE2, C.POINTER(OracleBlocks)]
            pk = getattr(self.lib, f"oracle_pack_{sfx}")
            pk.restype = None
            pk.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, _SIZE2, C.c_size_t, C.c_size_t, C.c_int]
            for nm in ("mtv", "vtm"):
                fv = getattr(self.lib, f"oracle_{nm}_{sfx}")
                fv.restype = None
                fv.argtypes = [C.c_void_p, C.c_void_p, _SIZE2, _SIZE2, C.c_void_p, C.c_int, C.c_size_t]
            ft = getattr(self.lib, f"oracle_transpose_{sfx}")
            ft.restype = None
            ft.argtypes = [C.c_void_p, _SIZE2, C.c_void_p, _SIZE2, _SIZE2]
            fi = getattr(self.lib, f"oracle_transpose_inplace_{sfx}")
            fi.restype = None
            fi.argtypes = [C.c_void_p, C.c_size_t]
        self.lib.oracle_exact_i64.restype = None
        self.lib.oracle_exact_i64.argtypes = [C.c_void_p, _SIZE2, C.c_void_p, _SIZE2, _SIZE2,
                                              C.c_void_p, _SIZE2, _SIZE2]
        self.lib.oracle_default_blocks.argtypes = [C.c_int, C.c_int, C.POINTER(OracleBlocks)]
        self.lib.oracle_max_threads.restype = C.c_int

    def threads(self) -> int:
        return int(self.lib.oracle_max_threads())

    def mtm(self, c: np.ndarray, a: np.ndarray, b: np.ndarray, blocks=None) -> None:
        """In place ``c += a @ b`` following amt::mtm_helper (mtm.hpp:116-206)."""
        sfx = _sfx(c.dtype)
        assert a.dtype == c.dtype == b.dtype
        nc, wc = _desc(c)
        na, wa = _desc(a)
        nb, wb = _desc(b)
        blk = None
        if blocks is not None:
            blk = C.byref(OracleBlocks(*blocks))
        rc = getattr(self.lib, f"oracle_mtm_{sfx}")(_ptr(c), nc, wc, _ptr(a), na, wa,
                                                      _ptr(b), nb, wb, blk)
        if rc:
            raise RuntimeError(f"oracle_mtm_{sfx} failed with status {rc}")

    def pack(self, out: np.ndarray, wo: int, inp: np.ndarray, wi, m: int, n: int, trans: bool,
             in_offset: int = 0, out_offset: int = 0) -> None:
        sfx = _sfx(out.dtype)
        it = out.dtype.itemsize
        getattr(self.lib, f"oracle_pack_{sfx}")(
            C.c_void_p(out.ctypes.data + out_offset * it), wo,
            C.c_void_p(inp.ctypes.data + in_offset * it), _SIZE2(*wi), m, n, int(trans))

    def _mv(self, name: str, c: np.ndarray, a: np.ndarray, b: np.ndarray, a_last_order, kb: int) -> None:
        sfx = _sfx(c.dtype)
        assert a.dtype == b.dtype == c.dtype and a.ndim == 2
        na, wa = _desc(a)
        if a_last_order is None:     # layout tag from the strides: smaller stride along k -> last_order
            a_last_order = bool(wa[1] < wa[0])
        getattr(self.lib, f"oracle_{name}_{sfx}")(_ptr_flat(c), _ptr(a), na, wa, _ptr_flat(b), int(a_last_order), kb)

    def mtv(self, c, a, b, a_last_order=None, kb: int = 0) -> None:
        """amt::mtv (mtv.hpp:102-168): c (op)= a @ b; accumulates for a first_order ``a``, assigns for last_order."""
        self._mv("mtv", c, a, b, a_last_order, kb)

    def vtm(self, c, a, b, a_last_order=None, kb: int = 0) -> None:
        """amt::vtm (mtv.hpp:170-236): c (op)= b @ a; assigns for a first_order ``a``, accumulates for last_order."""
        self._mv("vtm", c, a, b, a_last_order, kb)

    def transpose(self, c: np.ndarray, a: np.ndarray) -> None:
        """amt::transpose(c, a) (trans.hpp:94-141): c(j, i) = a(i, j), any strides on both."""
        sfx = _sfx(c.dtype)
        assert c.dtype == a.dtype and c.shape == a.shape[::-1]
        _, wc = _desc(c)
        na, wa = _desc(a)
        getattr(self.lib, f"oracle_transpose_{sfx}")(_ptr(c), wc, _ptr(a), na, wa)

    def transpose_inplace(self, a: np.ndarray) -> None:
        """amt::transpose(a) (trans.hpp:143-168): square, contiguous storage."""
        sfx = _sfx(a.dtype)
        assert a.shape[0] == a.shape[1] and (a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"])
        getattr(self.lib, f"oracle_transpose_inplace_{sfx}")(_ptr(a), a.shape[0])

    def exact_i64(self, c: np.ndarray, a: np.ndarray, b: np.ndarray) -> None:
        assert c.dtype == a.dtype == b.dtype == np.int64
        _, wc = _desc(c)
        na, wa = _desc(a)
        nb, wb = _desc(b)
        self.lib.oracle_exact_i64(_ptr(c), wc, _ptr(a), na, wa, _ptr(b), nb, wb)

    def default_blocks(self, dtype, c_last_order: bool):
        blk = OracleBlocks()
        self.lib.oracle_default_blocks(int(np.dtype(dtype) == np.float64), int(c_last_order),
                                       C.byref(blk))
        return (blk.mr, blk.nr, blk.kb, blk.mb, blk.nb)


class Reference:
    """The unmodified reference compiled behind C exports (oracle/_ref, ref_mtm.cpp)."""

    kind = "reference"

    def __init__(self, isa: str | None = None):
        if isa is None:
            isa = os.environ.get("B200_MTM_REF_ISA") or ("v4" if _cpu_has_avx512() else "v3")
        path = HERE / "_ref" / f"libref_mtm_{isa}.so"
        if not path.exists() and Path("/root/reference/include").is_dir():
            build()
        if not path.exists():
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` where /root/reference exists")
        self.isa = isa
        self.path = path
        self.lib = C.CDLL(str(path))
        L = self.lib
        L.ref_last_error.restype = C.c_char_p
        L.ref_max_threads.restype = C.c_int
        for sfx in ("f32", "f64"):
            fn = getattr(L, f"ref_mtm_{sfx}")
            fn.restype = None
            fn.argtypes = [C.c_void_p, _SIZE2, _SIZE2, C.c_void_p, _SIZE2, _SIZE2,
                           C.c_void_p, _SIZE2, _SIZE2, C.c_int]
            bn = getattr(L, f"ref_mtm_bench_{sfx}")
            bn.restype = C.c_double
            bn.argtypes = [C.c_int] + fn.argtypes
            tn = getattr(L, f"ref_mtm_tensor_{sfx}")
            tn.restype = C.c_int
            tn.argtypes = [C.c_int] * 3 + [C.c_size_t] * 6 + [C.c_void_p] * 3
            pk = getattr(L, f"ref_pack_{sfx}")
            pk.restype = None
            pk.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, _SIZE2, C.c_size_t, C.c_size_t, C.c_int]
            mv = getattr(L, f"ref_mtv_tensor_{sfx}")
            mv.restype = C.c_int
            mv.argtypes = [C.c_int, C.c_int] + [C.c_size_t] * 4 + [C.c_void_p] * 3
            tt = getattr(L, f"ref_transpose_tensor_{sfx}")
            tt.restype = C.c_int
            tt.argtypes = [C.c_int, C.c_int] + [C.c_size_t] * 4 + [C.c_void_p] * 2
            ti = getattr(L, f"ref_transpose_inplace_{sfx}")
            ti.restype = C.c_int
            ti.argtypes = [C.c_size_t, C.c_void_p]
        L.ref_block_sizes.restype = None
        L.ref_block_sizes.argtypes = [C.c_int, C.c_int, C.c_size_t * 5]

    def threads(self) -> int:
        return int(self.lib.ref_max_threads())

    @staticmethod
    def _c_last_order(c: np.ndarray) -> int:
        it = c.dtype.itemsize
        s0, s1 = (s // it for s in c.strides)
        if s1 == 1 and s0 != 1:
            return 1
        if s0 == 1 and s1 != 1:
            return 0
        return int(c.flags["C_CONTIGUOUS"] and not c.flags["F_CONTIGUOUS"])

    def mtm(self, c: np.ndarray, a: np.ndarray, b: np.ndarray) -> None:
        """In place ``c += a @ b`` through amt::mtm_helper (mtm.hpp:116-206)."""
        sfx = _sfx(c.dtype)
        assert a.dtype == c.dtype == b.dtype
        nc, wc = _desc(c)
        na, wa = _desc(a)
        nb, wb = _desc(b)
        getattr(self.lib, f"ref_mtm_{sfx}")(_ptr(c), nc, wc, _ptr(a), na, wa, _ptr(b), nb, wb,
                                            self._c_last_order(c))

    def bench_ns(self, iters: int, c: np.ndarray, a: np.ndarray, b: np.ndarray) -> float:
        """Mean ns/call over ``iters`` back-to-back calls (amt::benchmark, benchmark.hpp:34-52)."""
        sfx = _sfx(c.dtype)
        nc, wc = _desc(c)
        na, wa = _desc(a)
        nb, wb = _desc(b)
        return float(getattr(self.lib, f"ref_mtm_bench_{sfx}")(
            iters, _ptr(c), nc, wc, _ptr(a), na, wa, _ptr(b), nb, wb, self._c_last_order(c)))

    def mtm_tensor(self, layouts: str, a: np.ndarray, b: np.ndarray, c: np.ndarray,
                   c_shape=None) -> None:
        """Full front-end amt::mtm(c,a,b,nullopt)() on freshly built tensors.

        ``layouts`` is the reference's (C, A, B) tag triple, e.g. "FLF"; ``a``/``b``/``c`` are
        flat storage in that layout.  Raises RuntimeError with the reference's message on
        the validation throws (mtm.hpp:234-250).
        """
        sfx = _sfx(c.dtype)
        lc, la, lb = (int(ch == "L") for ch in layouts)
        (M, Ka), (Kb, N) = a.shape, b.shape
        Mc, Nc = c_shape if c_shape is not None else c.shape
        rc = getattr(self.lib, f"ref_mtm_tensor_{sfx}")(lc, la, lb, M, N, Ka, Kb, Mc, Nc,
                                                        _ptr(a), _ptr(b), _ptr(c))
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())

    def transpose_tensor(self, c: np.ndarray, a: np.ndarray, c_shape=None) -> None:
        """amt::transpose(c, a, nullopt)() on fresh tensors; contiguous arrays, order gives the layout."""
        sfx = _sfx(c.dtype)
        lc = int(c.flags["C_CONTIGUOUS"] and not c.flags["F_CONTIGUOUS"])
        la = int(a.flags["C_CONTIGUOUS"] and not a.flags["F_CONTIGUOUS"])
        Mc, Nc = c_shape if c_shape is not None else c.shape
        rc = getattr(self.lib, f"ref_transpose_tensor_{sfx}")(lc, la, a.shape[0], a.shape[1], Mc, Nc, _ptr(a), _ptr(c))
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())

    def transpose_inplace(self, a: np.ndarray) -> None:
        """amt::transpose(a, nullopt)() on a fresh n x n first_order tensor holding a's storage."""
        sfx = _sfx(a.dtype)
        assert a.shape[0] == a.shape[1]
        rc = getattr(self.lib, f"ref_transpose_inplace_{sfx}")(a.shape[0], _ptr(a))
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())

    def mtv_tensor(self, is_vtm: bool, a: np.ndarray, b: np.ndarray, c: np.ndarray, nb_len=None, nc_len=None) -> None:
        """amt::mtv / amt::vtm through the reference front-end on fresh tensors (test/test.mtv.cpp:36-65).
        ``a`` is a contiguous 2-D array whose order gives the layout; ``b``, ``c`` are 1-D."""
        sfx = _sfx(c.dtype)
        last = bool(a.flags["C_CONTIGUOUS"] and not a.flags["F_CONTIGUOUS"])
        rc = getattr(self.lib, f"ref_mtv_tensor_{sfx}")(int(is_vtm), int(last), a.shape[0], a.shape[1],
                                                        b.size if nb_len is None else nb_len,
                                                        c.size if nc_len is None else nc_len,
                                                        _ptr(a), _ptr_flat(b), _ptr_flat(c))
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())

    def pack(self, out: np.ndarray, wo: int, inp: np.ndarray, wi, m: int, n: int, trans: bool,
             in_offset: int = 0, out_offset: int = 0) -> None:
        sfx = _sfx(out.dtype)
        it = out.dtype.itemsize
        getattr(self.lib, f"ref_pack_{sfx}")(
            C.c_void_p(out.ctypes.data + out_offset * it), wo,
            C.c_void_p(inp.ctypes.data + in_offset * it), _SIZE2(*wi), m, n, int(trans))

    def block_sizes(self, dtype, c_last_order: bool):
        out = (C.c_size_t * 5)()
        self.lib.ref_block_sizes(int(np.dtype(dtype) == np.float64), int(c_last_order), out)
        return tuple(int(v) for v in out)


def have_reference() -> bool:
    return any((HERE / "_ref").glob("libref_mtm_*.so")) or Path("/root/reference/include").is_dir()
