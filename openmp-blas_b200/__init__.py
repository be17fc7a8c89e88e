"""B200-native matrix-times-matrix (``mtm``) — Python host mirror of the reference interface.

This package is a thin binding over ``libb200mtm.so`` (C ABI in ``include/b200_mtm.h``, CUDA
for sm_100a in ``csrc/``).  It mirrors the reference's operator interface for the mtm path:

    reference (C++):  auto fn = amt::mtm(c, a, b, std::nullopt);  fn();      include/mtm.hpp:208-267
    here (Python):    fn = mtm(c, a, b, None);                     fn()

* validation happens when ``mtm(...)`` is called, the work when the returned callable is invoked;
* each invocation ACCUMULATES: ``c += a @ b`` (simd_loop.hpp:169,187);
* layouts are carried by strides: ``order="F"`` == uBLAS ``first_order`` (column-major, the
  reference default, utils.hpp:21), ``order="C"`` == ``last_order`` (row-major);
* operands are numpy arrays (host path: staged to the GPU and back, synchronous) or torch CUDA
  tensors (device path: asynchronous on torch's current stream).

There is no CPU implementation here and no fallback: if the CUDA library is missing or no
B200 is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Callable, Optional

__all__ = [
    "mtm", "mtv", "vtm", "transpose", "transpose_inplace", "make_tensor", "lib", "library_path", "last_choice", "launch_count", "device_info",
    "num_configs", "config_name", "flags", "B200Error", "VARIANTS", "pinned_empty", "replicate_push", "replicate_push_2d", "flag_wait", "flag_signal",
]

HERE = Path(__file__).resolve().parent
_SIZE2 = C.c_size_t * 2

VARIANTS = {"auto": 0, "simt": 1, "3xtf32": 2, "dfma": 3, "dmma": 4}
_ERR_NAMES = {1: "invalid argument", 2: "dimension mismatch", 3: "layout", 4: "CUDA", 5: "out of memory"}

# Messages of the reference's two validation throws (include/mtm.hpp:234-250), kept verbatim so
# callers matching on them keep working (the text says "amt::mtv": a copy/paste in the reference).
_MSG_NOT_MATRIX = ("amt::mtv(boost::numeric::ublas::tensor_core<Out>& c, "
                   "boost::numeric::ublas::tensor_core<E1> const& a, "
                   "boost::numeric::ublas::tensor_core<E2> const& b) : "
                   "a, b, and c must be the matrices")
_MSG_DIM = ("amt::mtv(boost::numeric::ublas::tensor_core<Out>&, "
            "boost::numeric::ublas::tensor_core<E1> const&, "
            "boost::numeric::ublas::tensor_core<E2> const&) : "
            "dimension mismatch")


class B200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libb200mtm: {_ERR_NAMES.get(code, code)}: {message}")
        self.code = code


class _Choice(C.Structure):
    _fields_ = [("variant", C.c_int), ("config", C.c_int), ("launches", C.c_int),
                ("a_mode", C.c_int), ("b_mode", C.c_int), ("name", C.c_char * 64)]


class _DeviceInfo(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("cc_major", C.c_int), ("cc_minor", C.c_int),
                ("sm_count", C.c_int), ("sm_clock_khz", C.c_int), ("mem_clock_khz", C.c_int),
                ("mem_bus_bits", C.c_int), ("smem_per_sm", C.c_size_t),
                ("smem_per_block_optin", C.c_size_t), ("l2_bytes", C.c_size_t),
                ("hbm_bytes", C.c_size_t), ("peak_fp32_tflops", C.c_double),
                ("peak_fp64_tflops", C.c_double)]


_lib = None


def library_path() -> Path:
    return HERE / "libb200mtm.so"


def lib() -> C.CDLL:
    """Load libb200mtm.so.  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise ImportError(f"{path} not found: build it with `python openmp-blas_b200/build.py` "
                          "(or __graft_entry__.build()); the mtm path has no CPU fallback")
    L = C.CDLL(str(path))
    mat = [C.c_void_p, _SIZE2, _SIZE2]
    for sfx in ("f32", "f64"):
        fn = getattr(L, f"b200_mtm_{sfx}")
        fn.restype = C.c_int
        fn.argtypes = mat * 3 + [C.c_int]
        fm = getattr(L, f"b200_mtm_{sfx}_mgpu")
        fm.restype = C.c_int
        fm.argtypes = mat * 3 + [C.c_int, C.POINTER(C.c_int), C.c_int]
        fd = getattr(L, f"b200_mtm_{sfx}_dev")
        fd.restype = C.c_int
        fd.argtypes = mat * 3 + [C.c_int, C.c_void_p]
        fb = getattr(L, f"b200_mtm_bench_{sfx}_dev")
        fb.restype = C.c_int
        fb.argtypes = mat * 3 + [C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
    for sfx in ("f32", "f64"):
        mv = [C.c_void_p, C.c_void_p, _SIZE2, _SIZE2, C.c_void_p, C.c_int, C.c_int]
        getattr(L, f"b200_mtv_{sfx}").restype = C.c_int
        getattr(L, f"b200_mtv_{sfx}").argtypes = mv
        getattr(L, f"b200_mtv_{sfx}_dev").restype = C.c_int
        getattr(L, f"b200_mtv_{sfx}_dev").argtypes = mv + [C.c_void_p]
        getattr(L, f"b200_mtv_bench_{sfx}_dev").restype = C.c_int
        getattr(L, f"b200_mtv_bench_{sfx}_dev").argtypes = mv + [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
    for sfx in ("f32", "f64"):
        tr = [C.c_void_p, _SIZE2, _SIZE2, C.c_void_p, _SIZE2, _SIZE2, C.c_int]
        getattr(L, f"b200_transpose_{sfx}").argtypes = tr
        getattr(L, f"b200_transpose_{sfx}_dev").argtypes = tr + [C.c_void_p]
        getattr(L, f"b200_transpose_bench_{sfx}_dev").argtypes = tr + [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        getattr(L, f"b200_transpose_inplace_{sfx}").argtypes = [C.c_void_p, _SIZE2, C.c_int]
        getattr(L, f"b200_transpose_inplace_{sfx}_dev").argtypes = [C.c_void_p, _SIZE2, C.c_int, C.c_void_p]
    L.b200_last_error.restype = C.c_char_p
    L.b200_launch_count.restype = C.c_uint64
    L.b200_last_bench_enqueue_us.restype = C.c_double
    L.b200_mtm_last_choice.argtypes = [C.POINTER(_Choice)]
    L.b200_mtm_plan_f32.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.POINTER(_Choice)]
    L.b200_mtm_num_configs.argtypes = [C.c_int, C.c_int]
    L.b200_mtm_config_name.argtypes = [C.c_int, C.c_int, C.c_int]
    L.b200_mtm_config_name.restype = C.c_char_p
    L.b200_device_count.argtypes = [C.POINTER(C.c_int)]
    L.b200_get_device_info.argtypes = [C.c_int, C.POINTER(_DeviceInfo)]
    L.b200_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.b200_host_free.argtypes = [C.c_void_p]
    L.b200_malloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.b200_free.argtypes = [C.c_void_p]
    L.b200_replicate_push.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                      C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_void_p]
    L.b200_replicate_push_2d.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t,
                                         C.c_size_t, C.c_size_t, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint32,
                                         C.c_int, C.c_void_p]
    L.b200_flag_wait.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.b200_flag_signal.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise B200Error(rc, lib().b200_last_error().decode(errors="replace"))


def flags(variant="auto", config: Optional[int] = None, reserve_sms: int = 0, split_k: int = 0) -> int:
    v = VARIANTS[variant] if isinstance(variant, str) else int(variant)
    return (v | ((0 if config is None else int(config) + 1) << 8) | ((int(reserve_sms) & 0xff) << 16)
            | ((int(split_k) & 0x7f) << 24))


# ---- operand description -------------------------------------------------------------------------
def _is_torch(x) -> bool:
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda")


def _describe(x, what: str):
    """-> (pointer, extents, element strides, dtype suffix, on_device)"""
    if _is_torch(x):
        import torch
        if x.dim() != 2:
            raise RuntimeError(_MSG_NOT_MATRIX)
        sfx = {torch.float32: "f32", torch.float64: "f64"}.get(x.dtype)
        if sfx is None:
            raise TypeError(f"{what}: mtm supports float32/float64 only, got {x.dtype}")
        if not x.is_cuda:
            raise TypeError(f"{what}: torch operands must live on the GPU (use numpy arrays for the host path)")
        if any(s < 0 for s in x.stride()):
            raise ValueError(f"{what}: negative strides are not supported")
        return x.data_ptr(), tuple(x.shape), tuple(x.stride()), sfx, True
    import numpy as np
    if not isinstance(x, np.ndarray):
        raise TypeError(f"{what}: expected a numpy array or a torch CUDA tensor, got {type(x).__name__}")
    if x.ndim != 2:
        raise RuntimeError(_MSG_NOT_MATRIX)
    sfx = {"float32": "f32", "float64": "f64"}.get(x.dtype.name)
    if sfx is None:
        raise TypeError(f"{what}: mtm supports float32/float64 only, got {x.dtype}")
    it = x.dtype.itemsize
    if any(s < 0 or s % it for s in x.strides):
        raise ValueError(f"{what}: strides must be non-negative multiples of the element size")
    return x.ctypes.data, tuple(x.shape), tuple(s // it for s in x.strides), sfx, False


def mtm(c, a, b, num_threads: Optional[int] = None, *, variant="auto", config: Optional[int] = None,
        stream=None, reserve_sms: int = 0, devices=None, split_k: int = 0) -> Callable[[], None]:
    """Mirror of ``amt::mtm(c, a, b, num_threads)`` (include/mtm.hpp:208-267).

    Validates now (raising ``RuntimeError`` with the reference's messages), returns a nullary
    callable; every call performs ``c += a @ b`` on the B200.  ``num_threads`` is accepted for
    signature compatibility and ignored (the reference only ever raises the thread count to
    the maximum, thread_utils.hpp:47-56).  The callable borrows the operands' storage.
    ``devices`` (host arrays only): spread the call over several GPUs of the box through the multi-GPU C entry
    (``b200_mtm_*_mgpu``) — an int (that many of the visible devices, 0 = all) or a list of device indices.
    """
    del num_threads
    L = lib()
    pc, nc, wc, tc, dc = _describe(c, "c")
    pa, na, wa, ta, da = _describe(a, "a")
    pb, nb, wb, tb, db = _describe(b, "b")
    if not (ta == tb == tc):
        raise TypeError("both tensor type and result type must be of same value_type")  # mtm.hpp:224-228
    if not (dc == da == db):
        raise TypeError("c, a and b must all be host arrays or all be CUDA tensors")
    if not all(e >= 1 for e in (*nc, *na, *nb)):
        raise RuntimeError(_MSG_NOT_MATRIX)
    if not (na[0] == nc[0] and na[1] == nb[0] and nc[1] == nb[1]):
        raise RuntimeError(_MSG_DIM)
    if not _is_torch(c) and not c.flags.writeable:
        raise ValueError("c must be writeable")
    fl = flags(variant, config, reserve_sms, split_k)
    args = (C.c_void_p(pc), _SIZE2(*nc), _SIZE2(*wc), C.c_void_p(pa), _SIZE2(*na), _SIZE2(*wa),
            C.c_void_p(pb), _SIZE2(*nb), _SIZE2(*wb), fl)
    keep = (c, a, b)
    if dc:
        if devices is not None:
            raise TypeError("devices= applies to host arrays (device-resident operands live on one GPU)")
        fn = getattr(L, f"b200_mtm_{tc}_dev")

        def run_device() -> None:
            import torch
            _ = keep
            st = stream if stream is not None else torch.cuda.current_stream(c.device).cuda_stream
            with torch.cuda.device(c.device):
                _check(fn(*args, C.c_void_p(st)))
        return run_device
    if devices is not None:
        fn = getattr(L, f"b200_mtm_{tc}_mgpu")
        if isinstance(devices, int):
            dev_arr, n_dev = None, int(devices)
        else:
            devs = [int(d) for d in devices]
            dev_arr, n_dev = (C.c_int * len(devs))(*devs), len(devs)

        def run_mgpu() -> None:
            _ = keep
            _check(fn(*args, dev_arr, n_dev))
        return run_mgpu
    fn = getattr(L, f"b200_mtm_{tc}")

    def run_host() -> None:
        _ = keep
        _check(fn(*args))
    return run_host


_MSG_MTV_SHAPE = ("amt::mtv(boost::numeric::ublas::tensor_core<Out>& c, "
                  "boost::numeric::ublas::tensor_core<E1> const& a, "
                  "boost::numeric::ublas::tensor_core<E2> const& b) : "
                  "c and b must be vector, and a must be a matrix")


def _vec_describe(x, what: str):
    """1-D (or 1 x n / n x 1) contiguous vector -> (pointer, length, dtype suffix, on_device)."""
    if _is_torch(x):
        import torch
        sfx = {torch.float32: "f32", torch.float64: "f64"}.get(x.dtype)
        if sfx is None or not x.is_cuda:
            raise TypeError(f"{what}: expected a float32/float64 CUDA tensor")
        if x.dim() not in (1, 2) or (x.dim() == 2 and 1 not in x.shape) or not x.is_contiguous():
            raise RuntimeError(_MSG_MTV_SHAPE)
        return x.data_ptr(), x.numel(), sfx, True
    import numpy as np
    if not isinstance(x, np.ndarray):
        raise TypeError(f"{what}: expected a numpy array or a torch CUDA tensor")
    sfx = {"float32": "f32", "float64": "f64"}.get(x.dtype.name)
    if sfx is None:
        raise TypeError(f"{what}: mtv supports float32/float64 only, got {x.dtype}")
    if x.ndim not in (1, 2) or (x.ndim == 2 and 1 not in x.shape) or not (x.flags["C_CONTIGUOUS"] or x.flags["F_CONTIGUOUS"]):
        raise RuntimeError(_MSG_MTV_SHAPE)
    return x.ctypes.data, x.size, sfx, False


def _mtv_common(is_vtm: bool, c, a, b, stream, bench=None, layout=None):
    L = lib()
    pa, na, wa, ta, da = _describe(a, "a")
    pb, nb_len, tb, db = _vec_describe(b, "b")
    pc, nc_len, tc, dc = _vec_describe(c, "c")
    if not (ta == tb == tc):
        raise TypeError("both tensor type and result type must be of same value_type")
    if not (da == db == dc):
        raise TypeError("c, a and b must all be host arrays or all be CUDA tensors")
    if (is_vtm and (na[1] != nc_len or na[0] != nb_len)) or (not is_vtm and (na[1] != nb_len or na[0] != nc_len)):
        raise RuntimeError(_MSG_DIM)
    # The reference takes the layout as a template argument; here it is `layout` ("F" first_order /
    # "L" last_order) or, when omitted, read off the strides: smaller stride along k -> last_order.
    if layout is None:
        last_order = int(wa[1] < wa[0])
    else:
        last_order = int(str(layout).upper() in ("L", "C", "LAST_ORDER"))
    ext, strides = na, wa
    if is_vtm:   # c = b @ a == a^T b with the other layout's path (mtv.hpp:206-236)
        ext, strides, last_order = (na[1], na[0]), (wa[1], wa[0]), 1 - last_order
    args = (C.c_void_p(pc), C.c_void_p(pa), _SIZE2(*ext), _SIZE2(*strides), C.c_void_p(pb), last_order, 0)
    keep = (c, a, b)
    if bench is not None:
        import torch
        st = stream if stream is not None else torch.cuda.current_stream(a.device).cuda_stream
        out = C.c_double(0.0)
        with torch.cuda.device(a.device):
            _check(getattr(L, f"b200_mtv_bench_{ta}_dev")(*args, C.c_void_p(st), bench[0], bench[1], C.byref(out)))
        return out.value
    if da:
        fn = getattr(L, f"b200_mtv_{ta}_dev")

        def run_device() -> None:
            import torch
            _ = keep
            st = stream if stream is not None else torch.cuda.current_stream(a.device).cuda_stream
            with torch.cuda.device(a.device):
                _check(fn(*args, C.c_void_p(st)))
        return run_device
    fn = getattr(L, f"b200_mtv_{ta}")

    def run_host() -> None:
        _ = keep
        _check(fn(*args))
    return run_host


def mtv(c, a, b, num_threads: Optional[int] = None, *, layout=None, stream=None) -> Callable[[], None]:
    """Mirror of ``amt::mtv(c, a, b, num_threads)`` (include/mtv.hpp:102-168): ``c (op)= a @ b``.
    As in the reference, a first_order (column-major) ``a`` ACCUMULATES into ``c`` and a last_order
    (row-major) ``a`` ASSIGNS."""
    del num_threads
    return _mtv_common(False, c, a, b, stream, layout=layout)


def vtm(c, a, b, num_threads: Optional[int] = None, *, layout=None, stream=None) -> Callable[[], None]:
    """Mirror of ``amt::vtm(c, a, b, num_threads)`` (include/mtv.hpp:170-236): ``c (op)= b @ a``,
    computed as mtv on the transposed view with the other layout's path: a first_order ``a`` ASSIGNS,
    a last_order ``a`` ACCUMULATES."""
    del num_threads
    return _mtv_common(True, c, a, b, stream, layout=layout)


def bench_mtv_device(c, a, b, *, is_vtm=False, warmup=3, iters=10, stream=None) -> float:
    """Mean ms per mtv/vtm call with device-resident operands (CUDA events on the launching stream)."""
    return _mtv_common(is_vtm, c, a, b, stream, bench=(warmup, iters))


_MSG_TRANS_DIM = ("amt::transpose(boost::numeric::ublas::tensor_core<Out>& c, "
                  "boost::numeric::ublas::tensor_core<E> const& a) : dimension mismatch")


def transpose(c, a, num_threads: Optional[int] = None, *, stream=None, _bench=None):
    """Mirror of ``amt::transpose(c, a, num_threads)`` (include/trans.hpp:94-141): returns a callable
    performing ``c = a.T`` (out of place; any strides on both operands)."""
    del num_threads
    L = lib()
    pc, nc, wc, tc, dc = _describe(c, "c")
    pa, na, wa, ta, da = _describe(a, "a")
    if ta != tc:
        raise TypeError("input value type and result value type must be of same value type")   # trans.hpp:106-109
    if dc != da:
        raise TypeError("c and a must both be host arrays or both be CUDA tensors")
    if not (na[0] == nc[1] and na[1] == nc[0]):
        raise RuntimeError(_MSG_TRANS_DIM)
    args = (C.c_void_p(pc), _SIZE2(*nc), _SIZE2(*wc), C.c_void_p(pa), _SIZE2(*na), _SIZE2(*wa), 0)
    keep = (c, a)
    if _bench is not None:
        import torch
        st = stream if stream is not None else torch.cuda.current_stream(a.device).cuda_stream
        out = C.c_double(0.0)
        with torch.cuda.device(a.device):
            _check(getattr(L, f"b200_transpose_bench_{ta}_dev")(*args, C.c_void_p(st), _bench[0], _bench[1], C.byref(out)))
        return out.value
    if dc:
        fn = getattr(L, f"b200_transpose_{ta}_dev")

        def run_device() -> None:
            import torch
            _ = keep
            st = stream if stream is not None else torch.cuda.current_stream(a.device).cuda_stream
            with torch.cuda.device(a.device):
                _check(fn(*args, C.c_void_p(st)))
        return run_device
    fn = getattr(L, f"b200_transpose_{ta}")

    def run_host() -> None:
        _ = keep
        _check(fn(*args))
    return run_host


def transpose_inplace(a, num_threads: Optional[int] = None, *, stream=None):
    """Mirror of ``amt::transpose(a, num_threads)`` (include/trans.hpp:143-168): in place, square,
    contiguous storage (transposing the storage of a contiguous square matrix is layout-agnostic)."""
    del num_threads
    L = lib()
    pa, na, wa, ta, da = _describe(a, "a")
    if na[0] * na[1] and not ((wa[0] == 1 and wa[1] == na[0]) or (wa[1] == 1 and wa[0] == na[1]) or na[0] == 1):
        raise ValueError("transpose_inplace needs contiguous storage")
    args = (C.c_void_p(pa), _SIZE2(*na), 0)
    if da:
        fn = getattr(L, f"b200_transpose_inplace_{ta}_dev")

        def run_device() -> None:
            import torch
            st = stream if stream is not None else torch.cuda.current_stream(a.device).cuda_stream
            with torch.cuda.device(a.device):
                _check(fn(*args, C.c_void_p(st)))
        return run_device
    fn = getattr(L, f"b200_transpose_inplace_{ta}")

    def run_host() -> None:
        _check(fn(*args))
    return run_host


def bench_transpose_device(c, a, *, warmup=3, iters=10, stream=None) -> float:
    """Mean ms per out-of-place transpose with device-resident operands."""
    return transpose(c, a, stream=stream, _bench=(warmup, iters))


def bench_device(c, a, b, *, variant="auto", config=None, warmup=3, iters=10, stream=None, split_k=0) -> float:
    """Mean ms per ``c += a @ b`` over ``iters`` back-to-back device calls (CUDA events on the
    launching stream) — the device-side counterpart of amt::benchmark (benchmark.hpp:34-52)."""
    import torch
    L = lib()
    pc, nc, wc, tc, dc = _describe(c, "c")
    pa, na, wa, ta, _ = _describe(a, "a")
    pb, nb, wb, tb, _ = _describe(b, "b")
    assert dc and ta == tb == tc
    st = stream if stream is not None else torch.cuda.current_stream(c.device).cuda_stream
    out = C.c_double(0.0)
    with torch.cuda.device(c.device):
        _check(getattr(L, f"b200_mtm_bench_{tc}_dev")(
            C.c_void_p(pc), _SIZE2(*nc), _SIZE2(*wc), C.c_void_p(pa), _SIZE2(*na), _SIZE2(*wa),
            C.c_void_p(pb), _SIZE2(*nb), _SIZE2(*wb), flags(variant, config, 0, split_k), C.c_void_p(st),
            warmup, iters, C.byref(out)))
    return out.value


def make_tensor(dtype, M: int, N: int, layout: str = "F", val=None, device=None):
    """Mirror of ``amt::make_tensor<T, L>(M, N[, val])`` (include/utils.hpp:21-31): a zero- (or
    ``val``-) initialised M x N matrix, ``layout`` "F" = first_order (default, as in the reference)
    or "L"/"C" = last_order.  ``device=None`` gives a numpy array, otherwise a torch tensor."""
    import numpy as np
    order = "F" if layout.upper() == "F" else "C"
    if device is None:
        x = np.zeros((M, N), dtype=dtype, order=order)
        if val is not None:
            x[...] = val
        return x
    import torch
    tdt = {"float32": torch.float32, "float64": torch.float64}[np.dtype(dtype).name]
    if order == "C":
        x = torch.zeros((M, N), dtype=tdt, device=device)
    else:
        x = torch.zeros((N, M), dtype=tdt, device=device).t()
    if val is not None:
        x.fill_(val)
    return x


def pinned_empty(shape, dtype, order="C"):
    """numpy array backed by pinned host memory from ``b200_host_alloc`` (async, overlappable copies)."""
    import numpy as np
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    p = C.c_void_p()
    _check(lib().b200_host_alloc(C.byref(p), n * dtype.itemsize))
    buf = (C.c_char * (n * dtype.itemsize)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape, order=order)
    _pinned_keepalive[arr.ctypes.data] = p
    return arr


_pinned_keepalive: dict = {}


def pinned_free(arr) -> None:
    p = _pinned_keepalive.pop(arr.ctypes.data, None)
    if p is not None:
        _check(lib().b200_host_free(p))


# ---- operand replication over NVLink (include/b200_replicate.h) ------------------------------------
def replicate_push(dst_ptrs, src_ptr: int, nbytes: int, flag_ptrs, flag_value: int, *, multicast: bool,
                   flag_multicast: bool, ctas: int = 0, stream: int = 0) -> None:
    """``b200_replicate_push``: copy ``nbytes`` from ``src_ptr`` to every address in ``dst_ptrs`` (one
    multicast address when ``multicast``), then publish ``flag_value`` at ``flag_ptrs``."""
    d = (C.c_void_p * len(dst_ptrs))(*dst_ptrs)
    f = (C.c_void_p * max(1, len(flag_ptrs)))(*flag_ptrs)
    _check(lib().b200_replicate_push(d, len(dst_ptrs), int(bool(multicast)), C.c_void_p(src_ptr), nbytes,
                                     f, len(flag_ptrs), int(bool(flag_multicast)), flag_value & 0xffffffff,
                                     int(ctas), C.c_void_p(stream)))


def replicate_push_2d(dst_ptrs, src_ptr: int, rows: int, row_bytes: int, src_pitch: int, dst_pitch: int, flag_ptrs,
                      flag_value: int, *, multicast: bool, flag_multicast: bool, ctas: int = 0, stream: int = 0) -> None:
    """``b200_replicate_push_2d``: ``rows`` runs of ``row_bytes`` bytes with separate pitches (a column panel
    of a row-major matrix), then the flag(s)."""
    d = (C.c_void_p * len(dst_ptrs))(*dst_ptrs)
    f = (C.c_void_p * max(1, len(flag_ptrs)))(*flag_ptrs)
    _check(lib().b200_replicate_push_2d(d, len(dst_ptrs), int(bool(multicast)), C.c_void_p(src_ptr), rows, row_bytes,
                                        src_pitch, dst_pitch, f, len(flag_ptrs), int(bool(flag_multicast)),
                                        flag_value & 0xffffffff, int(ctas), C.c_void_p(stream)))


def flag_wait(flag_ptr: int, value: int, *, count: int = 1, stride: int = 1, skip: int = -1, stream: int = 0) -> None:
    """``b200_flag_wait``: the stream waits until the flag word(s) have reached ``value``."""
    _check(lib().b200_flag_wait(C.c_void_p(flag_ptr), value & 0xffffffff, count, stride, skip, C.c_void_p(stream)))


def flag_signal(flag_ptr: int, value: int, *, stream: int = 0) -> None:
    """``b200_flag_signal``: write ``value`` to the (peer) flag word after the stream's prior work."""
    _check(lib().b200_flag_signal(C.c_void_p(flag_ptr), value & 0xffffffff, C.c_void_p(stream)))


def last_choice() -> dict:
    ch = _Choice()
    _check(lib().b200_mtm_last_choice(C.byref(ch)))
    inv = {v: k for k, v in VARIANTS.items()}
    return {"variant": inv.get(ch.variant, ch.variant), "config": ch.config, "launches": ch.launches,
            "a_mode": ch.a_mode, "b_mode": ch.b_mode, "name": ch.name.decode()}


def plan_f32(M: int, N: int, K: int, sm_count: int = 0) -> dict:
    """What AUTO resolves an M x N x K fp32 problem to (kernel family, tile config) on a device with ``sm_count`` SMs
    (0: the current device) — ``b200_mtm_plan_f32``; needs no GPU when ``sm_count`` is given."""
    ch = _Choice()
    _check(lib().b200_mtm_plan_f32(M, N, K, sm_count, C.byref(ch)))
    inv = {v: k for k, v in VARIANTS.items()}
    return {"variant": inv.get(ch.variant, ch.variant), "config": ch.config, "name": ch.name.decode()}


def last_bench_enqueue_us() -> float:
    """Host microseconds per enqueued call of the last ``bench_device`` on this thread."""
    return float(lib().b200_last_bench_enqueue_us())


def launch_count() -> int:
    return int(lib().b200_launch_count())


def num_configs(variant, is_f64: bool) -> int:
    v = VARIANTS[variant] if isinstance(variant, str) else int(variant)
    return int(lib().b200_mtm_num_configs(v, int(is_f64)))


def config_name(variant, is_f64: bool, config: int) -> str:
    v = VARIANTS[variant] if isinstance(variant, str) else int(variant)
    return lib().b200_mtm_config_name(v, int(is_f64), config).decode()


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().b200_device_count(C.byref(n))
    return int(n.value) if rc == 0 else 0


def device_info(device: int = 0) -> dict:
    info = _DeviceInfo()
    _check(lib().b200_get_device_info(device, C.byref(info)))
    d = {k: getattr(info, k) for k, _ in _DeviceInfo._fields_}
    d["name"] = info.name.decode()
    return d
