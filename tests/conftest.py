"""pytest configuration for the mtm path.

`-m "not gpu"`: oracle vs the reference's golden vectors, host logic, C-ABI symbol checks (CPU only).
`-m gpu`      : parity of the CUDA path against the oracle, through the C ABI (needs a B200).
"""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    return oracle.Oracle()


@pytest.fixture(scope="session")
def reference_lib():
    import oracle
    try:
        return oracle.Reference()
    except FileNotFoundError as e:  # prebuilt oracle/_ref missing and no /root/reference to build from
        pytest.skip(str(e))


@pytest.fixture(scope="session")
def ob():
    import __graft_entry__ as ge
    ge.build_library()
    import openmp_blas_b200
    return openmp_blas_b200


LAYOUTS = ["FFF", "FFL", "FLF", "LFF", "FLL", "LFL", "LLF", "LLL"]  # (C, A, B) as in test/test.mtm.cpp


def order_of(tag: str) -> str:
    return "F" if tag == "F" else "C"


def int_matrix(rng, shape, dtype, tag):
    """rand()%100-style integer-valued matrix (test/test_utils.hpp:4-9) in the given layout."""
    return np.asarray(rng.integers(0, 100, size=shape).astype(dtype), order=order_of(tag))


def uniform_matrix(rng, shape, dtype, tag, lo=-1.0, hi=1.0):
    return np.asarray(rng.uniform(lo, hi, size=shape).astype(dtype), order=order_of(tag))


def exact_int(c0, a, b, oracle_lib, calls=1):
    """Exact C0 + calls*A@B for integer-valued inputs via the oracle's int64 triple loop."""
    ci = np.zeros(c0.shape, dtype=np.int64)
    oracle_lib.exact_i64(ci, np.ascontiguousarray(a.astype(np.int64)), np.ascontiguousarray(b.astype(np.int64)))
    return c0.astype(np.int64) + calls * ci
