"""Generate tests/golden/mtm_golden.npz by running the UNMODIFIED reference here.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
Each case stores the inputs, the initial C, the C produced by the reference's amt::mtm
(through oracle/_ref, i.e. /root/reference/include/mtm.hpp compiled as-is) and the block sizes
the reference derived on the generating host (they fix the K-block summation order).
Non-integer data on purpose: these vectors pin rounding behaviour, the integer cases of
test/test.mtm.cpp pin indexing.  The file is committed; the GPU box never needs /root/reference.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402

CASES = [
    # (layouts C,A,B ; dtype ; M, N, K)   — K > KB (512 f32 / 448, 426 f64) in the last two of each dtype
    ("FFF", "float32", 17, 23, 31), ("LLL", "float32", 40, 48, 64), ("FLF", "float32", 33, 20, 129),
    ("LFL", "float32", 24, 7, 600), ("LLF", "float32", 5, 30, 520),
    ("FFF", "float64", 19, 21, 33), ("LLL", "float64", 24, 32, 72), ("FLL", "float64", 17, 9, 500),
    ("LFF", "float64", 9, 20, 450),
]


def main():
    ref = oracle.Reference(isa="v3")
    rng = np.random.default_rng(20261017)
    out = {}
    for idx, (lay, dt, M, N, K) in enumerate(CASES):
        o = lambda ch: "F" if ch == "F" else "C"
        a = np.asarray(rng.uniform(-1, 1, (M, K)).astype(dt), order=o(lay[1]))
        b = np.asarray(rng.uniform(-1, 1, (K, N)).astype(dt), order=o(lay[2]))
        c0 = np.asarray(rng.uniform(-1, 1, (M, N)).astype(dt), order=o(lay[0]))
        c = c0.copy(order="K")
        ref.mtm(c, a, b)
        blk = ref.block_sizes(dt, lay[0] == "L")
        out[f"case{idx}_layout"] = np.array(lay)
        out[f"case{idx}_a"] = a
        out[f"case{idx}_b"] = b
        out[f"case{idx}_c0"] = c0
        out[f"case{idx}_c"] = c
        out[f"case{idx}_blocks"] = np.array(blk, dtype=np.int64)
    out["ncases"] = np.array(len(CASES))
    path = Path(__file__).with_name("mtm_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
