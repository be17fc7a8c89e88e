"""A few mtm calls of one (family, config) on an M x N x K row-major problem, for ncu captures:
    python tools/one_call_shape.py 3xtf32 2 4096 32768 32768"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import openmp_blas_b200 as ob  # noqa: E402

fam, cfg, m, n, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
dtype = torch.float64 if fam in ("dfma", "dmma") else torch.float32
a = torch.rand((m, k), device="cuda", dtype=dtype) * 2 - 1
b = torch.rand((k, n), device="cuda", dtype=dtype) * 2 - 1
c = torch.zeros((m, n), device="cuda", dtype=dtype)
fn = ob.mtm(c, a, b, None, variant=fam, config=cfg)
for _ in range(3):
    fn()
torch.cuda.synchronize()
print(fam, cfg, (m, n, k), ob.last_choice())
