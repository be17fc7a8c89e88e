"""A handful of mtm calls of one kernel family, for ncu captures:  python tools/one_call.py simt [n] [layout]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import openmp_blas_b200 as ob  # noqa: E402

fam = sys.argv[1] if len(sys.argv) > 1 else "simt"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
lay = sys.argv[3] if len(sys.argv) > 3 else "LLL"
cfg = int(sys.argv[4]) if len(sys.argv) > 4 else None
dtype = torch.float64 if fam in ("dfma", "dmma") else torch.float32
mk = lambda t: (torch.rand((n, n), device="cuda", dtype=dtype) * 2 - 1) if t == "L" else (torch.rand((n, n), device="cuda", dtype=dtype) * 2 - 1).t()
a, b = mk(lay[1]), mk(lay[2])
c = torch.zeros((n, n), device="cuda", dtype=dtype)
if lay[0] == "F":
    c = c.t()
fn = ob.mtm(c, a, b, None, variant=fam, config=cfg)
for _ in range(4):
    fn()
torch.cuda.synchronize()
print(fam, n, lay, ob.last_choice())
