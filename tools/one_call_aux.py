"""A few mtv / vtm / transpose calls for ncu captures:  python tools/one_call_aux.py mtv|vtm|trans [n] [F|L] [f32|f64]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import openmp_blas_b200 as ob  # noqa: E402

op = sys.argv[1] if len(sys.argv) > 1 else "mtv"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
order = sys.argv[3] if len(sys.argv) > 3 else "F"
dtype = torch.float64 if (len(sys.argv) > 4 and sys.argv[4] == "f64") else torch.float32
a = torch.rand((n, n), device="cuda", dtype=dtype) * 2 - 1
if order == "F":
    a = a.t()
if op in ("mtv", "vtm"):
    v = torch.rand(n, device="cuda", dtype=dtype)
    c = torch.zeros(n, device="cuda", dtype=dtype)
    fn = (ob.vtm if op == "vtm" else ob.mtv)(c, a, v)
else:
    c = torch.zeros((n, n), device="cuda", dtype=dtype)
    fn = ob.transpose(c, a)
for _ in range(4):
    fn()
torch.cuda.synchronize()
print(op, n, order, ob.last_choice())
