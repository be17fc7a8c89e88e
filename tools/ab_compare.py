"""A/B timing of kernel configs under the SAME thermal / power state: the candidates are
interleaved round-robin (several rounds, many launches each) so that power-cap clock drift, which
confounds back-to-back measurements of tensor-core kernels, hits all of them equally.

    python tools/ab_compare.py 3xtf32 0 2 [--n 8192] [--rounds 4] [--iters 30]
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import openmp_blas_b200 as ob  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("family")
ap.add_argument("configs", nargs="+", type=int)
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--m", type=int, default=0)
ap.add_argument("--k", type=int, default=0)
ap.add_argument("--rounds", type=int, default=4)
ap.add_argument("--iters", type=int, default=30)
args = ap.parse_args()
dtype = torch.float64 if args.family in ("dfma", "dmma") else torch.float32
n = args.n
m = args.m or n
k = args.k or n
a = torch.rand((m, k), device="cuda", dtype=dtype) * 2 - 1
b = torch.rand((k, n), device="cuda", dtype=dtype) * 2 - 1
c = torch.zeros((m, n), device="cuda", dtype=dtype)
fl = m * n * (2.0 * k - 1)
res = {cfg: [] for cfg in args.configs}
for cfg in args.configs:     # warm every candidate (and the clocks) first
    ob.bench_device(c, a, b, variant=args.family, config=cfg, warmup=3, iters=20)
for r in range(args.rounds):
    for cfg in args.configs:
        ms = ob.bench_device(c, a, b, variant=args.family, config=cfg, warmup=1, iters=args.iters)
        res[cfg].append(ms)
out = {ob.config_name(args.family, dtype == torch.float64, cfg): {"ms": [round(m, 4) for m in v],
       "tflops_mean": round(fl / (sum(v) / len(v)) / 1e9, 2)} for cfg, v in res.items()}
print(json.dumps(out))
