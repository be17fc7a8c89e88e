"""A/B timing of environment-selected variants of the 3xTF32 path under the SAME thermal / power state: the settings
are interleaved round-robin in one process (the knobs named below are read on every call).

    python tools/ab_env.py --shapes 512,1024,65536x1024x1024 --env "" B200_TF32_NO_PDL=1 B200_TF32_NO_PDL=2
"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import openmp_blas_b200 as ob  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="512,1024,2048,4096,8192,65536x1024x1024")
ap.add_argument("--env", nargs="+", default=["", "B200_TF32_NO_PDL=1"])
ap.add_argument("--family", default="3xtf32")
ap.add_argument("--rounds", type=int, default=4)
ap.add_argument("--layout", default="LLL")
ap.add_argument("--config", type=int, default=None)
ap.add_argument("--check", action="store_true", help="also compare the results of the settings bit for bit (integer data)")
args = ap.parse_args()
KNOBS = ("B200_TF32_NO_PDL", "B200_TF32_PEER_ARRIVE", "B200_TF32_GROUP")


def apply(setting):
    for k in KNOBS:
        os.environ.pop(k, None)
    for kv in setting.split(","):
        if kv:
            k, v = kv.split("=")
            os.environ[k] = v


def mat(rows, cols, layout, gen=None, ints=False):
    shape = (rows, cols) if layout == "L" else (cols, rows)
    t = torch.randint(0, 10, shape, device="cuda").float() if ints else torch.rand(shape, device="cuda") * 2 - 1
    return t if layout == "L" else t.t()


for sh in args.shapes.split(","):
    dims = [int(v) for v in sh.split("x")]
    m, n, k = dims if len(dims) == 3 else (dims[0],) * 3
    lc, la, lb = args.layout
    a, b, c = mat(m, k, la), mat(k, n, lb), mat(m, n, lc)
    fl = m * n * (2.0 * k - 1)
    iters = 200 if fl < 1e10 else (50 if fl < 5e11 else 15)
    res = {e: [] for e in args.env}
    names = {}
    for e in args.env:
        apply(e)
        ob.bench_device(c, a, b, variant=args.family, config=args.config, warmup=3, iters=iters)
        names[e] = ob.last_choice()["name"]
    for r in range(args.rounds):
        for e in args.env:
            apply(e)
            res[e].append(ob.bench_device(c, a, b, variant=args.family, config=args.config, warmup=1, iters=iters))
    out = {"shape": [m, n, k], "layout": args.layout}
    for e in args.env:
        best = min(res[e])
        out[e or "default"] = {"kernel": names[e], "us_best": round(best * 1e3, 2), "us_mean": round(sum(res[e]) / len(res[e]) * 1e3, 2),
                               "tflops_best": round(fl / best / 1e9, 2)}
    if args.check:
        ai, bi = mat(m, k, la, ints=True), mat(k, n, lb, ints=True)
        outs = []
        for e in args.env:
            apply(e)
            ci = torch.zeros_like(c)
            ob.mtm(ci, ai, bi, None, variant=args.family, config=args.config)()
            ob.mtm(ci, ai, bi, None, variant=args.family, config=args.config)()
            torch.cuda.synchronize()
            outs.append(ci)
        rows = torch.randint(0, m, (min(m, 64),), device="cuda")
        want = 2 * (ai[rows].double() @ bi.double())
        out["exact_vs_fp64_rows"] = [bool(torch.equal(o[rows].double(), want)) for o in outs]
        out["identical"] = all(torch.equal(outs[0], o) for o in outs[1:])
    print(json.dumps(out), flush=True)
    apply("")
