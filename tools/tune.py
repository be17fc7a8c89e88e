"""Time every tile config of every kernel family on the headline shapes (run on the B200 box).

    python tools/tune.py [--out gpurun_out/tune.json] [--n 8192]

Feeds the speed tables in csrc/mtm_api.cu (kSimtF32Speed ...) and DESIGN.md section 5.
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import openmp_blas_b200 as ob  # noqa: E402


def uniform(shape, dtype, layout, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    if layout == "L":
        return torch.rand(shape, device="cuda", dtype=dtype, generator=g) * 2 - 1
    return (torch.rand((shape[1], shape[0]), device="cuda", dtype=dtype, generator=g) * 2 - 1).t()


def sweep(args, fams_wanted):
    """Every config of the wanted fp32 families on a list of shapes (row-major): the table behind the AUTO choice."""
    shapes = [(int(x),) * 3 for x in args.sizes.split(",") if x] + \
             [tuple(int(v) for v in sh.split("x")) for sh in args.shapes.split(",") if sh]
    rows = []
    for (M, N, K) in shapes:
        a = uniform((M, K), torch.float32, "L", 1)
        b = uniform((K, N), torch.float32, "L", 2)
        fl = M * N * (2.0 * K - 1)
        for fam in ("simt", "3xtf32"):
            if fam not in fams_wanted:
                continue
            best = None
            splits = [0]
            if fam == "3xtf32" and M * N <= 2048 * 2048:
                splits = [0, 1, 2, 4, 8, 16]
            for cfg in [None] + list(range(ob.num_configs(fam, False))):
                for sk in (splits if cfg is not None else [0]):
                    c = torch.zeros((M, N), device="cuda", dtype=torch.float32)
                    try:
                        ms = ob.bench_device(c, a, b, variant=fam, config=cfg, warmup=3, split_k=sk,
                                             iters=max(args.iters, 20 if M * N * K < 2 ** 34 else 3))
                        row = {"shape": [M, N, K], "family": fam, "config": cfg, "split_k": sk, "name": ob.last_choice()["name"],
                               "ms": round(ms, 5), "tflops": round(fl / ms / 1e9, 2)}
                    except Exception as e:
                        row = {"shape": [M, N, K], "family": fam, "config": cfg, "split_k": sk, "error": str(e)[:200]}
                    print(json.dumps(row), flush=True)
                    rows.append(row)
                    del c
        del a, b
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps({"rows": rows}, indent=1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/tune.json")
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--families", default="simt,3xtf32,dfma,dmma")
    ap.add_argument("--sizes", default="", help="comma list of square sizes (overrides --n); LLL only beyond the first")
    ap.add_argument("--shapes", default="", help="extra MxNxK shapes, comma separated (LLL)")
    args = ap.parse_args()
    fams_wanted = set(args.families.split(","))
    if args.sizes or args.shapes:
        return sweep(args, fams_wanted)
    n = args.n
    info = ob.device_info(0)
    res = {"device": info["name"], "sm_count": info["sm_count"], "sm_clock_khz": info["sm_clock_khz"],
           "peak_fp32": info["peak_fp32_tflops"], "peak_fp64": info["peak_fp64_tflops"], "n": n, "rows": []}
    fl = n * n * (2.0 * n - 1)
    for dtype, fams in ((torch.float32, ["simt", "3xtf32"]), (torch.float64, ["dfma", "dmma"])):
        is64 = dtype == torch.float64
        for fam in fams:
            if fam not in fams_wanted:
                continue
            for cfg in range(ob.num_configs(fam, is64)):
                for lay in (("LLL", "LFL", "LLF", "FFF") if cfg == 0 else ("LLL",)):
                    a = uniform((n, n), dtype, lay[1], 1)
                    b = uniform((n, n), dtype, lay[2], 2)
                    c = torch.zeros((n, n), device="cuda", dtype=dtype)
                    if lay[0] == "F":
                        c = c.t()
                    try:
                        ms = ob.bench_device(c, a, b, variant=fam, config=cfg, warmup=2, iters=args.iters)
                        ch = ob.last_choice()
                        row = {"family": fam, "config": cfg, "name": ch["name"], "layout": lay, "ms": ms,
                               "tflops": fl / ms / 1e9, "a_mode": ch["a_mode"], "b_mode": ch["b_mode"]}
                    except Exception as e:  # keep going: one bad config must not lose the run
                        row = {"family": fam, "config": cfg, "layout": lay, "error": str(e)}
                    print(json.dumps(row), flush=True)
                    res["rows"].append(row)
                    del a, b, c
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
