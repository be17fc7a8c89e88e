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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/tune.json")
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    n = args.n
    info = ob.device_info(0)
    res = {"device": info["name"], "sm_count": info["sm_count"], "sm_clock_khz": info["sm_clock_khz"],
           "peak_fp32": info["peak_fp32_tflops"], "peak_fp64": info["peak_fp64_tflops"], "n": n, "rows": []}
    fl = n * n * (2.0 * n - 1)
    for dtype, fams in ((torch.float32, ["simt", "3xtf32"]), (torch.float64, ["dfma", "dmma"])):
        is64 = dtype == torch.float64
        for fam in fams:
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
