"""Where AUTO should switch from the CUDA-core kernels to the 3xTF32 path: both families (their own AUTO tile choice)
timed on small and thin shapes, row-major.   python tools/auto_crossover.py [--out gpurun_out/auto_crossover.jsonl]"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import openmp_blas_b200 as ob  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/auto_crossover.jsonl")
args = ap.parse_args()
shapes = [(n, n, n) for n in (64, 96, 128, 192, 256, 384, 512, 768)]
shapes += [(128, 128, 1024), (128, 128, 8192), (256, 256, 64), (256, 256, 128), (512, 512, 64), (512, 512, 128), (1024, 1024, 64), (1024, 1024, 128),
           (4096, 128, 128), (128, 4096, 128), (4096, 64, 512), (64, 4096, 512), (8192, 256, 256), (256, 8192, 256), (16384, 128, 1024),
           (4096, 4096, 32), (4096, 4096, 64), (2048, 2048, 128), (300, 260, 520), (1000, 1000, 1000), (33, 4096, 4096), (4096, 33, 4096)]
with open(args.out, "w") as f:
    for (m, n, k) in shapes:
        a = torch.rand((m, k), device="cuda") * 2 - 1
        b = torch.rand((k, n), device="cuda") * 2 - 1
        c = torch.zeros((m, n), device="cuda")
        row = {"shape": [m, n, k]}
        for fam in ("simt", "3xtf32", "auto"):
            try:
                best = min(ob.bench_device(c, a, b, variant=fam, warmup=3, iters=100) for _ in range(3))
                row[fam] = {"us": round(best * 1e3, 2), "kernel": ob.last_choice()["name"]}
            except Exception as e:
                row[fam] = {"error": str(e)[:100]}
        line = json.dumps(row)
        print(line, flush=True)
        f.write(line + "\n")
