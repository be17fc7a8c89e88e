#!/bin/bash
# Is the small-shape loop launch-bound?  Device ms per call vs host enqueue time per call.
timeout 300 python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
import openmp_blas_b200 as ob
for fam, cfgs in (("3xtf32", [None, 5, 1, 8]), ("simt", [None])):
    for n in (128, 256, 512, 768, 1024, 2048):
        for cfg in cfgs:
            a = torch.rand((n, n), device="cuda") * 2 - 1; b = torch.rand((n, n), device="cuda") * 2 - 1; c = torch.zeros((n, n), device="cuda")
            ms = ob.bench_device(c, a, b, variant=fam, config=cfg, warmup=5, iters=200)
            print(f"{fam} n={n} cfg={cfg} {ob.last_choice()['name']}: device {ms*1e3:.1f} us/call, host enqueue {ob.last_bench_enqueue_us():.1f} us/call, {n*n*(2.0*n-1)/ms/1e9:.1f} TFLOP/s")
PY
