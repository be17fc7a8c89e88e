#!/bin/bash
# Where does the fused 3xTF32 kernel lose its time?  Bring-up switches, timing only (results are wrong for dbg != 0).
mkdir -p gpurun_out
for dbg in 0 1 2 4 8 12 3 15; do
  B200_TF32_FUSED_DBG=$dbg timeout 300 python - <<PY
import torch, sys
sys.path.insert(0, ".")
import openmp_blas_b200 as ob
out = []
for (n, cfg) in ((1024, 8), (1024, 7), (2048, 6), (8192, 6)):
    a = torch.rand((n, n), device="cuda") * 2 - 1; b = torch.rand((n, n), device="cuda") * 2 - 1; c = torch.zeros((n, n), device="cuda")
    ms = ob.bench_device(c, a, b, variant="3xtf32", config=cfg, warmup=5, iters=30)
    out.append(f"{n}^3 cfg{cfg}: {ms*1e3:.1f} us")
print("dbg=$dbg", " | ".join(out))
PY
done
