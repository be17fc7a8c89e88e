#!/bin/bash
# Round-2 fourth single-GPU contact: the fused (in-kernel lo conversion) 3xTF32 configs.
mkdir -p gpurun_out
show() { grep -E "^==|^BAD|^FAIL|^HANG" "$1" | cut -c1-300 | head -${2:-8}; }
TF32_PROBE_CFGS=7,8,6 timeout 900 python tools/tf32_probe.py > gpurun_out/probe_fused.log 2>&1; echo "probe fused exit $?"; show gpurun_out/probe_fused.log 16
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python tools/tune.py --families 3xtf32 --sizes 256,512,768,1024,1536,2048,3072,4096,8192 --shapes 65536x1024x1024,512x512x8192,8192x8192x1024 --out gpurun_out/tune_fused.json > gpurun_out/tune_fused.log 2>&1; echo "tune exit $?"
python - <<'PY'
import json
from collections import defaultdict
try:
    rows = json.load(open("gpurun_out/tune_fused.json"))["rows"]
    t = defaultdict(list)
    for r in rows:
        if "tflops" in r and r["split_k"] in (0, 1): t[tuple(r["shape"])].append((r["tflops"], r["ms"], r["config"], r["split_k"], r["name"].replace("tf32x3_", "")))
    for sh, v in t.items():
        v.sort(reverse=True)
        auto = [x for x in v if x[2] is None]
        print(sh, "auto:", auto[0] if auto else None, "| best:", v[:4])
except Exception as e:
    print("parse failed", e)
PY
for cfg in 0 6 0 6; do
  timeout 300 python - <<PY
import torch, sys
sys.path.insert(0, ".")
import openmp_blas_b200 as ob
n = 8192
a = torch.rand((n, n), device="cuda") * 2 - 1; b = torch.rand((n, n), device="cuda") * 2 - 1; c = torch.zeros((n, n), device="cuda")
ms = ob.bench_device(c, a, b, variant="3xtf32", config=$cfg, warmup=5, iters=40)
print("8192^3 LLL cfg $cfg", ob.last_choice()["name"], round(ms, 4), "ms", round(n * n * (2.0 * n - 1) / ms / 1e9, 1), "TFLOP/s (40 back-to-back calls)")
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mtm_tf32x3" -s 2 -c 1 -f -o gpurun_out/prof_3xtf32_fused_8192 \
      python tools/one_call.py 3xtf32 8192 LLL 6 > gpurun_out/ncu_fused_8192.log 2>&1; echo "ncu fused 8192 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mtm_tf32x3" -s 2 -c 1 -f -o gpurun_out/prof_3xtf32_fused_1024 \
      python tools/one_call.py 3xtf32 1024 LLL 8 > gpurun_out/ncu_fused_1024.log 2>&1; echo "ncu fused 1024 exit $?"
