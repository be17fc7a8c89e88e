#!/bin/bash
# Round-2: tail split-K validation + A/B.
mkdir -p gpurun_out
show() { grep -E "^==|^BAD|^FAIL|^HANG" "$1" | cut -c1-300 | head -${2:-8}; }
timeout 900 python tools/tf32_probe.py > gpurun_out/probe_tail.log 2>&1; echo "probe exit $?"; show gpurun_out/probe_tail.log 10
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; grep "tail split\|stream-K" gpurun_out/pytest_gpu.log | cut -c1-200; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
for ts in 0 1 0 1; do
B200_TF32_TAIL_SPLIT=$ts timeout 600 python - <<PY
import torch, sys
sys.path.insert(0, ".")
import openmp_blas_b200 as ob
out = []
for (M, N, K) in ((2048,)*3, (2560,)*3, (3072,)*3, (3584,)*3, (4096,)*3, (5120,)*3, (6144,)*3, (7168,)*3, (8192,)*3, (4096, 32768, 8192), (65536, 1024, 1024)):
    a = torch.rand((M, K), device="cuda") * 2 - 1; b = torch.rand((K, N), device="cuda") * 2 - 1; c = torch.zeros((M, N), device="cuda")
    ms = ob.bench_device(c, a, b, variant="3xtf32", warmup=3, iters=20 if M * N * K < 2 ** 36 else 5)
    out.append(f"{M}x{N}x{K}: {M*N*(2.0*K-1)/ms/1e9:.1f} ({ob.last_choice()['name'].replace('tf32x3_', '')})")
    del a, b, c
print("TAIL_SPLIT=$ts |", " | ".join(out))
PY
done
