#!/bin/bash
# closing checks on the final kernel sources: probe, full GPU tests, unaligned-C epilogue timing, the ncu record the traffic stamp refers to
mkdir -p gpurun_out
show() { grep -E "^==|^BAD|^FAIL|^HANG" "$1" | cut -c1-300 | head -${2:-8}; }
timeout 900 python tools/tf32_probe.py > gpurun_out/probe_final.log 2>&1; echo "probe exit $?"; show gpurun_out/probe_final.log 10
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
import openmp_blas_b200 as ob
for (M, N, K) in ((4097, 4099, 4101), (4096, 4099, 4096), (4096, 4096, 4096)):
    a = torch.rand((M, K), device="cuda") * 2 - 1; b = torch.rand((K, N), device="cuda") * 2 - 1; c = torch.zeros((M, N), device="cuda")
    ms = ob.bench_device(c, a, b, variant="3xtf32", warmup=3, iters=10)
    ch = ob.last_choice()
    print(f"unaligned-C epilogue: {M}x{N}x{K}: {M*N*(2.0*K-1)/ms/1e9:.1f} TFLOP/s {ms:.3f} ms ({ch['name']}, modes {ch['a_mode']},{ch['b_mode']})")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mtm_tf32x3|split_kernel" -s 4 -c 2 -f -o gpurun_out/prof_3xtf32_final \
      python tools/one_call.py 3xtf32 8192 LLL > gpurun_out/ncu_3xtf32_final.log 2>&1; echo "ncu 3xtf32 exit $?"
