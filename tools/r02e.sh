#!/bin/bash
# Round-2 8-GPU confirmation (gpurun --gpus 8): every second here is charged 8x — bench first, then the C++ front-end.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus_n$N.txt 2>&1; nproc >> gpurun_out/gpus_n$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?"; grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" gpurun_out/bench_n$N.err | tail -5 | cut -c1-300
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1])
    print("  value", d["value"], "ms", d["ms_per_step"], d["config"].get("b_replication"), d["config"].get("calibration_ms_per_step"), d["config"].get("k_chunks"))
    print("  by rank", d["config"].get("ms_per_step_by_rank"))
    e = d["e2e"]; print("  e2e", e["value"], e["ms_per_step"], "| one call:", {k: v for k, v in (e.get("one_call_mgpu_c_abi") or {}).items() if k != "api"}, "| per rank:", (e.get("one_process_per_gpu") or {}).get("value"))
    print("  config5", {k: v for k, v in d["config5"].items() if k not in ("workload", "exact_check")})
    print("  summa", d.get("summa_2d"))
    print("  watchdog", d.get("watchdog"))
except Exception as ex:
    print("  parse failed", ex)
PY
timeout 300 python -m pytest tests -m gpu -q -s -k "mgpu" > gpurun_out/pytest_mgpu_n$N.log 2>&1; echo "pytest mgpu exit $?"; grep -E "^\[mgpu\]|passed|failed" gpurun_out/pytest_mgpu_n$N.log | cut -c1-200 | tail -8
/usr/bin/g++ -std=c++20 -O2 -fopenmp -Iinclude/compat -Iinclude tools/mtm_mgpu_check.cpp -o /tmp/mtm_mgpu_check -Lopenmp-blas_b200 -lb200mtm -Wl,-rpath,$PWD/openmp-blas_b200 || echo "compile failed"
for args in "--size 8192 --calls 3" "--size 32768 --calls 2"; do
  timeout 600 /tmp/mtm_mgpu_check $args >> gpurun_out/mgpu_cpp_n$N.jsonl 2>> gpurun_out/mgpu_cpp_n$N.err; echo "mtm_mgpu_check $args exit $?"; tail -1 gpurun_out/mgpu_cpp_n$N.jsonl | cut -c1-400
done
