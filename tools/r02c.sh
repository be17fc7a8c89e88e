#!/bin/bash
# Round-2 multi-GPU contact (run with gpurun --gpus N): the one-call mgpu C entry, the sharded drivers, bench at N.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus_n$N.txt 2>&1; nproc >> gpurun_out/gpus_n$N.txt
nvidia-smi topo -m >> gpurun_out/gpus_n$N.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -s -k "mgpu or replicate or slab" > gpurun_out/pytest_mgpu_n$N.log 2>&1; echo "pytest mgpu/replicate/slab exit $?"; grep -E "^\[mgpu\]|passed|failed|Error" gpurun_out/pytest_mgpu_n$N.log | cut -c1-250 | tail -12
/usr/bin/g++ -std=c++20 -O2 -fopenmp -Iinclude/compat -Iinclude tools/mtm_mgpu_check.cpp -o /tmp/mtm_mgpu_check -Lopenmp-blas_b200 -lb200mtm -Wl,-rpath,$PWD/openmp-blas_b200 || echo "compile failed"
for args in "--size 8192 --devices 1 --calls 3" "--size 8192 --calls 3" "--size 8192 --calls 3 --layout F" "--size 16384 --calls 2" "--size ${BIG:-32768} --calls 2" "--size ${BIG:-32768} --devices 1 --calls 2"; do
  timeout 900 /tmp/mtm_mgpu_check $args >> gpurun_out/mgpu_cpp_n$N.jsonl 2>> gpurun_out/mgpu_cpp_n$N.err; echo "mtm_mgpu_check $args exit $?"; tail -1 gpurun_out/mgpu_cpp_n$N.jsonl | cut -c1-400
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    tools/multi_gpu_check.py --size 4096 --big-size 0 --bcast nccl,nvlink --summa > gpurun_out/mgpu_check_n$N.out 2> gpurun_out/mgpu_check_n$N.err
echo "multi_gpu_check exit $?"; grep '^{' gpurun_out/mgpu_check_n$N.out | cut -c1-600; tail -3 gpurun_out/mgpu_check_n$N.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --gpus $N --steps ${STEPS:-20} --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?"; tail -5 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1])
    print("  value", d["value"], "ms", d["ms_per_step"], d["config"].get("b_replication"), d["config"].get("calibration_ms_per_step"), d["config"].get("k_chunks"))
    print("  by rank", d["config"].get("ms_per_step_by_rank"))
    e = d["e2e"]; print("  e2e", e["value"], e["ms_per_step"], "| one call:", e.get("one_call_mgpu_c_abi"), "| per rank:", (e.get("one_process_per_gpu") or {}).get("value"))
    print("  config5", d["config5"])
    print("  watchdog", d.get("watchdog"))
except Exception as ex:
    print("  parse failed", ex)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
echo "reference arm N=$N exit $?"; cut -c1-700 gpurun_out/bench_ref_n$N.json
