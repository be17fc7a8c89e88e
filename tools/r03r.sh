#!/bin/bash
# r03r (gpurun --gpus 8): the double-tile configs on 8 GPUs — one-call mgpu tests, bench at N = 8 (weak step, config 5, SUMMA, e2e).
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -k "mgpu" > gpurun_out/r03r_pytest_mgpu_n$N.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|Error" gpurun_out/r03r_pytest_mgpu_n$N.log | cut -c1-250 | tail -4
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r03r_bench_n$N.json 2> gpurun_out/r03r_bench_n$N.err
echo "bench N=$N exit $?"; tail -3 gpurun_out/r03r_bench_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r03r_bench_n$N.json") if l.startswith("{")][-1])
    print("  value", d["value"], "ms", d["ms_per_step"], d["config"].get("kernel"), d["config"].get("b_replication"), d["config"].get("calibration_ms_per_step"), d["config"].get("k_chunks"))
    print("  by rank", d["config"].get("ms_per_step_by_rank"))
    e = d["e2e"]; print("  e2e", e["value"], e["ms_per_step"])
    print("  config5", d["config5"])
    print("  summa", d.get("summa_2d"))
    print("  watchdog", d.get("watchdog"))
except Exception as ex:
    print("  parse failed", ex)
PY
