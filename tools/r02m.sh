#!/bin/bash
# bench at N GPUs only (scaling table)
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
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
