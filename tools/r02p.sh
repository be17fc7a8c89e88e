#!/bin/bash
# Round-2 closing single-GPU session on the final tree: the driver's sequence + the ncu capture the traffic stamp refers to.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "reference arm exit $?"; cut -c1-200 gpurun_out/bench_reference_arm.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
    print("  value", d["value"], "ms", d["ms_per_step"], "roof", d["roofline"]["frac"], d["roofline"]["ms_per_launch"], d["roofline"].get("sustained", {}).get("frac"), "traffic", d["roofline"]["traffic"], d["roofline"]["traffic_stale"], "clocks", d["clocks"])
    print("  e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "pageable", d["e2e"]["pageable"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "launches", d["gpu_launches"])
    print("  config5", d["config5"]["tflops_with_broadcast"], d["config5"]["exact"], "watchdog", d.get("watchdog"))
    for row in d["extras"]["config2_fp32_square_sweep_LLL"]:
        print("  n", row["n"], {k: (v["tflops"], v["kernel"]) for k, v in row.items() if k != "n"})
    for k, v in d["extras"]["config4_fp32_rect_and_transposed"].items():
        print("  ", k, v)
    print("  f2", d["extras"].get("f2_unaligned_and_strided_operands"))
except Exception as e:
    print("  parse failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu --config5-size 0 > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mtm_tf32x3|split_kernel" -s 4 -c 2 -f -o gpurun_out/prof_3xtf32_final \
      python tools/one_call.py 3xtf32 8192 LLL > gpurun_out/ncu_3xtf32_final.log 2>&1; echo "ncu 3xtf32 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mtm_tf32x3" -s 2 -c 1 -f -o gpurun_out/prof_3xtf32_4096_tailsplit \
      python tools/one_call.py 3xtf32 4096 LLL > gpurun_out/ncu_3xtf32_4096.log 2>&1; echo "ncu 4096 exit $?"
