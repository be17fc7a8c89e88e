#!/bin/bash
# Round-2 final single-GPU session: the driver's own sequence (tests, smoke, both bench arms) + ncu evidence + sanitizer.
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "reference arm exit $?"; cut -c1-300 gpurun_out/bench_reference_arm.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
    print("  value", d["value"], "ms", d["ms_per_step"], "roof", d["roofline"]["frac"], d["roofline"].get("sustained", {}).get("frac"), "traffic", d["roofline"]["traffic"], d["roofline"]["traffic_stale"], "clocks", d["clocks"])
    print("  e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "pageable", d["e2e"]["pageable"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "launches", d["gpu_launches"])
    print("  config5", d["config5"]["tflops_with_broadcast"], d["config5"]["exact"], "watchdog", d.get("watchdog"))
    for row in d["extras"]["config2_fp32_square_sweep_LLL"]:
        print("  n", row["n"], {k: (v["tflops"], v["kernel"]) for k, v in row.items() if k != "n"})
    for k, v in d["extras"]["config4_fp32_rect_and_transposed"].items():
        print("  ", k, v)
except Exception as e:
    print("  parse failed", e)
PY
timeout 600 python tools/tune.py --families simt --sizes 512,1024,1536 --out gpurun_out/tune_simt_small.json > gpurun_out/tune_simt_small.log 2>&1; echo "tune simt exit $?"; grep -o '"shape": \[[0-9, ]*\].*"config": [a-z0-9]*.*"tflops": [0-9.]*' gpurun_out/tune_simt_small.log | sed 's/"family": "simt", //' | cut -c1-160
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu --config5-size 0 > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mtm_tf32x3|split_kernel" -s 4 -c 2 -f -o gpurun_out/prof_3xtf32_final \
      python tools/one_call.py 3xtf32 8192 LLL > gpurun_out/ncu_3xtf32_final.log 2>&1; echo "ncu 3xtf32 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mtm_ffma_tma" -s 2 -c 1 -f -o gpurun_out/prof_simt_final \
      python tools/one_call.py simt 8192 LLL > gpurun_out/ncu_simt_final.log 2>&1; echo "ncu simt exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mtm_dmma_tma" -s 2 -c 1 -f -o gpurun_out/prof_dmma_final \
      python tools/one_call.py dmma 8192 LLL > gpurun_out/ncu_dmma_final.log 2>&1; echo "ncu dmma exit $?"
bash tools/sanitize.sh 2>&1 | tail -10
# tile-walk group height (L2 locality of the wave) A/B, interleaved twice
for rep in 1 2; do for g in 4 6 8 12 16; do
  B200_TF32_GROUP=$g timeout 300 python - <<PY
import torch, sys
sys.path.insert(0, ".")
import openmp_blas_b200 as ob
n = 8192
a = torch.rand((n, n), device="cuda") * 2 - 1; b = torch.rand((n, n), device="cuda") * 2 - 1; c = torch.zeros((n, n), device="cuda")
ms = ob.bench_device(c, a, b, variant="3xtf32", config=0, warmup=5, iters=30)
print("group $g rep $rep: 8192^3", round(ms, 4), "ms", round(n * n * (2.0 * n - 1) / ms / 1e9, 1), "TFLOP/s")
PY
done; done
