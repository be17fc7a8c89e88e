#!/bin/bash
# Round-2 second contact: MN-major operand feed with the 32-byte-atom swizzle; then tests and benches.
mkdir -p gpurun_out
show() { grep -E "^==|^BAD|^FAIL|^HANG" "$1" | cut -c1-300 | head -${2:-8}; }
TF32_PROBE_LAYOUTS=LL,FL,FF timeout 900 python tools/tf32_probe.py > gpurun_out/probe_default.log 2>&1; RC=$?; echo "probe default exit $RC"; show gpurun_out/probe_default.log 14
if [ "$RC" != "0" ]; then
  for v in "B200_TF32_MN_LBO=512 B200_TF32_MN_SBO=4096" "B200_TF32_MN_TMA_SWIZZLE=5" "B200_TF32_MN_LAYOUT=2" "B200_TF32_MN_TMA_SWIZZLE=3 B200_TF32_MN_LAYOUT=1" "B200_TF32_MN_SBO=1024" "B200_TF32_MN_LBO=4096 B200_TF32_MN_SBO=128"; do
    env $v TF32_PROBE_LAYOUTS=LL TF32_PROBE_CFGS=1 timeout 300 python tools/tf32_probe.py > "gpurun_out/probe_alt.log" 2>&1; echo "probe [$v] exit $?"; show gpurun_out/probe_alt.log 5
  done
fi
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; grep -E "tolerance|extreme" gpurun_out/pytest_gpu.log | cut -c1-220 | head -60; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
for env in "" "B200_TF32_NO_TMA_EPI=1" "B200_TF32_FORCE_PACKED=1" "B200_TF32_ROUND_HI=1"; do
  tag=${env:-default}
  env $env timeout 600 python bench.py --steps 30 --no-cpu --no-extras --config5-size 0 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench [$tag] exit $?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
    print("  value", d["value"], "ms", d["ms_per_step"], "roof", d["roofline"]["frac"], d["roofline"]["ms_per_launch"], "e2e", d["e2e"]["value"], d["e2e"].get("pageable"), d["config"]["kernel"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("  parse failed", e)
PY
done
timeout 900 python tools/tune.py --families 3xtf32 --sizes 512,1024,2048,3072,4096,8192 --shapes 65536x1024x1024,8192x8192x1024,1024x8192x8192 --out gpurun_out/tune.json > gpurun_out/tune.log 2>&1; echo "tune exit $?"; tail -80 gpurun_out/tune.log | cut -c1-250
timeout 900 python bench.py --steps 20 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
    print("  value", d["value"], "cpu", d["cpu_baseline"], "\n  config5", d["config5"])
    for row in d["extras"]["config2_fp32_square_sweep_LLL"]:
        print("  n", row["n"], {k: (v["tflops"], v["kernel"]) for k, v in row.items() if k != "n"})
    for k, v in d["extras"]["config4_fp32_rect_and_transposed"].items():
        print("  ", k, v)
except Exception as e:
    print("  parse failed", e)
PY
