#!/bin/bash
# Round-2 first contact: the reworked 3xTF32 path (raw operand as hi, MN-major descriptors, TMA reduce epilogue).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python tools/tf32_probe.py > gpurun_out/probe_default.log 2>&1; RC=$?; echo "probe default exit $RC"; grep -E "^==|^BAD|^FAIL|^HANG" gpurun_out/probe_default.log | cut -c1-330 | head -40
if [ "$RC" != "0" ]; then
  B200_TF32_MN_LBO=1024 B200_TF32_MN_SBO=4096 TF32_PROBE_LAYOUTS=LL TF32_PROBE_CFGS=1 timeout 300 python tools/tf32_probe.py > gpurun_out/probe_swapped.log 2>&1; echo "probe swapped LBO/SBO exit $?"; grep -E "^==|^BAD|^FAIL|^HANG" gpurun_out/probe_swapped.log | cut -c1-330 | head -12
  B200_TF32_NO_TMA_EPI=1 TF32_PROBE_LAYOUTS=LF,LL TF32_PROBE_CFGS=1,0 timeout 300 python tools/tf32_probe.py > gpurun_out/probe_noepi.log 2>&1; echo "probe no-TMA-epilogue exit $?"; grep -E "^==|^BAD|^FAIL|^HANG" gpurun_out/probe_noepi.log | cut -c1-330 | head -12
  B200_TF32_FORCE_PACKED=1 TF32_PROBE_LAYOUTS=LL TF32_PROBE_CFGS=1,0 timeout 300 python tools/tf32_probe.py > gpurun_out/probe_packed.log 2>&1; echo "probe packed exit $?"; grep -E "^==|^BAD|^FAIL|^HANG" gpurun_out/probe_packed.log | cut -c1-330 | head -12
  B200_TF32_ROUND_HI=1 TF32_PROBE_LAYOUTS=LL TF32_PROBE_CFGS=1 timeout 300 python tools/tf32_probe.py > gpurun_out/probe_roundhi.log 2>&1; echo "probe round-hi exit $?"; grep -E "^==|^BAD|^FAIL|^HANG" gpurun_out/probe_roundhi.log | cut -c1-330 | head -12
fi
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 -x -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
for env in "" "B200_TF32_NO_TMA_EPI=1" "B200_TF32_FORCE_PACKED=1" "B200_TF32_ROUND_HI=1"; do
  tag=${env:-default}
  env $env timeout 600 python bench.py --steps 20 --no-cpu --no-extras > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench [$tag] exit $?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
    print("  value", d["value"], "ms", d["ms_per_step"], "roof", d["roofline"]["frac"], d["roofline"]["ms_per_launch"], "e2e", d["e2e"]["value"], d["config"]["kernel"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("  parse failed", e)
PY
done
timeout 900 python bench.py --steps 20 --no-cpu > gpurun_out/bench_extras.json 2> gpurun_out/bench_extras.err; echo "bench extras exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_extras.json").read().strip().splitlines()[-1])
    for row in d["extras"]["config2_fp32_square_sweep_LLL"]:
        print("  n", row["n"], {k: (v["tflops"], v["kernel"]) for k, v in row.items() if k != "n"})
    for k, v in d["extras"]["config4_fp32_rect_and_transposed"].items():
        print("  ", k, v)
except Exception as e:
    print("  parse failed", e)
PY
