#!/bin/bash
# r03a: concatenated-B form of the 1-CTA tiles (2 MMAs per k step) — probe, A/B, full tests.
mkdir -p gpurun_out
timeout 300 python tools/tf32_probe.py > gpurun_out/r03a_tf32_probe.log 2>&1; echo "probe exit $?"; tail -4 gpurun_out/r03a_tf32_probe.log
show='
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d["shape"], d["layout"], {k[-9:]: (v["kernel"][7:], v["us_best"], v["tflops_best"]) for k, v in d.items() if isinstance(v, dict)}, d.get("exact_vs_fp64_rows"), d.get("identical"))
'
for lay in LLL LLF FFF; do
timeout 300 python tools/ab_env.py --check --rounds 3 --layout $lay --shapes 128,256,512,768,1024,1536,2048,1024x4096x1024,512x512x8192,300x260x520 \
  --env "" B200_TF32_CONCAT=0 2>> gpurun_out/r03a_ab.err | tee -a gpurun_out/r03a_ab_concat.jsonl | python -c "$show"
done
for cfg in 1 5; do
timeout 300 python tools/ab_env.py --check --rounds 3 --config $cfg --shapes 1024,2048,4096,8192x8192x1024 \
  --env "" B200_TF32_CONCAT=0 2>> gpurun_out/r03a_ab.err | tee -a gpurun_out/r03a_ab_concat.jsonl | python -c "$show"
done
tail -3 gpurun_out/r03a_ab.err
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r03a_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r03a_pytest_gpu.log
