#!/bin/bash
# r03c: separate accumulators per product (no back-to-back MMAs on one accumulator) — probe + A/B.
mkdir -p gpurun_out
for n in 2 3; do
B200_TF32_NACC=$n timeout 300 python tools/tf32_probe.py > gpurun_out/r03c_tf32_probe_nacc$n.log 2>&1; echo "probe nacc=$n exit $?"; grep -c "^BAD" gpurun_out/r03c_tf32_probe_nacc$n.log; grep "A/B layouts" gpurun_out/r03c_tf32_probe_nacc$n.log
done
show='
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d["shape"], d["layout"], {k[-6:]: (v["kernel"][7:], v["us_best"], v["tflops_best"]) for k, v in d.items() if isinstance(v, dict)}, d.get("exact_vs_fp64_rows"), d.get("identical"))
'
timeout 600 python tools/ab_env.py --check --rounds 3 --shapes 512,1024,1536,2048,4096,8192,65536x1024x1024 \
  --env B200_TF32_NACC=1 B200_TF32_NACC=2 B200_TF32_NACC=3 2>> gpurun_out/r03c_ab.err | tee -a gpurun_out/r03c_ab_nacc.jsonl | python -c "$show"
for cfg in 1 5 4; do
timeout 300 python tools/ab_env.py --rounds 3 --config $cfg --shapes 2048,4096 \
  --env B200_TF32_NACC=1 B200_TF32_NACC=2 B200_TF32_NACC=3 2>> gpurun_out/r03c_ab.err | tee -a gpurun_out/r03c_ab_nacc.jsonl | python -c "$show"
done
tail -3 gpurun_out/r03c_ab.err
