#!/bin/bash
mkdir -p gpurun_out
for co in 1 0; do
echo "== split carve-out max: $co"
B200_TF32_SPLIT_CARVEOUT=$co timeout 600 python tools/ab_env.py --check --rounds 3 --shapes 128,512,1024,2048,8192,65536x1024x1024,16384x1024x1024 \
  --env "" B200_TF32_ROW_CHUNKS=0 B200_TF32_ROW_CHUNKS=2 B200_TF32_ROW_CHUNKS=8 B200_TF32_NO_PDL=1 2>> gpurun_out/r02y_ab.err | tee -a gpurun_out/r02y_ab_carveout_$co.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['shape'], {k[-14:]: (v['kernel'][-12:], v['us_best']) for k, v in d.items() if isinstance(v, dict)}, d.get('exact_vs_fp64_rows'), d.get('identical'))
"
done
tail -3 gpurun_out/r02y_ab.err
