#!/bin/bash
# r03b: chained PDL + read-only bulk wait at exit + AUTO thresholds (3xTF32 from 96^3 on, FFMA-TMA 64x128 at ~1024^2): full validation.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r03b_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r03b_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03b_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r03b_smoke.log
timeout 300 python tools/ab_env.py --check --rounds 3 --shapes 128,256,512,768,1024,2048 --env "" B200_TF32_NO_PDL=1 2> gpurun_out/r03b_ab.err | tee gpurun_out/r03b_ab.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['shape'], {k[-9:]: (v['kernel'][7:], v['us_best'], v['tflops_best']) for k, v in d.items() if isinstance(v, dict)}, d.get('exact_vs_fp64_rows'), d.get('identical'))
"
timeout 900 python bench.py > gpurun_out/r03b_bench.json 2> gpurun_out/r03b_bench.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r03b_bench.json').readline())
print('value',d['value'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'],d['roofline'].get('sustained',{}).get('frac'),'stale',d['roofline'].get('traffic_stale'))
for r in d['extras']['config2_fp32_square_sweep_LLL']: print(r['n'], r['simt']['tflops'], r['simt']['kernel'], r['3xtf32']['tflops'], r['3xtf32']['kernel'])
print(d['extras']['config4_fp32_rect_and_transposed'])
"
