#!/bin/bash
# r02z: split pass chained to the previous call by PDL as well; AUTO crossover; small-shape tile sweep after the PDL change.
mkdir -p gpurun_out
timeout 600 python tools/ab_env.py --check --rounds 3 --shapes 128,256,512,768,1024,2048,4096,8192,65536x1024x1024 \
  --env "" B200_TF32_NO_PDL=2 B200_TF32_NO_PDL=1 2> gpurun_out/r02z_ab.err | tee gpurun_out/r02z_ab_pdl.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['shape'], {k[-14:]: (v['kernel'][-12:], v['us_best']) for k, v in d.items() if isinstance(v, dict)}, d.get('exact_vs_fp64_rows'), d.get('identical'))
"
tail -3 gpurun_out/r02z_ab.err
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02z_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r02z_pytest_gpu.log
timeout 600 python tools/auto_crossover.py --out gpurun_out/r02z_auto_crossover.jsonl > gpurun_out/r02z_auto.log 2>&1; echo "crossover exit $?"
timeout 900 python tools/tune.py --families 3xtf32 --sizes 512,768,1024,1536,2048 --shapes 1024x4096x1024,512x512x8192 --out gpurun_out/r02z_tune_small.json > gpurun_out/r02z_tune_small.log 2>&1; echo "tune exit $?"
