#!/bin/bash
# r02x: PDL between the split pass and the MMA kernel + row chunks for tall problems — parity, then A/B timing.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02x_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r02x_pytest_gpu.log
timeout 600 python tools/ab_env.py --check --shapes 128,256,512,768,1024,2048,4096,8192,65536x1024x1024,16384x1024x1024,32768x2048x512 \
  --env "" B200_TF32_NO_PDL=1 B200_TF32_ROW_CHUNKS=0 > gpurun_out/r02x_ab_pdl_chunks.jsonl 2> gpurun_out/r02x_ab.err; echo "ab exit $?"
cat gpurun_out/r02x_ab_pdl_chunks.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['shape'], {k: (v['kernel'][-22:], v['us_best'], v['tflops_best']) for k, v in d.items() if isinstance(v, dict)}, d.get('exact_vs_fp64_rows'), d.get('identical'))
"
tail -3 gpurun_out/r02x_ab.err
timeout 300 python tools/ab_env.py --shapes 65536x1024x1024 --env B200_TF32_ROW_CHUNKS=0 B200_TF32_ROW_CHUNKS=2 B200_TF32_ROW_CHUNKS=4 B200_TF32_ROW_CHUNKS=8 B200_TF32_ROW_CHUNKS=16 >> gpurun_out/r02x_ab_pdl_chunks.jsonl 2>> gpurun_out/r02x_ab.err
timeout 300 python tools/ab_env.py --shapes 8192,4096 --env B200_TF32_ROW_CHUNKS=0 B200_TF32_ROW_CHUNKS=2 B200_TF32_ROW_CHUNKS=4 >> gpurun_out/r02x_ab_pdl_chunks.jsonl 2>> gpurun_out/r02x_ab.err
tail -3 gpurun_out/r02x_ab_pdl_chunks.jsonl
