#!/bin/bash
# r03p: final single-GPU validation of the tree: full GPU tests, smoke, probe, bench (driver's command line), reference arm,
# launch list of a short bench, ncu --set full of the two big pair kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r03p_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r03p_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03p_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r03p_smoke.log
timeout 900 python bench.py > gpurun_out/r03p_bench.json 2> gpurun_out/r03p_bench.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r03p_bench.json').readline())
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'],d['roofline'].get('burst',{}).get('frac'),d['roofline']['kernel'],'stale',d['roofline'].get('traffic_stale'),'clocks',d['clocks'])
for r in d['extras']['config2_fp32_square_sweep_LLL']: print(r['n'], r['simt']['tflops'], r['simt']['kernel'], r['3xtf32']['tflops'], r['3xtf32']['kernel'])
for k,v in d['extras']['config4_fp32_rect_and_transposed'].items(): print(k,v)
print(d['extras']['f2_unaligned_and_strided_operands'])
print(d['config5'])
"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r03p_bench_reference_arm.json 2>/dev/null; echo "reference arm exit $?"; cut -c1-300 gpurun_out/r03p_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03p_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-extras --no-cpu --config5-size 0 > gpurun_out/r03p_bench_under_ncu.log 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mtm_tf32x3|split_kernel" -s 4 -c 2 -f -o gpurun_out/r03p_prof_3xtf32_cfg9 \
      python tools/one_call.py 3xtf32 8192 LLL 9 > gpurun_out/r03p_ncu_cfg9.log 2>&1; echo "ncu cfg9 exit $?"
timeout 600 ncu --set full --clock-control none -k regex:"mtm_tf32x3" -s 2 -c 1 -f -o gpurun_out/r03p_prof_3xtf32_cfg0 \
      python tools/one_call.py 3xtf32 8192 LLL 0 > gpurun_out/r03p_ncu_cfg0.log 2>&1; echo "ncu cfg0 exit $?"
timeout 600 ncu --set full --clock-control none -k regex:"mtm_tf32x3" -s 2 -c 1 -f -o gpurun_out/r03p_prof_3xtf32_cfg4 \
      python tools/one_call.py 3xtf32 4096 LLL 4 > gpurun_out/r03p_ncu_cfg4.log 2>&1; echo "ncu cfg4 exit $?"
