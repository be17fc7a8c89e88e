#!/bin/bash
# r03f: ncu --set full of the double-tile kernel (AUTO at 8192^3) and of the 256x256 kernel beside it; launch list of a short bench
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mtm_tf32x3|split_kernel" -s 4 -c 2 -f -o gpurun_out/r03f_prof_3xtf32_double \
      python tools/one_call.py 3xtf32 8192 LLL > gpurun_out/r03f_ncu_double.log 2>&1; echo "ncu double exit $?"; tail -2 gpurun_out/r03f_ncu_double.log
B200_TF32_NO_DOUBLE_TILES=1 timeout 600 ncu --set full --clock-control none -k regex:"mtm_tf32x3" -s 2 -c 1 -f -o gpurun_out/r03f_prof_3xtf32_single \
      python tools/one_call.py 3xtf32 8192 LLL > gpurun_out/r03f_ncu_single.log 2>&1; echo "ncu single exit $?"; tail -2 gpurun_out/r03f_ncu_single.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03f_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-extras --no-cpu --config5-size 0 > gpurun_out/r03f_bench_under_ncu.log 2>&1; echo "launch list exit $?"
timeout 900 python bench.py > gpurun_out/r03f_bench.json 2> gpurun_out/r03f_bench.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r03f_bench.json').readline())
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'],d['roofline'].get('burst',{}).get('frac'),d['roofline']['kernel'],'clocks',d['clocks'])
for r in d['extras']['config2_fp32_square_sweep_LLL']: print(r['n'], r['simt']['tflops'], r['simt']['kernel'], r['3xtf32']['tflops'], r['3xtf32']['kernel'])
for k,v in d['extras']['config4_fp32_rect_and_transposed'].items(): print(k,v)
print(d['config5'])
"
