#!/bin/bash
mkdir -p gpurun_out
timeout 120 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu --config5-size 0 > gpurun_out/r03u_bench_quick.json 2> gpurun_out/r03u_bench_quick.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r03u_bench_quick.json').readline())
print(d['value'], d['roofline'])
"
