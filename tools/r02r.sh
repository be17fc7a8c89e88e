#!/bin/bash
# closing capture: the ncu record the traffic stamp refers to, on the final kernel sources
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "tail_split or full_size or tolerance_at_baseline or split_k" > gpurun_out/pytest_subset.log 2>&1; echo "pytest subset exit $?"; tail -2 gpurun_out/pytest_subset.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mtm_tf32x3|split_kernel" -s 4 -c 2 -f -o gpurun_out/prof_3xtf32_final \
      python tools/one_call.py 3xtf32 8192 LLL > gpurun_out/ncu_3xtf32_final.log 2>&1; echo "ncu 3xtf32 exit $?"
