#!/bin/bash
# r03j: where the narrow tile configs spend their k-blocks (ncu, tensor pipe activity), cuBLAS yardsticks
mkdir -p gpurun_out
timeout 120 python tools/cublas_yardstick.py | tee gpurun_out/r03j_cublas_yardstick.json
for cfg in 4 5 1; do
timeout 300 ncu --set full --clock-control none -k regex:"mtm_tf32x3" -s 2 -c 1 -f -o gpurun_out/r03j_prof_cfg$cfg \
      python tools/one_call.py 3xtf32 4096 LLL $cfg > gpurun_out/r03j_ncu_cfg$cfg.log 2>&1; echo "ncu cfg $cfg exit $?"; tail -1 gpurun_out/r03j_ncu_cfg$cfg.log | cut -c1-200
done
