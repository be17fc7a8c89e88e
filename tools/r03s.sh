#!/bin/bash
# r03s: fresh ncu --set full records of the TMA-fed FFMA and DMMA kernels on the final tree (traffic stamps)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none -k regex:"mtm_ffma_tma" -s 2 -c 1 -f -o gpurun_out/r03s_prof_simt python tools/one_call.py simt 8192 LLL > gpurun_out/r03s_ncu_simt.log 2>&1; echo "ncu simt exit $?"
timeout 300 ncu --set full --clock-control none -k regex:"mtm_dmma_tma" -s 2 -c 1 -f -o gpurun_out/r03s_prof_dmma python tools/one_call.py dmma 8192 LLL > gpurun_out/r03s_ncu_dmma.log 2>&1; echo "ncu dmma exit $?"
