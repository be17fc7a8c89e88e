#!/bin/bash
# Where does the gated product's time go?  (a) ungated, default carve-out  (b) ungated, full carve-out
# (c) gated with every panel pre-arrived (mechanics only).  One GPU.
mkdir -p gpurun_out
GATED_TIMING=1 timeout 120 python tools/gated_check.py > gpurun_out/ab_default.out 2> gpurun_out/ab_default.err; echo "default carve-out exit $?"; grep "^#" gpurun_out/ab_default.err | cut -c1-330
B200_TF32_MAX_CARVEOUT=1 GATED_TIMING=1 timeout 120 python tools/gated_check.py > gpurun_out/ab_max.out 2> gpurun_out/ab_max.err; echo "max carve-out exit $?"; grep "^#" gpurun_out/ab_max.err | cut -c1-330
