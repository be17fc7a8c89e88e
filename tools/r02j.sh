#!/bin/bash
# Round-2: validation after the 4-stage narrow-tile ring and the fused clean-up; small-shape timings.
mkdir -p gpurun_out
show() { grep -E "^==|^BAD|^FAIL|^HANG" "$1" | cut -c1-300 | head -${2:-8}; }
timeout 900 python tools/tf32_probe.py > gpurun_out/probe_all.log 2>&1; echo "probe all cfgs exit $?"; show gpurun_out/probe_all.log 10
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
bash tools/r02i.sh 2>&1 | grep "us/call" | tee gpurun_out/small_shapes.txt | cut -c1-200
