#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "mgpu or slab or cpp_front_end_spreads" > gpurun_out/pytest_mgpu.log 2>&1; echo "pytest mgpu exit $?"; grep -c "^\[mgpu\]" gpurun_out/pytest_mgpu.log; tail -3 gpurun_out/pytest_mgpu.log | cut -c1-300
