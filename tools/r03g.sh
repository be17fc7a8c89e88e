#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/small_call_probe.py 2>&1 | tee gpurun_out/r03g_small_call_probe.txt
