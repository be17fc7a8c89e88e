#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    tools/multi_gpu_check.py --size 8192 --big-size 0 --bcast nccl > gpurun_out/r03v_mgpu_check_n2.out 2> gpurun_out/r03v_mgpu_check_n2.err
echo "multi_gpu_check exit $?"; grep '^{' gpurun_out/r03v_mgpu_check_n2.out | cut -c1-500; tail -2 gpurun_out/r03v_mgpu_check_n2.err | cut -c1-200
