#!/bin/bash
# N-GPU confirmation: NCCL broadcast vs own NVLink multicast push (SM kernel / copy engines), config 5, SUMMA auto grid.
N="${1:-8}"; BIG="${2:-32768}"; STEPS="${3:-30}"
mkdir -p gpurun_out
PORT=29600
for mode in "nccl 0" "nvlink 0" "nvlink -1"; do
  set -- $mode; PORT=$((PORT+1))
  tag="$1$2"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps $STEPS --warmup 3 --bcast $1 --push-ctas $2 > gpurun_out/nv8_bench_$tag.out 2> gpurun_out/nv8_bench_$tag.err
  echo "bench $tag exit $?"; grep '^{' gpurun_out/nv8_bench_$tag.out | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config'].get('b_replication'), d['config'].get('k_chunks'))"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    tools/multi_gpu_check.py --size 4096 --big-size $BIG --variants 3xtf32 --bcast nccl,nvlink --push-ctas 0,-1 --push-bw 32,-1 --summa --summa-auto-only \
    > gpurun_out/nv8_check.out 2> gpurun_out/nv8_check.err
echo "check exit $?"; grep '^{' gpurun_out/nv8_check.out | cut -c1-3000; grep -v '^#' gpurun_out/nv8_check.err | tail -5 | cut -c1-300
