#!/bin/bash
# 1-GPU: gated product check; then (N>1) fused vs chunked multi-GPU step.
N="${1:-1}"
mkdir -p gpurun_out
GATED_ONLY_PRESET=1 timeout 200 python tools/gated_check.py > gpurun_out/gated_check.out 2> gpurun_out/gated_check.err; RC=$?; echo "gated_check exit $RC"; tail -c 2500 gpurun_out/gated_check.out; grep "^#" gpurun_out/gated_check.err | cut -c1-400; tail -3 gpurun_out/gated_check.err | cut -c1-300
if [ "$RC" != "0" ]; then echo "gated check failed: skipping the multi-GPU part"; exit 1; fi
if [ "$N" != "1" ]; then
  for mode in "--fused" ""; do
    tag=$( [ -n "$mode" ] && echo fused || echo chunked )
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N \
        bench.py --gpus $N --steps 30 --warmup 3 --bcast nvlink $mode > gpurun_out/gated_bench_$tag.out 2> gpurun_out/gated_bench_$tag.err
    echo "bench $tag exit $?"; grep '^{' gpurun_out/gated_bench_$tag.out | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config'].get('b_replication'), d['gpu_launches'])"; tail -3 gpurun_out/gated_bench_$tag.err | cut -c1-300
  done
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 \
      tools/multi_gpu_check.py --size 4096 --big-size 16384 --variants 3xtf32 --bcast nvlink --fused > gpurun_out/gated_mgpu.out 2> gpurun_out/gated_mgpu.err
  echo "mgpu fused exit $?"; grep '^{' gpurun_out/gated_mgpu.out | cut -c1-1500; grep -v '^#' gpurun_out/gated_mgpu.err | tail -4 | cut -c1-300
fi
timeout 200 python tools/gated_check.py > gpurun_out/gated_check_full.out 2> gpurun_out/gated_check_full.err; echo "gated_check (all cases, concurrent arrival) exit $?"; grep "^#" gpurun_out/gated_check_full.err | cut -c1-330
