#!/bin/bash
# 2-GPU (or N-GPU) validation of the NVLink replicator and the SUMMA split:  bash tools/nvlink_round.sh [N] [big]
N="${1:-2}"; BIG="${2:-16384}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/nv_gpus.txt 2>&1
nvidia-smi topo -m > gpurun_out/nv_topo.txt 2>&1
timeout 300 python -m pytest tests/test_replicate_gpu.py -x -q -m gpu > gpurun_out/nv_pytest.log 2>&1; echo "pytest replicate exit $?"; tail -3 gpurun_out/nv_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
    tools/multi_gpu_check.py --size 4096 --big-size $BIG --variants 3xtf32 --bcast nccl,nvlink --push-ctas 0,16,64 --summa \
    > gpurun_out/nv_check.out 2> gpurun_out/nv_check.err
echo "check exit $?"; grep '^{' gpurun_out/nv_check.out; grep '^#' gpurun_out/nv_check.err | cut -c1-300; tail -5 gpurun_out/nv_check.err | cut -c1-300
for mode in nccl nvlink; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N \
      bench.py --gpus $N --steps 30 --warmup 3 --bcast $mode > gpurun_out/nv_bench_$mode.out 2> gpurun_out/nv_bench_$mode.err
  echo "bench $mode exit $?"; grep '^{' gpurun_out/nv_bench_$mode.out | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config'].get('b_replication'), d['config'].get('k_chunks'))"
done
