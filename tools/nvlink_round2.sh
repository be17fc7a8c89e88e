#!/bin/bash
# second N-GPU validation round: SUMMA diagnostics, push bandwidth (SM vs copy engines), replica depth.
N="${1:-2}"; BIG="${2:-16384}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mtm_gpu.py -x -q -m gpu -k "k_panel_sequence" > gpurun_out/nv2_pytest.log 2>&1; echo "pytest panels exit $?"; tail -12 gpurun_out/nv2_pytest.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 \
    tools/multi_gpu_check.py --size 4096 --big-size $BIG --variants 3xtf32 --bcast nvlink --push-ctas 16,8,-1 --push-bw 4,8,16,32,64,-1 --summa \
    > gpurun_out/nv2_check.out 2> gpurun_out/nv2_check.err
echo "check exit $?"; grep '^{' gpurun_out/nv2_check.out; grep '^#' gpurun_out/nv2_check.err | cut -c1-400; grep -v '^#' gpurun_out/nv2_check.err | tail -8 | cut -c1-300
for mode in nvlink nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2958$N \
      bench.py --gpus $N --steps 30 --warmup 3 --bcast $mode > gpurun_out/nv2_bench_$mode.out 2> gpurun_out/nv2_bench_$mode.err
  echo "bench $mode exit $?"; grep '^{' gpurun_out/nv2_bench_$mode.out | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config'].get('b_replication'), d['config'].get('k_chunks'))"
done
