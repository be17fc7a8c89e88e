#!/bin/bash
N="${1:-2}"
mkdir -p gpurun_out
P=29650
for mode in "--fused --push-ctas 0" "--fused --push-ctas 16" "--push-ctas 0"; do
  P=$((P+1)); tag=$(echo $mode | tr -d ' -')
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $N --steps 30 --warmup 3 --bcast nvlink $mode > gpurun_out/g2_bench_$tag.out 2> gpurun_out/g2_bench_$tag.err
  echo "bench $tag exit $?"; grep '^{' gpurun_out/g2_bench_$tag.out | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config'].get('ms_per_step_by_rank'), d['config'].get('b_replication'))"
done
