#!/bin/bash
# 1 -> N GPU scaling of bench.py (as the driver runs it) plus the config-5 check at the largest N.
# Usage: bash tools/scale_run.sh "1 2 4 8" [steps]
mkdir -p gpurun_out
NS="${1:-1 2 4 8}"; STEPS="${2:-30}"
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/scale_gpus.txt 2>&1
PORT=29520
for n in $NS; do
  PORT=$((PORT+1))
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps $STEPS --warmup 3 --no-extras --no-cpu > gpurun_out/scale_n1.out 2> gpurun_out/scale_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --gpus $n --steps $STEPS --warmup 3 > gpurun_out/scale_n$n.out 2> gpurun_out/scale_n$n.err
  fi
  echo "N=$n exit $?"; grep '^{' gpurun_out/scale_n$n.out | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'] if d.get('e2e') else None)"
done
last=$(echo $NS | awk '{print $NF}')
if [ "$last" != "1" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $last --master-addr 127.0.0.1 --master-port 29540 \
      tools/multi_gpu_check.py --size 4096 --big-size 32768 > gpurun_out/mgpu${last}_check.out 2> gpurun_out/mgpu${last}_check.err
  echo "check exit $?"; grep '^{' gpurun_out/mgpu${last}_check.out
fi
