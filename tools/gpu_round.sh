#!/bin/bash
# One batched GPU session: smoke, parity tests, tuning sweep, bench, ncu launch list + full captures.
# Usage (from the repo root on the GPU box):  bash tools/gpu_round.sh [stages...]   (default: all)
mkdir -p gpurun_out
STAGES="${@:-tf32probe tests smoke tune bench ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Flags" | cut -c1-400 >> gpurun_out/gpu.txt
for st in $STAGES; do
case $st in
smoke) timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt ;;
tf32probe) timeout 1200 python tools/tf32_probe.py > gpurun_out/tf32_probe.log 2>&1; echo "tf32probe exit $?" | tee -a gpurun_out/summary.txt; tail -40 gpurun_out/tf32_probe.log ;;
tests) timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -s -k "not 3xtf32 and not test_cpp_front_end and not full_size and not golden_vectors_gpu" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest(non-tf32) exit $?" | tee -a gpurun_out/summary.txt; tail -5 gpurun_out/pytest_gpu.log
       timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -s -k "3xtf32 or test_cpp_front_end or full_size or golden_vectors_gpu" > gpurun_out/pytest_gpu_tf32.log 2>&1; echo "pytest(tf32) exit $?" | tee -a gpurun_out/summary.txt; tail -8 gpurun_out/pytest_gpu_tf32.log ;;
tune)  timeout 900 python tools/tune.py --out gpurun_out/tune.json > gpurun_out/tune.log 2>&1; echo "tune exit $?" | tee -a gpurun_out/summary.txt ;;
bench) timeout 900 python bench.py --variant simt --no-cpu > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err; echo "bench simt exit $?" | tee -a gpurun_out/summary.txt; cut -c1-600 gpurun_out/bench_simt.json
       timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/summary.txt; cut -c1-1500 gpurun_out/bench.json ;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?" | tee -a gpurun_out/summary.txt
  for fam in ${NCU_FAMS:-simt dfma dmma 3xtf32}; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:mtm_ -s 2 -c 1 -f -o gpurun_out/prof_$fam \
      python tools/one_call.py $fam > gpurun_out/ncu_$fam.log 2>&1; echo "ncu $fam exit $?" | tee -a gpurun_out/summary.txt
  done ;;
esac
done
ls -la gpurun_out | head -40
