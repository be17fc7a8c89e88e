#!/bin/bash
# Round-end rehearsal on one GPU: what the driver runs (pytest -m gpu, smoke, both bench arms) plus the
# ncu launch list and one full capture per headline kernel.  Outputs under gpurun_out/.
mkdir -p gpurun_out
(time timeout 1800 python -m pytest tests/ -x -q -m gpu --durations=12) > gpurun_out/pytest_full.log 2>&1; echo "pytest full exit $?"; tail -22 gpurun_out/pytest_full.log | grep -E "passed|failed|s call|real"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.out 2> gpurun_out/bench_ref.err; echo "ref exit $?"
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"; tail -2 gpurun_out/bench_full.err
python bench.py --steps 20 --warmup 3 --variant simt --no-extras --no-cpu > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err; echo "bench simt exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?"
for spec in "3xtf32:mtm_tf32x3" "simt:mtm_ffma" "dmma:mtm_dmma_tma"; do
  fam=${spec%%:*}; pat=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -f -o gpurun_out/prof_final_$fam \
      python tools/one_call.py $fam > gpurun_out/ncu_final_$fam.log 2>&1; echo "ncu $fam exit $?"
done
