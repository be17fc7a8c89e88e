#!/bin/bash
# Short round-end check on one GPU: the GPU test suite, smoke(), and one bench line without the secondary sweeps.
mkdir -p gpurun_out
(time timeout 400 python -m pytest tests/ -x -q -m gpu --durations=8) > gpurun_out/pytest_full.log 2>&1; echo "pytest full exit $?"; tail -16 gpurun_out/pytest_full.log | grep -E "passed|failed|error|s call|real"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench_short.json
