#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 10 --no-extras --no-cpu --config5-size 0 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('bench quick', d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['pageable'])"
