#!/bin/bash
# r03t: closing gate on the final tree: full GPU tests (with the dependent-launch chain test) + smoke
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r03t_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r03t_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03t_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/r03t_smoke.log
