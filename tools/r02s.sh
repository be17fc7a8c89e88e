#!/bin/bash
# pageable staging after the copy-pool change: C++ front-end on make_tensor storage + the slab / mgpu tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "mgpu or slab or cpp" > gpurun_out/pytest_host_paths.log 2>&1; echo "pytest host paths exit $?"; tail -2 gpurun_out/pytest_host_paths.log
/usr/bin/g++ -std=c++20 -O2 -fopenmp -Iinclude/compat -Iinclude tools/mtm_mgpu_check.cpp -o /tmp/mtm_mgpu_check -Lopenmp-blas_b200 -lb200mtm -Wl,-rpath,$PWD/openmp-blas_b200 || echo "compile failed"
for args in "--size 8192 --devices 1 --calls 4" "--size 8192 --calls 4" "--size 16384 --calls 2"; do
  timeout 600 /tmp/mtm_mgpu_check $args; echo "exit $?"
done
timeout 600 python bench.py --steps 10 --no-extras --no-cpu --config5-size 0 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('bench quick', d['value'], d['e2e']['value'], d['e2e']['pageable'])"
