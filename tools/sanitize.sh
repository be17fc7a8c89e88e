#!/bin/bash
# compute-sanitizer over the C++ front-end tests (no Python in the sanitized process):
#   memcheck  — every kernel family, all the reference's small cases (sizes 2..31, 8 layouts, 2 dtypes, mtv/vtm, transpose)
#   racecheck — shared-memory hazards of the register-staged / DMMA / mtv / transpose kernels
#   synccheck — barrier usage of the TMA / tcgen05 kernels
# Usage (GPU box, repo root): bash tools/sanitize.sh      -> gpurun_out/sanitize_*.log, summary on stdout
mkdir -p gpurun_out /tmp/san
L=openmp-blas_b200
for t in test_mtm test_mtv test_trans; do
  /usr/bin/g++ -std=c++20 -O2 -Iinclude/compat -Iinclude tests/cpp/$t.cpp -o /tmp/san/$t -L$L -lb200mtm -Wl,-rpath,$PWD/$L || exit 1
done
CS="compute-sanitizer --error-exitcode 9 --print-limit 20"
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/sanitize_$name.log 2>&1; rc=$?; echo "$name exit $rc : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|checks,' gpurun_out/sanitize_$name.log | tr '\n' ' ')"; }
run memcheck_mtm_simt   $CS --tool memcheck /tmp/san/test_mtm 1
run memcheck_mtm_tf32   $CS --tool memcheck /tmp/san/test_mtm 2
run memcheck_mtm_dmma   $CS --tool memcheck /tmp/san/test_mtm 4
run memcheck_mtv        $CS --tool memcheck /tmp/san/test_mtv
run memcheck_trans      $CS --tool memcheck /tmp/san/test_trans
run racecheck_mtm_simt  $CS --tool racecheck /tmp/san/test_mtm 1
run racecheck_mtm_dmma  $CS --tool racecheck /tmp/san/test_mtm 4
run racecheck_trans     $CS --tool racecheck /tmp/san/test_trans
run synccheck_mtm_tf32  $CS --tool synccheck /tmp/san/test_mtm 2
