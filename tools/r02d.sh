#!/bin/bash
# Round-2 third single-GPU contact: split-K, pageable staging, ncu evidence.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; grep -E "split-K|\[mgpu\]" gpurun_out/pytest_gpu.log | cut -c1-200 | head -20; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python tools/tune.py --families 3xtf32 --sizes 256,512,768,1024,1536,2048 --shapes 512x512x8192,256x4096x4096,4096x256x2048 --out gpurun_out/tune_small.json > gpurun_out/tune_small.log 2>&1; echo "tune small exit $?"
python - <<'PY'
import json
from collections import defaultdict
try:
    rows = json.load(open("gpurun_out/tune_small.json"))["rows"]
    t = defaultdict(list)
    for r in rows:
        if "tflops" in r: t[tuple(r["shape"])].append((r["tflops"], r["ms"], r["config"], r["split_k"], r["name"]))
    for sh, v in t.items():
        v.sort(reverse=True)
        auto = [x for x in v if x[2] is None]
        print(sh, "auto:", auto[0] if auto else None, "| best:", v[:3])
except Exception as e:
    print("parse failed", e)
PY
/usr/bin/g++ -std=c++20 -O2 -fopenmp -Iinclude/compat -Iinclude tools/mtm_mgpu_check.cpp -o /tmp/mtm_mgpu_check -Lopenmp-blas_b200 -lb200mtm -Wl,-rpath,$PWD/openmp-blas_b200 || echo "compile failed"
for env in "" "B200_NO_PAGEABLE_STAGING=1" "B200_COPY_THREADS=3" "B200_COPY_THREADS=7" "B200_COPY_THREADS=11"; do
  env $env timeout 600 /tmp/mtm_mgpu_check --size 8192 --devices 1 --calls 4 > gpurun_out/pageable_tmp.json 2>&1; echo "pageable 8192 [$env] exit $?: $(cut -c1-330 gpurun_out/pageable_tmp.json)"
  cat gpurun_out/pageable_tmp.json >> gpurun_out/pageable_ab.jsonl
done
timeout 900 python bench.py --steps 30 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
    print("  value", d["value"], "ms", d["ms_per_step"], "roof", d["roofline"]["frac"], d["roofline"].get("sustained"), "clocks", d["clocks"])
    print("  e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "pageable", d["e2e"].get("pageable"))
    print("  cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "config5", d["config5"]["tflops_with_broadcast"], d["config5"]["exact"])
    for row in d["extras"]["config2_fp32_square_sweep_LLL"]:
        print("  n", row["n"], {k: (v["tflops"], v["kernel"]) for k, v in row.items() if k != "n"})
    for k, v in d["extras"]["config4_fp32_rect_and_transposed"].items():
        print("  ", k, v)
    print("  f64", d["extras"]["config3_fp64_8192"])
except Exception as e:
    print("  parse failed", e)
PY
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu --config5-size 0 > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?"
for spec in "3xtf32 8192 LLL" "3xtf32 8192 LLF" "simt 8192 LLL"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mtm_tf32x3|split_kernel|mtm_ffma" -s 4 -c 2 -f -o gpurun_out/prof_$1_$3 \
      python tools/one_call.py $1 $2 $3 > gpurun_out/ncu_$1_$3.log 2>&1; echo "ncu $spec exit $?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mtm_tf32x3|split_kernel" -s 4 -c 2 -f -o gpurun_out/prof_3xtf32_1024 \
      python tools/one_call.py 3xtf32 1024 LLL > gpurun_out/ncu_3xtf32_1024.log 2>&1; echo "ncu 1024 exit $?"
ls -la gpurun_out | head -50
