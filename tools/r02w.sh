#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "every_tile_config or edge_shapes or randomized" > gpurun_out/pytest_cfg.log 2>&1; echo "pytest configs exit $?"; tail -2 gpurun_out/pytest_cfg.log
timeout 600 python tools/tune.py --families simt --sizes 512,768,1024,1280,1536,2048 --shapes 1024x4096x1024,4096x1024x512 --out gpurun_out/tune_simt_small2.json > gpurun_out/tune_simt_small2.log 2>&1; echo "tune simt exit $?"
python - <<'PY'
import json
from collections import defaultdict
rows = json.load(open("gpurun_out/tune_simt_small2.json"))["rows"]
t = defaultdict(list)
for r in rows:
    if "tflops" in r: t[tuple(r["shape"])].append((r["tflops"], r["config"], r["name"]))
for sh, v in t.items():
    v.sort(reverse=True, key=lambda x: x[0])
    auto = [x for x in v if x[1] is None]
    print(sh, "auto:", auto[0] if auto else None, "| best:", v[:3])
PY
timeout 600 python bench.py --steps 10 --no-extras --no-cpu --config5-size 0 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('bench quick', d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('pcie_copy_floor_ms'), d['e2e'].get('frac_of_copy_floor'), d['roofline']['traffic_stale'])"
