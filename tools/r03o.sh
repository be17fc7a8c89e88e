#!/bin/bash
# r03o: tile-config sweep after the pair fix, for the AUTO table
mkdir -p gpurun_out
timeout 900 python tools/tune.py --families 3xtf32 --sizes 512,768,1024,1280,1536,2048,2560,3072 --shapes 1024x4096x1024,512x512x8192,4096x1024x2048,2048x2048x512,16384x1024x1024 --out gpurun_out/r03o_tune.json > gpurun_out/r03o_tune.log 2>&1; echo "tune exit $?"
python - <<'PY'
import json
rows = json.load(open("gpurun_out/r03o_tune.json"))["rows"]
shapes = []
for r in rows:
    if r["shape"] not in shapes: shapes.append(r["shape"])
for sh in shapes:
    r = [x for x in rows if x["shape"] == sh and "tflops" in x and "fused" not in x["name"] and "dyn" not in x["name"]]
    auto = [x for x in r if x["config"] is None]
    r.sort(key=lambda x: -x["tflops"])
    print(sh, "AUTO", [(x["name"][7:], round(x["ms"] * 1e3, 1)) for x in auto], "best", [(x["name"][7:], x["split_k"], round(x["ms"] * 1e3, 1)) for x in r[:6]])
PY
