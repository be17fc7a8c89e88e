#!/bin/bash
# r03n: same-box A/B of the peer producer's per-k-block remote arrive (old behaviour = B200_TF32_PEER_ARRIVE=1), per pair config
mkdir -p gpurun_out
show='
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d["shape"], {k[-14:]: (v["kernel"][7:], v["us_best"], v["us_mean"], v["tflops_best"]) for k, v in d.items() if isinstance(v, dict)}, d.get("exact_vs_fp64_rows"), d.get("identical"))
'
for cfg in 9 0 4; do
timeout 300 python tools/ab_env.py --check --rounds 4 --config $cfg --shapes 8192,4096,16384,65536x1024x1024,8192x8192x1024 --env "" B200_TF32_PEER_ARRIVE=1 2>> gpurun_out/r03n_ab.err | tee -a gpurun_out/r03n_ab_peer_arrive.jsonl | python -c "$show"
done
timeout 300 python tools/ab_env.py --rounds 4 --config 9 --shapes 8192 --env B200_TF32_GROUP=8 B200_TF32_GROUP=16 B200_TF32_GROUP=32 2>> gpurun_out/r03n_ab.err | tee -a gpurun_out/r03n_ab_peer_arrive.jsonl | python -c "$show"
timeout 300 python tools/ab_env.py --rounds 4 --config 0 --shapes 8192 --env B200_TF32_GROUP=8 B200_TF32_GROUP=16 B200_TF32_GROUP=4 2>> gpurun_out/r03n_ab.err | tee -a gpurun_out/r03n_ab_peer_arrive.jsonl | python -c "$show"
tail -3 gpurun_out/r03n_ab.err
