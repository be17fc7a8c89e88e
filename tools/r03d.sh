#!/bin/bash
# r03d: double tiles (256 x 512 per pair, config 9) — probe, per-config parity tests, timing against config 0.
mkdir -p gpurun_out
timeout 400 python tools/tf32_probe.py > gpurun_out/r03d_tf32_probe.log 2>&1; echo "probe exit $?"; grep -c "^BAD" gpurun_out/r03d_tf32_probe.log; grep "A/B layouts" gpurun_out/r03d_tf32_probe.log; grep "^BAD" gpurun_out/r03d_tf32_probe.log | head -5
timeout 600 python -m pytest tests -m gpu -q -x -k "every_tile_config or edge_shapes or tolerance" > gpurun_out/r03d_pytest_cfg.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r03d_pytest_cfg.log
for n in 8192 16384; do
timeout 300 python tools/ab_compare.py 3xtf32 0 9 --n $n --rounds 3 --iters 8 | tee -a gpurun_out/r03d_ab_double_tile.jsonl
done
timeout 300 python tools/ab_compare.py 3xtf32 0 9 --n 4096 --rounds 3 --iters 30 | tee -a gpurun_out/r03d_ab_double_tile.jsonl
timeout 300 python tools/ab_compare.py 3xtf32 0 9 --n 8192 --k 2048 --rounds 3 --iters 20 | tee -a gpurun_out/r03d_ab_double_tile.jsonl
for g in 4 16; do echo "group $g"; B200_TF32_GROUP=$g timeout 300 python tools/ab_compare.py 3xtf32 0 9 --n 8192 --rounds 2 --iters 8 | tee -a gpurun_out/r03d_ab_double_tile.jsonl; done
