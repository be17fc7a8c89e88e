#!/bin/bash
# r03m: pair configs without the per-k-block remote arrive of the peer producer — probe, parity, timing of every pair config
mkdir -p gpurun_out
timeout 400 python tools/tf32_probe.py > gpurun_out/r03m_tf32_probe.log 2>&1; echo "probe exit $?"; grep -c "^BAD" gpurun_out/r03m_tf32_probe.log; grep "A/B layouts" gpurun_out/r03m_tf32_probe.log
timeout 600 python -m pytest tests -m gpu -q -x -k "every_tile_config or edge_shapes or tolerance or split or stream or dynamic" > gpurun_out/r03m_pytest_cfg.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r03m_pytest_cfg.log
timeout 300 python tools/ab_compare.py 3xtf32 0 9 4 1 5 --n 4096 --rounds 3 --iters 20 | tee -a gpurun_out/r03m_ab_uniform_issue.jsonl
timeout 300 python tools/ab_compare.py 3xtf32 0 9 2 10 --n 8192 --rounds 3 --iters 8 | tee -a gpurun_out/r03m_ab_uniform_issue.jsonl
timeout 300 python tools/ab_compare.py 3xtf32 0 9 4 1 --n 2048 --rounds 3 --iters 50 | tee -a gpurun_out/r03m_ab_uniform_issue.jsonl
timeout 300 python tools/ab_compare.py 3xtf32 0 4 1 5 --n 1536 --rounds 3 --iters 50 | tee -a gpurun_out/r03m_ab_uniform_issue.jsonl
timeout 300 python tools/ab_compare.py 3xtf32 4 1 5 --n 1024 --rounds 3 --iters 100 | tee -a gpurun_out/r03m_ab_uniform_issue.jsonl
