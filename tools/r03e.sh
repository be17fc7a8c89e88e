#!/bin/bash
# r03e: double tiles in AUTO + K-split support — full tests; K dependence; raster group sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r03e_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r03e_pytest_gpu.log
for k in 256 512 1024 2048 4096; do
timeout 300 python tools/ab_compare.py 3xtf32 0 9 --n 8192 --k $k --rounds 3 --iters 20 | tee -a gpurun_out/r03e_ab_double_tile_k.jsonl
done
timeout 300 python tools/ab_compare.py 3xtf32 0 9 --m 65536 --n 1024 --k 1024 --rounds 3 --iters 20 | tee -a gpurun_out/r03e_ab_double_tile_k.jsonl
timeout 300 python tools/ab_compare.py 3xtf32 0 9 --m 1024 --n 8192 --k 8192 --rounds 3 --iters 20 | tee -a gpurun_out/r03e_ab_double_tile_k.jsonl
for n in 2560 3072 3584 5120 6144; do
timeout 300 python tools/ab_compare.py 3xtf32 0 9 --n $n --rounds 3 --iters 20 | tee -a gpurun_out/r03e_ab_double_tile_k.jsonl
done
for g in 8 12 16 24 32; do echo "group $g"; B200_TF32_GROUP=$g timeout 300 python tools/ab_compare.py 3xtf32 9 --n 8192 --rounds 3 --iters 8 | tee -a gpurun_out/r03e_ab_double_tile_group.jsonl; done
for g in 8 16 32; do echo "group $g 16384"; B200_TF32_GROUP=$g timeout 300 python tools/ab_compare.py 3xtf32 9 --n 16384 --rounds 2 --iters 3 | tee -a gpurun_out/r03e_ab_double_tile_group.jsonl; done
