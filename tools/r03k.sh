#!/bin/bash
# r03k: FFMA-TMA stage release by named-barrier arrive (only the refilling warp waits) — parity + timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "simt or every_tile_config or randomized or edge or full_size" > gpurun_out/r03k_pytest_simt.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r03k_pytest_simt.log
python - <<'PY'
import sys, json
sys.path.insert(0, ".")
import torch
import openmp_blas_b200 as ob
n_classic = 5
for (m, n, k) in ((8192, 8192, 8192), (4096, 4096, 4096), (2048, 2048, 2048), (1024, 1024, 1024), (65536, 1024, 1024)):
    a = torch.rand((m, k), device="cuda") * 2 - 1; b = torch.rand((k, n), device="cuda") * 2 - 1; c = torch.zeros((m, n), device="cuda")
    row = {}
    for cfg in (None, n_classic + 0, n_classic + 1, n_classic + 3):
        ms = min(ob.bench_device(c, a, b, variant="simt", config=cfg, warmup=2, iters=5 if m * n * k > 1e11 else 50) for _ in range(3))
        row[ob.last_choice()["name"] + ("(auto)" if cfg is None else "")] = round(m * n * (2.0 * k - 1) / ms / 1e9, 2)
    print((m, n, k), json.dumps(row), flush=True)
PY
