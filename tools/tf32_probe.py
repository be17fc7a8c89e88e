"""Bring-up probe for the tcgen05 3xTF32 path: structured inputs that localise layout / descriptor
bugs, each case in its own subprocess so a trapped kernel cannot poison the rest.

    python tools/tf32_probe.py            # driver: runs every case, prints a table
    python tools/tf32_probe.py CASE CFG M N K   # one case (child)
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CASES = ["ones", "rows", "cols", "kpat", "randint", "uniform"]
SHAPES = [(128, 128, 32), (128, 128, 64), (256, 256, 96), (300, 260, 520), (1024, 1024, 1024), (2048, 4096, 512)]


def child(case, cfg, M, N, K):
    import torch
    import openmp_blas_b200 as ob
    g = torch.Generator(device="cuda").manual_seed(1)
    m = torch.arange(M, device="cuda", dtype=torch.float32)[:, None]
    n = torch.arange(N, device="cuda", dtype=torch.float32)[None, :]
    k_r = torch.arange(K, device="cuda", dtype=torch.float32)[None, :]
    k_c = torch.arange(K, device="cuda", dtype=torch.float32)[:, None]
    if case == "ones":
        a, b = torch.ones(M, K, device="cuda"), torch.ones(K, N, device="cuda")
    elif case == "rows":
        a, b = (m % 7 + 1).expand(M, K).contiguous(), torch.ones(K, N, device="cuda")
    elif case == "cols":
        a, b = torch.ones(M, K, device="cuda"), (n % 5 + 1).expand(K, N).contiguous()
    elif case == "kpat":
        a, b = (k_r % 3).expand(M, K).contiguous(), (k_c % 5).expand(K, N).contiguous()
    elif case == "randint":
        a = torch.randint(0, 100, (M, K), device="cuda", generator=g).float()
        b = torch.randint(0, 100, (K, N), device="cuda", generator=g).float()
    else:
        a = torch.rand(M, K, device="cuda", generator=g) * 2 - 1
        b = torch.rand(K, N, device="cuda", generator=g) * 2 - 1
    c = torch.zeros(M, N, device="cuda")
    want = a.double() @ b.double()
    ob.mtm(c, a, b, None, variant="3xtf32", config=cfg)()
    torch.cuda.synchronize()
    err = (c.double() - want).abs()
    scale = (a.double().abs() @ b.double().abs()).clamp_min(1e-30)
    rel = (err / scale).max().item()
    bad = err > (1e-3 if case == "uniform" else 0)
    nbad = int(bad.sum().item())
    msg = f"{case:8s} cfg={cfg} {M}x{N}x{K}: max_abs={err.max().item():.3e} max_rel_to_|A||B|={rel:.3e} nbad={nbad}"
    if nbad and case != "uniform":
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        msg += f" bad_rows[{len(rows)}]={rows[:8].tolist()}.. bad_cols[{len(cols)}]={cols[:8].tolist()}.."
        i, j = bad.nonzero()[0].tolist()
        msg += f" first=({i},{j}) got={c[i, j].item()} want={want[i, j].item()}"
    print(msg, flush=True)
    return 0 if (nbad == 0 or case == "uniform") else 1


def main():
    if len(sys.argv) > 1:
        cases = CASES if sys.argv[1] == "all" else [sys.argv[1]]
        rc = 0
        for case in cases:
            rc |= child(case, int(sys.argv[2]), *map(int, sys.argv[3:6]))
        sys.exit(rc)
    fails = 0
    import os
    cfgs = [int(c) for c in os.environ.get("TF32_PROBE_CFGS", "1,0").split(",")]
    for cfg in cfgs:       # 1-CTA first, then the 2-CTA pair kernel
        for shape in SHAPES:  # one process per (cfg, shape): a trapped kernel only loses that group
            try:
                r = subprocess.run([sys.executable, __file__, "all", str(cfg), *map(str, shape)],
                                   capture_output=True, text=True, timeout=180)
                print(r.stdout.strip(), flush=True)
                if r.returncode != 0:
                    print(f"FAIL cfg={cfg} {shape} rc={r.returncode} " + r.stderr.strip()[-500:], flush=True)
                    fails += 1
            except subprocess.TimeoutExpired:
                print(f"HANG cfg={cfg} {shape}", flush=True)
                fails += 1
            if fails >= 4:
                print("too many failures, stopping")
                sys.exit(1)
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
