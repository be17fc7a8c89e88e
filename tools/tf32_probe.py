"""Bring-up probe for the tcgen05 3xTF32 path: structured inputs that localise layout / descriptor bugs
(K-major and MN-major operand feeds, TMA reduce epilogue, the hi = raw-operand split), one subprocess per
(A, B) layout pair so a trapped kernel cannot poison the rest.

    python tools/tf32_probe.py                      # driver: every layout pair, every config, prints a table
    python tools/tf32_probe.py child LAYOUT         # one layout pair (child), e.g. "LL" = A row-major, B row-major
Environment knobs of the library that this probe is meant to A/B: B200_TF32_MN_LBO / B200_TF32_MN_SBO (MN-major
descriptor strides), B200_TF32_NO_TMA_EPI, B200_TF32_FORCE_PACKED, B200_TF32_ROUND_HI.
"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CASES = ["ones", "rows", "cols", "kpat", "randint", "trunc", "uniform"]
SHAPES = [(128, 128, 32), (128, 64, 64), (256, 256, 96), (300, 260, 520), (1024, 1024, 1024), (2048, 4096, 512)]


def mk(t, layout):
    """Give a (rows x cols) tensor the requested storage order: L = row-major, F = column-major."""
    return t.contiguous() if layout == "L" else t.t().contiguous().t()


def child(layouts):
    import torch
    import openmp_blas_b200 as ob
    la, lb = layouts[0], layouts[1]
    ncfg = ob.num_configs("3xtf32", False)
    cfgs = [int(c) for c in os.environ.get("TF32_PROBE_CFGS", ",".join(map(str, range(ncfg)))).split(",")]
    rc = 0
    for cfg in cfgs:
        for (M, N, K) in SHAPES:
            g = torch.Generator(device="cuda").manual_seed(1)
            m = torch.arange(M, device="cuda", dtype=torch.float32)[:, None]
            n = torch.arange(N, device="cuda", dtype=torch.float32)[None, :]
            k_r = torch.arange(K, device="cuda", dtype=torch.float32)[None, :]
            k_c = torch.arange(K, device="cuda", dtype=torch.float32)[:, None]
            for case in CASES:
                if case == "ones":
                    a, b = torch.ones(M, K, device="cuda"), torch.ones(K, N, device="cuda")
                elif case == "rows":
                    a, b = (m % 7 + 1).expand(M, K), torch.ones(K, N, device="cuda")
                elif case == "cols":
                    a, b = torch.ones(M, K, device="cuda"), (n % 5 + 1).expand(K, N)
                elif case == "kpat":
                    a, b = (k_r % 3).expand(M, K), (k_c % 5).expand(K, N)
                elif case == "randint":
                    a = torch.randint(0, 100, (M, K), device="cuda", generator=g).float()
                    b = torch.randint(0, 100, (K, N), device="cuda", generator=g).float()
                elif case == "trunc":
                    # bits below TF32 precision: exact iff hi is what the tensor core really uses (truncation) and lo carries the rest
                    a = (1.0 + 2.0 ** -11 + 2.0 ** -12) * torch.ones(M, K, device="cuda") * (m % 2 * 2 - 1)
                    b = torch.ones(K, N, device="cuda") * (1.0 + 2.0 ** -12)
                    b = torch.where((n % 3 == 0).expand(K, N), torch.ones(K, N, device="cuda"), b)
                else:
                    a = torch.rand(M, K, device="cuda", generator=g) * 2 - 1
                    b = torch.rand(K, N, device="cuda", generator=g) * 2 - 1
                a, b = mk(a, la), mk(b, lb)
                c = torch.zeros(M, N, device="cuda")
                want = a.double() @ b.double()
                print(f"RUN  {case:8s} cfg={cfg} A{la}B{lb} {M}x{N}x{K}", flush=True)
                ob.mtm(c, a, b, None, variant="3xtf32", config=cfg)()
                torch.cuda.synchronize()
                ch = ob.last_choice()
                err = (c.double() - want).abs()
                scale = (a.double().abs() @ b.double().abs()).clamp_min(1e-30)
                rel = (err / scale).max().item()
                if case == "uniform":
                    bad = err > 4 * (K + 1) * 2.0 ** -24 * scale
                elif case == "trunc":
                    bad = err > 2.0 ** -21 * scale      # lo*lo (<= 2^-22 relative) is dropped by design; a wrong hi costs 2^-11
                else:
                    bad = err > 0
                nbad = int(bad.sum().item())
                msg = (f"{'OK  ' if nbad == 0 else 'BAD '} {case:8s} cfg={cfg} A{la}B{lb} modes=({ch['a_mode']},{ch['b_mode']}) {M}x{N}x{K}: "
                       f"max_abs={err.max().item():.3e} max_rel_to_|A||B|={rel:.3e} nbad={nbad}")
                if nbad:
                    rows = bad.any(1).nonzero().flatten()
                    cols = bad.any(0).nonzero().flatten()
                    msg += f" bad_rows[{len(rows)}]={rows[:8].tolist()}.. bad_cols[{len(cols)}]={cols[:8].tolist()}.."
                    i, j = bad.nonzero()[0].tolist()
                    msg += f" first=({i},{j}) got={c[i, j].item()} want={want[i, j].item()}"
                    rc = 1
                print(msg, flush=True)
    return rc


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        sys.exit(child(sys.argv[2]))
    fails = 0
    for layouts in os.environ.get("TF32_PROBE_LAYOUTS", "LF,LL,FL,FF").split(","):
        try:
            r = subprocess.run([sys.executable, __file__, "child", layouts], capture_output=True, text=True, timeout=600)
            lines = r.stdout.strip().splitlines()
            shown = [ln for ln in lines if not ln.startswith("RUN ")]
            bad = [ln for ln in shown if ln.startswith("BAD")]
            print(f"== A/B layouts {layouts}: {len(shown)} cases, {len(bad)} bad, rc={r.returncode}", flush=True)
            for ln in (bad[:30] if bad else shown[-3:]):
                print(ln, flush=True)
            if r.returncode != 0:
                last_run = [ln for ln in lines if ln.startswith("RUN ")][-1:] or ["(none)"]
                print(f"FAIL layouts={layouts} rc={r.returncode} last={last_run[0]} " + r.stderr.strip()[-600:], flush=True)
                fails += 1
        except subprocess.TimeoutExpired:
            print(f"HANG layouts={layouts}", flush=True)
            fails += 1
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
