"""Why a small-shape timing differs between runs: device time (library bench entry) and host wall time of the same loop,
fresh, after a heavy burst, and with fresh allocations each time."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import openmp_blas_b200 as ob  # noqa: E402


def one(n, tag, iters=200):
    a = torch.rand((n, n), device="cuda") * 2 - 1
    b = torch.rand((n, n), device="cuda") * 2 - 1
    c = torch.zeros((n, n), device="cuda")
    torch.cuda.synchronize()
    t = time.perf_counter()
    ms = ob.bench_device(c, a, b, variant="3xtf32", warmup=3, iters=iters)
    wall = (time.perf_counter() - t) / (iters + 3) * 1e6
    print(f"{tag} n={n}: device {ms * 1e3:.2f} us/call, host wall {wall:.2f} us/call, {ob.last_choice()['name']}", flush=True)


for rep in range(3):
    for n in (512, 1024):
        one(n, f"fresh#{rep}")
big = [torch.rand((8192, 8192), device="cuda") for _ in range(3)]
ob.bench_device(big[2], big[0], big[1], variant="3xtf32", warmup=1, iters=40)
for n in (512, 1024):
    one(n, "after 8192^3 x 40")
time.sleep(1.0)
for n in (512, 1024):
    one(n, "1 s later")
for n in (512, 1024):
    one(n, "iters=20", iters=20)
    one(n, "iters=2000", iters=2000)
