#!/bin/bash
# PCIe ceiling of the box vs what the host-pointer entry achieves
timeout 300 python - <<'PY'
import torch, time
n = 256 * 1024 * 1024          # 1 GiB of fp32
h = torch.empty(n, dtype=torch.float32, pin_memory=True).uniform_(-1, 1)
h2 = torch.empty(n // 4, dtype=torch.float32, pin_memory=True)
d = torch.empty(n, device="cuda"); d2 = torch.empty(n // 4, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(f, reps=5):
    f(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps
t = run(lambda: d.copy_(h, non_blocking=True)); print(f"H2D 1 GiB one copy: {4*n/t/1e9:.1f} GB/s")
t = run(lambda: h2.copy_(d2, non_blocking=True)); print(f"D2H 256 MiB one copy: {n/t/1e9:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
t = run(both); print(f"H2D 1 GiB || D2H 256 MiB: {t*1e3:.2f} ms -> H2D {4*n/t/1e9:.1f} GB/s")
def chunks():
    for i in range(32):
        d[i * (n // 32):(i + 1) * (n // 32)].copy_(h[i * (n // 32):(i + 1) * (n // 32)], non_blocking=True)
t = run(chunks); print(f"H2D 1 GiB in 32 copies of 32 MiB: {4*n/t/1e9:.1f} GB/s")
# 768 MiB up while 256 MiB down: the byte counts of one 8192^3 call
hu = h[: 3 * n // 4]; du = d[: 3 * n // 4]
def call_like():
    with torch.cuda.stream(s1): du.copy_(hu, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
t = run(call_like); print(f"768 MiB up || 256 MiB down: {t*1e3:.2f} ms (the copy floor of the 8192^3 host call)")
PY
