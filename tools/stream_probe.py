"""Does this box run kernels of two streams concurrently?  A long sleep kernel on stream A, a tiny op on
stream B enqueued AFTER it; B's event should complete while A is still running.  Prints one JSON line."""
import json
import os
import sys
import time

import torch

which = sys.argv[1] if len(sys.argv) > 1 else "legacy"
torch.zeros(1, device="cuda")
torch.cuda.synchronize()
a = torch.cuda.current_stream() if which == "legacy" else torch.cuda.Stream()
b = torch.cuda.Stream()
x = torch.zeros(1024, device="cuda")
ea, eb = torch.cuda.Event(), torch.cuda.Event()
torch.cuda.synchronize()
with torch.cuda.stream(a):
    torch.cuda._sleep(int(1.0e9))          # ~0.5 s
    ea.record()
with torch.cuda.stream(b):
    x.add_(1)
    eb.record()
t0 = time.time()
b_done_at = a_done_at = None
while time.time() - t0 < 5 and (b_done_at is None or a_done_at is None):
    if b_done_at is None and eb.query():
        b_done_at = time.time() - t0
    if a_done_at is None and ea.query():
        a_done_at = time.time() - t0
torch.cuda.synchronize()
print(json.dumps({"main": which, "b_done_s": b_done_at, "a_done_s": a_done_at,
                  "concurrent": bool(b_done_at is not None and a_done_at is not None and b_done_at < a_done_at - 0.05),
                  "env": {k: v for k, v in os.environ.items() if k.startswith(("CUDA", "NCCL", "TORCH_NCCL"))}}), flush=True)
