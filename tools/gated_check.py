"""Single-GPU check of the gated product (b200_mtm_f32_gated_dev): B "arrives" panel by panel from a second
stream (delay, copy the panel into the slot, raise the sequence flag) while the product is already running.
The slot starts as NaN, so a tile that reads a panel before its flag poisons C.  Prints one JSON line.
The product leaves 8 SMs free here: the arrival producer is LOCAL in this test, and the GPU dispatches grids in
order — a persistent grid that cannot be placed completely (one SM held by the waiting warp) would sit in front
of the very kernels that deliver the panels.  Across GPUs the panels arrive by remote stores and nothing local
is needed for progress.

    python tools/gated_check.py [M N K [config]]
"""
import json
import os
import sys
from pathlib import Path

if os.environ.get("GATED_FORCE_CONNECTIONS"):
    os.environ["CUDA_DEVICE_MAX_CONNECTIONS"] = os.environ["GATED_FORCE_CONNECTIONS"]   # the arrival stream must not share a work queue with the product

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import openmp_blas_b200 as ob  # noqa: E402


def run_case(M, N, K, config, preset=False, delay_cycles=150000, reps=2):
    if os.environ.get("GATED_TIMING") and preset:
        reps = 6
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randint(0, 10, (M, K), device="cuda", generator=g).float()
    B = torch.randint(0, 10, (K, N), device="cuda", generator=g).float()
    C0 = torch.randint(0, 10, (M, N), device="cuda", generator=g).float()
    c = C0.clone()
    slot = torch.empty((K, N), device="cuda")
    flag = torch.zeros(16, device="cuda", dtype=torch.int32)
    side = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    n_panels = -(-N // ob.GATE_PANEL)
    seq = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # Warm every kernel the arrival stream and the product will launch: CUDA loads kernels lazily at first
    # launch, and a load can need an idle context — never let that happen while a kernel is spinning.
    with torch.cuda.stream(side):
        torch.cuda._sleep(10)
        slot[:, :min(N, 8)].copy_(B[:, :min(N, 8)])
        ob.flag_signal(flag.data_ptr() + 32, 1, stream=side.cuda_stream)
    slot.fill_(0.0)
    warm_c = C0.clone()
    ob.mtm(warm_c, A, B, None, variant="3xtf32", config=config)()
    torch.cuda.synchronize()
    for rep in range(reps):
        slot.fill_(float("nan"))
        side.wait_stream(main)
        first = seq + 1
        if preset:
            # no concurrency: B and the final sequence number are in place before the product is launched
            slot.copy_(B)
            seq += n_panels
            ob.flag_signal(flag.data_ptr(), seq, stream=main.cuda_stream)
        else:
            # the arrival work is enqueued FIRST (it does not depend on the product; the reverse order can
            # deadlock if both streams share a hardware queue), and starts with a delay
            with torch.cuda.stream(side):
                for j in range(n_panels):
                    torch.cuda._sleep(delay_cycles)
                    c0, c1 = j * ob.GATE_PANEL, min(N, (j + 1) * ob.GATE_PANEL)
                    slot[:, c0:c1].copy_(B[:, c0:c1])
                    seq += 1
                    ob.flag_signal(flag.data_ptr(), seq, stream=side.cuda_stream)
        e0.record()
        ob.mtm_gated(c, A, slot, flag.data_ptr(), first, config=config,
                     reserve_sms=0 if (os.environ.get("GATED_TIMING") and preset) else 8)()   # waits in-kernel for the panels
        e1.record()
        main.wait_stream(side)
    torch.cuda.synchronize()
    want = C0.double() + reps * (A.double() @ B.double())
    bad = c.double() != want
    # same problem through the ordinary entry for the time of an ungated call
    c2 = C0.clone()
    fn = ob.mtm(c2, A, B, None, variant="3xtf32", config=config)
    fn()
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(5):
        fn()
    f1.record()
    torch.cuda.synchronize()
    return {"shape": [M, N, K], "config": config, "preset": preset, "panels": n_panels, "exact": not bool(bad.any()),
            "n_bad": int(bad.sum().item()), "nan": int(torch.isnan(c).sum().item()), "kernel": ob.last_choice()["name"],
            "ms_gated_last": e0.elapsed_time(e1), "ms_ungated": f0.elapsed_time(f1) / 5}


def main():
    if os.environ.get("GATED_MAIN_STREAM") == "new":
        with torch.cuda.stream(torch.cuda.Stream()):
            return _main()
    return _main()


def _main():
    if os.environ.get("GATED_TIMING"):
        # all panels pre-arrived: isolates the cost of the gating mechanics (side-stream split chain, polling,
        # full carve-out) from any waiting for the sender
        cases = [(8192, 8192, 8192, 0, True), (8192, 8192, 8192, 0, True), (8192, 8192, 8192, 2, True),
                 (1024, 1344, 2048, None, False), (2048, 4096, 4096, 0, False), (2048, 2048, 1000, 2, False)]
    elif os.environ.get("GATED_ONLY_PRESET"):
        cases = [(1024, 1344, 2048, None, True)]
    elif len(sys.argv) >= 4:
        cases = [(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else None, False)]
    else:
        cases = [(1024, 1344, 2048, None, True), (1024, 1344, 2048, None, False), (4096, 2176, 1024, 1, False),
                 (300, 520, 96, None, False), (2048, 4096, 4096, 0, False), (2048, 2048, 1000, 2, False),
                 (8192, 8192, 8192, None, False)]
    out = []
    for cs in cases:
        try:
            out.append(run_case(*cs))
        except Exception as e:      # a trapped kernel poisons the context: report and stop
            out.append({"shape": list(cs[:3]), "config": cs[3], "preset": cs[4], "exact": False, "error": str(e)[:300]})
            print("# case failed", json.dumps(out[-1]), file=sys.stderr, flush=True)
            break
        print("# case", json.dumps(out[-1]), file=sys.stderr, flush=True)
    print(json.dumps({"ok": all(o["exact"] for o in out) and len(out) == len(cases), "cases": out}), flush=True)
    return 0 if all(o["exact"] for o in out) and len(out) == len(cases) else 1


if __name__ == "__main__":
    sys.exit(main())
