#!/usr/bin/env python
"""bench.py — mtm (C += A*B) throughput on B200, beside the reference's OpenMP mtm on the host cores.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU implementation

One JSON line on stdout (rank 0).  A "step" is one call of the hot path, C += A*B, on the
workload named in `config.workload`:

  N = 1 : fp32 8192 x 8192 x 8192, all operands last_order (row-major) — the square-sweep point
          (BASELINE.json configs[1]) the north-star target is quoted on.
  N > 1 : the same problem per GPU, row-block sharded (BASELINE.json configs[4] layout): rank r owns
          8192 rows of A and C, B (8192 x 8192) lives on rank 0 and is broadcast over NCCL/NVLink
          INSIDE every timed step, chunked along K and overlapped with the K-chunk products
          (mtm accumulates, so C += A[:, chunk] * B[chunk, :] needs no reduction).  scaling = weak.

`value`   : whole-job TFLOP/s with operands resident in HBM, flops = M*N*(2K-1) as in src/mtm.cpp:203,
            CUDA events on the launching stream, max over ranks.
`e2e`     : same metric through the public host-buffer API (pinned host arrays -> b200_mtm_f32 ->
            host), H2D/D2H inside the timed region.
`roofline`: the dominant kernel against its bounding pipe.
`cpu_baseline`: the reference's amt::mtm (oracle/_ref, unmodified headers) on the host cores,
            bounded sample, rank 0 at N=1 only.
Inputs are larger than L2 (3 x 256 MiB vs 126 MB), so no explicit L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "mtm TFLOP/s (fp32/fp64) and % of peak at 1/2/4/8 B200 vs OpenMP host cores"
SIZE = 8192


def flops(M, N, K):
    return float(M) * float(N) * (2.0 * float(K) - 1.0)     # src/mtm.cpp:203


def kernel_source_hash(kernel_name: str = "tf32x3") -> str:
    """sha256 over the CODE of the sources a kernel family is compiled from (// comments and blank lines stripped, so
    that editing a comment does not invalidate a capture): the stamp that ties a committed ncu capture to the kernel
    it was taken on."""
    import hashlib
    import re
    main = ("mtm_tf32.cu" if kernel_name.startswith("tf32") else "mtm_ffma_tma.cu" if kernel_name.startswith("ffma") else
            "mtm_dmma_tma.cu" if kernel_name.startswith("dmma_tma") else "mtm_simt.cuh")
    h = hashlib.sha256()
    for f in (main, "sm100_ptx.cuh", "mtm_common.cuh"):
        text = (ROOT / "openmp-blas_b200" / "csrc" / f).read_text()
        for line in text.splitlines():
            line = re.sub(r"\s*//.*$", "", line).rstrip()      # (no string literal in these files contains //)
            if line:
                h.update(line.encode() + b"\n")
    return h.hexdigest()[:16]


def ncu_traffic(kernel_name: str):
    """(DRAM bytes read + written per launch of the dominant kernel, stale?) from the committed `ncu --set full`
    capture (profiles/ncu_traffic.json; same shape, same kernel).  The capture carries the source hash it was
    taken on: a kernel edited since then is reported as stale instead of silently keeping the old number."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    try:
        rec = json.loads(p.read_text()).get(kernel_name, {})
        val = rec.get("dram_bytes_per_launch")
        if val is None:
            return None, None
        return val, rec.get("source_hash") != kernel_source_hash(kernel_name)
    except Exception:
        return None, None


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text()), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, "fallback"


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [l for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l for _, l in self.lines]
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, budget_s: float = 20.0):
    """Time the reference's OpenMP amt::mtm (oracle/_ref; oracle port if _ref is absent) on a bounded
    sample of the workload: fp32 row-major, N = K = 8192, an M-slab sized to ~budget_s of CPU work."""
    import oracle
    try:
        lib = oracle.Reference()
        kind, desc = "reference", f"oracle/_ref ({lib.isa}: unmodified include/mtm.hpp, g++ -O3 -ffast-math -fopenmp)"
        threads = lib.threads()
        run = lambda c, a, b: lib.bench_ns(1, c, a, b) * 1e-9
        mb = lib.block_sizes(np.float32, True)[3]
    except (FileNotFoundError, OSError):
        lib = oracle.Oracle()
        kind, desc = "port", "oracle/oracle_mtm.c (plain-C restatement, OpenMP)"
        threads = lib.threads()
        mb = lib.default_blocks(np.float32, True)[3]

        def run(c, a, b):
            t = time.perf_counter()
            lib.mtm(c, a, b)
            return time.perf_counter() - t
    rng = np.random.default_rng(0xB200)
    N = K = SIZE
    b = rng.uniform(-1, 1, (K, N)).astype(np.float32)
    # calibrate on one MB-block per thread, then size the slab for the time budget
    m_unit = mb * threads                       # every thread gets one M-block per (j,k) step (mtm.hpp:182)
    m_cal = min(SIZE, m_unit)
    a = rng.uniform(-1, 1, (m_cal, K)).astype(np.float32)
    c = np.zeros((m_cal, N), np.float32)
    run(c, a, b)                                # first call sizes the static pack buffers (mtm.hpp:147-151)
    t_cal = run(c, a, b)
    total_calls = max(1, steps + warmup)
    per_call = max(0.5, budget_s / total_calls)
    mult = max(1, int(per_call / max(t_cal, 1e-6)))
    M = int(min(SIZE, m_cal * mult))
    if M != m_cal:
        a = rng.uniform(-1, 1, (M, K)).astype(np.float32)
        c = np.zeros((M, N), np.float32)
    for _ in range(warmup):
        run(c, a, b)
    times = [run(c, a, b) for _ in range(steps)]
    mean = float(np.mean(times))
    tf = flops(M, N, K) / mean / 1e12
    sample = (f"M-slab of the 8192^3 problem: M={M}, N=K=8192 fp32 row-major, {steps} calls after {warmup} warm-up, "
              f"{desc}, {threads} OpenMP threads")
    return {"value": tf, "unit": "TFLOP/s", "cores": threads, "kind": kind, "sample": sample,
            "ms_per_call": mean * 1e3, "M": M}


def cpu_extras(budget_s: float = 12.0):
    """More CPU lines beside the GPU figures: config 1 as the reference runs it, an fp64
    sample of config 3, and numpy's bundled OpenBLAS (the only host BLAS that installs offline here;
    MKL / BLIS / Eigen, which the reference also compares against, are not installed)."""
    import oracle
    out = {}
    try:
        lib = oracle.Reference()
    except (FileNotFoundError, OSError):
        return {"unavailable": "oracle/_ref not built"}
    threads = lib.threads()
    rng = np.random.default_rng(1)

    def ref_gflops(dtype, M, N, K, iters):
        a = rng.uniform(-1, 1, (M, K)).astype(dtype)
        b = rng.uniform(-1, 1, (K, N)).astype(dtype)
        c = np.zeros((M, N), dtype)
        t_warm, n_warm = 0.0, 0                                   # warm-up: the first calls size the static pack
        while n_warm < 3 or (t_warm < 2.0 and n_warm < 40):       # buffers (mtm.hpp:147-151), spin up the OpenMP
            t_warm += lib.bench_ns(1, c, a, b) * 1e-9             # team and wake the cores (measured: the first
            n_warm += 1                                           # ~1 s of calls runs 10x slower than steady state)
        ns = lib.bench_ns(iters, c, a, b)                         # amt::benchmark<iters>
        return flops(M, N, K) / ns, (a, b)

    g, _ = ref_gflops(np.float32, 1024, 1024, 1024, 4)
    out["config1_f32_1024^3_LLL_reference"] = {"gflops": round(g, 1), "protocol": "amt::benchmark<4> after warm-up to steady state", "threads": threads}
    mb64 = lib.block_sizes(np.float64, True)[3]
    m64 = int(min(SIZE, mb64 * threads))
    g, _ = ref_gflops(np.float64, m64, SIZE, SIZE, 2)
    out["config3_f64_sample_reference"] = {"gflops": round(g, 1), "sample": f"M={m64}, N=K=8192 fp64 row-major", "threads": threads}
    mb32 = lib.block_sizes(np.float32, True)[3]
    m32 = int(min(SIZE, mb32 * threads))
    a = rng.uniform(-1, 1, (m32, SIZE)).astype(np.float32)
    b = rng.uniform(-1, 1, (SIZE, SIZE)).astype(np.float32)
    a @ b
    t = time.perf_counter()
    for _ in range(2):
        a @ b
    dt = (time.perf_counter() - t) / 2
    out["openblas_f32_sample_numpy"] = {"gflops": round(flops(m32, SIZE, SIZE) / dt / 1e9, 1),
                                        "sample": f"numpy {np.__version__} matmul (bundled OpenBLAS), M={m32}, N=K=8192 fp32"}
    out["not_installed"] = "MKL, BLIS, Eigen, standalone OpenBLAS (the reference's other comparators) cannot be installed offline"
    return out


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1 into every rank; the reference reads that variable when its library is
    loaded (include/thread_utils.hpp:13-17).  The CPU arm has to run on every core this process may use, so set
    it BEFORE the oracle library (and libgomp) is loaded."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    return n


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    use_all_host_cores()
    r = cpu_reference_run(args.steps, args.warmup, budget_s=60.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_call"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic uniform(-1,1), seed 0xB200",
        "config": {"workload": f"mtm fp32 {SIZE * max(1, args.gpus)}x8192x8192 last_order (row-major), C += A*B (the GPU arm's "
                               f"problem at N={max(1, args.gpus)}: 8192 rows per GPU); CPU sample: {r['sample']}"},
        "cpu_baseline": {"value": r["value"], "unit": "TFLOP/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def bind_to_gpu_numa_node(torch, local_rank):
    """N > 1: pin this rank's host threads to the CPUs next to its GPU (NVML's ideal affinity), so the pinned
    staging buffers of the e2e leg are allocated on the GPU's own NUMA node instead of crossing the socket
    link with 8 ranks at once.  Not used at N = 1, where the CPU baseline needs every host core."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            pr = torch.cuda.get_device_properties(local_rank)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------
def time_device(ob, torch, c, a, b, variant, config, warmup, iters):
    """Mean ms/call with CUDA events on the launching (torch current) stream."""
    fn = ob.mtm(c, a, b, None, variant=variant, config=config)
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def dev_uniform(torch, shape, dtype, layout, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    if layout == "L":
        return torch.rand(shape, device="cuda", dtype=dtype, generator=g) * 2 - 1
    return (torch.rand((shape[1], shape[0]), device="cuda", dtype=dtype, generator=g) * 2 - 1).t()


def extras_single_gpu(ob, torch, info, peaks, quick):
    """Secondary measurements reported beside the headline (BASELINE.json configs 2-4)."""
    out = {}
    p32, p64 = info["peak_fp32_tflops"], info["peak_fp64_tflops"]
    tf32_peak = peaks["bf16_tflops"] / 2.0

    def point(dtype, M, N, K, lay, variant, iters=5, warm=3):
        a = dev_uniform(torch, (M, K), dtype, lay[1], 1)
        b = dev_uniform(torch, (K, N), dtype, lay[2], 2)
        c = torch.zeros((M, N), device="cuda", dtype=dtype) if lay[0] == "L" else torch.zeros((N, M), device="cuda", dtype=dtype).t()
        # timed by the library's own benchmark entry (b200_mtm_bench_*_dev: back-to-back calls issued from C, CUDA events
        # on the launching stream — the counterpart of amt::benchmark, benchmark.hpp:34-52): a Python loop adds 10-20 us
        # of host time per call, more than a small problem takes on the device
        fl = flops(M, N, K)
        iters = max(iters, 200 if fl < 1e10 else (50 if fl < 5e11 else iters))
        ms = ob.bench_device(c, a, b, variant=variant, config=None, warmup=warm, iters=iters)
        name = ob.last_choice()["name"]
        del a, b, c
        return ms, fl / ms / 1e9, name

    sweep = []
    sizes = [512, 1024, 2048, 4096, 8192] + ([] if quick else [16384])
    for n in sizes:
        row = {"n": n}
        for v in ("simt", "3xtf32"):
            if ob.num_configs(v, False) == 0:
                continue
            ms, tf, name = point(torch.float32, n, n, n, "LLL", v, iters=3 if n >= 16384 else 10)
            row[v] = {"tflops": round(tf, 2), "ms": round(ms, 4), "kernel": name,
                      "frac_fp32_simt_peak": round(tf / p32, 4)}
            if v == "3xtf32":
                row[v]["frac_tf32_peak_div3"] = round(tf / (tf32_peak / 3.0), 4)
        sweep.append(row)
    out["config2_fp32_square_sweep_LLL"] = sweep
    f64 = {}
    for v in ("dfma", "dmma"):
        if ob.num_configs(v, True) == 0:
            continue
        ms, tf, name = point(torch.float64, 8192, 8192, 8192, "LLL", v, iters=5)
        f64[v] = {"tflops": round(tf, 2), "ms": round(ms, 3), "kernel": name, "frac_fp64_peak": round(tf / p64, 4)}
    out["config3_fp64_8192"] = f64
    c4 = {}
    for lay, shape in (("LLL", (65536, 1024, 1024)), ("FLF", (65536, 1024, 1024)), ("FLF", (8192, 8192, 8192)),
                       ("FFF", (8192, 8192, 8192))):
        for v in ("simt", "3xtf32"):
            if ob.num_configs(v, False) == 0:
                continue
            ms, tf, name = point(torch.float32, *shape, lay, v, iters=5)
            c4[f"{lay}_{shape[0]}x{shape[1]}x{shape[2]}_{v}"] = {"tflops": round(tf, 2), "ms": round(ms, 3), "kernel": name}
    out["config4_fp32_rect_and_transposed"] = c4
    # operands that are NOT TMA-legal — odd extents / leading dimensions (packed 3xTF32 feed, scalar
    # loaders of the CUDA-core kernels) and strided sub-views (every second column of a wider matrix)
    try:
        odd = {}
        M, N, K = 4097, 4099, 4101
        for v in ("simt", "3xtf32"):
            if ob.num_configs(v, False) == 0:
                continue
            ms, tf, name = point(torch.float32, M, N, K, "LLL", v, iters=5)
            ch = ob.last_choice()
            odd[f"odd_{M}x{N}x{K}_LLL_{v}"] = {"tflops": round(tf, 2), "ms": round(ms, 3), "kernel": name,
                                                 "operand_modes": [ch["a_mode"], ch["b_mode"]]}
        n = 4096
        wide_a = dev_uniform(torch, (n, 2 * n), torch.float32, "L", 21)
        wide_b = dev_uniform(torch, (n, 2 * n), torch.float32, "L", 22)
        for v in ("simt", "3xtf32"):
            if ob.num_configs(v, False) == 0:
                continue
            c = torch.zeros((n, n), device="cuda", dtype=torch.float32)
            ms = time_device(ob, torch, c, wide_a[:, ::2], wide_b[:, ::2], v, None, 3, 5)
            ch = ob.last_choice()
            odd[f"strided_views_{n}^3_{v}"] = {"tflops": round(flops(n, n, n) / ms / 1e9, 2), "ms": round(ms, 3), "kernel": ch["name"],
                                                "operand_modes": [ch["a_mode"], ch["b_mode"]]}
            del c
        del wide_a, wide_b
        out["f2_unaligned_and_strided_operands"] = odd
    except Exception as e:
        out["f2_unaligned_and_strided_operands"] = {"error": str(e)[:200]}
    # The reference's own harness inputs (src/mtm.cpp:204-206: all-ones A and B, zero C) for a like-for-like
    # line; the headline uses uniform(-1,1) so that the numbers carry no data-dependent power artefact.
    try:
        ones = {}
        n = 8192
        a1 = torch.ones((n, n), device="cuda", dtype=torch.float32)
        b1 = torch.ones((n, n), device="cuda", dtype=torch.float32)
        for v in ("simt", "3xtf32"):
            if ob.num_configs(v, False) == 0:
                continue
            c1 = torch.zeros((n, n), device="cuda", dtype=torch.float32)
            ms = time_device(ob, torch, c1, a1, b1, v, None, 3, 5)
            # 8 calls of C += 1*1 summed over K = 8192: every entry is exactly 8 * 8192
            ones[v] = {"tflops": round(flops(n, n, n) / ms / 1e9, 2), "ms": round(ms, 3), "kernel": ob.last_choice()["name"],
                       "exact_after_8_calls": bool((c1 == 8.0 * n).all().item())}
            del c1
        del a1, b1
        out["config2_all_ones_8192_like_src_mtm_cpp"] = ones
    except Exception as e:      # a secondary line must never cost the headline
        out["config2_all_ones_8192_like_src_mtm_cpp"] = {"error": str(e)[:200]}
    # matrix-times-vector: HBM-bound, bytes = rows * cols * sizeof(T)
    mtv = {}
    for dtype, n in ((torch.float32, 32768), (torch.float64, 16384)):
        for order in ("F", "L"):
            a = dev_uniform(torch, (n, n), dtype, order, 5)
            v = dev_uniform(torch, (1, n), dtype, "L", 6).reshape(n)
            for is_vtm in (False, True):
                c = torch.zeros(n, device="cuda", dtype=dtype)
                ms = ob.bench_mtv_device(c, a, v, is_vtm=is_vtm, warmup=2, iters=5)
                gbs = n * n * a.element_size() / ms / 1e6
                mtv[f"{'vtm' if is_vtm else 'mtv'}_{'f32' if dtype == torch.float32 else 'f64'}_{n}_{order}"] = {
                    "GB/s": round(gbs, 1), "ms": round(ms, 4), "kernel": ob.last_choice()["name"],
                    "frac_hbm_measured": round(gbs / peaks["hbm_gbs"], 4)}
            del a, v
    out["mtv_vtm_hbm_bound"] = mtv
    # transpose: HBM-bound copy, bytes = 2 * rows * cols * sizeof(T)
    tr = {}
    for dtype, n in ((torch.float32, 16384), (torch.float64, 16384)):
        a = dev_uniform(torch, (n, n), dtype, "L", 7)
        for name, av, c_first in (("LtoL", a, False), ("LtoF_copy", a, True), ("FtoF", a.t(), True)):
            c = torch.zeros((n, n), device="cuda", dtype=dtype)
            cv = c.t() if c_first else c
            ms = ob.bench_transpose_device(cv, av, warmup=2, iters=5)
            gbs = 2.0 * n * n * a.element_size() / ms / 1e6
            tr[f"transpose_{'f32' if dtype == torch.float32 else 'f64'}_{n}_{name}"] = {
                "GB/s": round(gbs, 1), "ms": round(ms, 4), "frac_hbm_measured": round(gbs / peaks["hbm_gbs"], 4)}
            del c
        del a
    out["transpose_hbm_bound"] = tr
    out["peaks"] = {"fp32_simt_tflops": round(p32, 2), "fp64_tflops": round(p64, 2),
                    "tf32_dense_tflops_from_measured_bf16_div2": round(tf32_peak, 1),
                    "note": "fp32/fp64 peaks = SMs * {128,64} lanes * 2 * max SM clock (cudaDevAttrClockRate)"}
    return out


def max_over_ranks(torch, dist, world, x: float) -> float:
    if world == 1:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed_steps(torch, dist, world, fn, warmup, steps):
    """ms per step: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return max_over_ranks(torch, dist, world, e0.elapsed_time(e1) / steps)


def config5_record(ob, torch, dist, rank, world, headline, args):
    """BASELINE.json configs[4]: fp32 S^3 (S = 32768) row-block sharded over the N GPUs of this run, B replicated
    from rank 0 INSIDE the timed region.  Strong scaling: the total problem is fixed, rank r owns S/N rows of A
    and C.  Reports the throughput with the exchange, compute-only (B already resident), the single-GPU figure
    of the same problem measured on rank 0 of the same box, and an exactness flag from one integer-data step
    checked on every rank against fp64 on sampled rows (MIN over ranks)."""
    from openmp_blas_b200.sharded import RowBlockMtm
    S = int(args.config5_size)
    drv = RowBlockMtm(S, S, S, torch.float32, variant=headline, bcast=args.bcast, push_ctas=args.push_ctas)
    r0, r1 = drv.my_rows
    rows = r1 - r0
    root = rank == 0
    # -- exactness: integers in [0, 9], K * 81 < 2^24: every summation order is exact in fp32 -----------------
    gb = torch.Generator(device="cuda").manual_seed(0xC5)           # the same B on every rank (each checks against its own copy)
    ga = torch.Generator(device="cuda").manual_seed(0xA5 + rank)
    b = torch.randint(0, 10, (S, S), device="cuda", generator=gb, dtype=torch.float32)
    a = torch.randint(0, 10, (rows, S), device="cuda", generator=ga, dtype=torch.float32)
    c = torch.zeros((rows, S), device="cuda", dtype=torch.float32)
    drv.step(c, a, b if root else None)
    torch.cuda.synchronize()
    ok = True
    if rows > 0:
        sample = torch.unique(torch.linspace(0, rows - 1, min(rows, 48), device="cuda").long())
        a_s = a[sample].double()
        for j0 in range(0, S, 4096):
            want = a_s @ b[:, j0:j0 + 4096].double()
            ok = ok and bool(torch.equal(c[sample, j0:j0 + 4096].double(), want))
            del want
    if world > 1:
        t = torch.tensor([int(ok)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
    # -- timing on uniform(-1, 1) data --------------------------------------------------------------------------
    a.uniform_(-1, 1)
    c.zero_()
    if root:
        b.uniform_(-1, 1)
    else:
        del b
        b = None
    cal = drv.calibrate(a, b, steps=1) if world > 1 else None
    steps = max(2, min(args.steps, 3))
    fl = flops(S, S, S)
    ms_bcast = timed_steps(torch, dist, world, lambda: drv.step(c, a, b), 1, steps)
    b_local = b if root else drv.b_buf
    local = ob.mtm(c, a, b_local, None, variant=drv.variant, config=drv.config)
    ms_comp = timed_steps(torch, dist, world, local, 1, steps) if rows > 0 or world > 1 else ms_bcast
    rec = {"workload": f"mtm fp32 {S}x{S}x{S} last_order, row-block sharded over {world} GPU(s), B replicated from rank 0 every step",
           "scaling": "strong", "rows_per_gpu": rows if world == 1 else -(-S // world), "exact": ok,
           "exact_check": "one step on integer data in [0,9]; 48 sampled rows per rank x all columns vs fp64, MIN over ranks",
           "tflops_with_broadcast": round(fl / ms_bcast / 1e9, 2), "ms_with_broadcast": round(ms_bcast, 3),
           "tflops_compute_only": round(fl / ms_comp / 1e9, 2), "ms_compute_only": round(ms_comp, 3),
           "k_chunks": drv.chunks, "kernel": ob.last_choice()["name"], "steps": steps,
           "b_replication": ("none (single GPU)" if world == 1 else ("own NVLink multicast push" if drv.use_nvlink else "NCCL broadcast")),
           "calibration_ms_per_step": cal}
    # -- the same problem on ONE GPU of this box (rank 0), for the strong-scaling efficiency -------------------
    if world == 1:
        n1 = fl / ms_bcast / 1e9
    else:
        n1 = 0.0
        if root:
            del c, a
            a1 = torch.empty((S, S), device="cuda", dtype=torch.float32).uniform_(-1, 1)
            c1 = torch.zeros((S, S), device="cuda", dtype=torch.float32)
            n1 = fl / time_device(ob, torch, c1, a1, b, headline, None, 1, 2) / 1e9
            del a1, c1
        n1 = max_over_ranks(torch, dist, world, n1)
    rec["n1_tflops_same_box"] = round(n1, 2)
    rec["efficiency_vs_n1"] = round(rec["tflops_with_broadcast"] / (world * n1), 4) if n1 > 0 else None
    rec["efficiency_compute_only_vs_n1"] = round(rec["tflops_compute_only"] / (world * n1), 4) if n1 > 0 else None
    return rec


def summa_record(ob, torch, dist, rank, world, headline, size=8192):
    """The optional 2-D split (SummaMtm, Pr x Pc grid chosen by choose_grid) on the N GPUs of this run: one
    accumulate step on integer data checked EXACTLY on every rank's C block against fp64 (MIN over ranks), then
    timed on uniform data with the panel broadcasts inside the timed region."""
    from openmp_blas_b200.sharded import SummaMtm
    M, N, K = size, size + 128, size - 96                 # ragged on purpose: blocks and panels are not all equal
    drv = SummaMtm(M, N, K, torch.float32, variant=headline)
    r0, r1, c0, c1 = drv.my_block
    (ka0, ka1), (kb0, kb1) = drv.my_a_cols, drv.my_b_rows
    g = torch.Generator(device="cuda").manual_seed(11)    # same stream of numbers on every rank
    A = torch.randint(0, 10, (M, K), device="cuda", generator=g).float()
    B = torch.randint(0, 10, (K, N), device="cuda", generator=g).float()
    C0 = torch.randint(0, 10, (M, N), device="cuda", generator=g).float()
    c = C0[r0:r1, c0:c1].clone()
    a = A[r0:r1, ka0:ka1].clone()
    b = B[kb0:kb1, c0:c1].clone()
    drv.step(c, a, b)
    drv.step(c, a, b)
    torch.cuda.synchronize()
    want = C0[r0:r1, c0:c1].double() + 2 * (A[r0:r1].double() @ B[:, c0:c1].double())
    ok = torch.tensor([int(torch.equal(c.double(), want))], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    del A, B, C0, want
    a.uniform_(-1, 1)
    b.uniform_(-1, 1)
    c.zero_()
    ms = timed_steps(torch, dist, world, lambda: drv.step(c, a, b), 1, 3)
    return {"workload": f"mtm fp32 {M}x{N}x{K} last_order, 2-D SUMMA split, every operand stored once across the grid",
            "grid": [drv.Pr, drv.Pc], "panels": len(drv.panels), "exact": bool(ok.item()),
            "tflops": round(flops(M, N, K) / ms / 1e9, 2), "ms_per_step": round(ms, 3)}


def e2e_sharded(ob, torch, dist, rank, world, sharded, M, N, K, steps):
    """N > 1 end to end THROUGH THE SHARDED PATH: every rank's rows of A and C start in pinned host memory, B in
    the root's; per step each rank uploads its rows (its own PCIe link), the root uploads B once and replicates
    it over NVLink inside RowBlockMtm.step, and each rank reads its rows of C back."""
    root = rank == 0
    ha = torch.empty((M, K), dtype=torch.float32, pin_memory=True).uniform_(-1, 1)
    hc = torch.zeros((M, N), dtype=torch.float32, pin_memory=True)
    hb = torch.empty((K, N), dtype=torch.float32, pin_memory=True).uniform_(-1, 1) if root else None
    da = torch.empty((M, K), device="cuda", dtype=torch.float32)
    dc = torch.empty((M, N), device="cuda", dtype=torch.float32)
    db = torch.empty((K, N), device="cuda", dtype=torch.float32) if root else None
    side = torch.cuda.Stream()

    def step():
        cur = torch.cuda.current_stream()
        if root:
            db.copy_(hb, non_blocking=True)          # B first: its replication is the critical path of every rank
        with torch.cuda.stream(side):                # A and C ride a second stream (same link, queued behind B on the root)
            side.wait_stream(cur)
            da.copy_(ha, non_blocking=True)
            dc.copy_(hc, non_blocking=True)
        cur.wait_stream(side)
        sharded.step(dc, da, db)
        hc.copy_(dc, non_blocking=True)
        torch.cuda.synchronize()

    step()
    dist.barrier()
    t = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t) / steps
    dist.barrier()
    dt = max_over_ranks(torch, dist, world, dt)
    chk = float(hc[0, 0])
    del da, dc, db
    return dt, chk


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="auto", help="auto | simt | 3xtf32 (headline fp32 kernel family)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary sweeps")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--push-ctas", type=int, default=0, help="NVLink push kernel CTAs (0 = 32, -1 = copy engines)")
    ap.add_argument("--watchdog-s", type=int, default=600, help="seconds the secondary measurements may take before the line is printed without them")
    ap.add_argument("--config5-size", type=int, default=32768, help="size of the strong-scaled configs[4] record (0 = skip it)")
    ap.add_argument("--bcast", default="auto", choices=["nccl", "nvlink", "auto"],
                    help="N>1: how B is replicated (NCCL broadcast | this library's NVLink multicast push kernels)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference_arm(args)

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's banner off stdout (one JSON line there)
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        ge.build_library()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the mtm path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    bound_cpus = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    import openmp_blas_b200 as ob
    ob.lib()
    info = ob.device_info(local_rank)
    peaks, peak_src = measured_peaks()

    headline = args.variant
    if headline == "auto":
        headline = "3xtf32" if ob.num_configs("3xtf32", False) > 0 else "simt"

    M = N = K = SIZE
    a = dev_uniform(torch, (M, K), torch.float32, "L", 0xB200 + rank)
    c = torch.zeros((M, N), device="cuda", dtype=torch.float32)
    if world == 1:
        b = dev_uniform(torch, (K, N), torch.float32, "L", 0xB201)
        step_fn = ob.mtm(c, a, b, None, variant=headline)
        launches_per_step = None
    else:
        from openmp_blas_b200.sharded import RowBlockMtm
        b_root = dev_uniform(torch, (K, N), torch.float32, "L", 0xB201) if rank == 0 else None
        sharded = RowBlockMtm(M_total=M * world, N=N, K=K, dtype=torch.float32, variant=headline, bcast=args.bcast,
                              push_ctas=args.push_ctas)
        step_fn = lambda: sharded.step(c, a, b_root)

    calibration = None
    if world > 1:
        calibration = sharded.calibrate(a, b_root, steps=2)      # own NVLink push vs NCCL broadcast: keep the faster
        bcast_used = ("nccl broadcast (K-chunked)" if not sharded.use_nvlink else
                      "own NVLink push (" + ("copy engines" if args.push_ctas < 0 else f"{args.push_ctas or 32} CTAs") + "), "
                      + ("NVSwitch multicast" if sharded.replicator.multicast else "unicast to each peer") + " (K-chunked, arrival flags)")
    for _ in range(args.warmup):
        step_fn()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ob.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step_fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t1 = time.time()
    ms_total = e0.elapsed_time(e1)
    launches = ob.launch_count() - launches0
    kernel_name = ob.last_choice()["name"]
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        per_rank = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(per_rank, t)
        per_rank_ms = [round(float(x.item()) / args.steps, 4) for x in per_rank]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_step = ms_total / args.steps
    total_flops = flops(M * world, N, K)
    value = total_flops / (ms_step * 1e-3) / 1e12
    if not torch.isfinite(c).all():
        raise SystemExit("non-finite values in C after the timed steps")

    # ---- roofline of the dominant kernel, timed live (per launch, same stream, same data) -----------
    b_local = b if world == 1 else dev_uniform(torch, (K, N), torch.float32, "L", 0xB201)
    # "timed alone" = after the GPU has idled for a moment (as MEASURED_PEAKS.json's burst figure is taken), a few
    # launches only; the rate inside the long loop above is reported next to it as `sustained`
    time.sleep(2.0)
    ms_kernel = time_device(ob, torch, c, a, b_local, headline, None, 1, 5)
    kname = ob.last_choice()["name"]
    alg_tflops = flops(M, N, K) / (ms_kernel * 1e-3) / 1e12
    if headline == "3xtf32":
        tf32_peak = peaks["bf16_tflops"] / 2.0
        tf32_sust = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) / 2.0
        step_alg_tflops = flops(M, N, K) / (ms_step * 1e-3) / 1e12 if world == 1 else None
        burst = {"achieved": round(3.0 * alg_tflops, 2), "peak": round(tf32_peak, 1), "frac": round(3.0 * alg_tflops / tf32_peak, 4),
                 "ms_per_launch": round(ms_kernel, 4),
                 "note": f"the same call (split pass + MMA kernel) timed alone: 5 launches after 2 s of idling, against the {peak_src} "
                         f"cuBLAS bf16 BURST figure ({peaks['bf16_tflops']}) / 2"}
        if step_alg_tflops is not None:
            # N = 1: the dominant kernel is what the timed region consists of, so its average launch duration is the
            # step time, and the peak that goes with a kernel timed inside a long step is the SUSTAINED one
            roofline = {"bound": "tensor", "achieved": round(3.0 * step_alg_tflops, 2), "peak": round(tf32_sust, 1),
                        "unit": "TFLOP/s", "frac": round(3.0 * step_alg_tflops / tf32_sust, 4), "traffic": None,
                        "kernel": kname, "ms_per_launch": round(ms_step, 4),
                        "note": f"3xTF32 issues 3 tcgen05 kind::tf32 MMAs per algorithmic MAC: achieved = 3 * {step_alg_tflops:.1f} "
                                f"algorithmic TFLOP/s; ms_per_launch = average duration of one call (operand-split pass + MMA kernel) "
                                f"over the {args.steps} steps of the timed region (CUDA events on the launching stream); peak = TF32 "
                                f"dense = {peak_src} cuBLAS bf16 SUSTAINED figure ({peaks.get('bf16_tflops_sustained')}) / 2, the one "
                                f"that goes with a kernel timed inside a long step (both run power-capped); `burst` = timed alone "
                                f"against the burst figure",
                        "burst": burst}
        else:
            roofline = {"bound": "tensor", "achieved": burst["achieved"], "peak": burst["peak"], "unit": "TFLOP/s",
                        "frac": burst["frac"], "traffic": None, "kernel": kname, "ms_per_launch": burst["ms_per_launch"],
                        "note": "3xTF32 issues 3 tcgen05 kind::tf32 MMAs per algorithmic MAC; " + burst["note"]}
    else:
        p32 = info["peak_fp32_tflops"]
        roofline = {"bound": "fp32-fma", "achieved": round(alg_tflops, 2), "peak": round(p32, 2), "unit": "TFLOP/s",
                    "frac": round(alg_tflops / p32, 4), "traffic": None, "kernel": kname,
                    "ms_per_launch": round(ms_kernel, 4),
                    "note": "CUDA-core kernel: bound is the FP32 FMA pipe (148 SMs * 128 lanes * 2 * max SM clock), "
                            "not HBM or the tensor pipe; MEASURED_PEAKS.json has no FP32-SIMT figure"}
    roofline["traffic"], roofline["traffic_stale"] = ncu_traffic(kname)
    roofline["kernel_source_hash"] = kernel_source_hash(kname)
    try:
        # what the committed ncu capture of this kernel (same stamp as `traffic`) says the bounding pipe was doing, and the fraction
        # of the NOMINAL pipe at the maximum SM clock: the measured-cuBLAS proxy above is itself power-capped, so a fraction of it
        # can exceed 1 — these two say how far the kernel is from the hardware
        rec = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text()).get(kname.split("_tailsplit")[0].split("_splitk")[0], {})
        if rec.get("pipe_active_pct") is not None:
            roofline["ncu_pipe_active_pct"] = rec["pipe_active_pct"]
            roofline["ncu_pipe"] = rec.get("pipe")
            roofline["ncu_sm_ghz"] = rec.get("sm_ghz")
        if headline == "3xtf32":
            nominal = info["sm_count"] * 2048 * 2 * (info["sm_clock_khz"] * 1e3) / 1e12     # TF32 dense: 2048 MAC / clk / SM
            roofline["nominal_tf32_dense_tflops"] = round(nominal, 1)
            roofline["frac_of_nominal"] = round(roofline["achieved"] / nominal, 4)
    except Exception:
        pass
    alg_bytes = 4.0 * (M * K + K * N + 2.0 * M * N)
    roofline["hbm_check"] = {"algorithmic_bytes": alg_bytes, "achieved_gbs": round(alg_bytes / (ms_kernel * 1e-3) / 1e9, 1),
                             "peak_gbs": peaks["hbm_gbs"], "source": peak_src}
    if world > 1:
        del b_local

    # ---- e2e: host buffers through the public API, copies inside the timed region ------------------
    e2e = None
    e2e_steps = max(2, min(args.steps, 5))
    if world == 1:
        ha = ob.pinned_empty((M, K), np.float32)
        hb = ob.pinned_empty((K, N), np.float32)
        hc = ob.pinned_empty((M, N), np.float32)
        rng = np.random.default_rng(0xB200 + rank)
        ha[...] = rng.uniform(-1, 1, (M, K)).astype(np.float32)
        hb[...] = np.random.default_rng(0xB201).uniform(-1, 1, (K, N)).astype(np.float32)
        hc[...] = 0
        fn = ob.mtm(hc, ha, hb, None, variant=headline)
        fn()
        t = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        dt = (time.perf_counter() - t) / e2e_steps
        e2e = {"value": round(total_flops / dt / 1e12, 3), "unit": "TFLOP/s",
               "h2d_bytes_per_step": int(4 * (M * K + K * N + M * N)), "d2h_bytes_per_step": int(4 * M * N),
               "ms_per_step": round(dt * 1e3, 3),
               "api": "openmp_blas_b200.mtm(c, a, b)() on pinned numpy arrays -> b200_mtm_f32 (host-pointer C ABI)",
               "checksum_c00": float(hc[0, 0])}
        # what the link allows: the call's bytes (A, B, C up; C down) as bare pinned copies on two streams, nothing else
        try:
            up = torch.empty(M * K + K * N + M * N, dtype=torch.float32, pin_memory=True)
            down = torch.empty(M * N, dtype=torch.float32, pin_memory=True)
            dup, ddown = torch.empty(up.numel(), device="cuda"), torch.empty(down.numel(), device="cuda")
            s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()

            def bare_copies():
                with torch.cuda.stream(s_up):
                    dup.copy_(up, non_blocking=True)
                with torch.cuda.stream(s_down):
                    down.copy_(ddown, non_blocking=True)
            bare_copies()
            torch.cuda.synchronize()
            t = time.perf_counter()
            for _ in range(3):
                bare_copies()
            torch.cuda.synchronize()
            floor = (time.perf_counter() - t) / 3
            e2e["pcie_copy_floor_ms"] = round(floor * 1e3, 3)
            e2e["frac_of_copy_floor"] = round(floor / dt, 4)
            del up, down, dup, ddown
        except Exception as ex:          # a context line must never cost the headline
            e2e["pcie_copy_floor_ms"] = None
            e2e["pcie_copy_floor_error"] = str(ex)[:120]
        # the same call on ordinary (pageable) numpy arrays: what the reference's make_tensor storage is (src/mtm.cpp:204-208)
        pa, pb, pc = np.array(ha), np.array(hb), np.zeros((M, N), np.float32)
        for h in (ha, hb, hc):
            ob.pinned_free(h)
        fnp = ob.mtm(pc, pa, pb, None, variant=headline)
        fnp()
        t = time.perf_counter()
        for _ in range(e2e_steps):
            fnp()
        dtp = (time.perf_counter() - t) / e2e_steps
        e2e["pageable"] = {"value": round(total_flops / dtp / 1e12, 3), "unit": "TFLOP/s", "ms_per_step": round(dtp * 1e3, 3),
                           "api": "same call on pageable numpy arrays (np.zeros / np.array storage)"}
        del pa, pb, pc
    else:
        dt, chk = e2e_sharded(ob, torch, dist, rank, world, sharded, M, N, K, e2e_steps)
        by_ranks = {"value": round(total_flops / dt / 1e12, 3), "unit": "TFLOP/s", "ms_per_step": round(dt * 1e3, 3),
                    "api": ("openmp_blas_b200.sharded.RowBlockMtm.step on pinned host shards: every rank uploads its rows of A and C "
                            "over its own PCIe link, the root uploads B ONCE and replicates it over NVLink, every rank reads its rows "
                            "of C back"), "checksum_c00": chk}
        # The reference-facing call: ONE b200_mtm_f32_mgpu call from ONE process (rank 0) holding the whole
        # (8192 N) x 8192 x 8192 problem in pinned host memory, spread over the N GPUs by the library itself.  The
        # other ranks wait on the rendezvous store (a host wait: an NCCL barrier would keep a spinning kernel on
        # the very GPUs rank 0's call is using).
        store = dist.distributed_c10d._get_default_store()
        one_call = None
        if rank == 0:
            try:
                Mt = M * world
                ha, hb, hc = ob.pinned_empty((Mt, K), np.float32), ob.pinned_empty((K, N), np.float32), ob.pinned_empty((Mt, N), np.float32)
                torch.from_numpy(ha).uniform_(-1, 1)
                torch.from_numpy(hb).uniform_(-1, 1)
                hc[...] = 0
                torch.cuda.synchronize()
                fn = ob.mtm(hc, ha, hb, None, variant=headline, devices=world)
                fn()
                t = time.perf_counter()
                for _ in range(e2e_steps):
                    fn()
                dtm = (time.perf_counter() - t) / e2e_steps
                one_call = {"value": round(total_flops / dtm / 1e12, 3), "unit": "TFLOP/s", "ms_per_step": round(dtm * 1e3, 3),
                            "api": (f"ONE openmp_blas_b200.mtm(c, a, b, devices={world})() call on pinned numpy arrays -> b200_mtm_f32_mgpu "
                                    "(C ABI; one process, a host thread per GPU, B uploaded once as N slices and forwarded over NVLink)"),
                            "kernel": ob.last_choice()["name"], "launches": ob.last_choice()["launches"], "checksum_c00": float(hc[0, 0])}
                for h in (ha, hb, hc):
                    ob.pinned_free(h)
            except Exception as ex:
                one_call = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}
            store.set("b200_bench_mgpu_done", "1")
        else:
            import datetime
            store.wait(["b200_bench_mgpu_done"], datetime.timedelta(seconds=600))
        dist.barrier()
        h2d = int(4 * (M * K + M * N)) * world + int(4 * K * N)
        best_is_call = rank == 0 and one_call is not None and "value" in one_call and one_call["value"] >= by_ranks["value"]
        head = one_call if best_is_call else by_ranks
        e2e = {"value": head["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(4 * M * N) * world,
               "ms_per_step": head["ms_per_step"], "api": head["api"], "checksum_c00": head.get("checksum_c00"),
               "one_call_mgpu_c_abi": one_call, "one_process_per_gpu": by_ranks}

    # The headline, roofline and e2e are measured: assemble the line now.  What follows (config 5, extras, the CPU
    # baseline) only ADDS to it, and a watchdog prints the line as it stands if that part hangs (a multi-rank
    # collective that one rank left early would otherwise take the whole record with it).
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic uniform(-1,1), fixed seeds, generated on device",
            "config": {
                "workload": (f"mtm fp32 {M * world}x{N}x{K} last_order (row-major), C += A*B, variant {headline}"
                             + ("" if world == 1 else f", row-block sharded {M} rows/GPU, B broadcast from rank 0 every step")),
                "kernel": kernel_name, "flops_per_step": total_flops, "flop_count": "M*N*(2K-1) (src/mtm.cpp:203)",
                "l2": "inputs (3 x 256 MiB per GPU) exceed the 126 MB L2; no flush between steps",
                "device": info["name"], "sm_count": info["sm_count"],
                **({} if world == 1 else {"b_replication": bcast_used, "k_chunks": sharded.chunks, "calibration_ms_per_step": calibration,
                                           "ms_per_step_by_rank": per_rank_ms,
                                           "host_threads_bound_to_gpu_numa_cpus": bound_cpus}),
            },
            "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "peaks_source": peak_src, "config5": None,
        }
    printed = threading.Lock()

    def emit(final: bool):
        if not printed.acquire(blocking=False):
            return
        if rank == 0:
            if not final:
                line["watchdog"] = f"the secondary measurements exceeded {args.watchdog_s} s and were cut off"
            print(json.dumps(line), flush=True)
        if not final:
            os._exit(0)

    watchdog = threading.Timer(args.watchdog_s + (0 if rank == 0 else 5), emit, args=(False,))
    watchdog.daemon = True
    watchdog.start()

    # ---- BASELINE configs[4]: 32768^3 strong-scaled over the N GPUs, with an in-run exactness flag ----
    if args.config5_size > 0:
        del a, c
        if world == 1:
            del b
        else:
            del b_root
        step_fn = None
        torch.cuda.empty_cache()
        try:
            config5 = config5_record(ob, torch, dist, rank, world, headline, args)
        except Exception as e:          # (collective code: every rank fails alike or the watchdog ends the run)
            config5 = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        if rank == 0:
            line["config5"] = config5
        torch.cuda.empty_cache()
        if world > 1:
            try:
                summa = summa_record(ob, torch, dist, rank, world, headline)
            except Exception as e:
                summa = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
            if rank == 0:
                line["summa_2d"] = summa
            torch.cuda.empty_cache()

    if rank == 0 and world == 1:
        if not args.no_extras:
            if args.config5_size <= 0:
                del a, b, c
            torch.cuda.empty_cache()
            line["extras"] = extras_single_gpu(ob, torch, info, peaks, args.quick)
        if not args.no_cpu:
            r = cpu_reference_run(steps=4, warmup=1, budget_s=20.0)   # amt::benchmark<4> protocol, src/mtm.cpp:373
            line["cpu_baseline"] = {"value": round(r["value"], 4), "unit": "TFLOP/s", "cores": r["cores"], "kind": r["kind"],
                                    "sample": r["sample"]}
            if "extras" in line:
                line["extras"]["cpu"] = cpu_extras()

    watchdog.cancel()
    emit(True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
