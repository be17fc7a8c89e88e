"""Row-block sharded mtm across the GPUs of one box: one process per GPU, torch.distributed (NCCL).

The reference is single-process shared-memory (its only parallelism is the OpenMP team of
include/mtm.hpp:156-201, which work-shares M-blocks).  The multi-GPU analogue keeps that
decomposition: C[rows_r, :] += A[rows_r, :] * B — rank r owns a block of rows of A and C, row
blocks are independent given all of B, and K is never split across ranks, so no reduction is
needed and every rank's result is bit-identical to the single-GPU kernel on the same rows.

The one exchange step is replicating B (K x N) from the root.  It is issued as K-chunk
broadcasts (row slabs of a row-major B are contiguous) on NCCL's stream, and because mtm
ACCUMULATES (C += A*B, simd_loop.hpp:169,187) the product is issued as one mtm call per K-chunk,
    C += A[:, k0:k1] * B[k0:k1, :],
each waiting only for its own chunk: the broadcast of chunk i+1 overlaps the product of chunk i.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

__all__ = ["row_partition", "k_chunks", "plan_chunks", "RowBlockMtm", "split_even", "choose_grid",
           "summa_panels", "SummaMtm"]


def row_partition(M: int, world: int, align: int = 128) -> List[Tuple[int, int]]:
    """[begin, end) row range of every rank: ceil(M/world) rows rounded up to the CTA tile height,
    trailing ranks may get fewer (or zero) rows."""
    per = -(-M // world)
    per = -(-per // align) * align
    out = []
    for r in range(world):
        b = min(M, r * per)
        e = min(M, (r + 1) * per)
        out.append((b, e))
    return out


def k_chunks(K: int, n_chunks: int = 3, align: int = 32, ratio: Optional[float] = None) -> List[Tuple[int, int]]:
    """Split [0, K) into contiguous slabs (lengths multiples of `align`) for the pipelined broadcast.

    Every chunk costs one extra read-modify-write pass over C and one more set of launches, so few
    chunks are better; what has to be hidden is the broadcast of chunk i behind the products of the
    chunks before it.  The slabs therefore GROW geometrically: a short first one (its broadcast is the
    only exposed transfer), then longer ones.

    With ``ratio`` = (time to broadcast all of B) / (time of the local product) the schedule is derived
    from the pipeline model: chunk i+1 has landed by the time chunk i's product ends iff the cumulative
    fractions satisfy F[i+1] <= F[i] / ratio (+ the first chunk), so F grows by 0.8 / ratio per step from
    F[0] = 1/32.  Without ``ratio``: n_chunks = 2 -> 1/8, 7/8; 3 -> 1/16, 5/16, 10/16.
    """
    units = -(-K // align)
    if ratio is not None:
        ratio = min(max(ratio, 1e-3), 0.7)
        growth = max(1.5, 0.8 / ratio)
        cum, f = [], max(1.0 / 32.0, 1.0 / units)
        while f < 1.0 and len(cum) < 7:
            cum.append(f)
            f *= growth
        cum.append(1.0)
        bounds = sorted({min(units, max(1, int(round(c * units)))) for c in cum})
        if bounds[-1] != units:
            bounds.append(units)
        out, k = [], 0
        for b in bounds:
            k1 = min(K, b * align)
            if k1 > k:
                out.append((k, k1))
            k = k1
        return out
    n_chunks = max(1, min(n_chunks, units))
    if n_chunks == 1:
        return [(0, K)]
    weights = {2: [1, 7], 3: [1, 5, 10]}.get(n_chunks)
    if weights is None:
        weights = [1] + [max(1, round(15 * (i + 1) * 2 / (n_chunks * (n_chunks - 1)))) for i in range(n_chunks - 1)]
    total = float(sum(weights))
    out, k = [], 0
    for i, w in enumerate(weights):
        if i == len(weights) - 1:
            k1 = K
        else:
            k1 = min(K, k + max(align, int(round(units * w / total)) * align))
        if k1 > k:
            out.append((k, k1))
        k = k1
    if out[-1][1] < K:
        out[-1] = (out[-1][0], K)
    return out


def plan_chunks(K: int, t_bcast: float, t_comp: float, t_chunk_overhead: float, align: int = 32) -> List[Tuple[int, int]]:
    """Pick the K-chunk schedule that minimises the modelled step time.

    Model (validated against profiles/r01f, r01e: 19.34 ms predicted vs 19.36 measured at 16384^3 on
    2 GPUs): broadcasts run back to back from t = 0, chunk i has landed at t_bcast * F[i] (F = cumulative
    fraction of K); its product starts when it has landed and the previous product is done and takes
    (F[i] - F[i-1]) * t_comp + t_chunk_overhead (the extra read-modify-write of C, the operand re-layout
    launches and the short-K inefficiency every additional call pays).  Candidates: geometric schedules
    F[i] = f1 * g^i.  One chunk means no overlap at all."""
    units = max(1, -(-K // align))
    best, best_t = [1.0], t_bcast + t_comp + t_chunk_overhead
    for f1 in (1 / 32, 1 / 16, 1 / 8, 1 / 4, 1 / 2):
        if f1 * units < 1 or (f1 * K < 512 and f1 < 0.5):   # very short-K calls are epilogue-dominated
            continue
        for g in (2.0, 3.0, 4.0, 6.0, 8.0, 16.0, 32.0):
            cum, f = [], f1
            while f < 1.0 and len(cum) < 6:
                cum.append(f)
                f *= g
            cum.append(1.0)
            end, prev = 0.0, 0.0
            for F in cum:
                end = max(t_bcast * F, end) + (F - prev) * t_comp + t_chunk_overhead
                prev = F
            if end < best_t - 1e-9:
                best, best_t = cum, end
    bounds = sorted({min(units, max(1, int(round(c * units)))) for c in best})
    if bounds[-1] != units:
        bounds.append(units)
    out, k = [], 0
    for b in bounds:
        k1 = min(K, b * align)
        if k1 > k:
            out.append((k, k1))
        k = k1
    return out


class NvlinkReplicator:
    """B's replica on every GPU, filled by the root with this library's own NVLink kernels
    (include/b200_replicate.h) instead of a collective library's copy kernels.

    Two symmetric (peer-mapped) allocations from torch.distributed's symmetric memory: the replica
    (``depth`` slots used round-robin by successive steps, so the root does not have to wait for the
    receivers of the step before) and a page of uint32 flag words.  With NVSwitch multicast the root
    streams each K-chunk into the multicast address once (the switch replicates the stores to every
    GPU: root egress is 1x B whatever the world size, and the receivers run NO communication kernel —
    their SMs stay with the product); without it the same kernel stores to each peer's mapped buffer in
    turn.  ``ctas`` < 0 moves the data with the root's copy engines instead of SMs.
    Flag layout (uint32 words): [0] = chunk sequence number that has landed in THIS GPU's replica
    (written by the root after the chunk's data); [8 + r] on the ROOT = number of steps receiver r
    has finished reading.  Setup is collective over the default process group; raises if symmetric
    memory cannot be established (the caller then stays on the NCCL broadcast).
    """
    ARRIVED, CONSUMED = 0, 8

    def __init__(self, shape, dtype, root: int, device, ctas: int = 0, depth: int = 2):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.rank, self.world, self.root = dist.get_rank(), dist.get_world_size(), root
        if self.world > 8:
            raise ValueError("NvlinkReplicator: one NVSwitch box (<= 8 GPUs)")
        self.ctas, self.depth = ctas, max(1, int(depth))
        self.shape = tuple(shape)
        self.slot_elems = 1
        for d in self.shape:
            self.slot_elems *= int(d)
        self.slot_elems = -(-self.slot_elems // 64) * 64          # keep every slot 256-byte aligned
        # Stage 1 is local (allocation): agree on its outcome BEFORE anyone enters the collective stages below, so a
        # rank that cannot allocate makes every rank abandon together instead of leaving the others in a rendezvous.
        err = None
        try:
            self.store = symm.empty((self.depth * self.slot_elems,), dtype=dtype, device=device)
            self.flags = symm.empty((64,), dtype=torch.int32, device=device)
            self.flags.zero_()
            torch.cuda.synchronize(device)
        except Exception as e:
            err = e
        ok = torch.tensor([0 if err is not None else 1], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not bool(ok.item()):
            raise RuntimeError(f"NvlinkReplicator: symmetric allocation failed on a rank ({err or 'a peer'})")
        self.h_buf = symm.rendezvous(self.store, dist.group.WORLD)
        self.h_flags = symm.rendezvous(self.flags, dist.group.WORLD)
        bp, fp = list(self.h_buf.buffer_ptrs), list(self.h_flags.buffer_ptrs)
        off_b, off_f = self.store.data_ptr() - bp[self.rank], self.flags.data_ptr() - fp[self.rank]
        self.peer_buf = [p + off_b for p in bp]
        self.peer_flags = [p + off_f for p in fp]
        mc_b, mc_f = int(self.h_buf.multicast_ptr or 0), int(self.h_flags.multicast_ptr or 0)
        self.multicast = bool(mc_b) and bool(mc_f)
        self.mc_buf = mc_b + off_b if self.multicast else 0
        self.mc_flags = mc_f + off_f if self.multicast else 0
        self.seq = 0            # chunk sequence number (same on every rank: all walk the same schedule)
        self.steps_done = 0
        self.stream = torch.cuda.Stream(device=device) if self.rank == root else None
        dist.barrier()          # every rank's flag page is zeroed and mapped before anyone writes to it
        if self.ctas < 0 and self.multicast and not self._probe_copy_engine_multicast(device):
            self.multicast_data = False     # DMA engines cannot target the multicast mapping here: unicast copies
        else:
            self.multicast_data = self.multicast

    @property
    def buf(self):
        """This step's replica slot as a tensor of the operand's shape."""
        k = self.steps_done % self.depth
        n = 1
        for d in self.shape:
            n *= int(d)
        return self.store[k * self.slot_elems: k * self.slot_elems + n].view(self.shape)

    def _slot_byte_off(self) -> int:
        return (self.steps_done % self.depth) * self.slot_elems * self.store.element_size()

    def _probe_copy_engine_multicast(self, device) -> bool:
        """Collective: can cudaMemcpyAsync write through the multicast mapping?  The root copies a small
        pattern into slot 0 of every replica; every rank checks what landed."""
        import torch
        import torch.distributed as dist
        from . import B200Error, replicate_push
        n = min(4096, self.slot_elems)
        pat = torch.arange(1, n + 1, device=device).to(self.store.dtype)
        self.store[:n].zero_()
        torch.cuda.synchronize(device)
        dist.barrier()
        ok = 1
        if self.rank == self.root:
            try:
                replicate_push([self.mc_buf], pat.data_ptr(), n * pat.element_size(), [], 0, multicast=True,
                               flag_multicast=False, ctas=-1, stream=torch.cuda.current_stream(device).cuda_stream)
                torch.cuda.synchronize(device)
            except (B200Error, RuntimeError):
                ok = 0
        dist.barrier()
        if ok and not torch.equal(self.store[:n], pat):
            ok = 0
        t = torch.tensor([ok], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        self.store[:n].zero_()
        torch.cuda.synchronize(device)
        dist.barrier()
        return bool(t.item())

    def push_chunk(self, src, byte_off: int, nbytes: int) -> None:
        """Root: replicate `nbytes` of `src` (device tensor, contiguous) at `byte_off` of this step's slot,
        then publish the next sequence number.  Runs on the replicator's side stream."""
        from . import replicate_push
        self.seq += 1
        off = self._slot_byte_off() + byte_off
        peers = [r for r in range(self.world) if r != self.root]
        dst = [self.mc_buf + off] if self.multicast_data else [self.peer_buf[r] + off for r in peers]
        if self.multicast:
            fl = [self.mc_flags + 4 * self.ARRIVED]
        else:
            fl = [self.peer_flags[r] + 4 * self.ARRIVED for r in peers]
        replicate_push(dst, src.data_ptr(), nbytes, fl, self.seq, multicast=self.multicast_data,
                       flag_multicast=self.multicast, ctas=self.ctas, stream=self.stream.cuda_stream)

    def push_panel(self, src, col0: int, cols: int) -> None:
        """Root: replicate columns [col0, col0 + cols) of the row-major matrix `src` (all rows)
        into the same columns of this step's slot, then publish the next sequence number."""
        from . import replicate_push_2d
        self.seq += 1
        esz = src.element_size()
        rows, ld = int(src.shape[0]), int(src.stride(0))
        off = self._slot_byte_off() + col0 * esz
        peers = [r for r in range(self.world) if r != self.root]
        dst = [self.mc_buf + off] if self.multicast_data else [self.peer_buf[r] + off for r in peers]
        if self.multicast:
            fl = [self.mc_flags + 4 * self.ARRIVED]
        else:
            fl = [self.peer_flags[r] + 4 * self.ARRIVED for r in peers]
        replicate_push_2d(dst, src.data_ptr() + col0 * esz, rows, cols * esz, ld * esz, int(self.shape[1]) * esz, fl,
                          self.seq, multicast=self.multicast_data, flag_multicast=self.multicast, ctas=self.ctas,
                          stream=self.stream.cuda_stream)

    def wait_receivers(self) -> None:
        """Root, side stream: every receiver has finished reading the slot this step overwrites
        (i.e. has completed step `steps_done - depth`)."""
        from . import flag_wait
        flag_wait(self.peer_flags[self.root] + 4 * self.CONSUMED, self.steps_done - self.depth + 1, count=self.world,
                  stride=1, skip=self.root, stream=self.stream.cuda_stream)

    def wait_chunk(self, stream: int) -> None:
        """Receiver: `stream` waits until the next chunk of the schedule has landed."""
        from . import flag_wait
        self.seq += 1
        flag_wait(self.peer_flags[self.rank] + 4 * self.ARRIVED, self.seq, stream=stream)

    def signal_consumed(self, stream: int) -> None:
        """Receiver: tell the root (after `stream`'s prior work) that this step's replica has been read."""
        from . import flag_signal
        flag_signal(self.peer_flags[self.root] + 4 * (self.CONSUMED + self.rank), self.steps_done + 1, stream=stream)

    def replicate(self, src=None, chunks=None) -> None:
        """One whole replication step without a product (timing / tests): the root pushes `src` chunk by
        chunk (row ranges `chunks` of its first dimension, default one chunk), receivers wait for them."""
        import torch
        chunks = chunks or [(0, self.shape[0])]
        row_bytes = self.store.element_size()
        for d in self.shape[1:]:
            row_bytes *= int(d)
        cur = torch.cuda.current_stream(self.store.device)
        if self.rank == self.root:
            ev = torch.cuda.Event()
            ev.record(cur)
            self.stream.wait_event(ev)
            self.wait_receivers()
            for (k0, k1) in chunks:
                self.push_chunk(src[k0:k1], k0 * row_bytes, (k1 - k0) * row_bytes)
            done = torch.cuda.Event()
            done.record(self.stream)
            cur.wait_event(done)
        else:
            for _ in chunks:
                self.wait_chunk(cur.cuda_stream)
            self.signal_consumed(cur.cuda_stream)
        self.steps_done += 1


def chunk_tile_config(k_len: int, dynamic: bool, has_double: bool = True) -> int:
    """3xTF32 tile config of one K-chunk call of the sharded drivers: the 256 x 256 CTA-pair tiles (config 0, or 2 with the
    dynamic tile scheduler for ranks whose SMs are shared with a collective's kernels) for chunks shorter than 6144, the
    256 x 512 double tiles (9 / 10) for longer ones — a quarter less L2 -> SM and DRAM traffic per flop pays on long K
    (16384^3: +11 %), their exposed accumulator drain costs on short K (8192^2 x 1024: -9 %)."""
    double = has_double and k_len >= 6144
    if dynamic:
        return 10 if double else 2
    return 9 if double else 0


class RowBlockMtm:
    """C_local += A_local * B with B broadcast from `root` inside every step.

    Operands are row-major (last_order): ``a_local`` is (rows_r x K), ``c_local`` (rows_r x N),
    ``b_root`` (K x N) is only read on the root rank (pass None elsewhere).
    ``local_mtm(c, a, b)`` performs one in-place accumulate; by default it is the CUDA path
    (``openmp_blas_b200.mtm``).  Tests inject a CPU checker to exercise the partition / chunking /
    broadcast plumbing over gloo — the product default never leaves the GPU.
    """

    def __init__(self, M_total: int, N: int, K: int, dtype, variant: str = "auto", n_chunks: Optional[int] = None,
                 root: int = 0, group=None, local_mtm: Optional[Callable] = None, device=None,
                 bcast_ctas: int = 0, config: Optional[int] = None, bcast: str = "auto", push_ctas: int = 0,
                 replica_depth: int = 2):
        """``bcast``: how B reaches the other GPUs — "nccl" (chunked ncclBroadcast), "nvlink" (this
        library's multicast push kernels, NvlinkReplicator; raises if unavailable) or "auto" (nvlink when
        every rank can set it up, else nccl)."""
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.root = root
        self.N, self.K = N, K
        self.rows = row_partition(M_total, self.world)
        # K-chunk schedules, one per replication path (plan_chunks: growing chunks from the pipeline model, fed
        # with the measured sustained 3xTF32 / FFMA / DMMA throughput and the replication bandwidth of the path on
        # this NVSwitch box — NCCL broadcast ~600 GB/s between 2 ranks, ~350 GB/s across 8; own multicast push
        # 520 GB/s into 2 GPUs, 390 GB/s into 8 (profiles/r01r_*, r01s_*); calibrate() replaces the bandwidths by
        # what it measures).  A chunk call costs one more accumulate pass over the C shard plus launches.
        # Planned from rank 0's row count on EVERY rank: all ranks must walk the same schedule.
        self._dtype_is64 = str(dtype).endswith("float64")
        self._plan_variant = variant
        self._fixed_chunks = None if (n_chunks is None and self.world > 1) else k_chunks(K, n_chunks or 1)
        self._plans = {"nccl": self._plan("nccl"), "nvlink": self._plan("nvlink")}
        self.use_nvlink = False            # set below once the replicator exists; calibrate() may flip it
        self.calibration = None
        self.variant = variant
        self.config = config
        # While NCCL's broadcast kernels occupy some SMs, the tensor-core kernel's dynamic tile scheduler
        # (config 2) lets the CTA groups that do run take over the tiles of those that cannot start yet:
        # measured +4.6% on 2 GPUs at 16384^3 (profiles/r01m_mgpu2_static_vs_dynamic.json); on a GPU the
        # kernel has to itself static assignment is as fast or faster, so this is only chosen here.
        my_rows = self.rows[self.rank][1] - self.rows[self.rank][0]
        self._auto_sched = (config is None and self.world > 1 and variant in ("auto", "3xtf32")
                            and str(dtype).endswith("float32") and my_rows >= 1024 and N >= 1024 and K >= 1024)
        # double tiles (configs 9 static / 10 dynamic, 256 x 512 per CTA pair) for the long K-chunks when the library has them:
        # a quarter less L2 -> SM and DRAM traffic per flop (16384^3: +11 %); short chunks run the 256 x 256 tiles (configs 0 / 2),
        # whose epilogue is fully hidden (8192^2 x 1024: +9 %) — profiles/r03n_ab_peer_arrive.jsonl
        try:
            from . import num_configs as _num_configs
            has_double = _num_configs("3xtf32", False) > 10
        except Exception:                     # (host-logic tests run without the CUDA library)
            has_double = False
        self._has_double = has_double
        self._dynamic = True
        if self._auto_sched:
            self.variant, self.config = "3xtf32", 2
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.device = device
        self.replicator = None
        if bcast not in ("nccl", "nvlink", "auto"):
            raise ValueError(f"bcast must be nccl, nvlink or auto, got {bcast!r}")
        if (bcast != "nccl" and self.world > 1 and group is None and dist.is_initialized()
                and dist.get_backend() == "nccl" and device.type == "cuda"):
            esz = 8 if str(dtype).endswith("float64") else 4
            ok, err = all((k1 - k0) * N * esz % 16 == 0 and k0 * N * esz % 16 == 0 for k0, k1 in self._plans["nvlink"]), None
            if ok and self.world > 8:
                ok = False
            if ok:
                try:
                    import torch.distributed._symmetric_memory  # noqa: F401
                except Exception as e:
                    ok, err = False, e
            # The replicator's constructor is collective: agree on feasibility BEFORE entering it ...
            agree = torch.tensor([int(ok)], device=device)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN)
            ok = bool(agree.item())
            rep = None
            if ok:
                try:
                    rep = NvlinkReplicator((K, N), dtype, root, device, ctas=push_ctas, depth=replica_depth)
                except Exception as e:      # no symmetric memory on this box / torch build (raised on every rank alike)
                    ok, err = False, e
            # ... and on its outcome: every rank must take the same path
            agree = torch.tensor([int(ok)], device=device)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN)
            if bool(agree.item()):
                self.replicator = rep
                self.use_nvlink = True
                self._apply_scheduler()
            elif bcast == "nvlink":
                raise RuntimeError(f"bcast='nvlink' unavailable on rank {self.rank}: {err or 'a peer failed or chunks are not 16-byte multiples'}")
        elif bcast == "nvlink" and self.world > 1:
            raise RuntimeError("bcast='nvlink' needs the default NCCL process group and CUDA tensors")
        # Replica of B on the non-root ranks (the root multiplies straight out of b_root).
        if self.replicator is not None:
            self.b_buf = None if self.rank == root else self.replicator.buf
        else:
            self.b_buf = None if self.rank == root else torch.empty((K, N), dtype=dtype, device=device)
        # The broadcast runs concurrently with the products and NCCL's copy kernels occupy SMs.
        # `bcast_ctas` > 0 gives the broadcast its own communicator capped at that many CTAs and makes
        # the persistent tensor-core kernel leave as many SMs free.  Measured on 2 GPUs at 16384^3
        # (profiles/r01f_mgpu2_bcast_ctas.json): uncapped 454 TFLOP/s, cap 8 -> 428, 4 -> 374, 2 -> 292
        # (the capped broadcast is too slow to hide), so the default is the uncapped shared communicator.
        self.bcast_group = group
        self.reserve_sms = 0
        if (self.world > 1 and bcast_ctas and dist.is_initialized() and dist.get_backend(group) == "nccl"):
            try:
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = int(bcast_ctas)
                opts.config.min_ctas = 1
                ranks = dist.get_process_group_ranks(group) if group is not None else list(range(self.world))
                self.bcast_group = dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)
                self.reserve_sms = int(bcast_ctas)
            except Exception:       # older NCCL / torch without ncclConfig support: share the default group
                self.bcast_group = group
        if local_mtm is None:
            from . import mtm as _mtm

            def local_mtm(c, a, b):
                cfg = self.config
                if self._auto_sched:
                    cfg = chunk_tile_config(int(a.shape[1]), self._dynamic, self._has_double)
                _mtm(c, a, b, None, variant=self.variant, config=cfg, reserve_sms=self.reserve_sms)()
        self.local_mtm = local_mtm

    def _plan(self, path: str, bw: Optional[float] = None) -> List[Tuple[int, int]]:
        if self._fixed_chunks is not None:
            return self._fixed_chunks
        rows = max(self.rows[0][1] - self.rows[0][0], 1)
        esz = 8 if self._dtype_is64 else 4
        rate = 30e12 if self._dtype_is64 else (60e12 if self._plan_variant == "simt" else 240e12)
        t_comp = 2.0 * rows * self.N * self.K / rate
        if bw is None:
            if path == "nccl":
                bw = 600e9 if self.world <= 2 else (450e9 if self.world <= 4 else 350e9)
            else:
                bw = 520e9 if self.world <= 2 else (450e9 if self.world <= 4 else 390e9)
        t_bcast = self.K * self.N * esz / bw
        t_over = rows * self.N * esz / 2.5e12 + 3e-5      # the chunk's extra accumulate pass over C + two launches
        return plan_chunks(self.K, t_bcast, t_comp, t_over)

    def _apply_scheduler(self) -> None:
        """Tile hand-out of the tensor-core kernel for the active replication path: with NCCL every rank shares
        its SMs with the broadcast's copy kernels -> dynamic scheduler everywhere; with the own NVLink push the
        receivers run no communication kernel at all -> static assignment there (0-5 % faster on a GPU the kernel
        has to itself), dynamic only on the root, whose push kernels run next to its product.  Same tiles, same K
        order: the result bits do not depend on the scheduler."""
        if not self._auto_sched:
            return
        self._dynamic = (not self.use_nvlink) or self.rank == self.root
        self.config = 2 if self._dynamic else 0

    @property
    def chunks(self) -> List[Tuple[int, int]]:
        """K-chunk schedule of the active replication path."""
        return self._plans["nvlink" if self.use_nvlink else "nccl"]

    def calibrate(self, a_local, b_root=None, steps: int = 2) -> dict:
        """Collective.  Times `steps` full steps of each available replication path (own NVLink push kernels,
        NCCL broadcast) on a scratch C and keeps the faster one — the default then follows the measurement on
        THIS box instead of constants.  Returns / stores the per-path step times (ms, max over
        ranks)."""
        import torch
        if self.world == 1:
            self.calibration = {"paths": {}, "chosen": "local"}
            return self.calibration
        lo, hi = self.my_rows
        scratch = torch.zeros((hi - lo, self.N), dtype=a_local.dtype, device=a_local.device)
        paths = ["nccl"] + (["nvlink"] if self.replicator is not None else [])
        res = {}
        on_gpu = a_local.is_cuda
        for path in paths:
            self.use_nvlink = path == "nvlink"
            self._apply_scheduler()
            for _ in range(1):
                self.step(scratch, a_local, b_root)
            if on_gpu:
                torch.cuda.synchronize()
            self.dist.barrier(self.group)
            if on_gpu:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    self.step(scratch, a_local, b_root)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
            else:                       # host tensors (the gloo tests of the driver logic)
                import time
                t0 = time.perf_counter()
                for _ in range(steps):
                    self.step(scratch, a_local, b_root)
                ms = (time.perf_counter() - t0) * 1e3 / steps
            t = torch.tensor([ms], device=a_local.device, dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
            res[path] = float(t.item())
        chosen = min(res, key=res.get)
        self.use_nvlink = chosen == "nvlink"
        self._apply_scheduler()
        self.calibration = {"paths": {k: round(v, 4) for k, v in res.items()}, "chosen": chosen}
        del scratch
        return self.calibration

    @property
    def my_rows(self) -> Tuple[int, int]:
        return self.rows[self.rank]

    def step_first_order(self, c_local, a_root, b_local) -> None:
        """Column-major (first_order) operands: C[:, cols_r] += A * B[:, cols_r].

        The shard is taken along C's slow dimension so it stays contiguous: rank r owns a
        block of COLUMNS of C and B, and A (M x K, only read on the root) is the replicated operand.
        This is the row-block problem on the transposes, C^T[cols_r, :] += B^T[cols_r, :] * A^T, and a
        column-major matrix viewed transposed is row-major, so it reuses step() unchanged.
        Construct the driver with M_total = number of columns of C, N = rows of C, K = K.
        """
        self.step(c_local.t(), b_local.t(), None if a_root is None else a_root.t())

    def step(self, c_local, a_local, b_root=None, b_ready=None) -> None:
        """One pass: broadcast B chunk-wise, accumulate chunk products as the chunks land.

        ``b_ready`` (root, NVLink replicator only): a CUDA event after which b_root holds this step's B.
        Without it B is assumed to be produced by the work already enqueued on the current stream, so
        the replication starts after the root's previous product; with it (and replica_depth >= 2) the
        push of step s+1 overlaps the products of step s."""
        if self.world == 1:
            self.local_mtm(c_local, a_local, b_root)
            return
        if self.replicator is not None and self.rank != self.root:
            self.b_buf = self.replicator.buf          # this step's slot of the replica
        b = b_root if self.rank == self.root else self.b_buf
        if b is None:
            raise ValueError("b_root must be given on the root rank")
        if self.replicator is not None and self.use_nvlink:
            self._step_nvlink(c_local, a_local, b, b_ready)
            return
        works = [self.dist.broadcast(b[k0:k1], src=self.root, group=self.bcast_group, async_op=True)
                 for (k0, k1) in self.chunks]
        for (k0, k1), w in zip(self.chunks, works):
            w.wait()  # CUDA: makes the compute stream wait for this chunk only; the host does not block
            if c_local.shape[0] > 0:
                self.local_mtm(c_local, a_local[:, k0:k1], b[k0:k1])

    def _step_nvlink(self, c_local, a_local, b, b_ready=None) -> None:
        """step() with the NVLink replicator: the root pushes the K-chunks from a side stream while it
        multiplies; a receiver's compute stream waits on each chunk's arrival flag right before the
        chunk's product and reports back to the root after the last one."""
        import torch
        rep = self.replicator
        cur = torch.cuda.current_stream(self.device)
        esz = b.element_size()
        if self.rank == self.root:
            if not b.is_contiguous():
                raise ValueError("b_root must be a contiguous row-major (K x N) tensor")
            if b_ready is None:
                b_ready = torch.cuda.Event()
                b_ready.record(cur)              # B's producer is whatever the current stream holds
            rep.stream.wait_event(b_ready)
            rep.wait_receivers()
            for (k0, k1) in self.chunks:
                rep.push_chunk(b[k0:k1], k0 * self.N * esz, (k1 - k0) * self.N * esz)
            for (k0, k1) in self.chunks:
                if c_local.shape[0] > 0:
                    self.local_mtm(c_local, a_local[:, k0:k1], b[k0:k1])
            b.record_stream(rep.stream)
        else:
            for (k0, k1) in self.chunks:
                rep.wait_chunk(cur.cuda_stream)
                if c_local.shape[0] > 0:
                    self.local_mtm(c_local, a_local[:, k0:k1], b[k0:k1])
            rep.signal_consumed(cur.cuda_stream)
        rep.steps_done += 1


# ---- optional 2-D split (SUMMA) --------------------------------------------------------------------
def split_even(n: int, parts: int, align: int = 1) -> List[Tuple[int, int]]:
    """[begin, end) of `parts` consecutive pieces of [0, n), piece length a multiple of `align`
    (trailing pieces may be shorter or empty)."""
    per = -(-n // parts)
    per = -(-per // align) * align
    return [(min(n, i * per), min(n, (i + 1) * per)) for i in range(parts)]


def choose_grid(world: int, M: int, N: int) -> Tuple[int, int]:
    """Pr x Pc = world with C blocks as square as possible (minimises the panel traffic
    M*K/Pr... per rank: a rank receives K*(rows_r + cols_c) elements per step)."""
    best, best_cost = (world, 1), None
    for pr in range(1, world + 1):
        if world % pr:
            continue
        pc = world // pr
        cost = M / pr + N / pc
        if best_cost is None or cost < best_cost - 1e-9:
            best, best_cost = (pr, pc), cost
    return best


def summa_panels(K: int, Pr: int, Pc: int, panel: Optional[int] = None, align: int = 32) -> List[Tuple[int, int, int, int]]:
    """K-panels of the SUMMA loop as (k0, k1, owner_col_of_A, owner_row_of_B).

    A's K range is split over the Pc grid columns, B's over the Pr grid rows (`split_even`, aligned);
    a panel never crosses either boundary, so each one has exactly one owner per grid row (for A) and
    per grid column (for B).  `panel` caps the panel width (default: no further cut — every extra panel
    costs one more read-modify-write pass over the C block)."""
    a_parts = split_even(K, Pc, align)
    b_parts = split_even(K, Pr, align)
    cuts = sorted({p for rng in a_parts + b_parts for p in rng if 0 <= p <= K} | {0, K})
    out = []
    for k0, k1 in zip(cuts, cuts[1:]):
        if k1 <= k0:
            continue
        step = k1 - k0 if not panel else max(align, -(-panel // align) * align)
        k = k0
        while k < k1:
            e = min(k1, k + step)
            oa = next(i for i, (s, t) in enumerate(a_parts) if s <= k < t)
            ob = next(i for i, (s, t) in enumerate(b_parts) if s <= k < t)
            out.append((k, e, oa, ob))
            k = e
    return out


class SummaMtm:
    """2-D block partition of C over a Pr x Pc grid of ranks (SUMMA), for shapes where the row-block
    split runs out of rows per rank or a full replica of B does not fit beside the shard.

    Rank (pr, pc) = divmod(rank, Pc) owns the C block rows_pr x cols_pc, the A block rows_pr x ka_pc
    (A's K range split over grid columns) and the B block kb_pr x cols_pc (B's K range split over
    grid rows) — every operand is stored exactly once across the grid.  For each K-panel the owner
    column broadcasts its A panel along the grid row, the owner row broadcasts its B panel along the
    grid column, and every rank accumulates C_block += A_panel * B_panel.  mtm accumulates
    (simd_loop.hpp:169,187), so the panel loop needs no temporaries; K is walked in ascending order on
    every rank, so a rank's result equals the single-GPU kernel called panel by panel.  Panel t+1 is
    in flight while panel t is multiplied (two panel buffers per operand).

    Operands are row-major (last_order) blocks; ``local_mtm`` as in RowBlockMtm.
    """

    def __init__(self, M: int, N: int, K: int, dtype, grid: Optional[Tuple[int, int]] = None,
                 panel: Optional[int] = None, variant: str = "auto", group=None,
                 local_mtm: Optional[Callable] = None, device=None, config: Optional[int] = None):
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.Pr, self.Pc = grid if grid is not None else choose_grid(self.world, M, N)
        if self.Pr * self.Pc != self.world:
            raise ValueError(f"grid {self.Pr}x{self.Pc} does not match world size {self.world}")
        self.pr, self.pc = divmod(self.rank, self.Pc)
        self.M, self.N, self.K = M, N, K
        self.row_parts = row_partition(M, self.Pr)
        self.col_parts = split_even(N, self.Pc, 128)
        self.a_parts = split_even(K, self.Pc, 32)
        self.b_parts = split_even(K, self.Pr, 32)
        self.panels = summa_panels(K, self.Pr, self.Pc, panel)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.device = device
        # Sub-communicators: one per grid row (A panels travel along it) and one per grid column
        # (B panels).  Every rank has to take part in the creation of every group, in the same order.
        ranks = dist.get_process_group_ranks(group) if (group is not None and dist.is_initialized()) else list(range(self.world))
        self.row_group = self.col_group = None
        self.row_ranks = [ranks[self.pr * self.Pc + j] for j in range(self.Pc)]
        self.col_ranks = [ranks[i * self.Pc + self.pc] for i in range(self.Pr)]
        if self.world > 1:
            for i in range(self.Pr):
                g = dist.new_group(ranks=[ranks[i * self.Pc + j] for j in range(self.Pc)]) if self.Pc > 1 else None
                if i == self.pr:
                    self.row_group = g
            for j in range(self.Pc):
                g = dist.new_group(ranks=[ranks[i * self.Pc + j] for i in range(self.Pr)]) if self.Pr > 1 else None
                if j == self.pc:
                    self.col_group = g
        r0, r1 = self.row_parts[self.pr]
        c0, c1 = self.col_parts[self.pc]
        wmax = max((k1 - k0 for k0, k1, _, _ in self.panels), default=0)
        # Two panel buffers per operand; a rank that owns the panel multiplies straight out of its
        # block when the grid dimension is 1 (nothing to send), otherwise the owner stages its slice
        # into the buffer so that the broadcast always moves one contiguous matrix.
        self.a_buf = [torch.empty((r1 - r0, wmax), dtype=dtype, device=device) for _ in range(2)] if self.Pc > 1 else None
        self.b_buf = [torch.empty((wmax, c1 - c0), dtype=dtype, device=device) for _ in range(2)] if self.Pr > 1 else None
        self.variant, self.config = variant, config
        if local_mtm is None:
            from . import mtm as _mtm

            def local_mtm(c, a, b):
                _mtm(c, a, b, None, variant=self.variant, config=self.config)()
        self.local_mtm = local_mtm

    @property
    def my_block(self) -> Tuple[int, int, int, int]:
        """(row0, row1, col0, col1) of this rank's block of C."""
        return (*self.row_parts[self.pr], *self.col_parts[self.pc])

    @property
    def my_a_cols(self) -> Tuple[int, int]:
        """K range of the A block this rank stores (rows = my_block rows)."""
        return self.a_parts[self.pc]

    @property
    def my_b_rows(self) -> Tuple[int, int]:
        """K range of the B block this rank stores (cols = my_block cols)."""
        return self.b_parts[self.pr]

    def _send_panel(self, t: int, a_local, b_local):
        """Issue the two broadcasts of panel t; returns (a_panel, b_panel, works)."""
        k0, k1, oa, ob = self.panels[t]
        w = k1 - k0
        works = []
        if self.Pc > 1:
            a_panel = self.a_buf[t & 1].flatten()[: a_local.shape[0] * w].view(a_local.shape[0], w)
            if self.pc == oa:
                ka = self.a_parts[oa][0]
                a_panel.copy_(a_local[:, k0 - ka:k1 - ka])
            if a_panel.numel() > 0:       # an empty block is empty on the whole grid row: nothing to send
                works.append(self.dist.broadcast(a_panel, src=self.row_ranks[oa], group=self.row_group, async_op=True))
        else:
            a_panel = a_local[:, k0:k1]
        if self.Pr > 1:
            b_panel = self.b_buf[t & 1][:w]
            if self.pr == ob:
                kb = self.b_parts[ob][0]
                b_panel.copy_(b_local[k0 - kb:k1 - kb])
            if b_panel.numel() > 0:
                works.append(self.dist.broadcast(b_panel, src=self.col_ranks[ob], group=self.col_group, async_op=True))
        else:
            b_panel = b_local[k0:k1]
        return a_panel, b_panel, works

    def step(self, c_local, a_local, b_local) -> None:
        """One pass C_block += sum over panels of A_panel * B_panel."""
        if not self.panels:
            return
        nxt = self._send_panel(0, a_local, b_local)
        for t in range(len(self.panels)):
            a_panel, b_panel, works = nxt
            for w in works:
                w.wait()      # CUDA: the compute stream waits for this panel only
            if t + 1 < len(self.panels):
                # Buffer (t+1)&1 was last read by the product of panel t-1, already enqueued on the
                # compute stream: NCCL's stream orders the new broadcast after it.
                nxt = self._send_panel(t + 1, a_local, b_local)
            if c_local.shape[0] > 0 and c_local.shape[1] > 0:
                self.local_mtm(c_local, a_panel, b_panel)
