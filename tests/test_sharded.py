"""Multi-process (world_size 2, gloo, CPU) tests of the row-block sharded driver's host logic:
partitioning, K-chunking and the chunked broadcast of B.  The per-rank compute is injected
(the CPU oracle, test-only) — the product default is the CUDA kernel and has no CPU path."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_row_partition_covers_and_aligns():
    from openmp_blas_b200.sharded import row_partition
    for M in (1, 127, 128, 1000, 8192, 32768, 65536 + 5):
        for world in (1, 2, 4, 8):
            parts = row_partition(M, world)
            assert len(parts) == world
            assert parts[0][0] == 0 and parts[-1][1] == M
            for (b0, e0), (b1, e1) in zip(parts, parts[1:]):
                assert e0 == b1 and b0 <= e0
            assert all(b % 128 == 0 for b, _ in parts if b < M)
    assert row_partition(32768, 8) == [(i * 4096, (i + 1) * 4096) for i in range(8)]


def test_k_chunks_cover():
    from openmp_blas_b200.sharded import k_chunks
    for K in (1, 31, 32, 1000, 8192, 32768):
        for n in (1, 3, 8):
            ch = k_chunks(K, n)
            assert ch[0][0] == 0 and ch[-1][1] == K
            for (a0, a1), (b0, b1) in zip(ch, ch[1:]):
                assert a1 == b0 and a0 < a1
            assert all(a0 % 32 == 0 for a0, _ in ch)


def _worker(rank, world, port, M, N, K, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, str(ROOT))
        import oracle
        from openmp_blas_b200.sharded import RowBlockMtm
        orc = oracle.Oracle()

        def cpu_checker_mtm(c, a, b):     # test-only stand-in for the CUDA kernel
            cn = c.numpy()
            orc.mtm(cn, a.numpy(), b.numpy())

        rng = np.random.default_rng(1234)                       # same data on every rank
        A = rng.integers(0, 100, (M, K)).astype(np.float32)
        B = rng.integers(0, 100, (K, N)).astype(np.float32)
        C0 = rng.integers(0, 100, (M, N)).astype(np.float32)
        drv = RowBlockMtm(M, N, K, torch.float32, n_chunks=4, local_mtm=cpu_checker_mtm,
                          device=torch.device("cpu"))
        r0, r1 = drv.my_rows
        c_local = torch.from_numpy(C0[r0:r1].copy())
        a_local = torch.from_numpy(A[r0:r1].copy())
        b_root = torch.from_numpy(B.copy()) if rank == 0 else None
        drv.step(c_local, a_local, b_root)
        cal = drv.calibrate(a_local, b_root, steps=1)           # times each available path on a scratch C; gloo: nccl-style only
        drv.step(c_local, a_local, b_root)                      # accumulates, B re-broadcast
        want = C0.astype(np.int64) + 2 * (A.astype(np.int64) @ B.astype(np.int64))
        ok = np.array_equal(c_local.numpy().astype(np.int64), want[r0:r1])
        ok = ok and cal["chosen"] == "nccl" and set(cal["paths"]) == {"nccl"} and not drv.use_nvlink
        # column-major operands: shard the columns of C and B, broadcast A (step_first_order)
        drv2 = RowBlockMtm(N, M, K, torch.float32, n_chunks=3, local_mtm=cpu_checker_mtm,
                           device=torch.device("cpu"))
        c0, c1 = drv2.my_rows                                     # column range of this rank
        cf = torch.from_numpy(np.ascontiguousarray(C0[:, c0:c1].T)).t()      # (M x cols) column-major
        bf = torch.from_numpy(np.ascontiguousarray(B[:, c0:c1].T)).t()       # (K x cols) column-major
        af = torch.from_numpy(np.ascontiguousarray(A.T)).t() if rank == 0 else None   # (M x K) column-major
        drv2.step_first_order(cf, af, bf)
        want1 = C0.astype(np.int64) + A.astype(np.int64) @ B.astype(np.int64)
        ok = ok and np.array_equal(cf.numpy().astype(np.int64), want1[:, c0:c1])
        q.put((rank, bool(ok), (r0, r1), len(drv.chunks)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(300, 70, 200), (129, 33, 64)])
def test_row_block_mtm_gloo_world2(shape):
    import torch.multiprocessing as mp
    M, N, K = shape
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, M, N, K, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    results = sorted(q.get(timeout=5) for _ in range(2))
    assert all(p.exitcode == 0 for p in procs)
    assert [r[1] for r in results] == [True, True], results
    assert results[0][2][0] == 0 and results[1][2][1] == M


# ---- 2-D SUMMA split ----------------------------------------------------------------------------------
def test_summa_panels_cover_and_have_single_owners():
    from openmp_blas_b200.sharded import choose_grid, split_even, summa_panels
    assert choose_grid(8, 32768, 32768) in ((2, 4), (4, 2))
    assert choose_grid(4, 8192, 8192) == (2, 2)
    assert choose_grid(8, 65536, 1024) == (8, 1)
    assert choose_grid(1, 5, 5) == (1, 1)
    for K in (1, 31, 64, 1000, 8192):
        for Pr, Pc in ((1, 1), (1, 2), (2, 1), (2, 2), (2, 4), (4, 2), (3, 2)):
            for panel in (None, 96, 4096):
                ps = summa_panels(K, Pr, Pc, panel)
                assert ps[0][0] == 0 and ps[-1][1] == K
                a_parts, b_parts = split_even(K, Pc, 32), split_even(K, Pr, 32)
                for (k0, k1, oa, ob), nxt in zip(ps, ps[1:] + [None]):
                    assert k0 < k1
                    assert a_parts[oa][0] <= k0 and k1 <= a_parts[oa][1]
                    assert b_parts[ob][0] <= k0 and k1 <= b_parts[ob][1]
                    if nxt is not None:
                        assert nxt[0] == k1
                    if panel:
                        assert k1 - k0 <= -(-panel // 32) * 32


def _summa_worker(rank, world, port, grid, M, N, K, panel, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, str(ROOT))
        import oracle
        from openmp_blas_b200.sharded import SummaMtm
        orc = oracle.Oracle()
        calls = []

        def cpu_checker_mtm(c, a, b):     # test-only stand-in for the CUDA kernel
            calls.append((tuple(a.shape), tuple(b.shape)))
            an = np.ascontiguousarray(a.numpy())
            orc.mtm(c.numpy(), an, b.numpy())

        rng = np.random.default_rng(4321)                       # same data on every rank
        A = rng.integers(0, 100, (M, K)).astype(np.float32)
        B = rng.integers(0, 100, (K, N)).astype(np.float32)
        C0 = rng.integers(0, 100, (M, N)).astype(np.float32)
        drv = SummaMtm(M, N, K, torch.float32, grid=grid, panel=panel, local_mtm=cpu_checker_mtm,
                       device=torch.device("cpu"))
        r0, r1, c0, c1 = drv.my_block
        ka0, ka1 = drv.my_a_cols
        kb0, kb1 = drv.my_b_rows
        c_local = torch.from_numpy(C0[r0:r1, c0:c1].copy())
        a_local = torch.from_numpy(A[r0:r1, ka0:ka1].copy())
        b_local = torch.from_numpy(B[kb0:kb1, c0:c1].copy())
        drv.step(c_local, a_local, b_local)
        drv.step(c_local, a_local, b_local)                     # accumulates
        want = C0.astype(np.int64) + 2 * (A.astype(np.int64) @ B.astype(np.int64))
        ok = np.array_equal(c_local.numpy().astype(np.int64), want[r0:r1, c0:c1])
        q.put((rank, bool(ok), (r0, r1, c0, c1), len(drv.panels), len(calls)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,grid,shape,panel", [
    (2, (1, 2), (150, 300, 200), None),
    (2, (2, 1), (300, 70, 129), 64),
    (4, (2, 2), (300, 260, 200), None),
    (4, (2, 2), (129, 33, 70), 32),
    (4, (4, 1), (600, 40, 100), None),
    (4, (1, 4), (40, 600, 150), None),
])
def test_summa_mtm_gloo(world, grid, shape, panel):
    import torch.multiprocessing as mp
    M, N, K = shape
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_summa_worker, args=(r, world, port, grid, M, N, K, panel, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    results = sorted(q.get(timeout=5) for _ in range(world))
    assert all(p.exitcode == 0 for p in procs)
    assert [r[1] for r in results] == [True] * world, results
    # the blocks tile C exactly
    area = sum((r[2][1] - r[2][0]) * (r[2][3] - r[2][2]) for r in results)
    assert area == M * N


# ---- chunk planning and the replicator's host-side arithmetic (no GPU) -------------------------------
def test_plan_chunks_is_identical_on_every_rank_and_covers_k():
    """Every rank must walk the same chunk schedule (the receivers count sequence numbers): the plan may
    depend on the world size and the shape, never on the rank's own row count."""
    from openmp_blas_b200.sharded import plan_chunks, row_partition
    for K in (1024, 8192, 32768):
        for t_b, t_c in ((0.5e-3, 4.4e-3), (11e-3, 40e-3), (1e-5, 1e-3), (5e-3, 1e-3)):
            ch = plan_chunks(K, t_b, t_c, 1e-4)
            assert ch[0][0] == 0 and ch[-1][1] == K
            assert all(a1 == b0 and a0 < a1 for (a0, a1), (b0, _) in zip(ch, ch[1:]))
            assert all(a0 % 32 == 0 for a0, _ in ch)
    # a cheap transfer is not worth a second chunk; an expensive one is pipelined
    assert len(plan_chunks(8192, 1e-6, 4e-3, 1e-4)) == 1
    assert len(plan_chunks(8192, 2e-3, 4e-3, 1e-4)) >= 2
    # ragged partition: the last rank has fewer rows, the plan is made from rank 0's
    rows = row_partition(8192 * 8 - 1000, 8)
    assert rows[0][1] - rows[0][0] >= rows[-1][1] - rows[-1][0]


def test_chunk_tile_config():
    """K-chunk calls of the sharded drivers: single 256 x 256 pair tiles below K = 6144, double tiles from there on; the dynamic
    scheduler variants where a collective shares the SMs; no double tiles with a library that lacks them."""
    from openmp_blas_b200.sharded import chunk_tile_config
    assert chunk_tile_config(1024, dynamic=True) == 2 and chunk_tile_config(1024, dynamic=False) == 0
    assert chunk_tile_config(6144, dynamic=True) == 10 and chunk_tile_config(32768, dynamic=False) == 9
    assert chunk_tile_config(32768, dynamic=True, has_double=False) == 2
    assert chunk_tile_config(32768, dynamic=False, has_double=False) == 0
