"""Multi-GPU correctness + timing of the row-block sharded driver (run under torchrun on N GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py [--size 4096] [--big-size 32768] [--bcast-ctas 4,0] [--bcast nccl,nvlink] [--summa]

1. exactness: integer-valued operands; every rank's C block equals the fp64 product of its rows
   (and therefore the single-GPU result, which the single-GPU tests pin to the oracle);
2. config 5 timing: fp32 `big`^3 row-block sharded, B broadcast inside the step, max over ranks,
   for each NCCL CTA cap of the broadcast communicator (0 = share the default communicator).
"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import openmp_blas_b200 as ob  # noqa: E402
from openmp_blas_b200.sharded import RowBlockMtm, SummaMtm  # noqa: E402


def check_summa(variant, n, rank, grid, panel=None):
    """2-D SUMMA split: every rank's C block equals the fp64 product (integer-valued operands)."""
    M, N, K = n, n + 128, n - 96
    drv = SummaMtm(M, N, K, torch.float32, grid=grid, panel=panel, variant=variant)
    r0, r1, c0, c1 = drv.my_block
    (ka0, ka1), (kb0, kb1) = drv.my_a_cols, drv.my_b_rows
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.randint(0, 10, (M, K), device="cuda", generator=g).float()
    B = torch.randint(0, 10, (K, N), device="cuda", generator=g).float()
    C0 = torch.randint(0, 10, (M, N), device="cuda", generator=g).float()
    c = C0[r0:r1, c0:c1].clone()          # (.contiguous() would alias C0 when the block spans all columns)
    a = A[r0:r1, ka0:ka1].clone()
    b = B[kb0:kb1, c0:c1].clone()
    drv.step(c, a, b)
    drv.step(c, a, b)
    torch.cuda.synchronize()
    want = C0[r0:r1, c0:c1].double() + 2 * (A[r0:r1].double() @ B[:, c0:c1].double())
    flag = torch.tensor([int(torch.equal(c.double(), want))], device="cuda")
    diag = None
    if not bool(flag.item()):
        bad = (c.double() != want)
        idx = bad.nonzero()
        diag = {"rank": rank, "n_bad": int(bad.sum().item()), "first": idx[0].tolist(), "last": idx[-1].tolist(),
                "bad_rows": int(bad.any(dim=1).sum().item()), "bad_cols": int(bad.any(dim=0).sum().item()),
                "max_abs": float((c.double() - want).abs().max().item()), "block": [r0, r1, c0, c1]}
        # same panels through the CUDA-core kernels: separates the driver from the tensor-core path
        c2 = C0[r0:r1, c0:c1].clone()
        drv2 = SummaMtm(M, N, K, torch.float32, grid=grid, panel=panel, variant="simt")
    else:
        drv2 = SummaMtm(M, N, K, torch.float32, grid=grid, panel=panel, variant="simt")
        c2 = C0[r0:r1, c0:c1].clone()
    drv2.step(c2, a, b)
    drv2.step(c2, a, b)
    torch.cuda.synchronize()
    simt_ok = bool(torch.equal(c2.double(), want))
    if diag is not None:
        diag["simt_ok"] = simt_ok
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, diag)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    # timing of one step (panels double-buffered, broadcasts inside)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        drv.step(c, a, b)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 3], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return {"grid": [drv.Pr, drv.Pc], "panels": len(drv.panels), "exact": bool(flag.item()), "simt_exact": simt_ok,
            "diag": [g for g in gathered if g], "ms": ms.item(),
            "tflops": float(M) * N * (2.0 * K - 1) / ms.item() / 1e9}




def check_exact(variant, n, rank, bcast="nccl"):
    drv = RowBlockMtm(n, n, n, torch.float32, variant=variant, bcast=bcast, n_chunks=3 if bcast != "nccl" else None)
    r0, r1 = drv.my_rows
    g = torch.Generator(device="cuda").manual_seed(7)            # same stream of numbers on every rank
    A = torch.randint(0, 10, (n, n), device="cuda", generator=g).float()
    B = torch.randint(0, 10, (n, n), device="cuda", generator=g).float()
    C0 = torch.randint(0, 10, (n, n), device="cuda", generator=g).float()
    c = C0[r0:r1].clone()
    a = A[r0:r1].contiguous()
    drv.step(c, a, B if rank == 0 else None)
    drv.step(c, a, B if rank == 0 else None)
    torch.cuda.synchronize()
    want = C0[r0:r1].double() + 2 * (A[r0:r1].double() @ B.double())
    flag = torch.tensor([int(torch.equal(c.double(), want))], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


def time_config5(variant, N, bcast_ctas, rank, iters=3, config=None, bcast="nccl", push_ctas=0):
    drv = RowBlockMtm(N, N, N, torch.float32, variant=variant, bcast_ctas=bcast_ctas, config=config, bcast=bcast,
                      push_ctas=push_ctas)
    r0, r1 = drv.my_rows
    a = torch.rand((r1 - r0, N), device="cuda") * 2 - 1
    c = torch.zeros((r1 - r0, N), device="cuda")
    b = (torch.rand((N, N), device="cuda") * 2 - 1) if rank == 0 else None
    for _ in range(2):
        drv.step(c, a, b)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        drv.step(c, a, b)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # same shard product with B already resident everywhere: compute only
    b_all = b if rank == 0 else drv.b_buf
    fn = ob.mtm(c, a, b_all, None, variant=variant, config=config)
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms2 = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    fl = float(N) * N * (2.0 * N - 1)
    res = {"N": N, "chunks": drv.chunks, "reserve_sms": drv.reserve_sms, "kernel": ob.last_choice()["name"],
           "bcast": "nvlink" + ("_multicast" if drv.replicator.multicast else "_unicast") if drv.replicator is not None else "nccl",
           "ms_with_broadcast": ms.item(), "tflops_with_broadcast": fl / ms.item() / 1e9,
           "ms_compute_only": ms2.item(), "tflops_compute_only": fl / ms2.item() / 1e9}
    del a, b, c, drv
    torch.cuda.empty_cache()
    return res


def time_push(n, rank, ctas_list, iters=5):
    """Replication alone: the root pushes an n x n fp32 matrix (one chunk), receivers wait for it."""
    from openmp_blas_b200.sharded import NvlinkReplicator
    res = {}
    src = torch.rand((n, n), device="cuda") if rank == 0 else None
    for ctas in ctas_list:
        rep = NvlinkReplicator((n, n), torch.float32, 0, torch.device("cuda", torch.cuda.current_device()), ctas=ctas)
        for _ in range(2):
            rep.replicate(src)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            rep.replicate(src)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ok = torch.tensor([1], device="cuda")
        if rank != 0:
            pass
        chk = src if rank == 0 else rep.store[((rep.steps_done - 1) % rep.depth) * rep.slot_elems:][: n * n].view(n, n)
        ref = chk.clone()
        dist.broadcast(ref, src=0)
        ok[0] = int(torch.equal(ref, chk))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        res[f"ctas{ctas}"] = {"ms": ms.item(), "GBps": 4.0 * n * n / ms.item() / 1e6, "ok": bool(ok.item()),
                             "multicast_data": rep.multicast_data}
        del rep
        torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=4096)
    ap.add_argument("--big-size", dest="big", type=int, default=32768)
    ap.add_argument("--variants", default="simt,3xtf32")
    ap.add_argument("--configs", default="", help="comma list of tile configs to force (empty = library default)")
    ap.add_argument("--bcast-ctas", default="0",
                    help="comma list of NCCL CTA caps for the broadcast communicator (0 = default group)")
    ap.add_argument("--bcast", default="nccl", help="comma list of nccl / nvlink (this library's multicast push kernels)")
    ap.add_argument("--push-ctas", default="0", help="comma list of CTA counts of the push kernel (0 = default 32)")
    ap.add_argument("--push-bw", default="", help="comma list of push CTA counts to time alone (-1 = copy engines)")
    ap.add_argument("--summa-auto-only", action="store_true")
    ap.add_argument("--summa", action="store_true", help="also check the 2-D SUMMA split (grids 1xP, Px1 and the default)")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"world": world}
    for variant in args.variants.split(","):
        if ob.num_configs(variant, False) == 0:
            continue
        cfgs = [int(v) for v in args.configs.split(",")] if args.configs else [None]
        for mode in args.bcast.split(","):
            sfx = "" if mode == "nccl" else f"_{mode}"
            out[f"exact_{variant}{sfx}"] = check_exact(variant, args.n, rank, bcast=mode)
            torch.cuda.empty_cache()
            if args.big <= 0:
                continue
            for bc in ([int(v) for v in args.bcast_ctas.split(",")] if mode == "nccl" else [0]):
                for pc in ([int(v) for v in args.push_ctas.split(",")] if mode != "nccl" else [0]):
                    for cfg in cfgs:
                        key = (f"config5_{variant}{sfx}_bcastctas{bc}" + (f"_pushctas{pc}" if pc else "")
                               + ("" if cfg is None else f"_cfg{cfg}"))
                        out[key] = time_config5(variant, args.big, bc, rank, config=cfg, bcast=mode, push_ctas=pc)
                        if rank == 0:
                            print("#", key, json.dumps(out[key]), file=sys.stderr, flush=True)
        if args.push_bw and variant == args.variants.split(",")[0]:
            out["push_bandwidth_8192"] = time_push(8192, rank, [int(v) for v in args.push_bw.split(",")])
            if rank == 0:
                print("# push_bandwidth", json.dumps(out["push_bandwidth_8192"]), file=sys.stderr, flush=True)
        if args.summa:
            grids = {None} if args.summa_auto_only else {(1, world), (world, 1), None}
            for grid in sorted(grids, key=str):
                out[f"summa_{variant}_{'auto' if grid is None else 'x'.join(map(str, grid))}"] = check_summa(
                    variant, args.n, rank, grid, panel=None if grid is None else 1024)
            torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
