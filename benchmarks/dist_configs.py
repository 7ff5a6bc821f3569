#!/usr/bin/env python
"""cfg4 on N GPUs: BlockBandedMatrix x BlockBandedMatrix with the row blocks of A and C sharded over
the ranks and only the halo block rows of B exchanged (NCCL point-to-point); `--mode matvec` runs the sharded
block-skyline matvec y = A x on the same operand instead (x halo over NCCL, SURVEY.md 8e first row).  Launch:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P benchmarks/dist_configs.py [--mode matmul|matvec] [--nblk 2048] [--bs 256] [--reps 5]

Matrix entries are a deterministic function of the global (row, col) index, so every rank builds its
own panels without a global parent and any block can be regenerated on the host for the parity check
(sampled blocks of C against float64 numpy products).  Rank 0 prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def entry(i, j, salt):
    """Deterministic pseudo-random entry in [-1/32, 1/32) from global indices (numpy or torch int64)."""
    h = (i * 1315423911 + j * 2654435761 + salt * 97531) % 1000003
    return (h.astype(np.float64) if isinstance(h, np.ndarray) else h.double()) / 1000003.0 / 16.0 - 1.0 / 32.0


def fill_local(torch, M, t, row0_blk, col0_blk, bs, salt, dev, only_rows=None):
    """Fill the local skyline storage `t` of M (uniform block size bs): local block (Kp,Jp) is global
    block (row0_blk+Kp, col0_blk+Jp).  only_rows=(lo,hi): global block rows to fill, others get NaN."""
    ar = torch.arange(bs, device=dev, dtype=torch.int64)
    for Jp in range(len(M.cols)):
        S = int(M.block_strides[Jp])
        if S == 0:
            continue
        K0p = int(M.klo[Jp])
        gi = (row0_blk + K0p) * bs + torch.arange(S, device=dev, dtype=torch.int64)
        gj = (col0_blk + Jp) * bs + ar
        vals = entry(gi[None, :], gj[:, None], salt)            # [j][i] == column-major panel
        if only_rows is not None:
            blk = gi // bs
            bad = (blk < only_rows[0]) | (blk >= only_rows[1])
            vals = torch.where(bad[None, :], torch.full_like(vals, float("nan")), vals)
        o = int(M.panel_offsets[Jp])
        t[o:o + S * bs] = vals.reshape(-1)


def run_matvec(args, torch, dist, B, D, h, dev, rows, kb, world, rank):
    """y[own rows] = A[own rows, :] x with A = cfg4's (2,2) operand; every rank owns a block-row range of A, x, y."""
    nblk, bs = args.nblk, args.bs
    mv = D.DistBlockBandedMatvec(h, dist, torch, rows, rows, 2, 2, kb, rank)
    lay = mv.layout
    fill_local(torch, mv.A, mv.tA, lay["K0"], mv.Jlo, bs, 1, dev)
    gcol = lay["col_offset"] + torch.arange(lay["n_local"], device=dev, dtype=torch.int64)
    xg = entry(gcol, gcol * 0 + 7, 3)                                   # x[j] from its global index
    mv.x_local.fill_(float("nan"))
    o = lay["own_offset"]
    mv.x_local[o:o + lay["n_own"]] = xg[o:o + lay["n_own"]]            # only the exchange may fill the halo
    torch.cuda.synchronize()

    def step():
        mv.exchange()
        mv.mul_(1.0, 0.0)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    e0.record()
    for _ in range(args.reps):
        mv.mul_(1.0, 0.0)
    e1.record()
    torch.cuda.synchronize()
    kms = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev, dtype=torch.float64)
    dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    # parity: sampled owned rows against float64 numpy from the generating function
    ok = not bool(torch.isnan(mv.x_local[:lay["n_local"]]).any().item()) and not bool(torch.isnan(mv.y_own[:lay["m_own"]]).any().item())
    rng = np.random.default_rng(21 + rank)
    worst = 0.0
    y = mv.y_own.cpu().numpy()
    for _ in range(16):
        i = int(rng.integers(lay["row_offset"], lay["row_offset"] + lay["m_own"]))
        K = i // bs
        j = np.arange(max(0, K - 2) * bs, min(nblk, K + 3) * bs, dtype=np.int64)
        ref = float(np.dot(entry(np.full_like(j, i), j, 1), entry(j, j * 0 + 7, 3)))
        worst = max(worst, abs(y[i - lay["row_offset"]] - ref) / max(abs(ref), 1e-300))
    tol = 64 * 2.220446049250313e-16 * 5 * bs * 8      # per-entry check: allow for cancellation in a 1280-term sum
    flag = torch.tensor([1.0 if (ok and worst <= tol) else 0.0, worst], device=dev, dtype=torch.float64)
    dist.all_reduce(flag[:1], op=dist.ReduceOp.MIN)
    dist.all_reduce(flag[1:], op=dist.ReduceOp.MAX)
    nA = B.BlockSkylineMatrix.numentries(rows, rows, 2, 2)
    nbytes = 8 * (nA + 2 * nblk * bs)
    if rank == 0:
        print(json.dumps({
            "config": f"distributed skyline matvec: BlockBandedMatrix {nblk} blocks of {bs}, (2,2), Float64, row blocks over {world} GPUs",
            "n_gpus": world, "ms_per_step": float(ms.item()), "ms_kernel_only": float(kms.item()), "algorithmic_bytes": nbytes,
            "GBps_aggregate": nbytes / (float(ms.item()) * 1e-3) / 1e9,
            "halo": "x block columns inside the l+u halo over NCCL isend/irecv (torch.distributed batch_isend_irecv)",
            "parity_ok": bool(flag[0].item() == 1.0), "max_rel_err_sampled_rows": float(flag[1].item())}))
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nblk", type=int, default=2048)
    ap.add_argument("--bs", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--mode", default="matmul", choices=["matmul", "matvec"])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bbm_b200 as B
    from bbm_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream()
    h = B.Handle(local_rank, stream=stream.cuda_stream)

    nblk, bs = args.nblk, args.bs
    rows = [bs] * nblk
    kb = D.partition_rows(rows, world)
    if args.mode == "matvec":
        return run_matvec(args, torch, dist, B, D, h, dev, rows, kb, world, rank)
    prod = D.DistBlockBandedProduct(h, dist, torch, rows, rows, 2, 2, rows, 2, 2, kb, rank)
    P = prod.plan
    fill_local(torch, prod.A, prod.tA, P["K0"], prod.JloA, bs, 1, dev)
    fill_local(torch, prod.B, prod.tB, P["S0"], prod.JloB, bs, 2, dev, only_rows=(P["O0"], P["O1"]))
    prod.tC.fill_(float("nan"))
    torch.cuda.synchronize()

    def step():
        prod.exchange()
        prod.mul_(1.0, 0.0)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # exchange alone
    e0.record()
    for _ in range(args.reps):
        prod.exchange()
    e1.record()
    torch.cuda.synchronize()
    xms = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev, dtype=torch.float64)
    dist.all_reduce(xms, op=dist.ReduceOp.MAX)

    # parity: halo filled, no NaN in C, sampled blocks of the owned rows of C vs numpy
    ok = not bool(torch.isnan(prod.tB).any().item()) and not bool(torch.isnan(prod.tC[:max(1, prod.C.data.shape[0])]).any().item())
    rng = np.random.default_rng(11 + rank)
    worst = 0.0
    ar = np.arange(bs, dtype=np.int64)
    for _ in range(4):
        K = int(rng.integers(P["K0"], P["K1"]))
        J = int(rng.integers(max(0, K - 4), min(nblk - 1, K + 4) + 1))
        ref = np.zeros((bs, bs))
        for N in range(max(0, J - 2), min(nblk - 1, J + 2) + 1):
            if abs(K - N) <= 2:
                Ab = entry((K * bs + ar)[:, None], (N * bs + ar)[None, :], 1)
                Bb = entry((N * bs + ar)[:, None], (J * bs + ar)[None, :], 2)
                ref += Ab @ Bb
        Kp, Jp = K - P["K0"], J - prod.JloC
        s, ld = prod.C.block_start(Kp, Jp), int(prod.C.block_strides[Jp])
        got = prod.tC[s:s + ld * bs].cpu().numpy().reshape((ld, bs), order="F")[:bs]
        worst = max(worst, float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300)))
    tol = 64 * 2.220446049250313e-16 * 5 * bs
    flag = torch.tensor([1.0 if (ok and worst <= tol) else 0.0, worst], device=dev, dtype=torch.float64)
    dist.all_reduce(flag[:1], op=dist.ReduceOp.MIN)
    dist.all_reduce(flag[1:], op=dist.ReduceOp.MAX)

    triples = 0
    for J in range(nblk):
        for N in range(max(0, J - 2), min(nblk - 1, J + 2) + 1):
            triples += min(nblk - 1, N + 2) - max(0, N - 2) + 1
    flops = 2.0 * bs ** 3 * triples
    if rank == 0:
        print(json.dumps({
            "config": f"cfg4 distributed: BlockBandedMatrix {nblk} blocks of {bs}, (2,2)x(2,2)->(4,4), Float64, row blocks over {world} GPUs",
            "n_gpus": world, "ms_per_step": float(ms.item()), "exchange_ms": float(xms.item()),
            "TFLOPs_aggregate": flops / (float(ms.item()) * 1e-3) / 1e12, "flops": flops,
            "halo": "block rows of B inside the l_A+u_A halo, packed blocks over NCCL isend/irecv",
            "parity_ok": bool(flag[0].item() == 1.0), "max_rel_err_sampled_blocks": float(flag[1].item()), "tolerance": tol}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
