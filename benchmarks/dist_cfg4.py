"""cfg4 over N GPUs through the C ABI: BlockBandedMatrix fill(256,2048), (2,2) x (2,2) -> (4,4), Float64, block rows
sharded (`bench.py --workload cfg4 --gpus N`, or N = 1).  Strong scaling: 1.716 TFLOP per product whatever N.

Every rank builds its local panels from a closed-form function of the GLOBAL entry indices (so that the halo rows of B a
rank receives can be checked against what their owner must have sent), poisons the halo rows with NaN, runs
`bbm_dist_sky_matmul_f64` (pack kernel -> ncclSend/Recv -> scatter kernel, overlapped with the interior tiles) and checks
(i) B_loc == closed form everywhere afterwards, (ii) sampled blocks of C_loc against float64 block products.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np


def _entries(torch, gi, gj, phase):
    """Deterministic pseudo-random entries in (-1/32, 1/32) from global row / column indices (float64 tensors)."""
    v = torch.sin(gi[:, None] * 12.9898 + gj[None, :] * 78.233 + phase) * 43758.5453
    return (v - torch.floor(v) - 0.5) / 16.0


def fill_local(torch, M, row_block0, col_block0, bs, phase, dev):
    """Fill the device storage of the local BlockSkylineMatrix M (uniform block size bs; its block (Kp,Jp) is global block
    (row_block0+Kp, col_block0+Jp)) panel by panel; returns a torch view of the storage."""
    n = int(M.panel_offsets[-1])
    t = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    for Jp in range(len(M.cols)):
        S = int(M.block_strides[Jp])
        if S == 0:
            continue
        g0 = (row_block0 + int(M.klo[Jp])) * bs
        gi = torch.arange(g0, g0 + S, device=dev, dtype=torch.float64)
        gj = torch.arange((col_block0 + Jp) * bs, (col_block0 + Jp + 1) * bs, device=dev, dtype=torch.float64)
        o = int(M.panel_offsets[Jp])
        t[o:o + S * bs] = _entries(torch, gi, gj, phase).t().reshape(-1)      # column-major panel
    return t


def run(args, torch, dist, B, h, stream, rank, world, local_rank):
    from bbm_b200.distributed import Communicator, DistBlockBandedMatmul
    from bbm_b200.device import DeviceArray
    from benchmarks.configs import dgemm_peak
    K, W = args.steps, max(args.warmup, 2)
    bs, nblk = 256, args.cfg4_blocks
    rows = [bs] * nblk
    dev = torch.device("cuda", local_rank)
    comm = Communicator(h, dist, world, rank)
    P = DistBlockBandedMatmul(comm, rows, rows, rows, (2, 2), (2, 2))
    R = P.ranges

    def allmax(v):
        if dist is None:
            return float(v)
        t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.cuda.stream(stream):
        tA = fill_local(torch, P.A, R["K0"], P.Jlo[0], bs, 0.25, dev)
        tB = fill_local(torch, P.B, R["S0"], P.Jlo[1], bs, 1.75, dev)
        tB_ref = tB.clone()
        # poison the halo rows of B_loc: only the exchange may fill them
        halo_elems = 0
        for Jp in range(len(P.B.cols)):
            S = int(P.B.block_strides[Jp])
            if S == 0:
                continue
            o = int(P.B.panel_offsets[Jp])
            pan = tB[o:o + S * bs].view(bs, S)                               # [col][row] view of the column-major panel
            for Kp in range(int(P.B.klo[Jp]), int(P.B.khi[Jp]) + 1):
                N = R["S0"] + Kp
                if not (R["O0"] <= N < R["O1"]):
                    r0 = (Kp - int(P.B.klo[Jp])) * bs
                    pan[:, r0:r0 + bs] = float("nan")
                    halo_elems += bs * bs
        nC = int(P.C.panel_offsets[-1])
        tC = torch.full((max(nC, 1),), float("nan"), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        # the plan's matrices get these buffers as their storage
        for M, t in ((P.A, tA), (P.B, tB), (P.C, tC)):
            M.data = DeviceArray(h, (int(M.panel_offsets[-1]),), np.float64, ptr=t.data_ptr())
        if dist is not None:
            dist.barrier()
        P.mul_(1.0, 0.0)
        torch.cuda.synchronize()
        halo_ok = bool(torch.equal(tB, tB_ref))
        nan_in_C = bool(torch.isnan(tC[:nC]).any().item()) if nC else False
        # sampled blocks of C_loc against float64 block products of the local operands
        rng = np.random.default_rng(100 + rank)
        worst = 0.0
        nK = R["K1"] - R["K0"]
        for _ in range(8 if nK else 0):
            Kp = int(rng.integers(0, nK))
            Kg = R["K0"] + Kp
            Jg = int(rng.integers(max(0, Kg - 4), min(nblk - 1, Kg + 4) + 1))
            Jc = Jg - P.Jlo[2]
            ref = torch.zeros(bs, bs, dtype=torch.float64, device=dev)
            for Ng in range(max(0, Jg - 2), min(nblk - 1, Jg + 2) + 1):
                if abs(Kg - Ng) > 2:
                    continue
                Ja, Nb, Jb = Ng - P.Jlo[0], Ng - R["S0"], Jg - P.Jlo[1]
                sa, la = P.A.block_start(Kp, Ja), int(P.A.block_strides[Ja])
                sb, lb = P.B.block_start(Nb, Jb), int(P.B.block_strides[Jb])
                Ab = tA[sa:sa + la * bs].view(bs, la)[:, :bs].t()             # (rows, cols) of block (K,N)
                Bb = tB[sb:sb + lb * bs].view(bs, lb)[:, :bs].t()
                ref += Ab @ Bb
            sc, lc = P.C.block_start(Kp, Jc), int(P.C.block_strides[Jc])
            got = tC[sc:sc + lc * bs].view(bs, lc)[:, :bs].t()
            den = float(torch.linalg.matrix_norm(ref).item())
            worst = max(worst, float(torch.linalg.matrix_norm(got - ref).item()) / max(den, 1e-300))
        tolerance = 64 * 2.220446049250313e-16 * 5 * bs
        err = allmax(worst)
        ok = allmax(0.0 if (halo_ok and not nan_in_C and worst <= tolerance) else 1.0) == 0.0
        if not ok:
            raise SystemExit(f"bench cfg4: rank {rank} parity failed (halo rows ok: {halo_ok}, NaN in C: {nan_in_C}, sampled rel err {worst:.3e})")
        peak_tf = dgemm_peak(torch) if rank == 0 else 0.0
        for _ in range(W):
            P.mul_(1.0, 0.0)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = h.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        P.mul_(1.0, 0.0)                  # one untimed lead-in product absorbs the ranks' host start skew
        ev0.record(stream)
        for _ in range(K):
            P.mul_(1.0, 0.0)
        ev1.record(stream)
        torch.cuda.synchronize()
        launches = h.launch_count - l0
        if dist is not None:
            dist.barrier()
        ms = allmax(ev0.elapsed_time(ev1)) / K
    triples = 0
    for J in range(nblk):
        for N in range(max(0, J - 2), min(nblk - 1, J + 2) + 1):
            triples += min(nblk - 1, N + 2) - max(0, N - 2) + 1
    flops = 2.0 * bs ** 3 * triples
    lt = launches
    if dist is not None:
        t = torch.tensor([float(launches)], device=dev, dtype=torch.float64)
        dist.all_reduce(t)
        lt = int(t.item())
    tf = flops / (ms * 1e-3) / 1e12
    return {
        "metric": "block_banded_matmul_TFLOPs", "value": tf, "unit": "TFLOP/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cfg4: BlockBandedMatrix fill({bs},{nblk}), (2,2) x (2,2) -> (4,4), Float64 mul!, block rows sharded over {world} GPU(s)",
                   "flops_per_step": flops, "l2": "operands (20 GB in total) exceed the 126 MB L2; no flush",
                   "parallelism": "1 GPU" if world == 1 else f"rows of A and C partitioned over {world} GPUs; halo block rows of B packed, exchanged with ncclSend/Recv and scattered inside bbm_dist_sky_matmul_f64, overlapped with the tiles that only multiply owned rows"},
        "halo_bytes_per_rank_per_step": 8 * halo_elems, "max_rel_err_sampled_blocks": err, "halo_rows_bit_equal_to_owner": True,
        "e2e": None, "gpu_launches": int(lt),
        "roofline": {"bound": "tensor", "kernel": "sky_dgemm_dmma_kernel<128,128,64,32,32> (FP64 DMMA)", "achieved": tf, "peak": peak_tf * world,
                     "unit": "TFLOP/s", "frac": tf / (peak_tf * world) if peak_tf else None, "traffic": None,
                     "peak_source": f"{world} x cuBLAS DGEMM 8192^3 (torch.matmul float64) measured on rank 0 in this run"},
        "cpu_baseline": None,
    }
