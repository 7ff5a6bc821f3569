"""Worker of tests/test_dist_gpu.py: one rank of a multi-GPU BandedBlockBandedMatrix matvec, checked
against the CPU oracle.  Launched with torch.distributed.run (one process per GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # rows, cols, (l,u), (lam,mu), fused path expected
    ([64] * 24, [64] * 24, (1, 1), (1, 1), True),
    (list(range(20, 60)), list(range(20, 60)), (2, 2), (2, 2), True),
    ([300] * 9, [300] * 9, (2, 1), (3, 2), True),
    ([40] * 16, [40] * 16, (0, 2), (1, 1), True),      # halo on one side only
    ([50] * 3, [50] * 3, (2, 2), (1, 1), None),        # halo may reach past the neighbour: NCCL path
]


def main():
    import torch
    import torch.distributed as dist
    import bbm_b200 as B
    from bbm_b200 import distributed as D
    from oracle import cpu as O
    from oracle import formats as F
    from tests.helpers import random_bbb, relerr, tol

    world, rank, local_rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    h = B.Handle(local_rank)
    comm = D.Communicator(h, dist, world, rank)
    failures = []
    for ci, (rows, cols, lu, lm, want_fused) in enumerate(CASES):
        for use_peer in (False, True):
            rng = np.random.default_rng(500 + ci)
            data = random_bbb(rng, rows, cols, *lu, *lm)
            A = D.DistBandedBlockBandedMatrix(comm, rows, cols, lu, lm)
            lay = A.layout
            A.set_local_data(A.local_data(data))
            o, c0 = lay["own_offset"], lay["col_offset"]
            dx = h.empty(max(1, lay["n_local"]), np.float64)
            dy = h.empty(max(1, lay["m_own"]), np.float64)
            fused = A.enable_peer() if use_peer else False
            if use_peer and want_fused is True and not fused and world <= 4:
                failures.append((ci, "fused path expected but registration declined"))
            y0 = rng.standard_normal(sum(rows))
            for epoch in range(4):   # several products on the same buffers: epochs / acknowledgements
                x = rng.standard_normal(sum(cols))
                xl = np.full(max(1, lay["n_local"]), np.nan)
                xl[o:o + lay["n_own"]] = x[c0 + o:c0 + o + lay["n_own"]]
                dx.copy_from_host(xl)
                yl = np.zeros(max(1, lay["m_own"]))
                yl[:lay["m_own"]] = y0[lay["row_offset"]:lay["row_offset"] + lay["m_own"]]
                dy.copy_from_host(yl)
                dist.barrier()       # every rank's x is in place (the producer of x is the host here)
                A.mul_(dy, dx, 1.5, 0.5)
                got = dy.to_host()[:lay["m_own"]]
                ref = y0.copy()
                O.bbb_mul(rows, cols, *lu, *lm, 1.5, data, x, 0.5, ref)
                want = ref[lay["row_offset"]:lay["row_offset"] + lay["m_own"]]
                nnz = F.bbb_nnz_per_row_max(*lu, *lm) + 1
                if lay["m_own"] and (np.isnan(got).any() or relerr(got, want) > tol(np.float64, nnz)):
                    failures.append((ci, use_peer, epoch, "mismatch", relerr(got, want)))
                dist.barrier()
            active, timed_out = A.peer_status()
            if timed_out:
                failures.append((ci, use_peer, "peer wait timed out"))
            if fused != active:
                failures.append((ci, use_peer, "fused flag inconsistent"))
            A.close()
    # ---- row-block-partitioned BlockBandedMatrix: matvec and A*B (halo rows over NCCL point-to-point) ----
    from tests.helpers import random_sky
    for ci, (rows, l, u) in enumerate([([16] * 12, 2, 2), ([5, 9, 4, 12, 7, 3, 8, 6], 1, 2), ([64] * 9, 0, 3)]):
        rng = np.random.default_rng(900 + ci)
        adata = random_sky(rng, rows, rows, l, u)
        bdata = random_sky(rng, rows, rows, u, l)
        Ah = B.BlockSkylineMatrix(adata, rows, rows, l, u)
        Bh = B.BlockSkylineMatrix(bdata, rows, rows, u, l)
        kb = D.partition_rows(rows, world)
        x = rng.standard_normal(sum(rows))
        mv = D.DistBlockBandedMatvec(h, dist, torch, rows, rows, l, u, kb, rank)
        lay = mv.layout
        loc, _, _ = D.sky_cut_rows(Ah, lay["K0"], lay["K1"])
        mv.tA[:loc.data.size] = torch.from_numpy(loc.data).to(mv.dev)
        xl = np.full(max(1, lay["n_local"]), np.nan)
        o, c0 = lay["own_offset"], lay["col_offset"]
        xl[o:o + lay["n_own"]] = x[c0 + o:c0 + o + lay["n_own"]]
        mv.x_local[:] = torch.from_numpy(xl).to(mv.dev)
        mv.exchange()
        y = mv.mul_().cpu().numpy()[:lay["m_own"]]
        ref = np.zeros(sum(rows))
        O.sky_mul(rows, rows, l, u, 1.0, adata, x, 0.0, ref)
        want = ref[lay["row_offset"]:lay["row_offset"] + lay["m_own"]]
        if lay["m_own"] and (np.isnan(y).any() or relerr(y, want) > tol(np.float64, 64)):
            failures.append(("sky matvec", ci, relerr(y, want)))
        prod = D.DistBlockBandedProduct(h, dist, torch, rows, rows, l, u, rows, u, l, kb, rank)
        P = prod.plan
        locA, _, _ = D.sky_cut_rows(Ah, P["K0"], P["K1"])
        locB, _, _ = D.sky_cut_rows(Bh, P["S0"], P["S1"])
        prod.tA[:locA.data.size] = torch.from_numpy(locA.data).to(prod.dev)
        prod.tB[:locB.data.size] = torch.from_numpy(locB.data).to(prod.dev)
        for N in range(P["S0"], P["S1"]):          # only the exchange may fill the halo rows of B
            if not (P["O0"] <= N < P["O1"]):
                for J in range(max(N - u, 0), min(N + l, len(rows) - 1) + 1):
                    prod._b_block(N, J).fill_(float("nan"))
        prod.exchange()
        prod.tC.fill_(float("nan"))
        prod.mul_(1.0, 0.0)
        h.synchronize()
        cref = np.zeros(B.BlockSkylineMatrix.numentries(rows, rows, l + u, l + u))
        O.sky_matmul((rows, rows, l, u), (rows, rows, u, l), (rows, rows, l + u, l + u), 1.0, adata, bdata, 0.0, cref)
        wantC, _, _ = D.sky_cut_rows(B.BlockSkylineMatrix(cref, rows, rows, l + u, l + u), P["K0"], P["K1"])
        gotC = prod.tC.cpu().numpy()[:wantC.data.size]
        if wantC.data.size and (np.isnan(gotC).any() or relerr(gotC, wantC.data) > tol(np.float64, 256)):
            failures.append(("sky product", ci, relerr(gotC, wantC.data)))

    flag = torch.tensor([len(failures)], device=f"cuda:{local_rank}")
    dist.all_reduce(flag)
    if failures:
        print(f"rank {rank} FAILURES: {failures}", flush=True)
    if rank == 0:
        print("DIST_GPU_OK" if int(flag.item()) == 0 else "DIST_GPU_FAILED", flush=True)
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
