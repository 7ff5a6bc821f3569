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
    import faulthandler
    import time as _time
    _t0 = _time.time()
    # a rank that is still here after this long dumps its Python stack and leaves: a hang names itself
    import threading
    _deadline = float(os.environ.get("BBM_DIST_TEST_DEADLINE_S", "200"))
    faulthandler.dump_traceback_later(_deadline, exit=False)
    _killer = threading.Timer(_deadline + 5.0, lambda: os._exit(3))     # every rank has had time to print its stack
    _killer.daemon = True
    _killer.start()

    def trace(*what):
        if os.environ.get("BBM_DIST_TRACE"):
            print(f"[rank {rank} +{_time.time() - _t0:6.1f}s]", *what, file=sys.stderr, flush=True)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    h = B.Handle(local_rank)
    comm = D.Communicator(h, dist, world, rank)
    failures = []
    for ci, (rows, cols, lu, lm, want_fused) in enumerate(CASES):
        for use_peer in (False, True):
            trace("bbb case", ci, "peer", use_peer)
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
            # the host-vector call (pipelined upload / exchange / product / download) on the same operands
            if not use_peer:
                trace("bbb case", ci, "host vectors")
                for beta in (0.0, 0.5):
                    x = rng.standard_normal(sum(cols))
                    xo = np.ascontiguousarray(x[c0 + o:c0 + o + lay["n_own"]])
                    yo = y0[lay["row_offset"]:lay["row_offset"] + lay["m_own"]].copy()
                    dist.barrier()
                    A.mul_host_(yo, xo, 1.5, beta)
                    ref = y0.copy()
                    O.bbb_mul(rows, cols, *lu, *lm, 1.5, data, x, beta, ref)
                    want = ref[lay["row_offset"]:lay["row_offset"] + lay["m_own"]]
                    if lay["m_own"] and (np.isnan(yo).any() or relerr(yo, want) > tol(np.float64, F.bbb_nnz_per_row_max(*lu, *lm) + 1)):
                        failures.append((ci, "host-vector product", beta, relerr(yo, want)))
            active, timed_out = A.peer_status()
            if timed_out:
                failures.append((ci, use_peer, "peer wait timed out"))
            if fused != active:
                failures.append((ci, use_peer, "fused flag inconsistent"))
            A.close()
    # ---- row-block-partitioned BlockBandedMatrix through the C ABI: matvec (bbm_dist_sky_mul_*) and A*B
    # (bbm_dist_sky_matmul_*: halo rows of B packed, exchanged with ncclSend/Recv and scattered inside the library) ----
    from tests.helpers import random_sky
    for ci, (rows, l, u) in enumerate([([16] * 12, 2, 2), ([5, 9, 4, 12, 7, 3, 8, 6], 1, 2), ([64] * 9, 0, 3), ([130] * 11, 2, 1)]):
        rng = np.random.default_rng(900 + ci)
        trace("sky case", ci)
        adata = random_sky(rng, rows, rows, l, u)
        bdata = random_sky(rng, rows, rows, u, l)
        Ah = B.BlockSkylineMatrix(adata, rows, rows, l, u)
        Bh = B.BlockSkylineMatrix(bdata, rows, rows, u, l)
        kb = D.partition_rows(rows, world)
        for nrhs in (1, 3):
            X = rng.standard_normal((sum(rows), nrhs))
            mv = D.DistBlockBandedMatrix(comm, rows, rows, l, u, Kbounds=kb)
            lay = mv.layout
            loc, Jlo, Jhi = D.sky_cut_rows(Ah, lay["K0"], lay["K1"])
            if (Jlo, Jhi) != (mv.Jlo, mv.Jhi) or loc.data.shape != (mv.numentries,) or list(loc.l) != list(mv.local.l) or list(loc.u) != list(mv.local.u):
                failures.append(("sky matvec structure", ci))
                continue
            mv.set_local_data(loc.data)
            nl, mo = max(1, lay["n_local"]), max(1, lay["m_own"])
            xl = np.full((nl, nrhs), np.nan, order="F")
            o, c0 = lay["own_offset"], lay["col_offset"]
            xl[o:o + lay["n_own"], :] = X[c0 + o:c0 + o + lay["n_own"], :]
            dxs = h.to_device(xl)
            y0 = np.random.default_rng(5000 + rank).standard_normal((mo, nrhs))   # (rank-local sizes: must not advance the shared generator)
            dys = h.to_device(np.asfortranarray(y0))
            for epoch in range(2):
                dist.barrier()
                mv.mul_(dys, dxs, 1.5, 0.5 if epoch == 0 else 0.0, nrhs=nrhs)
                got = dys.to_host()[:lay["m_own"], :]
                ref = np.zeros((sum(rows), nrhs), order="F")
                O.sky_mul(rows, rows, l, u, 1.5, adata, np.asfortranarray(X), 0.0, ref)
                want = ref[lay["row_offset"]:lay["row_offset"] + lay["m_own"], :] + (0.5 * y0[:lay["m_own"], :] if epoch == 0 else 0.0)
                if lay["m_own"] and (np.isnan(got).any() or relerr(got, want) > tol(np.float64, 64)):
                    failures.append(("sky matvec", ci, nrhs, epoch, relerr(got, want)))
            mv.close()
        trace("sky product", ci)
        prod = D.DistBlockBandedMatmul(comm, rows, rows, rows, (l, u), (u, l), Kbounds=kb)
        P = prod.ranges
        locA, _, _ = D.sky_cut_rows(Ah, P["K0"], P["K1"])
        locB, JloB, _ = D.sky_cut_rows(Bh, P["S0"], P["S1"])
        if locA.data.shape != prod.A.data.shape or locB.data.shape != prod.B.data.shape:
            failures.append(("sky product structure", ci))
            continue
        bloc = locB.data.copy()
        for N in range(P["S0"], P["S1"]):          # only the exchange may fill the halo rows of B
            if not (P["O0"] <= N < P["O1"]):
                for J in [J for J in range(len(locB.cols)) if locB.has_block(N - P["S0"], J)]:
                    s, ld = locB.block_start(N - P["S0"], J), int(locB.block_strides[J])
                    for j in range(int(locB.cols[J])):
                        bloc[s + j * ld:s + j * ld + int(locB.rows[N - P["S0"]])] = np.nan
        if locA.data.size:
            prod.A.data.copy_from_host(locA.data)
        if bloc.size:
            prod.B.data.copy_from_host(bloc)
        prod.C.data.fill_bytes(0xFF)
        dist.barrier()
        for rep in range(2):                          # the second product re-sends the same halos
            prod.mul_(1.0, 0.0)
        h.synchronize()
        if bloc.size and not np.array_equal(prod.B.data.to_host(), locB.data):
            failures.append(("sky product halo rows of B", ci))
        cref = np.zeros(B.BlockSkylineMatrix.numentries(rows, rows, l + u, l + u))
        O.sky_matmul((rows, rows, l, u), (rows, rows, u, l), (rows, rows, l + u, l + u), 1.0, adata, bdata, 0.0, cref)
        wantC, _, _ = D.sky_cut_rows(B.BlockSkylineMatrix(cref, rows, rows, l + u, l + u), P["K0"], P["K1"])
        gotC = prod.C.data.to_host()
        if gotC.shape != wantC.data.shape:
            failures.append(("sky product C structure", ci))
        elif wantC.data.size and (np.isnan(gotC).any() or relerr(gotC, wantC.data) > tol(np.float64, 256)):
            failures.append(("sky product", ci, relerr(gotC, wantC.data)))
        prod.close()

    # ---- a rank that is late by more than the configured wait: the fused path must not return wrong numbers silently ----
    rows = [64] * (4 * world)
    rng = np.random.default_rng(77)
    data = random_bbb(rng, rows, rows, 1, 1, 1, 1)
    for timeout_ms, delay_s, expect_ok in ((30000, 1.5, True), (200, 1.0, False)):
        trace("late rank", timeout_ms, delay_s)
        A = D.DistBandedBlockBandedMatrix(comm, rows, rows, (1, 1), (1, 1))
        lay = A.layout
        A.set_local_data(A.local_data(data))
        h.set_option("dist.peer_timeout_ms", timeout_ms)
        if not A.enable_peer():
            A.close()
            continue
        x = rng.standard_normal(sum(rows))
        o, c0 = lay["own_offset"], lay["col_offset"]
        xl = np.full(lay["n_local"], np.nan)
        xl[o:o + lay["n_own"]] = x[c0 + o:c0 + o + lay["n_own"]]
        dx, dy = h.to_device(xl), h.empty(lay["m_own"], np.float64)
        dist.barrier()
        if rank == world - 1:
            import time
            time.sleep(delay_s)                    # GC pause, I/O, a first-call plan build ...
        A.mul_(dy, dx)
        got = dy.to_host()
        ref = np.zeros(sum(rows))
        O.bbb_mul(rows, rows, 1, 1, 1, 1, 1.0, data, x, 0.0, ref)
        want = ref[lay["row_offset"]:lay["row_offset"] + lay["m_own"]]
        _active, timed_out = A.peer_status()
        if expect_ok:
            if timed_out or np.isnan(got).any() or relerr(got, want) > tol(np.float64, 9):
                failures.append(("late rank within the time-out", timed_out, relerr(got, want)))
        else:
            # ranks whose neighbour was late see a time-out: their halo rows are NaN (never stale numbers) and the next
            # product is refused; everything that is not NaN is still right
            ok = ~np.isnan(got)
            if ok.any() and relerr(got[ok], want[ok]) > tol(np.float64, 9):
                failures.append(("late rank: wrong finite values", relerr(got[ok], want[ok])))
            if rank == world - 2 and not (timed_out and np.isnan(got).any()):
                failures.append(("late rank: the waiting neighbour saw no time-out", timed_out, int(np.isnan(got).sum())))
            if timed_out:
                try:
                    A.mul_(dy, dx)
                    failures.append(("late rank: product after a time-out was not refused",))
                except B.BBMError:
                    pass
        dist.barrier()
        h.synchronize()
        A.close()
    h.set_option("dist.peer_timeout_ms", 30000)

    trace("done, failures:", failures)
    flag = torch.tensor([len(failures)], device=f"cuda:{local_rank}")
    dist.all_reduce(flag)
    if failures:
        print(f"rank {rank} FAILURES: {failures}", flush=True)
    if rank == 0:
        print("DIST_GPU_OK" if int(flag.item()) == 0 else "DIST_GPU_FAILED", flush=True)
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    faulthandler.cancel_dump_traceback_later()
    _killer.cancel()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
