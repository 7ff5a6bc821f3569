"""Multi-rank path on CPU: the C library's host-only partition / layout planning drives a
world_size-2 (and 3) gloo run in which every rank exchanges its x halo exactly as the plan says and
multiplies its sub-views with the CPU oracle; the assembled result must equal the global oracle.

This covers everything of the N>1 path except the CUDA kernel itself (parity-tested on one GPU) and
NCCL: partition, ownership, halo send/recv lists, interior/boundary split, sub-view bandwidth shift.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    # rows, cols, l, u, lam, mu
    ([8] * 12, [8] * 12, 1, 1, 1, 1),
    (list(range(1, 15)), list(range(1, 15)), 2, 2, 2, 2),
    ([5, 3, 7, 2, 6, 4, 5], [4, 6, 2, 5, 3, 7, 2, 4], 1, 2, 2, 1),   # rectangular, Nc > Nr
    ([4] * 9, [4] * 6, 2, 1, 1, 1),                                   # Nr > Nc
    ([6] * 10, [6] * 10, -1, 2, 1, 0),                                # band off the diagonal
    ([6] * 10, [6] * 10, 3, 0, 0, 2),
    ([3] * 4, [3] * 4, 2, 2, 1, 1),                                   # halo reaches past the neighbour (world 3)
    ([2, 0, 3, 0, 0, 4, 1], [2, 0, 3, 0, 0, 4, 1], 1, 1, 1, 1),       # empty blocks
]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bbm_b200 import distributed as D
        from oracle import cpu as O
        from oracle import formats as F
        from tests.helpers import random_bbb, relerr, tol

        failures = []
        for ci, (rows, cols, l, u, lam, mu) in enumerate(CASES):
            rng = np.random.default_rng(100 + ci)  # same stream on every rank
            data = random_bbb(rng, rows, cols, l, u, lam, mu)
            x = rng.standard_normal(sum(cols))
            y0 = rng.standard_normal(sum(rows))
            lay, sends, recvs = D.layout(rows, cols, l, u, world, rank)
            C = F.prefix(cols)
            R = F.prefix(rows)
            # --- local buffers: column slab of data, x_local with only the owned part filled
            c0 = lay["col_offset"]
            slab = np.asfortranarray(data[:, c0:c0 + lay["n_local"]])
            xl = np.full(lay["n_local"], np.nan)
            o = lay["own_offset"]
            xl[o:o + lay["n_own"]] = x[c0 + o:c0 + o + lay["n_own"]]
            # --- halo exchange exactly as planned
            ops, bufs = [], []
            for peer, off, cnt in recvs:
                t = torch.empty(cnt, dtype=torch.float64)
                bufs.append((off, cnt, t))
                ops.append(dist.P2POp(dist.irecv, t, peer))
            for peer, off, cnt in sends:
                assert o <= off and off + cnt <= o + lay["n_own"], "a rank may only send what it owns"
                ops.append(dist.P2POp(dist.isend, torch.from_numpy(xl[off:off + cnt].copy()), peer))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            for off, cnt, t in bufs:
                assert not (o <= off < o + lay["n_own"]), "halo must not overwrite owned entries"
                xl[off:off + cnt] = t.numpy()
            # --- multiply the sub-views (interior, left boundary, right boundary) with the oracle
            y_own = y0[lay["row_offset"]:lay["row_offset"] + lay["m_own"]].copy()
            covered = np.zeros(lay["m_own"], dtype=int)
            for lo, hi, Jl, Jh, ls, us in D.piece_structures(rows, cols, l, u, lay):
                r_sub, c_sub = rows[lo:hi], cols[Jl:Jh]
                d_sub = np.asfortranarray(slab[:, C[Jl] - c0:C[Jh] - c0])
                x_sub = xl[C[Jl] - c0:C[Jh] - c0]
                y_sub = y_own[R[lo] - R[lay["K0"]]:R[hi] - R[lay["K0"]]]
                covered[R[lo] - R[lay["K0"]]:R[hi] - R[lay["K0"]]] += 1
                if F.bbb_data_shape(c_sub, ls, us, lam, mu)[0] != slab.shape[0] and slab.shape[0] > 0:
                    failures.append((ci, "sub-view P mismatch"))
                    continue
                tmp = np.ascontiguousarray(y_sub)
                O.bbb_mul(r_sub, c_sub, ls, us, lam, mu, 2.0, d_sub, np.ascontiguousarray(x_sub), -0.5, tmp)
                y_sub[:] = tmp
            if lay["m_own"] and not (covered == 1).all():
                failures.append((ci, "pieces do not tile the owned rows"))
            # --- gather and compare on rank 0
            parts = [None] * world
            dist.all_gather_object(parts, (lay["row_offset"], y_own))
            if rank == 0:
                y = np.full(sum(rows), np.nan)
                for off, part in parts:
                    y[off:off + len(part)] = part
                ref = y0.copy()
                O.bbb_mul(rows, cols, l, u, lam, mu, 2.0, data, x, -0.5, ref)
                nnz = F.bbb_nnz_per_row_max(l, u, lam, mu) + 1
                if np.isnan(y).any() or relerr(y, ref) > tol(np.float64, nnz):
                    failures.append((ci, f"result mismatch, relerr {relerr(y, ref)}"))
        q.put((rank, failures))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, [("exception", traceback.format_exc() + repr(e))]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_matvec_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, failures in results:
        assert not failures, f"rank {rank}: {failures}"


def test_partition_rows_balanced():
    sys.path.insert(0, ROOT)
    from bbm_b200 import distributed as D
    kb = D.partition_rows([4096] * 4096, 8)
    assert list(kb) == [512 * i for i in range(9)]
    kb = D.partition_rows(list(range(1, 4001)), 8)   # cfg3: rows ~ K, split near N*sqrt(p/G)
    assert kb[0] == 0 and kb[-1] == 4000 and all(np.diff(kb) > 0)
    for p in range(1, 8):
        assert abs(kb[p] - 4000 * np.sqrt(p / 8)) < 3
    kb = D.partition_rows([], 3)
    assert list(kb) == [0, 0, 0, 0]
    kb = D.partition_rows([5], 4)   # fewer blocks than ranks: some ranks own nothing
    assert kb[0] == 0 and kb[-1] == 1 and all(np.diff(kb) >= 0)


def test_layout_cfg2_halo_sizes():
    """cfg2 on 8 GPUs: each interior rank receives one 4096-entry x block from each neighbour."""
    sys.path.insert(0, ROOT)
    from bbm_b200 import distributed as D
    rows = [4096] * 4096
    for rank in range(8):
        lay, sends, recvs = D.layout(rows, rows, 1, 1, 8, rank)
        assert lay["K1"] - lay["K0"] == 512 and lay["m_own"] == 512 * 4096
        nb = (rank > 0) + (rank < 7)
        assert len(sends) == nb and len(recvs) == nb
        assert all(cnt == 4096 for _, _, cnt in sends + recvs)
        assert lay["n_local"] == (512 + nb) * 4096
        assert lay["Ka"] == lay["K0"] + (rank > 0) and lay["Kb"] == lay["K1"] - (rank < 7)


# --------------------------------------------------------------------------------------------------
# row-block-partitioned BlockBandedMatrix product C = A*B (cfg4's sharding) on gloo
# --------------------------------------------------------------------------------------------------
MM_CASES = [
    # rows(=cols of A = rows of B), colsB, (lA,uA), (lB,uB)
    ([4] * 10, [4] * 10, (2, 2), (2, 2)),
    ([3, 5, 2, 4, 6, 1, 3, 4], [3, 5, 2, 4, 6, 1, 3, 4], (1, 2), (2, 1)),
    ([4] * 9, [4] * 9, (0, 3), (3, 0)),
    ([2] * 5, [2] * 5, (2, 2), (1, 1)),      # halos reach past the neighbour with 3 ranks
]


def _mm_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bbm_b200 as B
        from bbm_b200 import distributed as D
        from oracle import cpu as O
        from oracle import formats as F
        from tests.helpers import random_sky, relerr, tol
        failures = []
        for ci, (rows, colsB, (lA, uA), (lB, uB)) in enumerate(MM_CASES):
            rng = np.random.default_rng(300 + ci)
            adata = random_sky(rng, rows, rows, lA, uA)
            bdata = random_sky(rng, rows, colsB, lB, uB)
            A = B.BlockSkylineMatrix(adata, rows, rows, lA, uA)
            Bm = B.BlockSkylineMatrix(bdata, rows, colsB, lB, uB)
            kb = D.partition_rows(rows, world)
            prod = D.DistBlockBandedProduct(None, dist, torch, rows, rows, lA, uA, colsB, lB, uB, kb, rank)
            P = prod.plan
            locA, _, _ = D.sky_cut_rows(A, P["K0"], P["K1"])
            locB, _, _ = D.sky_cut_rows(Bm, P["S0"], P["S1"])
            prod.A.data[:] = locA.data
            prod.B.data[:] = locB.data
            # poison the halo rows of B: only the exchange may fill them
            for N in range(P["S0"], P["S1"]):
                if not (P["O0"] <= N < P["O1"]):
                    for J in range(max(N - lB, 0), min(N + uB, len(colsB) - 1) + 1):
                        if rows[N] and colsB[J]:
                            prod._b_block(N, J).fill_(float("nan"))
            prod.exchange()
            if np.isnan(prod.B.data).any():
                failures.append((ci, "halo rows of B not filled by the exchange"))
                continue
            if not np.array_equal(prod.B.data, locB.data):
                failures.append((ci, "exchange delivered wrong blocks"))
            # local product on the local (ragged) structures with the oracle
            c0 = rng.standard_normal(prod.C.data.shape)
            prod.C.data[:] = c0
            O.sky_matmul((prod.A.rows, prod.A.cols, prod.A.l, prod.A.u), (prod.B.rows, prod.B.cols, prod.B.l, prod.B.u),
                         (prod.C.rows, prod.C.cols, prod.C.l, prod.C.u), 1.5, prod.A.data, prod.B.data, 0.0, prod.C.data)
            # compare the owned rows of C with the global oracle product
            ol, ou = lA + lB, uA + uB
            cref = np.zeros(B.BlockSkylineMatrix.numentries(rows, colsB, ol, ou))
            O.sky_matmul((rows, rows, lA, uA), (rows, colsB, lB, uB), (rows, colsB, ol, ou), 1.5, adata, bdata, 0.0, cref)
            Cref = B.BlockSkylineMatrix(cref, rows, colsB, ol, ou)
            want, _, _ = D.sky_cut_rows(Cref, P["K0"], P["K1"])
            if want.data.shape != prod.C.data.shape:
                failures.append((ci, "local C structure mismatch"))
            elif relerr(prod.C.data, want.data) > tol(np.float64, 64):
                failures.append((ci, f"C rows mismatch {relerr(prod.C.data, want.data)}"))
        q.put((rank, failures))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, [("exception", traceback.format_exc() + repr(e))]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_block_banded_product_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mm_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, failures in results:
        assert not failures, f"rank {rank}: {failures}"


def _mv_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bbm_b200 as B
        from bbm_b200 import distributed as D
        from oracle import cpu as O
        from tests.helpers import random_sky, relerr, tol
        failures = []
        for ci, (rows, cols, l, u) in enumerate([([4] * 10, [4] * 10, 2, 2), ([3, 5, 2, 4, 6, 1, 3, 4], [3, 5, 2, 4, 6, 1, 3, 4], 1, 2),
                                                 ([5, 3, 7, 2, 6], [4, 6, 2, 5, 3, 7], 1, 2), ([4] * 9, [4] * 9, 0, 3)]):
            rng = np.random.default_rng(700 + ci)
            adata = random_sky(rng, rows, cols, l, u)
            x = rng.standard_normal(sum(cols))
            A = B.BlockSkylineMatrix(adata, rows, cols, l, u)
            kb = D.partition_rows(rows, world)
            mv = D.DistBlockBandedMatvec(None, dist, torch, rows, cols, l, u, kb, rank)
            lay = mv.layout
            loc, _, _ = D.sky_cut_rows(A, lay["K0"], lay["K1"])
            mv.A.data[:] = loc.data
            xl = mv.x_local.numpy()
            xl[:] = np.nan
            o, c0 = lay["own_offset"], lay["col_offset"]
            xl[o:o + lay["n_own"]] = x[c0 + o:c0 + o + lay["n_own"]]
            mv.exchange()
            n_loc = mv.A.shape[1]
            xs = np.ascontiguousarray(xl[mv.x_off:mv.x_off + n_loc])
            if np.isnan(xs).any():
                failures.append((ci, "x halo not filled"))
                continue
            y = np.zeros(mv.A.shape[0])
            O.sky_mul(mv.A.rows, mv.A.cols, mv.A.l, mv.A.u, 1.0, mv.A.data, xs, 0.0, y)
            ref = np.zeros(sum(rows))
            O.sky_mul(rows, cols, l, u, 1.0, adata, x, 0.0, ref)
            want = ref[lay["row_offset"]:lay["row_offset"] + lay["m_own"]]
            if relerr(y, want) > tol(np.float64, 64):
                failures.append((ci, f"mismatch {relerr(y, want)}"))
        q.put((rank, failures))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, [("exception", traceback.format_exc() + repr(e))]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_block_banded_matvec_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mv_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, failures in results:
        assert not failures, f"rank {rank}: {failures}"


# --------------------------------------------------------------------------------------------------
# host-only planning of the C ABI's sharded BlockBandedMatrix paths against the Python planners above
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [(10, 10, 2, 2), (8, 8, 1, 2), (9, 9, 0, 3), (5, 6, 1, 2), (7, 5, 3, 0), (6, 6, -1, 2), (4, 4, 5, 5)])
def test_c_rowrange_structure_matches_python(case):
    sys.path.insert(0, ROOT)
    from bbm_b200 import distributed as D
    Nr, Nc, l, u = case
    rows, cols = [3] * Nr, [2] * Nc
    for S0 in range(Nr + 1):
        for S1 in range(S0, Nr + 1):
            Jlo, Jhi, ll, uu = D.sky_rowrange_structure(rows, cols, l, u, S0, S1)
            cJlo, cJhi, cl, cu = D.rowrange_structure_c(Nr, Nc, l, u, S0, S1)
            assert (Jlo, Jhi) == (cJlo, cJhi)
            assert list(ll) == list(cl) and list(uu) == list(cu)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_c_mm_layout_matches_python(world):
    sys.path.insert(0, ROOT)
    from bbm_b200 import distributed as D
    for rows, colsB, (lA, uA), (lB, uB) in MM_CASES + [([4, 0, 3, 5, 0, 2, 6, 1], [4, 0, 3, 5, 0, 2, 6, 1], (1, 1), (1, 2))]:
        kb = D.partition_rows(rows, world)
        for rank in range(world):
            P = D.matmul_plan(rows, rows, lA, uA, colsB, lB, uB, kb, rank)
            rng, sb, rb = D.mm_layout_c(rows, colsB, len(rows), lA, uA, lB, uB, kb, rank)
            assert rng == {k: P[k] for k in ("K0", "K1", "S0", "S1", "O0", "O1")}
            for q in range(world):
                assert sb[q] == len(P["sends"].get(q, [])) and rb[q] == len(P["recvs"].get(q, []))
        # what p sends to q is what q receives from p
        lay = [D.mm_layout_c(rows, colsB, len(rows), lA, uA, lB, uB, kb, r) for r in range(world)]
        for p in range(world):
            for q in range(world):
                assert lay[p][1][q] == lay[q][2][p]
