"""Parity of the CUDA BandedBlockBandedMatrix product (through the C ABI) with the CPU oracle.

Tolerance (BASELINE.json north_star): normwise relative error <= 64*eps(T)*(nonzeros per row)."""
import numpy as np
import pytest

import bbm_b200 as B
from oracle import cpu as O
from oracle import formats as F
from tests.helpers import BBB_CASES, random_bbb, relerr, tol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    h = B.Handle(0)
    yield h
    h.close()


def run_case(handle, rows, cols, l, u, lam, mu, dtype, nrhs, alpha, beta, seed=0, poison=True, data=None):
    rng = np.random.default_rng(seed)
    if data is None:
        data = random_bbb(rng, rows, cols, l, u, lam, mu, dtype, poison=poison)
    A = B.BandedBlockBandedMatrix(data, rows, cols, (l, u), (lam, mu))
    Ad = A.to_device(handle)
    m, n = A.shape
    xs = (n,) if nrhs is None else (n, nrhs)
    ys = (m,) if nrhs is None else (m, nrhs)
    X = np.asfortranarray(rng.standard_normal(xs).astype(dtype))
    Y0 = np.asfortranarray(rng.standard_normal(ys).astype(dtype))
    Yin = Y0.copy(order="F")
    if beta == 0:
        Yin[...] = np.nan  # beta == 0 overwrites and never reads (test/test_blockskyline.jl:34-39)
    dX, dY = handle.to_device(X), handle.to_device(Yin)
    out = B.mul_(dY, Ad, dX, alpha, beta)
    assert out is dY
    got = dY.to_host()
    ref = Y0.copy(order="F")
    O.bbb_mul(rows, cols, l, u, lam, mu, alpha, data, X, beta, ref)
    return got, ref


@pytest.mark.parametrize("kernel", [1, 2], ids=["walk", "generic"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", BBB_CASES, ids=lambda c: f"r{len(c[0])}c{len(c[1])}_l{c[2]}u{c[3]}_lam{c[4]}mu{c[5]}")
def test_bbb_mul_matches_oracle(handle, case, dtype, kernel):
    rows, cols, l, u, lam, mu = case
    handle.set_option("bbb.kernel", kernel)
    try:
        for nrhs, alpha, beta in [(None, 1.0, 0.0), (None, 2.5, -0.5), (3, 1.0, 1.0)]:
            got, ref = run_case(handle, rows, cols, l, u, lam, mu, dtype, nrhs, alpha, beta)
            assert got.shape == ref.shape
            assert not np.isnan(got).any()
            assert relerr(got, ref) <= tol(dtype, F.bbb_nnz_per_row_max(l, u, lam, mu) + 1)
    finally:
        handle.set_option("bbb.kernel", 0)


@pytest.mark.parametrize("tile", [32, 64, 128, 256])
@pytest.mark.parametrize("stages", [2, 3, 5])
def test_bbb_walk_tiles_and_stages(handle, tile, stages):
    handle.set_option("bbb.kernel", 1)
    handle.set_option("bbb.tile", tile)
    handle.set_option("bbb.stages", stages)
    handle.set_option("bbb.items_per_sm", 1)
    try:
        for rows, cols, lu, lm in [([300, 17, 513, 64, 100, 257], [290, 40, 500, 70, 90, 300], (1, 2), (3, 2)),
                                   ([257] * 9, [257] * 9, (1, 1), (1, 1)), (list(range(1, 70)), list(range(1, 70)), (2, 2), (2, 2))]:
            got, ref = run_case(handle, rows, cols, *lu, *lm, np.float64, None, 1.0, 0.0, seed=tile + stages)
            assert relerr(got, ref) <= tol(np.float64, 40)
    finally:
        for k in ("bbb.kernel", "bbb.tile", "bbb.stages", "bbb.items_per_sm"):
            handle.set_option(k, 0)


@pytest.mark.parametrize("maxseg", [1, 2, 7])
def test_bbb_walk_short_segments(handle, maxseg):
    """Segment boundaries: every block row is produced by exactly one work item."""
    handle.set_option("bbb.kernel", 1)
    handle.set_option("bbb.maxseg", maxseg)
    try:
        got, ref = run_case(handle, [40] * 23, [40] * 23, 2, 1, 1, 2, np.float64, 2, 1.5, 0.25, seed=maxseg)
        assert relerr(got, ref) <= tol(np.float64, 20)
    finally:
        handle.set_option("bbb.kernel", 0)
        handle.set_option("bbb.maxseg", 0)


def test_bbb_laplacian_cfg1(handle):
    """BASELINE.json configs[0]: 100x100 blocks of size 100, (1,1),(1,1), Float64."""
    n = 100
    data = F.laplacian_bbb_data(n)
    got, ref = run_case(handle, [n] * n, [n] * n, 1, 1, 1, 1, np.float64, None, 1.0, 0.0, data=data)
    assert relerr(got, ref) <= tol(np.float64, 9)
    # independent check: it is the 5-point stencil
    rng = np.random.default_rng(0)
    X = rng.standard_normal(n * n)
    assert relerr(ref, F.laplacian_apply(X, n)) <= tol(np.float64, 9)


def test_bbb_host_vectors_roundtrip(handle):
    rows = cols = [64] * 12
    rng = np.random.default_rng(5)
    data = random_bbb(rng, rows, cols, 1, 1, 1, 1)
    Ad = B.BandedBlockBandedMatrix(data, rows, cols, (1, 1), (1, 1)).to_device(handle)
    x = rng.standard_normal(sum(cols))
    y = Ad @ x  # host vectors: H2D, kernel, D2H inside the C-ABI call
    ref = np.zeros_like(y)
    O.bbb_mul(rows, cols, 1, 1, 1, 1, 1.0, data, x, 0.0, ref)
    assert relerr(y, ref) <= tol(np.float64, 9)
    y2 = y.copy()
    B.mul_(y2, Ad, x, 2.0, 1.0)
    assert relerr(y2, 3 * ref) <= tol(np.float64, 10)


def test_bbb_dimension_mismatch(handle):
    Ad = B.BandedBlockBandedMatrix.zeros(np.float64, [4, 4], [4, 4], (1, 1), (1, 1)).to_device(handle)
    with pytest.raises(B.DimensionMismatch):
        B.mul_(handle.empty(8, np.float64), Ad, handle.empty(7, np.float64))
    # and the library itself rejects a short leading dimension
    lib = handle._lib
    import ctypes as C
    rc = lib.bbm_bbb_mul_f64(handle.h, Ad.descriptor(handle), 1.0, C.c_void_p(Ad.data.ptr), C.c_void_p(Ad.data.ptr), 3, 1,
                             0.0, C.c_void_p(Ad.data.ptr), 8)
    assert rc == 2 and b"DimensionMismatch" in lib.bbm_last_error(handle.h)


def test_bbb_linearity_large(handle):
    """Size-independent property at a size the oracle is not run on: A(ax+by) = aAx + bAy."""
    n = 1024
    rows = cols = [n] * 64
    rng = np.random.default_rng(9)
    data = np.asfortranarray(rng.standard_normal((9, n * 64)))
    Ad = B.BandedBlockBandedMatrix(data, rows, cols, (1, 1), (1, 1)).to_device(handle)
    x1, x2 = rng.standard_normal(n * 64), rng.standard_normal(n * 64)
    y1, y2, y12 = Ad @ x1, Ad @ x2, Ad @ (2.0 * x1 - 3.0 * x2)
    assert relerr(y12, 2.0 * y1 - 3.0 * y2) <= tol(np.float64, 9) * 4


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_dist_plan_virtual_ranks(handle, world):
    """The multi-GPU plan on ONE device: every virtual rank multiplies its interior / boundary
    sub-views (communicator without NCCL, halos filled by hand); assembled result == oracle."""
    from bbm_b200 import distributed as D
    for ci, (rows, cols, l, u, lam, mu) in enumerate([([64] * 24, [64] * 24, 1, 1, 1, 1),
                                                     (list(range(1, 40)), list(range(1, 40)), 2, 2, 2, 2),
                                                     ([5, 3, 7, 2, 6, 4, 5], [4, 6, 2, 5, 3, 7, 2, 4], 1, 2, 2, 1),
                                                     ([300] * 9, [300] * 9, 2, 1, 3, 2)]):
        rng = np.random.default_rng(40 + ci)
        data = random_bbb(rng, rows, cols, l, u, lam, mu)
        x = rng.standard_normal(sum(cols))
        y0 = rng.standard_normal(sum(rows))
        ref = y0.copy()
        O.bbb_mul(rows, cols, l, u, lam, mu, 1.5, data, x, 0.5, ref)
        y = np.full(sum(rows), np.nan)
        for rank in range(world):
            comm = D.Communicator(handle, None, world, rank)
            A = D.DistBandedBlockBandedMatrix(comm, rows, cols, (l, u), (lam, mu))
            lay = A.layout
            A.set_local_data(A.local_data(data))
            c0 = lay["col_offset"]
            dx = handle.to_device(np.ascontiguousarray(x[c0:c0 + lay["n_local"]]))
            dy = handle.to_device(np.ascontiguousarray(y0[lay["row_offset"]:lay["row_offset"] + lay["m_own"]]))
            A.mul_(dy, dx, 1.5, 0.5)
            y[lay["row_offset"]:lay["row_offset"] + lay["m_own"]] = dy.to_host()
            A.close()
            comm.close()
        assert relerr(y, ref) <= tol(np.float64, F.bbb_nnz_per_row_max(l, u, lam, mu) + 1)


@pytest.mark.parametrize("kernel", [1, 2], ids=["walk", "generic"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("uplo,unit", [("U", False), ("U", True), ("L", False), ("L", True)])
def test_bbb_triangular_mul(handle, uplo, unit, dtype, kernel):
    """test/test_triblockbanded.jl:17-53: U*b == Matrix(U)*b for Upper/UnitUpper/Lower/UnitLower
    triangular BandedBlockBandedMatrix; the other triangle of `data` is poisoned with NaN."""
    handle.set_option("bbb.kernel", kernel)
    try:
        for rows, lu, lm in [(list(range(1, 11)), (1, 1), (1, 1)), ([40] * 9, (2, 1), (2, 3)), ([300, 17, 64], (1, 1), (1, 1))]:
            rng = np.random.default_rng(21)
            data = random_bbb(rng, rows, rows, *lu, *lm, dtype)
            A = B.BandedBlockBandedMatrix(data, rows, rows, lu, lm)
            D = A.to_dense().astype(np.float64)
            Tm = np.triu(D) if uplo == "U" else np.tril(D)
            if unit:
                np.fill_diagonal(Tm, 1.0)
            # poison the triangle that must not be read (and the diagonal for the unit variants)
            pois = data.copy(order="F")
            mask_other = (np.tril(np.ones_like(D), -1) if uplo == "U" else np.triu(np.ones_like(D), 1)).astype(bool)
            if unit:
                mask_other |= np.eye(D.shape[0], dtype=bool)
            for rho, gi, gj in A._slot_map():
                sel = mask_other[gi, gj]
                pois[rho, gj[sel]] = np.nan
            Ad = B.BandedBlockBandedMatrix(pois, rows, rows, lu, lm).to_device(handle)
            T = {"U": (B.UnitUpperTriangular if unit else B.UpperTriangular),
                 "L": (B.UnitLowerTriangular if unit else B.LowerTriangular)}[uplo](Ad)
            x = rng.standard_normal(sum(rows)).astype(dtype)
            y0 = rng.standard_normal(sum(rows)).astype(dtype)
            got = (T @ handle.to_device(x)).to_host()
            ref = Tm.astype(np.longdouble) @ x.astype(np.longdouble)
            assert not np.isnan(got).any()
            assert relerr(got, ref) <= tol(dtype, F.bbb_nnz_per_row_max(*lu, *lm) + 1)
            dy = handle.to_device(y0)
            B.mul_(dy, T, handle.to_device(x), 2.0, -1.0)
            assert relerr(dy.to_host(), 2.0 * ref - y0.astype(np.longdouble)) <= 2 * tol(dtype, F.bbb_nnz_per_row_max(*lu, *lm) + 1)
        assert T.blockbandwidths == ((0, lu[1]) if uplo == "U" else (lu[0], 0))   # src/triblockbanded.jl:10-25
    finally:
        handle.set_option("bbb.kernel", 0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_bbb_block_range_view_mul(handle, dtype):
    """test/test_triblockbanded.jl:55-78: c .= MulAdd(V, b) for block-range sub-views V of a device matrix."""
    rows = list(range(1, 11))
    rng = np.random.default_rng(51)
    data = random_bbb(rng, rows, rows, 1, 1, 1, 1, dtype)
    A = B.BandedBlockBandedMatrix(data, rows, rows, (1, 1), (1, 1))
    Dn = A.to_dense().astype(np.float64)
    Ad = A.to_device(handle)
    R = np.concatenate([[0], np.cumsum(rows)])
    for br, bc in [(2, range(2, 5)), (range(1, 4), 2), (range(3, 9), range(2, 10)), (range(0, 2), range(6, 9)), (range(0, 10), range(0, 10))]:
        V = Ad.view(br, bc)
        br = range(br, br + 1) if isinstance(br, int) else br
        bc = range(bc, bc + 1) if isinstance(bc, int) else bc
        sub = Dn[R[br.start]:R[br.stop], R[bc.start]:R[bc.stop]]
        assert V.shape == sub.shape
        b = rng.standard_normal(sub.shape[1]).astype(dtype)
        c = handle.empty(sub.shape[0], dtype).fill_bytes(0xFF)
        B.mul_(c, V, handle.to_device(b))
        assert relerr(c.to_host(), sub @ b.astype(np.float64)) <= tol(dtype, 10)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("base", [0, 1])
def test_bbb_sparse_export(handle, dtype, base):
    """sparse(A) (ext/BlockBandedMatricesSparseArraysExt.jl:10-27; test/test_bandedblockbanded.jl "sparse"):
    the device CSC arrays equal the restated per-block triplet construction and reproduce Matrix(A)."""
    import scipy.sparse as sp
    for rows, cols, lu, lm in [(list(range(1, 11)), list(range(1, 11)), (1, 1), (1, 1)), ([5, 0, 3, 7], [4, 2, 0, 6, 3], (1, 2), (2, 1)),
                               ([300, 17, 64], [300, 17, 64], (1, 1), (0, 2)), ([4] * 5, [4] * 5, (-1, 2), (1, 1)), ([3, 3], [2, 2], (0, 0), (-1, 0)),
                               ([], [], (1, 1), (1, 1)), ([0, 0], [0], (1, 1), (1, 1))]:
        rng = np.random.default_rng(81)
        data = random_bbb(rng, rows, cols, *lu, *lm, dtype)
        A = B.BandedBlockBandedMatrix(data, rows, cols, lu, lm)
        colptr, rowval, nzval = B.sparse(A.to_device(handle), index_base=base)
        cp, rv, nz = colptr.to_host(), rowval.to_host(), nzval.to_host()
        rcp, rrv, rnz = F.bbb_sparse(data, rows, cols, *lu, *lm)
        assert np.array_equal(cp - base, rcp) and np.array_equal(rv - base, rrv)
        assert np.array_equal(nz, rnz) and not np.isnan(nz).any()
        S = sp.csc_matrix((nz, rv - base, cp - base), shape=A.shape)
        assert np.array_equal(S.toarray(), A.to_dense())


def _well_conditioned_tri(rng, rows, lu, lm, dtype):
    """random band storage whose triangles are comfortably invertible: small off-diagonal, diagonal ~ 2"""
    data = random_bbb(rng, rows, rows, *lu, *lm, dtype)
    lam, mu = lm
    w = lam + mu + 1
    data *= dtype(0.5 / (F.bbb_nnz_per_row_max(*lu, *lm) + 1))
    data[lu[1] * w + mu, :] = (2.0 + rng.random(data.shape[1])).astype(dtype)
    return data


TRI_STRUCTURES = [
    (list(range(1, 11)), (1, 1), (1, 1)),        # test/test_triblockbanded.jl:80-99
    ([40] * 9, (2, 1), (2, 3)),
    ([300, 17, 64], (1, 1), (1, 1)),
    ([5, 0, 3, 7, 1, 1, 12], (3, 2), (0, 2)),    # empty block, lam = 0
    ([6] * 700, (1, 2), (1, 1)),                 # more block rows than one CTA of the cluster owns
    ([33] * 4, (0, 0), (3, 3)),                  # block diagonal
    ([9, 4, 11] * 10, (2, 2), (2, 2)),           # cfg3's bandwidths
    ([5, 3] * 200, (2, 2), (2, 2)),              # ... across several lane groups / CTAs of the cluster
]


@pytest.mark.parametrize("kernel,pf,lean,push", [(0, 4, 1, 1), (0, 4, 1, 0), (0, 4, 0, 0), (0, 2, 0, 0), (1, 6, 0, 0), (1, 2, 0, 0), (2, 0, 0, 0)],
                         ids=["lean-push", "lean-dsmem", "wave4", "wave2", "staged6", "staged2", "direct"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("uplo,unit", [("U", False), ("U", True), ("L", False), ("L", True)])
def test_bbb_triangular_ldiv(handle, uplo, unit, dtype, kernel, pf, lean, push):
    """test/test_triblockbanded.jl:80-99: U \\ b for Upper/UnitUpper/Lower/UnitLower triangular
    BandedBlockBandedMatrix against the block-substitution restatement and a dense solve; the triangle
    that must not be read (and the diagonal of the unit variants) is NaN."""
    handle.set_option("tri.pf", pf)
    handle.set_option("tri.kernel", kernel)
    handle.set_option("tri.lean", lean)
    handle.set_option("tri.push", push)
    try:
        _check_triangular_ldiv(handle, uplo, unit, dtype)
    finally:
        handle.set_option("tri.pf", 4)
        handle.set_option("tri.kernel", 0)
        handle.set_option("tri.lean", 1)
        handle.set_option("tri.push", 1)


def _check_triangular_ldiv(handle, uplo, unit, dtype):
    for rows, lu, lm in TRI_STRUCTURES:
        rng = np.random.default_rng(41)
        data = _well_conditioned_tri(rng, rows, lu, lm, dtype)
        A = B.BandedBlockBandedMatrix(data, rows, rows, lu, lm)
        m = sum(rows)
        pois = data.copy(order="F")
        for rho, gi, gj in A._slot_map():
            sel = (gi > gj) if uplo == "U" else (gi < gj)
            if unit:
                sel = sel | (gi == gj)
            pois[rho, gj[sel]] = np.nan
        Ad = B.BandedBlockBandedMatrix(pois, rows, rows, lu, lm).to_device(handle)
        T = {"U": (B.UnitUpperTriangular if unit else B.UpperTriangular),
             "L": (B.UnitLowerTriangular if unit else B.LowerTriangular)}[uplo](Ad)
        b = rng.standard_normal(m).astype(dtype)
        db = handle.to_device(b)
        B.ldiv_(T, db)
        got = db.to_host()
        ref = b.copy()
        O.bbb_trsm(rows, *lu, *lm, uplo, unit, data, ref)
        assert not np.isnan(got).any()
        nnz = F.bbb_nnz_per_row_max(*lu, *lm) + 1
        assert relerr(got, ref) <= tol(dtype, nnz)
        if m <= 1200:
            Dn = A.to_dense().astype(np.float64)
            Tm = np.triu(Dn) if uplo == "U" else np.tril(Dn)
            if unit:
                np.fill_diagonal(Tm, 1.0)
            assert relerr(got, np.linalg.solve(Tm, b.astype(np.float64))) <= tol(dtype, nnz)
        # several right-hand sides with a padded leading dimension; T * (T \ B) == B on the GPU
        Bm = np.asfortranarray(rng.standard_normal((m, 3)).astype(dtype))
        dB = handle.to_device(Bm)
        X = B.ldiv(T, dB)
        Xh = X.to_host()
        for q in range(3):
            refq = Bm[:, q].copy()
            O.bbb_trsm(rows, *lu, *lm, uplo, unit, data, refq)
            assert relerr(Xh[:, q], refq) <= tol(dtype, nnz)
            back = (T @ handle.to_device(np.ascontiguousarray(Xh[:, q]))).to_host()
            assert relerr(back, Bm[:, q]) <= 4 * tol(dtype, nnz)
        assert relerr(dB.to_host(), Bm) == 0.0      # ldiv (not ldiv!) leaves B alone


@pytest.mark.parametrize("kernel", [0, 1, 2], ids=["wave", "staged", "direct"])
@pytest.mark.parametrize("uplo", ["U", "L"])
def test_bbb_triangular_ldiv_more_block_rows_than_lanes(handle, uplo, kernel):
    """more than 4096 block rows: every thread of the cluster owns two block rows (round-robin passes)"""
    rows, lu, lm = [3, 2] * 2300, (1, 1), (1, 1)
    rng = np.random.default_rng(45)
    data = _well_conditioned_tri(rng, rows, lu, lm, np.float64)
    Ad = B.BandedBlockBandedMatrix(data, rows, rows, lu, lm).to_device(handle)
    T = (B.UpperTriangular if uplo == "U" else B.LowerTriangular)(Ad)
    b = rng.standard_normal(sum(rows))
    handle.set_option("tri.kernel", kernel)
    try:
        got = B.ldiv(T, handle.to_device(b)).to_host()
    finally:
        handle.set_option("tri.kernel", 0)
    ref = b.copy()
    O.bbb_trsm(rows, *lu, *lm, uplo, False, np.nan_to_num(data), ref)
    assert relerr(got, ref) <= tol(np.float64, 10)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("uplo,unit", [("U", False), ("L", False), ("L", True)])
def test_bbb_triangular_prepared_plan_and_many_rhs(handle, uplo, unit, dtype):
    """bbm_bbb_trsv_prepare_* / _solve_*: the matrix planes are gathered once and re-used by every later solve (the
    Gauss-Seidel loop); 20 right-hand sides = two batches of concurrently running clusters.  Bit-equal to the one-call
    solve, within tolerance of the oracle's substitution."""
    for rows, lu, lm in [([40] * 9, (2, 1), (2, 3)), ([300, 17, 64], (1, 1), (1, 1)), ([5, 3] * 200, (2, 2), (2, 2)), ([64] * 40, (1, 1), (1, 1))]:
        rng = np.random.default_rng(47)
        data = _well_conditioned_tri(rng, rows, lu, lm, dtype)
        m = sum(rows)
        Ad = B.BandedBlockBandedMatrix(np.nan_to_num(data), rows, rows, lu, lm).to_device(handle)
        T = {"U": B.UpperTriangular, "L": (B.UnitLowerTriangular if unit else B.LowerTriangular)}[uplo](Ad)
        plan = B.prepare_ldiv(T)
        nnz = F.bbb_nnz_per_row_max(*lu, *lm) + 1
        for nrhs in (1, 20):
            Bm = np.asfortranarray(rng.standard_normal((m, nrhs)).astype(dtype))
            d1, d2 = handle.to_device(Bm), handle.to_device(Bm)
            plan.ldiv_(d1)
            B.ldiv_(T, d2)
            X1, X2 = d1.to_host(), d2.to_host()
            assert np.array_equal(X1, X2)
            for q in (0, nrhs - 1, nrhs // 2):
                refq = Bm[:, q].copy()
                O.bbb_trsm(rows, *lu, *lm, uplo, unit, np.nan_to_num(data), refq)
                assert relerr(X1[:, q], refq) <= tol(dtype, nnz)
        plan.close()


def test_bbb_triangular_ldiv_gauss_seidel_sweep(handle):
    """examples/finitedifference_2d.jl:17-47 flavour: (D+L) \\ b on the 2-D Laplacian, 200 x 200 blocks."""
    n = 200
    A = B.BandedBlockBandedMatrix.finitedifference_2d(n)
    rng = np.random.default_rng(43)
    b = rng.standard_normal(n * n)
    Ad = A.to_device(handle)
    for uplo, T in (("L", B.LowerTriangular(Ad)), ("U", B.UpperTriangular(Ad))):
        db = handle.to_device(b)
        B.ldiv_(T, db)
        ref = b.copy()
        O.bbb_trsm([n] * n, 1, 1, 1, 1, uplo, False, A.data, ref)
        assert relerr(db.to_host(), ref) <= tol(np.float64, 9)


def test_gauss_seidel_example_on_device(handle):
    """examples/finitedifference_2d.jl:17-47: x <- L \\ (b - U x) sweeps for A = I - dt*Laplacian."""
    from benchmarks.gauss_seidel import gauss_seidel_device, gauss_seidel_host, heat_step_matrix
    n = 48
    A = heat_step_matrix(B, n)
    b = np.random.default_rng(61).standard_normal(n * n)
    x = gauss_seidel_device(B, handle, A.to_device(handle), handle.to_device(b), 5).to_host()
    ref = gauss_seidel_host(A, b, 5)
    assert relerr(x, ref) <= 5 * tol(np.float64, 9)
    assert np.linalg.norm(A.to_dense() @ x - b) < 1e-3 * np.linalg.norm(b)     # and it converges


def test_bbb_triangular_ldiv_errors(handle):
    A = B.BandedBlockBandedMatrix.zeros(np.float64, [3, 4], [4, 3], (1, 1), (1, 1)).to_device(handle)
    with pytest.raises(B.BBMError):       # no matching diagonal blocks: the reference falls back to a generic solve
        B.ldiv_(B.UpperTriangular(A), handle.empty(7, np.float64))
    A2 = B.BandedBlockBandedMatrix.zeros(np.float64, [3, 4], [3, 4], (1, 1), (1, 1)).to_device(handle)
    with pytest.raises(B.DimensionMismatch):
        B.ldiv_(B.LowerTriangular(A2), handle.empty(9, np.float64))


def test_bbb_triangular_mul_needs_matching_blocks(handle):
    A = B.BandedBlockBandedMatrix.zeros(np.float64, [3, 4], [4, 3], (1, 1), (1, 1)).to_device(handle)
    with pytest.raises(B.BBMError):
        B.mul(B.UpperTriangular(A), handle.empty(7, np.float64))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_axpy_and_scalar_mul_on_storage(handle, dtype):
    """axpy!(a, A, B) for identical structure and lmul!(a, A) act on `data`
    (src/BandedBlockBandedMatrix.jl:605-612); checked densely like test/test_broadcasting.jl."""
    rows, cols, lu, lm = list(range(1, 12)), list(range(1, 12)), (1, 2), (2, 1)
    rng = np.random.default_rng(31)
    a_data = random_bbb(rng, rows, cols, *lu, *lm, dtype, poison=False)
    b_data = random_bbb(rng, rows, cols, *lu, *lm, dtype, poison=False)
    A = B.BandedBlockBandedMatrix(a_data, rows, cols, lu, lm)
    Bm = B.BandedBlockBandedMatrix(b_data, rows, cols, lu, lm)
    Ad, Bd = A.to_device(handle), Bm.to_device(handle)
    B.axpy_(2.5, Ad, Bd)
    assert relerr(Bd.to_dense(), 2.5 * A.to_dense().astype(np.float64) + Bm.to_dense()) <= tol(dtype, 2)
    B.lmul_(-3.0, Ad)
    assert relerr(Ad.to_dense(), -3.0 * A.to_dense().astype(np.float64)) <= tol(dtype, 1)
    # odd length / unaligned tail, vectors
    x = rng.standard_normal(1001).astype(dtype)
    y = rng.standard_normal(1001).astype(dtype)
    dx, dy = handle.to_device(x), handle.to_device(y)
    B.axpy_(0.5, dx, dy)
    assert relerr(dy.to_host(), 0.5 * x.astype(np.float64) + y) <= tol(dtype, 2)
    with pytest.raises(TypeError):
        B.axpy_(1.0, Ad, B.BandedBlockBandedMatrix.zeros(dtype, rows, cols, (1, 1), lm).to_device(handle))


BBB_MM_CASES = [
    # rows (= cols), (lA,uA),(lamA,muA), (lB,uB),(lamB,muB)
    (list(range(1, 11)), (1, 1), (1, 1), (1, 1), (1, 1)),          # test/test_linalg.jl:91-104
    (list(range(1, 6)), (-1, 1), (-1, 1), (-1, 1), (-1, 1)),       # test/test_linalg.jl:160-171 (A^2)
    (list(range(1, 6)), (1, -1), (1, -1), (1, -1), (1, -1)),       # B^2
    (list(range(1, 6)), (-1, 1), (-1, 1), (1, -1), (1, -1)),       # A*B -> (0,0),(0,0)
    ([7, 3, 12, 5, 9], (2, 1), (1, 2), (0, 2), (2, 0)),
    ([40] * 6, (1, 1), (2, 2), (1, 0), (0, 3)),
    ([3, 0, 4, 2], (1, 1), (1, 1), (1, 1), (1, 1)),                 # zero-size block
    ([300, 70, 513, 64, 129], (1, 1), (1, 1), (1, 1), (1, 1)),      # wide ragged blocks: the staged (shared-memory) kernel, several tiles per block
    ([256] * 5, (1, 1), (1, 1), (1, 1), (1, 1)),                    # exactly one full tile per block column
]


@pytest.mark.parametrize("kernel", [0, 1, 2], ids=["auto", "column", "entry"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", BBB_MM_CASES, ids=lambda c: f"n{len(c[0])}_A{c[1]}{c[2]}_B{c[3]}{c[4]}".replace(" ", ""))
def test_bbb_matmul_matches_oracle(handle, case, dtype, kernel):
    rows, luA, lmA, luB, lmB = case
    handle.set_option("bbb.mm_kernel", kernel)
    try:
        rng = np.random.default_rng(61)
        adata = random_bbb(rng, rows, rows, *luA, *lmA, dtype)
        bdata = random_bbb(rng, rows, rows, *luB, *lmB, dtype)
        A = B.BandedBlockBandedMatrix(adata, rows, rows, luA, lmA)
        Bm = B.BandedBlockBandedMatrix(bdata, rows, rows, luB, lmB)
        Ad, Bd = A.to_device(handle), Bm.to_device(handle)
        Cd = B.mul(Ad, Bd)                                              # similar(::MulAdd) + mul!
        luC, lmC = (luA[0] + luB[0], luA[1] + luB[1]), (lmA[0] + lmB[0], lmA[1] + lmB[1])
        assert Cd.blockbandwidths == luC and Cd.subblockbandwidths == lmC    # src/linalg.jl:50-71
        nnz = max(1, (luA[0] + luA[1] + 1) * (lmA[0] + lmA[1] + 1)) + 1
        dense = A.to_dense().astype(np.longdouble) @ Bm.to_dense().astype(np.longdouble)
        got = Cd.to_dense()
        assert not np.isnan(got).any()
        assert relerr(got, dense) <= tol(dtype, nnz)
        # against the oracle's block-triple loop, storage slot by storage slot (unused slots: beta*old)
        shape = B.BandedBlockBandedMatrix.data_shape(rows, luC, lmC)
        c0 = np.asfortranarray(rng.standard_normal(shape).astype(dtype))
        Cd2 = B.BandedBlockBandedMatrix(c0.copy(order="F"), rows, rows, luC, lmC).to_device(handle)
        B.mul_(Cd2, Ad, Bd, 1.5, -0.5)
        ref = c0.copy(order="F")
        O.bbb_matmul((rows, rows, *luA, *lmA), (rows, rows, *luB, *lmB), (rows, rows, *luC, *lmC), 1.5,
                     np.nan_to_num(adata), np.nan_to_num(bdata), -0.5, ref)
        got2 = Cd2.data.to_host() if ref.size else ref
        mask = F.bbb_valid_mask(rows, rows, *luC, *lmC) if ref.size else np.zeros(shape, dtype=bool)
        if ref.size:
            assert relerr(got2[mask], ref[mask]) <= tol(dtype, nnz)
            assert relerr(got2[~mask], -0.5 * c0[~mask]) <= tol(dtype, 1)   # _fill_lmul! scales every stored slot
    finally:
        handle.set_option("bbb.mm_kernel", 0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_bbb_matmul_staged_kernel_is_bit_equal(handle, dtype):
    """The staged kernel keeps the ascending (N, i) summation order of the column kernels: identical bits, also into a
    wider C with beta != 0 and NaN in the unused slots of the operands."""
    rows = [300, 70, 513, 64, 129, 1000]
    rng = np.random.default_rng(65)
    adata = random_bbb(rng, rows, rows, 1, 1, 1, 1, dtype)
    bdata = random_bbb(rng, rows, rows, 1, 1, 1, 1, dtype)
    Ad = B.BandedBlockBandedMatrix(adata, rows, rows, (1, 1), (1, 1)).to_device(handle)
    Bd = B.BandedBlockBandedMatrix(bdata, rows, rows, (1, 1), (1, 1)).to_device(handle)
    for luC, lmC in (((2, 2), (2, 2)), ((3, 2), (2, 4))):
        shape = B.BandedBlockBandedMatrix.data_shape(rows, luC, lmC)
        c0 = np.asfortranarray(rng.standard_normal(shape).astype(dtype))
        out = []
        for staged in (1, 0):
            handle.set_option("bbb.mm_staged", staged)
            Cd = B.BandedBlockBandedMatrix(c0.copy(order="F"), rows, rows, luC, lmC).to_device(handle)
            B.mul_(Cd, Ad, Bd, 1.5, -0.5)
            out.append(Cd.data.to_host())
        handle.set_option("bbb.mm_staged", 1)
        assert np.array_equal(out[0], out[1], equal_nan=True)
        mask = F.bbb_valid_mask(rows, rows, *luC, *lmC)
        assert not np.isnan(out[0][mask]).any()


def test_bbb_matmul_errors(handle):
    rows = [3, 4, 5]
    A = B.BandedBlockBandedMatrix.zeros(np.float64, rows, rows, (1, 1), (1, 1)).to_device(handle)
    Cn = B.BandedBlockBandedMatrix.zeros(np.float64, rows, rows, (1, 1), (2, 2)).to_device(handle)
    with pytest.raises(B.BandError):
        B.mul_(Cn, A, A)
    Cw = B.BandedBlockBandedMatrix.zeros(np.float64, rows, rows, (2, 2), (3, 3)).to_device(handle)   # wider C is fine
    B.mul_(Cw, A, A)
    rng = np.random.default_rng(63)
    Ah = B.BandedBlockBandedMatrix(random_bbb(rng, rows, rows, 1, 1, 1, 1), rows, rows, (1, 1), (1, 1))
    for luC, lmC in (((2, 2), (3, 3)), ((3, 2), (2, 4))):               # products landing inside a wider C
        Cw = B.BandedBlockBandedMatrix.zeros(np.float64, rows, rows, luC, lmC).to_device(handle).data.fill_bytes(0xFF)
        Cw = B.BandedBlockBandedMatrix(Cw, rows, rows, luC, lmC)
        B.mul_(Cw, Ah.to_device(handle), Ah.to_device(handle))
        assert relerr(Cw.to_dense(), Ah.to_dense() @ Ah.to_dense()) <= tol(np.float64, 10)
    with pytest.raises(B.DimensionMismatch):
        B.mul_(Cw, A, B.BandedBlockBandedMatrix.zeros(np.float64, [3, 4, 5, 1], rows, (1, 1), (1, 1)).to_device(handle))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", BBB_CASES, ids=lambda c: f"r{len(c[0])}c{len(c[1])}_l{c[2]}u{c[3]}_lam{c[4]}mu{c[5]}")
def test_bbb_adjoint_mul_matches_oracle(handle, case, dtype):
    """mul!(y, A', x): reversed bandwidths (src/adjtransblockbanded.jl:2-3, test/test_adjtransblockbanded.jl)
    and the transposed product against the oracle's gbmv-'T' block loop and a dense check."""
    rows, cols, l, u, lam, mu = case
    rng = np.random.default_rng(71)
    data = random_bbb(rng, rows, cols, l, u, lam, mu, dtype)
    A = B.BandedBlockBandedMatrix(data, rows, cols, (l, u), (lam, mu))
    Ad = A.to_device(handle)
    At = B.Adjoint(Ad)
    assert At.blockbandwidths == (u, l) and At.subblockbandwidths == (mu, lam)
    m, n = A.shape
    for nrhs, alpha, beta in [(None, 1.0, 0.0), (3, 2.0, -0.5)]:
        xs = (m,) if nrhs is None else (m, nrhs)
        ys = (n,) if nrhs is None else (n, nrhs)
        X = np.asfortranarray(rng.standard_normal(xs).astype(dtype))
        Y0 = np.asfortranarray(rng.standard_normal(ys).astype(dtype))
        Yin = Y0.copy(order="F")
        if beta == 0:
            Yin[...] = np.nan
        dY = handle.to_device(Yin)
        B.mul_(dY, At, handle.to_device(X), alpha, beta)
        got = dY.to_host()
        ref = Y0.copy(order="F")
        O.bbb_mul_t(rows, cols, l, u, lam, mu, alpha, np.nan_to_num(data), X, beta, ref)
        nnz = F.bbb_nnz_per_row_max(l, u, lam, mu) + 1
        assert not np.isnan(got).any()
        assert relerr(got, ref) <= tol(dtype, nnz)
        if m and n:
            dense = alpha * (A.to_dense().astype(np.longdouble).T @ X.astype(np.longdouble)) + beta * Y0.astype(np.longdouble)
            assert relerr(got, dense) <= tol(dtype, nnz)
    with pytest.raises(B.DimensionMismatch):
        B.mul_(handle.empty(n + 1, dtype), At, handle.empty(m, dtype))


def test_bbb_mul_cached_plans_of_two_descriptors(handle):
    """Two matrices with the same bandwidths (one kernel instantiation) but different block sizes (different launch
    plans, different shared-memory footprints), multiplied alternately: the later plan must not invalidate the earlier
    one (the dynamic shared-memory limit is a per-kernel attribute)."""
    rng = np.random.default_rng(321)
    mats = []
    for rows in ([2048] * 6, [40] * 12, list(range(20, 60))):
        data = random_bbb(rng, rows, rows, 1, 1, 1, 1)
        A = B.BandedBlockBandedMatrix(data, rows, rows, (1, 1), (1, 1)).to_device(handle)
        mats.append((rows, data, A))
    for _ in range(3):
        for rows, data, A in mats:
            x = rng.standard_normal(sum(rows))
            y = B.mul(A, handle.to_device(x)).to_host()
            ref = np.zeros(sum(rows))
            O.bbb_mul(rows, rows, 1, 1, 1, 1, 1.0, data, x, 0.0, ref)
            assert relerr(y, ref) <= tol(np.float64, 9)
