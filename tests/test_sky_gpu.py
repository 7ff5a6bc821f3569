"""Parity of the CUDA BlockSkylineMatrix paths (matvec, block-banded x block-banded product on the
FP64 tensor cores, block triangular solve) with the CPU oracle, through the C ABI.

Tolerance (BASELINE.json north_star): normwise relative error <= 64*eps(T)*(nonzeros per row)."""
import numpy as np
import pytest

import bbm_b200 as B
from oracle import cpu as O
from oracle import formats as F
from tests.helpers import random_sky, relerr, tol

pytestmark = pytest.mark.gpu

SKY_CASES = [
    # rows, cols, l, u   (scalars = BlockBandedMatrix, vectors = ragged BlockSkylineMatrix)
    (list(range(1, 11)), list(range(1, 11)), 1, 1),                               # test/test_linalg.jl:35-57
    ([256] * 6, [256] * 6, 2, 2),                                                 # cfg4-like
    ([1, 2, 3, 4], [1, 2, 3, 4], 1, 1),                                           # test/test_blockbanded.jl:59-90
    ([5, 3, 7, 2, 6], [4, 6, 2, 5, 3, 7], 1, 2),                                  # rectangular
    ([3] * 7, [3] * 7, [2, 1, 2, 3, 1, 0, 0], [0, 1, 1, 2, 3, 1, 2]),             # ragged (test/test_blockskyline.jl:17-19)
    ([3] * 7, [3] * 7, [-1, 1, 2, 0, 1, 0, 0], [1, 1, 0, 2, -1, 1, 2]),           # ragged with negative entries
    ([2, 2], [2, 2, 2], 0, 1),                                                    # test/test_blockbanded.jl:213-221
    ([], [], 1, 1),
    ([4, 0, 5, 3], [4, 0, 5, 3], 1, 1),                                           # empty blocks
    ([130, 7, 260, 129], [64, 200, 31, 150], 2, 1),                               # blocks spanning several tiles
    ([6] * 5, [6] * 5, -1, 2),                                                    # band above the diagonal
]


@pytest.fixture(scope="module")
def handle():
    h = B.Handle(0)
    yield h
    h.close()


def ids(c):
    return f"r{len(c[0])}c{len(c[1])}"


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", SKY_CASES, ids=ids)
def test_sky_mul_matches_oracle(handle, case, dtype):
    rows, cols, l, u = case
    rng = np.random.default_rng(3)
    data = random_sky(rng, rows, cols, l, u, dtype)
    A = B.BlockSkylineMatrix(data, rows, cols, l, u)
    Ad = A.to_device(handle)
    m, n = A.shape
    nnz = int(max(A.block_strides.max() if len(A.cols) else 0, 1)) * 3
    for nrhs, alpha, beta in [(None, 1.0, 0.0), (None, 2.5, -0.5), (3, 1.0, 1.0), (11, -1.0, 0.0)]:
        xs = (n,) if nrhs is None else (n, nrhs)
        ys = (m,) if nrhs is None else (m, nrhs)
        X = np.asfortranarray(rng.standard_normal(xs).astype(dtype))
        Y0 = np.asfortranarray(rng.standard_normal(ys).astype(dtype))
        Yin = Y0.copy(order="F")
        if beta == 0:
            Yin[...] = np.nan  # test/test_blockskyline.jl:34-39
        dX, dY = handle.to_device(X), handle.to_device(Yin)
        B.mul_(dY, Ad, dX, alpha, beta)
        got = dY.to_host()
        ref = Y0.copy(order="F")
        O.sky_mul(rows, cols, l, u, alpha, data, X, beta, ref)
        assert not np.isnan(got).any()
        assert relerr(got, ref) <= tol(dtype, nnz)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", SKY_CASES, ids=ids)
def test_sky_transposed_mul_matches_oracle(handle, case, dtype):
    """mul!(y, A', x) / transpose(A)*x for a BlockSkylineMatrix (src/adjtransblockbanded.jl:2-7)."""
    rows, cols, l, u = case
    rng = np.random.default_rng(5)
    data = random_sky(rng, rows, cols, l, u, dtype)
    A = B.BlockSkylineMatrix(data, rows, cols, l, u)
    Ad = A.to_device(handle)
    m, n = A.shape
    nnz = int(max(A.block_strides.max() if len(A.cols) else 0, 1))
    Dn = A.to_dense().astype(np.float64)
    for nrhs, alpha, beta in [(None, 1.0, 0.0), (None, 2.5, -0.5), (3, 1.0, 1.0), (9, -1.0, 0.0)]:
        xs = (m,) if nrhs is None else (m, nrhs)
        ys = (n,) if nrhs is None else (n, nrhs)
        X = np.asfortranarray(rng.standard_normal(xs).astype(dtype))
        Y0 = np.asfortranarray(rng.standard_normal(ys).astype(dtype))
        Yin = Y0.copy(order="F")
        if beta == 0:
            Yin[...] = np.nan
        dX, dY = handle.to_device(X), handle.to_device(Yin)
        B.mul_(dY, B.Adjoint(Ad), dX, alpha, beta)
        got = dY.to_host()
        ref = Y0.copy(order="F")
        O.sky_mul_t(rows, cols, l, u, alpha, data, X, beta, ref)
        assert not np.isnan(got).any()
        assert relerr(got, ref) <= tol(dtype, nnz)
        dense = alpha * (Dn.T @ X.astype(np.float64)) + beta * Y0.astype(np.float64)
        assert relerr(got, dense) <= tol(dtype, nnz)


def product_structure(A, Bm):
    l, u = B.add_bandwidths(A, Bm)
    return l, u


@pytest.mark.parametrize("kernel", [0, 2], ids=["dmma", "simt"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", [c for c in SKY_CASES if len(c[0]) == len(c[1])], ids=ids)
def test_sky_matmul_matches_oracle(handle, case, dtype, kernel):
    rows, cols, l, u = case
    if list(rows) != list(cols):
        pytest.skip("A*A needs matching axes")
    handle.set_option("sky.mm_kernel", kernel)
    try:
        rng = np.random.default_rng(4)
        adata = random_sky(rng, rows, cols, l, u, dtype)
        bdata = random_sky(rng, rows, cols, l, u, dtype)
        A = B.BlockSkylineMatrix(adata, rows, cols, l, u)
        Bm = B.BlockSkylineMatrix(bdata, rows, cols, l, u)
        Ad, Bd = A.to_device(handle), Bm.to_device(handle)
        # A*B allocates the result with the reference's structure rule (src/linalg.jl:8-48)
        Cd = B.mul(Ad, Bd)
        cl, cu = product_structure(A, Bm)
        ol, ou = F.add_bandwidths(A.l, A.u, Bm.l, Bm.u, constant=A.constant_bandwidths and Bm.constant_bandwidths)
        assert list(np.broadcast_to(cl, len(cols))) == list(ol) and list(np.broadcast_to(cu, len(cols))) == list(ou)
        got = Cd.data.to_host()
        ref = np.zeros_like(got)
        O.sky_matmul((rows, cols, l, u), (rows, cols, l, u), (rows, cols, ol, ou), 1.0, adata, bdata, 0.0, ref)
        inner = int(max(A.block_strides.max() if len(cols) else 0, 1)) * 2
        assert not np.isnan(got).any()
        assert relerr(got, ref) <= tol(dtype, inner)
        # dense cross-check (independent of the oracle's loops)
        if A.shape[0] and A.shape[0] <= 600:
            dense = A.to_dense().astype(np.longdouble) @ Bm.to_dense().astype(np.longdouble)
            assert relerr(Cd.to_dense(), dense) <= tol(dtype, inner)
        # mul! with alpha, beta into the existing C
        c0 = rng.standard_normal(got.shape).astype(dtype)
        Cd2 = B.BlockSkylineMatrix(handle.to_device(c0), rows, cols, ol, ou)
        B.mul_(Cd2, Ad, Bd, -1.5, 0.75)
        ref2 = c0.copy()
        O.sky_matmul((rows, cols, l, u), (rows, cols, l, u), (rows, cols, ol, ou), -1.5, adata, bdata, 0.75, ref2)
        assert relerr(Cd2.data.to_host(), ref2) <= tol(dtype, inner)
    finally:
        handle.set_option("sky.mm_kernel", 0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("rows,cols,l,u", [(list(range(1, 5)), list(range(1, 5)), 1, 1),          # test/test_linalg.jl:38-57
                                           ([5, 3, 7, 2, 6], [4, 6, 2, 5, 3, 7], 1, 2),
                                           ([3] * 7, [3] * 7, [2, 1, 2, 3, 1, 0, 0], [0, 1, 1, 2, 3, 1, 2]),
                                           ([2, 2, 2], [2, 2, 2], [0, -1, 0], [0, -1, 0]),        # an empty block column
                                           ([128, 64, 200, 36], [128, 64, 200, 36], 1, 1),
                                           ([256] * 5, [256] * 5, 2, 2)])
@pytest.mark.parametrize("p", [1, 10, 150])
def test_sky_rmul_dense_times_block_banded(handle, rows, cols, l, u, dtype, p):
    """`X*A` and `mul!(Y, X, A, alpha, beta)` for a dense X (test/test_linalg.jl:56-57: `X*A ≈ (similar(X) .= MulAdd(X,A)) ≈ Matrix(X)*A`)
    against the restated block loop; beta = 0 never reads the NaN-filled destination."""
    rng = np.random.default_rng(31)
    data = random_sky(rng, rows, cols, l, u, dtype)
    A = B.BlockSkylineMatrix(data, rows, cols, l, u).to_device(handle)
    m, n = sum(rows), sum(cols)
    X = np.asfortranarray(rng.standard_normal((p, m)).astype(dtype))
    dX = handle.to_device(X)
    nnz = int(max(A.block_strides.max() if len(cols) else 0, 1))
    got = B.mul(dX, A).to_host()
    ref = np.zeros((p, n), dtype=dtype, order="F")
    O.sky_rmul(rows, cols, l, u, 1.0, X, data, 0.0, ref)
    assert got.shape == (p, n) and not np.isnan(got).any()
    assert relerr(got, ref) <= tol(dtype, nnz)
    Y0 = np.asfortranarray(rng.standard_normal((p, n)).astype(dtype))
    dY = handle.to_device(Y0)
    B.mul_(dY, dX, A, 2.5, -0.5)
    ref = Y0.copy(order="F")
    O.sky_rmul(rows, cols, l, u, 2.5, X, data, -0.5, ref)
    assert relerr(dY.to_host(), ref) <= tol(dtype, nnz + 1)
    dY = handle.to_device(np.full((p, n), np.nan, dtype=dtype, order="F"))
    B.mul_(dY, dX, A, 1.0, 0.0)
    assert not np.isnan(dY.to_host()).any()
    with pytest.raises(B.DimensionMismatch):
        B.mul_(handle.empty((p, n + 1), dtype), dX, A)


@pytest.mark.parametrize("f32_kernel", [0, 1], ids=["register-tiled", "first-version"])
@pytest.mark.parametrize("rows", [[128, 256, 384, 128, 256, 128], [132, 60, 260, 4, 96, 200], [100, 37, 250, 129, 3, 64]],
                         ids=["full-tiles", "16-byte-aligned-ragged", "unaligned-ragged"])
def test_sky_matmul_float32_kernels(handle, rows, f32_kernel):
    """Float32 products (both SIMT kernels; 16-byte cp.async path and element-wise path; tiles cut by the block edges; several
    k-chunks and product segments per tile; alpha/beta; a wider C; ragged per-column bandwidths) against the oracle."""
    cases = [((1, 2), (2, 1), None), ((1, 1), (1, 1), (3, 3)), (([1, 0, 2, 1, 0, 1], [0, 1, 1, 2, 1, 0]), ([0, 1, 1, 0, 2, 1], [1, 1, 0, 1, 1, 2]), None)]
    handle.set_option("sky.mm_f32_kernel", f32_kernel)
    try:
        for (la, ua), (lb, ub), wide in cases:
            rng = np.random.default_rng(15)
            adata, bdata = random_sky(rng, rows, rows, la, ua, np.float32), random_sky(rng, rows, rows, lb, ub, np.float32)
            A = B.BlockSkylineMatrix(adata, rows, rows, la, ua)
            Bm = B.BlockSkylineMatrix(bdata, rows, rows, lb, ub)
            if wide is None:
                ol, ou = F.add_bandwidths(A.l, A.u, Bm.l, Bm.u, constant=A.constant_bandwidths and Bm.constant_bandwidths)
            else:
                ol, ou = wide
            n = B.BlockSkylineMatrix.numentries(rows, rows, ol, ou)
            c0 = rng.standard_normal(n).astype(np.float32)
            for alpha, beta in ((1.0, 0.0), (-1.5, 0.75)):
                cin = c0.copy()
                if beta == 0:
                    cin[:] = np.nan
                Cd = B.BlockSkylineMatrix(handle.to_device(cin), rows, rows, ol, ou)
                B.mul_(Cd, A.to_device(handle), Bm.to_device(handle), alpha, beta)
                ref = c0.copy()
                O.sky_matmul((rows, rows, la, ua), (rows, rows, lb, ub), (rows, rows, ol, ou), alpha, adata, bdata, beta, ref)
                got = Cd.data.to_host()
                assert not np.isnan(got).any()
                assert relerr(got, ref) <= tol(np.float32, 3 * max(rows))
    finally:
        handle.set_option("sky.mm_f32_kernel", 0)


@pytest.mark.parametrize("ws", [1, 0], ids=["bulk-copy-pipeline", "cp.async"])
def test_sky_matmul_full_tiles_warp_specialised(handle, ws):
    """Blocks that are multiples of 128 with 16-byte-aligned panels take the warp-specialised DMMA kernel (one producer thread
    feeding a bulk-copy ring, "sky.mm_ws"); same oracle, alpha/beta, a wider C and ragged per-column bandwidths."""
    rows = [128, 256, 384, 128, 256, 128]
    cases = [((1, 2), (2, 1), None), ((1, 1), (1, 1), (3, 3)), (([1, 0, 2, 1, 0, 1], [0, 1, 1, 2, 1, 0]), ([0, 1, 1, 0, 2, 1], [1, 1, 0, 1, 1, 2]), None)]
    handle.set_option("sky.mm_ws", ws)
    try:
        for (la, ua), (lb, ub), wide in cases:
            rng = np.random.default_rng(14)
            adata, bdata = random_sky(rng, rows, rows, la, ua), random_sky(rng, rows, rows, lb, ub)
            A = B.BlockSkylineMatrix(adata, rows, rows, la, ua)
            Bm = B.BlockSkylineMatrix(bdata, rows, rows, lb, ub)
            if wide is None:
                ol, ou = F.add_bandwidths(A.l, A.u, Bm.l, Bm.u, constant=A.constant_bandwidths and Bm.constant_bandwidths)
            else:
                ol, ou = wide
            n = B.BlockSkylineMatrix.numentries(rows, rows, ol, ou)
            c0 = rng.standard_normal(n)
            for alpha, beta in ((1.0, 0.0), (-1.5, 0.75)):
                cin = c0.copy()
                if beta == 0:
                    cin[:] = np.nan
                Cd = B.BlockSkylineMatrix(handle.to_device(cin), rows, rows, ol, ou)
                B.mul_(Cd, A.to_device(handle), Bm.to_device(handle), alpha, beta)
                ref = c0.copy()
                O.sky_matmul((rows, rows, la, ua), (rows, rows, lb, ub), (rows, rows, ol, ou), alpha, adata, bdata, beta, ref)
                got = Cd.data.to_host()
                assert not np.isnan(got).any()
                assert relerr(got, ref) <= tol(np.float64, 3 * 384)
    finally:
        handle.set_option("sky.mm_ws", 0)


def test_sky_matmul_into_wider_c_and_band_error(handle):
    """test/test_blockbanded.jl:213-221: mul!(C, A, B) into a preallocated wider-banded C; a C that is
    too narrow is a BandError (src/BlockSkylineMatrix.jl:491)."""
    rows, cols = [2, 2], [2, 2]
    A = B.BlockSkylineMatrix(np.ones(B.BlockSkylineMatrix.numentries(rows, cols, 0, 1)), rows, cols, 0, 1).to_device(handle)
    Bm = B.BlockSkylineMatrix(np.ones(B.BlockSkylineMatrix.numentries(rows, cols, 1, 0)), rows, cols, 1, 0).to_device(handle)
    Cw = B.BlockSkylineMatrix(np.full(B.BlockSkylineMatrix.numentries(rows, cols, 1, 1), np.nan), rows, cols, 1, 1).to_device(handle)
    B.mul_(Cw, A, Bm)
    dense = A.to_dense() @ Bm.to_dense()
    assert np.array_equal(Cw.to_dense(), dense)
    Cn = B.BlockSkylineMatrix.zeros(np.float64, rows, cols, 0, 0).to_device(handle)
    with pytest.raises(B.BandError):
        B.mul_(Cn, A, Bm)
    with pytest.raises(B.DimensionMismatch):
        B.mul_(Cw, A, B.BlockSkylineMatrix.zeros(np.float64, [2, 2, 2], cols, 1, 1).to_device(handle))


def test_sky_matmul_ragged_all_ones_exact(handle):
    """test/test_blockskyline.jl:42-81: ragged M*M on all-ones data is exact, product bandwidths follow
    add_bandwidths."""
    rows = cols = [3] * 7
    l, u = [2, 1, 2, 3, 1, 0, 0], [0, 1, 1, 2, 3, 1, 2]
    n = B.BlockSkylineMatrix.numentries(rows, cols, l, u)
    M = B.BlockSkylineMatrix(np.ones(n), rows, cols, l, u)
    Md = M.to_device(handle)
    MM = B.mul(Md, Md)
    dense = M.to_dense() @ M.to_dense()
    assert np.array_equal(MM.to_dense(), dense)


TRI_CASES = [
    # rows, l, u
    (list(range(1, 11)), 1, 1),        # test/test_linalg.jl:59-76
    ([64] * 8, 0, 3),                  # cfg5-like
    ([64] * 8, 3, 0),
    ([40, 7, 100, 33, 1, 65], 2, 2),
    ([512, 300], 1, 1),                # diagonal blocks spanning many 32-row sub-blocks
    ([5, 0, 6, 3], 1, 2),
    ([3] * 7, [2, 1, 2, 3, 1, 0, 0], [0, 1, 1, 2, 3, 1, 2]),
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("uplo,unit", [("U", False), ("U", True), ("L", False), ("L", True)])
@pytest.mark.parametrize("case", TRI_CASES, ids=lambda c: f"N{len(c[0])}")
def test_sky_trsm_matches_oracle(handle, case, uplo, unit, dtype):
    rows, l, u = case
    rng = np.random.default_rng(8)
    A = B.BlockSkylineMatrix.zeros(dtype, rows, rows, l, u)
    m = A.shape[0]
    # well conditioned: N(0,1)/sqrt(m) off the diagonal, +10 on the global diagonal
    data = (rng.standard_normal(A.data.shape) / np.sqrt(max(m, 1))).astype(dtype)
    A = B.BlockSkylineMatrix(data, rows, rows, l, u)
    for K in range(len(rows)):
        s, ld = A.block_start(K, K), int(A.block_strides[K])
        for i in range(rows[K]):
            data[s + i + i * ld] += 10
    Ad = A.to_device(handle)
    T = {"U": (B.UnitUpperTriangular if unit else B.UpperTriangular),
         "L": (B.UnitLowerTriangular if unit else B.LowerTriangular)}[uplo](Ad)
    for nrhs in (None, 5, 64):
        shape = (m,) if nrhs is None else (m, nrhs)
        B0 = np.asfortranarray(rng.standard_normal(shape).astype(dtype))
        dB = handle.to_device(B0)
        out = B.ldiv_(T, dB)
        assert out is dB
        got = dB.to_host()
        ref = B0.copy(order="F")
        O.sky_trsm(rows, l, u, uplo, unit, data, ref)
        assert not np.isnan(got).any()
        assert relerr(got, ref) <= tol(dtype, max(m, 1))
    # residual check against the dense triangle (independent of the oracle)
    if m:
        D = A.to_dense().astype(np.float64)
        Tm = np.triu(D) if uplo == "U" else np.tril(D)
        if unit:
            np.fill_diagonal(Tm, 1.0)
        X = B.ldiv(T, handle.to_device(B0)).to_host().astype(np.float64)
        assert relerr(Tm @ X, B0.astype(np.float64)) <= tol(dtype, m) * 4


@pytest.mark.parametrize("uplo", ["U", "L"])
@pytest.mark.parametrize("kernel,path,flow_bk", [(0, 1, 64), (0, 1, 32), (3, 2, 64), (2, 0, 64)])
def test_sky_trsm_many_rhs_every_path(handle, uplo, kernel, path, flow_bk):
    """64 right-hand sides, aligned block sizes: the data-flow kernel (one persistent launch, default), the launch chain
    ("trsm.kernel" = 3) and the substitution kernels (= 2) against the oracle; several block steps deep so that every
    counter of the data-flow kernel is exercised (far updates of 3 block columns, ragged multiples of 2)."""
    rows = [128, 64, 192, 66, 0, 130, 64, 256, 32, 96, 64, 128]
    l, u = (0, 3) if uplo == "U" else (3, 0)
    rng = np.random.default_rng(81)
    m = sum(rows)
    data = rng.standard_normal(B.BlockSkylineMatrix.numentries(rows, rows, l, u)) / np.sqrt(m)
    A = B.BlockSkylineMatrix(data, rows, rows, l, u)
    for K in range(len(rows)):
        if rows[K]:
            s, ld = A.block_start(K, K), int(A.block_strides[K])
            i = np.arange(rows[K])
            data[s + i + i * ld] += 10
    Ad = A.to_device(handle)
    T = (B.UpperTriangular if uplo == "U" else B.LowerTriangular)(Ad)
    handle.set_option("trsm.kernel", kernel)
    handle.set_option("trsm.flow_bk", flow_bk)
    try:
        for nrhs in (64, 40, 16):
            B0 = np.asfortranarray(rng.standard_normal((m, nrhs)))
            dB = handle.to_device(B0)
            handle.set_option("trsm.last_path", 0)
            B.ldiv_(T, dB)
            got = dB.to_host()
            assert handle.get_option("trsm.last_path") == path
            ref = B0.copy(order="F")
            O.sky_trsm(rows, l, u, uplo, False, data, ref)
            assert not np.isnan(got).any()
            assert relerr(got, ref) <= tol(np.float64, 4 * 256)
    finally:
        handle.set_option("trsm.kernel", 0)
        handle.set_option("trsm.flow_bk", 64)


def test_sky_trsm_condition_guard(handle):
    """Ill-conditioned diagonal blocks (cond about 1e7): the inverse-based paths are not backward stable, so the solve must
    fall back to substitution -- like the reference's trsv -- and leave a residual at the n*eps level; with the guard
    switched off the explicit inverses show the cond*eps residual the guard exists to avoid."""
    rows = [64] * 8
    l, u = 0, 2
    rng = np.random.default_rng(82)
    m = sum(rows)
    data = rng.standard_normal(B.BlockSkylineMatrix.numentries(rows, rows, l, u)) / np.sqrt(m)
    A = B.BlockSkylineMatrix(data, rows, rows, l, u)
    for K in range(len(rows)):
        s, ld = A.block_start(K, K), int(A.block_strides[K])
        dg = np.logspace(0, -6, 64)
        R = np.diag(dg) @ (np.eye(64) + 0.01 * np.triu(rng.standard_normal((64, 64)), 1))        # graded rows: cond about 1e6
        for j in range(64):
            data[s + j * ld:s + j * ld + 64] = R[:, j]
    D = np.triu(A.to_dense())
    cond = max(np.linalg.cond(D[64 * K:64 * K + 64, 64 * K:64 * K + 64], np.inf) for K in range(8))
    assert 1e5 < cond < 1e9
    Ad = A.to_device(handle)
    T = B.UpperTriangular(Ad)
    B0 = np.asfortranarray(rng.standard_normal((m, 32)))

    def residual(X):
        return np.linalg.norm(D @ X - B0) / (np.linalg.norm(D, 2) * np.linalg.norm(X) + np.linalg.norm(B0))
    dB = handle.to_device(B0)
    B.ldiv_(T, dB)
    assert handle.get_option("trsm.last_path") == 3                      # substitution
    assert handle.get_option("trsm.last_cond_x1000") / 1000.0 >= 0.01 * cond
    X = dB.to_host()
    assert residual(X) <= 64 * 2.22e-16 * 192
    ref = B0.copy(order="F")
    O.sky_trsm(rows, l, u, "U", False, data, ref)
    assert relerr(X, ref) <= 64 * 2.22e-16 * 192 * cond                  # forward error: cond-scaled, as for any backward-stable solve
    # the guard off: the explicit-inverse path still runs (and is what the guard protects against)
    handle.set_option("trsm.cond_check", 0)
    try:
        dB2 = handle.to_device(B0)
        B.ldiv_(T, dB2)
        assert handle.get_option("trsm.last_path") == 1
        assert np.isfinite(dB2.to_host()).all()
    finally:
        handle.set_option("trsm.cond_check", 1)
    # a well-conditioned matrix keeps the fast path
    data2 = data.copy()
    A2 = B.BlockSkylineMatrix(data2, rows, rows, l, u)
    for K in range(len(rows)):
        s, ld = A2.block_start(K, K), int(A2.block_strides[K])
        for j in range(64):
            data2[s + j * ld:s + j * ld + 64] = (rng.standard_normal(64) / 8) * (np.arange(64) <= j)
            data2[s + j * ld + j] += 10
    dB3 = handle.to_device(B0)
    B.ldiv_(B.UpperTriangular(A2.to_device(handle)), dB3)
    assert handle.get_option("trsm.last_path") == 1


def test_sky_trsm_errors(handle):
    A = B.BlockSkylineMatrix.zeros(np.float64, [3, 4], [4, 3], 1, 1).to_device(handle)
    with pytest.raises(B.BBMError):   # no matching square diagonal blocks: not taken by the device path
        B.ldiv_(B.UpperTriangular(A), handle.empty(7, np.float64))
    A2 = B.BlockSkylineMatrix.zeros(np.float64, [3, 3, 3], [3, 3, 3], -1, 2).to_device(handle)
    with pytest.raises(B.BandError):  # diagonal blocks not stored
        B.ldiv_(B.UpperTriangular(A2), handle.empty(9, np.float64))
    A3 = B.BlockSkylineMatrix.zeros(np.float64, [3, 3], [3, 3], 1, 1).to_device(handle)
    with pytest.raises(B.DimensionMismatch):
        B.ldiv_(B.UpperTriangular(A3), handle.empty(5, np.float64))


def test_sky_layout_matches_host_container(handle):
    """Device descriptor's panel table == the host container's block_starts / block_strides."""
    import ctypes as C
    rows, cols = [1, 2, 3, 4], [1, 2, 3, 4]
    A = B.BlockBandedMatrix(np.arange(1, 71, dtype=np.float64), rows, cols, (1, 1)).to_device(handle)
    d = A.descriptor(handle)
    off = np.zeros(5, dtype=np.int64); S = np.zeros(4, dtype=np.int64)
    klo = np.zeros(4, dtype=np.int64); khi = np.zeros(4, dtype=np.int64)
    assert handle._lib.bbm_sky_desc_layout(d, off.ctypes.data_as(C.c_void_p), S.ctypes.data_as(C.c_void_p),
                                           klo.ctypes.data_as(C.c_void_p), khi.ctypes.data_as(C.c_void_p)) == 0
    assert list(S) == list(A.block_strides) == [3, 6, 9, 7]      # strides (1,7) of Block(4,4): test/test_linalg.jl:42-47
    assert list(off) == list(A.panel_offsets)
    assert off[3] + (A.R[3] - A.R[klo[3]]) == 45                 # first element of Block(4,4) is data[46] (1-based)


QR_GPU_CASES = [
    (list(range(1, 6)), list(range(1, 6)), 2, 1),      # test/test_blockskylineqr.jl:13-35 "Square QR"
    (list(range(1, 7)), list(range(1, 6)), 2, 1),      # :37-58 "Thin QR" (lmul!(Q', b) only)
    ([4] * 6, [4] * 6, 1, 1),
    ([3, 5, 2, 4], [3, 5, 2, 4], 1, 2),
    ([40] * 5, [40] * 5, 1, 1),                        # one doubling level of the T factors
    ([100, 70, 129, 33], [100, 70, 129, 33], 2, 1),    # ragged, three doubling levels, odd leading dimensions
    ([64] * 12, [64] * 12, 1, 2),
    ([7], [7], 0, 0),
    ([4, 0, 3, 5], [4, 0, 3, 5], 1, 1),                # zero-size block
    ([0, 3, 2], [0, 3, 2], 1, 1),
]


def _factor_on_host(rng, rows, cols, l, u):
    m, n = sum(rows), sum(cols)
    Ad = F.sky_to_dense(F.sky_from_dense(rng.standard_normal((m, n)) + 3.0 * np.eye(m, n), rows, cols, l, u), rows, cols, l, u)
    data = F.sky_from_dense(Ad, rows, cols, l, l + u)
    tau = O.sky_qr(rows, cols, l, l + u, data)                  # the reference computes qr! on the host too
    return Ad, data, tau


@pytest.mark.parametrize("nrhs", [1, 5, 64])
@pytest.mark.parametrize("rows,cols,l,u", QR_GPU_CASES)
def test_block_banded_qr_lmul_adj_and_ldiv(handle, rows, cols, l, u, nrhs):
    """lmul!(F.Q', B) and F \\ B (test/test_blockskylineqr.jl:29-34,52-57) against the restated per-panel
    reflector sweep and dense LAPACK."""
    rng = np.random.default_rng(71)
    Ad, data, tau = _factor_on_host(rng, rows, cols, l, u)
    m = sum(rows)
    Fd = B.BlockBandedQR(B.BlockSkylineMatrix(data, rows, cols, l, l + u), tau).to_device(handle)
    B0 = np.asfortranarray(rng.standard_normal((m, nrhs))) if nrhs > 1 else rng.standard_normal(m)
    dB = handle.to_device(B0)
    B.lmul_adj_(Fd, dB)
    ref = np.asfortranarray(B0.copy())
    O.sky_qr_lmul_adj(rows, cols, l, l + u, data, tau, ref)
    nnz = max(rows) * (l + 1) + max(cols)
    assert relerr(dB.to_host(), ref) <= tol(np.float64, nnz)
    Q, _ = np.linalg.qr(Ad, mode="complete")
    assert relerr(dB.to_host(), Q.T @ B0) <= tol(np.float64, nnz)
    if rows == cols:
        dX = B.ldiv(Fd, handle.to_device(B0))
        ref = np.asfortranarray(B0.copy())
        O.sky_qr_ldiv(rows, l, l + u, data, tau, ref)
        cond = np.linalg.cond(Ad)
        assert relerr(dX.to_host(), ref) <= tol(np.float64, nnz) * max(1.0, cond)
        assert relerr(Ad @ dX.to_host(), B0) <= tol(np.float64, nnz) * max(1.0, cond)


@pytest.mark.parametrize("flow", [1, 0], ids=["data-flow", "launch-chain"])
@pytest.mark.parametrize("rows,cols,l,u", [([64] * 12, [64] * 12, 1, 2), ([128, 64, 192, 64, 128], [128, 64, 192, 64, 128], 2, 1),
                                           ([64, 128, 64, 64], [64, 128, 64], 1, 1), ([192] * 6, [192] * 6, 3, 0)])
@pytest.mark.parametrize("nrhs", [8, 33, 64])
def test_block_banded_qr_sweep_kernels(handle, rows, cols, l, u, nrhs, flow):
    """The sweep of lmul!(F.Q', B) as one persistent data-flow kernel (block sizes multiples of 64, >= 8 right-hand sides;
    option qr.flow) and as the launch chain: same oracle, partial column tiles, thin factors, several block rows per panel."""
    rng = np.random.default_rng(75)
    Ad, data, tau = _factor_on_host(rng, rows, cols, l, u)
    m = sum(rows)
    Fd = B.BlockBandedQR(B.BlockSkylineMatrix(data, rows, cols, l, l + u), tau).to_device(handle)
    B0 = np.asfortranarray(rng.standard_normal((m, nrhs)))
    handle.set_option("qr.flow", flow)
    try:
        for _ in range(2):                                  # the second call re-uses the plan and re-zeroes the counters
            dB = handle.to_device(B0)
            B.lmul_adj_(Fd, dB)
            got = dB.to_host()
        if rows == cols:
            dX = B.ldiv(Fd, handle.to_device(B0)).to_host()
    finally:
        handle.set_option("qr.flow", 1)
    ref = np.asfortranarray(B0.copy())
    O.sky_qr_lmul_adj(rows, cols, l, l + u, data, tau, ref)
    nnz = max(rows) * (l + 1) + max(cols)
    assert not np.isnan(got).any()
    assert relerr(got, ref) <= tol(np.float64, nnz)
    if rows == cols:
        assert relerr(Ad @ dX, B0) <= tol(np.float64, nnz) * max(1.0, np.linalg.cond(Ad))


def test_block_banded_qr_errors(handle):
    rng = np.random.default_rng(72)
    rows, cols = [3, 3], [3, 3, 3]
    data = np.zeros(F.sky_numentries(rows, cols, 1, 2))
    Fw = B.BlockBandedQR(B.BlockSkylineMatrix(data, rows, cols, 1, 2), np.zeros(6)).to_device(handle)
    with pytest.raises(B.BBMError):                   # wide QR: not taken by the device path
        B.lmul_adj_(Fw, handle.empty(6, np.float64))
    _, data, tau = _factor_on_host(rng, [4] * 3, [4] * 3, 1, 1)
    Fd = B.BlockBandedQR(B.BlockSkylineMatrix(data, [4] * 3, [4] * 3, 1, 2), tau).to_device(handle)
    with pytest.raises(B.DimensionMismatch):
        B.ldiv_(Fd, handle.empty(11, np.float64))


@pytest.mark.parametrize("rows,cols,l,u", QR_GPU_CASES + [([300, 20, 150], [300, 20, 150], 1, 1), ([50] * 4, [50] * 4, 3, 0)])
def test_block_banded_qr_factorisation_on_device(handle, rows, cols, l, u):
    """qr!(A) (test/test_blockskylineqr.jl:19-20,43-44): F.factors and F.tau against the restated
    _blockbanded_qr! and against LAPACK's dense geqrf; then F \\ b with the device factors."""
    rng = np.random.default_rng(73)
    m, n = sum(rows), sum(cols)
    A = B.BlockSkylineMatrix(F.sky_from_dense(rng.standard_normal((m, n)) + 3.0 * np.eye(m, n), rows, cols, l, u), rows, cols, l, u)
    Aw = A.with_bandwidths(l, l + u)                          # `_qr`, src/blockskylineqr.jl:75
    Ad = A.to_dense()
    assert np.array_equal(Aw.to_dense(), Ad)
    ref = Aw.data.copy()
    tau_ref = O.sky_qr(rows, cols, l, l + u, ref)
    Fd = B.qr_(Aw.to_device(handle))
    got, tau = Fd.factors.data.to_host(), Fd.tau.to_host()
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 64 * 2.3e-16 * (max(rows) * (l + 1)) * scale
    assert np.abs(tau - tau_ref).max() <= 64 * 2.3e-16 * (max(rows) * (l + 1))
    h, t = np.linalg.qr(Ad, mode="raw")
    fac = F.sky_to_dense(got, rows, cols, l, l + u)
    assert np.abs(fac - h.T).max() <= 1e-12 * np.abs(h).max()
    if rows == cols:
        b = rng.standard_normal(m)
        x = B.ldiv(Fd, handle.to_device(b)).to_host()
        assert relerr(Ad @ x, b) <= tol(np.float64, max(rows) * (2 * l + u + 1)) * max(1.0, np.linalg.cond(Ad))


@pytest.mark.parametrize("panel,lookahead", [(0, 1), (0, 0), (1, 1)])
@pytest.mark.parametrize("rows,l,u", [([256] * 4, 1, 1),              # 512-row slabs: 32 warps x 16 rows per lane
                                      ([280, 260, 300], 2, 1),        # 840-row slabs: 16 warps x 32 rows per lane
                                      ([600, 500, 90], 1, 1),         # 1100-row slabs: shared-memory slab kernel
                                      ([37, 5, 64, 129], 1, 2)])
def test_block_banded_qr_panel_kernels(handle, rows, l, u, panel, lookahead):
    """Every panel-step kernel of qr! (register-resident slabs of <= 512 / <= 1024 rows, shared-memory slab; option
    qr.panel = 1 forces the latter), with and without the look-ahead split of the trailing update (qr.lookahead), gives
    LAPACK's factors and tau (test/test_blockskylineqr.jl:19-20)."""
    rng = np.random.default_rng(74)
    m = sum(rows)
    A = B.BlockSkylineMatrix(F.sky_from_dense(rng.standard_normal((m, m)) + 3.0 * np.eye(m), rows, rows, l, u), rows, rows, l, u)
    Aw = A.with_bandwidths(l, l + u)
    handle.set_option("qr.panel", panel)
    handle.set_option("qr.lookahead", lookahead)
    try:
        Fd = B.qr_(Aw.to_device(handle))
        got, tau = Fd.factors.data.to_host(), Fd.tau.to_host()
    finally:
        handle.set_option("qr.panel", 0)
        handle.set_option("qr.lookahead", 0)
    h, t = np.linalg.qr(A.to_dense(), mode="raw")
    fac = F.sky_to_dense(got, rows, rows, l, l + u)
    assert np.abs(fac - h.T).max() <= 1e-12 * np.abs(h).max()
    assert np.abs(tau - t).max() <= 1e-12
    b = rng.standard_normal(m)
    x = B.ldiv(Fd, handle.to_device(b)).to_host()
    assert relerr(A.to_dense() @ x, b) <= tol(np.float64, max(rows) * (2 * l + u + 1)) * max(1.0, np.linalg.cond(A.to_dense()))


QL_GPU_CASES = [
    (list(range(1, 6)), 2, 1),                  # test/test_blockskylineqr.jl:81-100 "Square QL"
    ([3, 2, 1, 3], 1, 1),                       # :125-140 "Off-diagonals"
    ([4] * 6, 1, 1),
    ([100, 70, 129, 33], 2, 1),
    ([64] * 10, 1, 2),
    ([7], 0, 0),
    ([3, 2, 0, 4], 2, 1),                       # zero-size block
]


@pytest.mark.parametrize("rows,l,u", QL_GPU_CASES)
def test_block_banded_ql_on_device(handle, rows, l, u):
    """ql!(A), lmul!(F.Q', B) and F \\ B for the QL factorisation (test/test_blockskylineqr.jl:81-100): factors and
    tau against the restated geql2 loop and against LAPACK through the reversal identity; solves against dense."""
    rng = np.random.default_rng(91)
    m = sum(rows)
    A = B.BlockSkylineMatrix(F.sky_from_dense(rng.standard_normal((m, m)) + 3.0 * np.eye(m), rows, rows, l, u), rows, rows, l, u)
    Aw = A.with_bandwidths(l + u, u)                          # `_ql`, src/blockskylineqr.jl:76
    Ad = A.to_dense()
    ref = Aw.data.copy()
    tau_ref = O.sky_ql(rows, rows, l + u, u, ref)
    Fd = B.ql_(Aw.to_device(handle))
    got, tau = Fd.factors.data.to_host(), Fd.tau.to_host()
    bound = 64 * 2.3e-16 * (max(rows) * (u + 1))
    assert np.abs(got - ref).max() <= bound * np.abs(ref).max()
    assert np.abs(tau - tau_ref).max() <= bound
    h, t = np.linalg.qr(Ad[::-1, ::-1], mode="raw")            # geqlf(A) is geqrf of the reversed matrix, reversed
    fac = F.sky_to_dense(got, rows, rows, l + u, u)
    assert np.abs(fac - h.T[::-1, ::-1]).max() <= 1e-12 * np.abs(h).max()
    for nrhs in (1, 7):
        B0 = np.asfortranarray(rng.standard_normal((m, nrhs))) if nrhs > 1 else rng.standard_normal(m)
        # host-computed factors through the device solve, and Q'B alone
        Fh = B.BlockBandedQL(B.BlockSkylineMatrix(ref, rows, rows, l + u, u), tau_ref).to_device(handle)
        dQ = handle.to_device(B0)
        B.lmul_adj_(Fh, dQ)
        rq = np.asfortranarray(B0.copy())
        O.sky_ql_lmul_adj(rows, l + u, u, ref, tau_ref, rq)
        nnz = max(rows) * (l + u + 1)
        assert relerr(dQ.to_host(), rq) <= tol(np.float64, nnz)
        cond = max(1.0, np.linalg.cond(Ad))
        for Fx in (Fh, Fd):
            x = B.ldiv(Fx, handle.to_device(B0)).to_host()
            assert relerr(Ad @ x, B0) <= tol(np.float64, nnz) * cond
        rs = np.asfortranarray(B0.copy())
        O.sky_ql_ldiv(rows, l + u, u, ref, tau_ref, rs)
        assert relerr(x, rs) <= tol(np.float64, nnz) * cond


@pytest.mark.parametrize("rows,l,u", [([4] * 6, 1, 1), ([3, 5, 2, 6, 4], 2, 1), ([40, 33, 64, 17], 1, 2)])
def test_block_banded_qr_ql_float32(handle, rows, l, u):
    """The factorisations and their solves for Float32 (src/blockskylineqr.jl:17-67 is generic in T): qr! / ql!, lmul!(Q', B) and
    F \\ B through bbm_sky_q{r,l}_*_f32 against the Float32 restatement of the reference's loops, Float32 tolerance."""
    rng = np.random.default_rng(93)
    m = sum(rows)
    f32 = np.float32
    Ad = (F.sky_to_dense(F.sky_from_dense(rng.standard_normal((m, m)) + 3.0 * np.eye(m), rows, rows, l, u), rows, rows, l, u)).astype(f32)
    nnz = max(rows) * (l + u + 1)
    B0 = np.asfortranarray(rng.standard_normal((m, 5)).astype(f32))
    for kind in ("qr", "ql"):
        lw, uw = (l, l + u) if kind == "qr" else (l + u, u)
        Aw = B.BlockSkylineMatrix(F.sky_from_dense(Ad, rows, rows, lw, uw).astype(f32), rows, rows, lw, uw)
        ref = Aw.data.copy()
        tau_ref = (O.sky_qr if kind == "qr" else O.sky_ql)(rows, rows, lw, uw, ref)
        Fd = (B.qr_ if kind == "qr" else B.ql_)(Aw.to_device(handle))
        got, tau = Fd.factors.data.to_host(), Fd.tau.to_host()
        assert got.dtype == f32 and tau.dtype == f32
        assert np.abs(got - ref).max() <= tol(f32, nnz) * np.abs(ref).max()
        assert np.abs(tau - tau_ref).max() <= tol(f32, nnz)
        dB = handle.to_device(B0)
        B.lmul_adj_(Fd, dB)
        r2 = np.asfortranarray(B0.copy())
        if kind == "qr":
            O.sky_qr_lmul_adj(rows, rows, lw, uw, ref, tau_ref, r2)
        else:
            O.sky_ql_lmul_adj(rows, lw, uw, ref, tau_ref, r2)
        assert relerr(dB.to_host(), r2) <= tol(f32, nnz)
        dX = B.ldiv(Fd, handle.to_device(B0))
        cond = np.linalg.cond(Ad.astype(np.float64))
        assert relerr(Ad.astype(np.float64) @ dX.to_host().astype(np.float64), B0) <= tol(f32, nnz) * max(1.0, cond)


def test_block_banded_ql_wide_is_an_argument_error(handle):
    A = B.BlockSkylineMatrix(np.zeros(F.sky_numentries([3, 3], [3, 3, 3], 2, 1)), [3, 3], [3, 3, 3], 2, 1).to_device(handle)
    with pytest.raises(ValueError):            # test/test_blockskylineqr.jl:121-123 "Wide QL"
        B.ql_(A)
