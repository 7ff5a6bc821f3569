"""The C oracle (built-in loops and OpenBLAS-backed) against dense long-double arithmetic --
the oracle pattern of the reference's own tests ("structured ~= dense")."""
import numpy as np
import pytest

from oracle import cpu as O
from oracle import formats as F
from tests.helpers import BBB_CASES, dense_mul_ld, random_bbb, random_sky, relerr, tol


@pytest.fixture(params=["builtin", "openblas"])
def blas(request):
    if request.param == "openblas":
        if not O.use_openblas(threads=1):
            pytest.skip("OpenBLAS not found")
    else:
        O.use_builtin()
    yield request.param
    O.use_builtin()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", BBB_CASES, ids=lambda c: f"r{len(c[0])}c{len(c[1])}_l{c[2]}u{c[3]}_lam{c[4]}mu{c[5]}")
def test_bbb_mul_vs_dense(case, dtype, blas):
    rows, cols, l, u, lam, mu = case
    rng = np.random.default_rng(1234)
    data = random_bbb(rng, rows, cols, l, u, lam, mu, dtype)
    A = F.bbb_to_dense(data, rows, cols, l, u, lam, mu)
    assert not np.isnan(A).any()
    m, n = A.shape
    for nrhs, alpha, beta in [(None, 1.0, 0.0), (None, 2.5, -0.5), (3, 1.0, 1.0)]:
        shape_x = (n,) if nrhs is None else (n, nrhs)
        shape_y = (m,) if nrhs is None else (m, nrhs)
        X = np.asfortranarray(rng.standard_normal(shape_x).astype(dtype))
        Y0 = np.asfortranarray(rng.standard_normal(shape_y).astype(dtype))
        Y = Y0.copy(order="F")
        if beta == 0.0:
            Y[...] = np.nan  # beta == 0 must overwrite, never read (test/test_blockskyline.jl:34-39)
        O.bbb_mul(rows, cols, l, u, lam, mu, alpha, data, X, beta, Y)
        ref = alpha * dense_mul_ld(A, X) + (beta * Y0.astype(np.longdouble) if beta != 0 else 0)
        assert relerr(Y, ref) <= tol(dtype, F.bbb_nnz_per_row_max(l, u, lam, mu) + 1)


def test_bbb_laplacian_is_the_five_point_stencil(blas):
    n = 24
    data = F.laplacian_bbb_data(n)
    x = np.random.default_rng(0).standard_normal(n * n)
    y = np.full(n * n, np.nan)
    O.bbb_mul([n] * n, [n] * n, 1, 1, 1, 1, 1.0, data, x, 0.0, y)
    assert relerr(y, F.laplacian_apply(x, n)) <= tol(np.float64, 9)


def test_bbb_dimension_mismatch():
    data = np.zeros(F.bbb_data_shape([2, 2], 1, 1, 1, 1), order="F")
    with pytest.raises(ValueError):
        O.bbb_mul([2, 2], [2, 2], 1, 1, 1, 1, 1.0, data, np.zeros(5), 0.0, np.zeros(4))


SKY_CASES = [
    (list(range(1, 5)), list(range(1, 5)), 1, 1),                       # test/test_linalg.jl:35-57
    ([3, 4, 5], [2, 3, 4, 3], 0, 1),                                    # test/test_triblockbanded.jl:101-107
    ([3, 1, 2, 1, 2, 1, 2, 1, 2, 1, 3], [3, 1, 2, 1, 2, 1, 2, 1, 2, 1, 3],
     [1, 2, 1, 2, 1, 2, 1, 2, 1, 2, 1], [1, 2, 1, 2, 1, 2, 1, 2, 1, 2, 1]),  # test/test_blockskyline.jl:43-44
    ([9, 4, 1, 10, 6], [9, 4, 1, 10, 6], [-2, 2, 0, 2, -1], [-1, 2, 1, 0, -1]),  # :75-76 ragged negative
    ([16] * 8, [16] * 8, 2, 2),                                         # cfg4-like
    ([40, 7, 130, 3], [20, 50, 9, 64], 1, 2),
    ([], [], 0, 0),
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", SKY_CASES, ids=lambda c: f"r{len(c[0])}c{len(c[1])}")
def test_sky_mul_vs_dense(case, dtype, blas):
    rows, cols, l, u = case
    rng = np.random.default_rng(7)
    data = random_sky(rng, rows, cols, l, u, dtype)
    A = F.sky_to_dense(data, rows, cols, l, u)
    m, n = A.shape
    for nrhs, alpha, beta in [(None, 1.0, 0.0), (4, -1.5, 0.75)]:
        X = np.asfortranarray(rng.standard_normal((n,) if nrhs is None else (n, nrhs)).astype(dtype))
        Y0 = np.asfortranarray(rng.standard_normal((m,) if nrhs is None else (m, nrhs)).astype(dtype))
        Y = Y0.copy(order="F")
        if beta == 0.0:
            Y[...] = np.nan
        O.sky_mul(rows, cols, l, u, alpha, data, X, beta, Y)
        ref = alpha * dense_mul_ld(A, X) + (beta * Y0.astype(np.longdouble) if beta != 0 else 0)
        assert relerr(Y, ref) <= tol(dtype, n + 1)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sky_matmul_vs_dense(dtype, blas):
    rng = np.random.default_rng(11)
    for rows, l, u, const in [([16] * 8, 2, 2, True), ([3, 1, 2, 1, 2, 1, 2], [1, 2, 1, 2, 1, 2, 1], [1, 2, 1, 2, 1, 2, 1], False),
                              ([5, 9, 2, 7], 1, 0, True)]:
        Ad = random_sky(rng, rows, rows, l, u, dtype)
        Bd = random_sky(rng, rows, rows, l, u, dtype)
        lv = np.full(len(rows), l) if np.isscalar(l) else np.array(l)
        uv = np.full(len(rows), u) if np.isscalar(u) else np.array(u)
        Cl, Cu = F.add_bandwidths(lv, uv, lv, uv, constant=const)
        n = F.sky_numentries(rows, rows, Cl, Cu)
        C0 = rng.standard_normal(n).astype(dtype)
        for alpha, beta in [(1.0, 0.0), (2.0, -1.0)]:
            Cd = C0.copy()
            if beta == 0:
                Cd[:] = np.nan
            O.sky_matmul((rows, rows, l, u), (rows, rows, l, u), (rows, rows, Cl, Cu), alpha, Ad, Bd, beta, Cd)
            A = F.sky_to_dense(Ad, rows, rows, l, u).astype(np.longdouble)
            B = F.sky_to_dense(Bd, rows, rows, l, u).astype(np.longdouble)
            ref = alpha * (A @ B) + (beta * F.sky_to_dense(C0, rows, rows, Cl, Cu).astype(np.longdouble) if beta else 0)
            got = F.sky_to_dense(Cd, rows, rows, Cl, Cu)
            assert relerr(got, ref) <= tol(dtype, sum(rows))
            # nothing outside the product's band was needed: the stored part holds the full product
            assert np.count_nonzero(np.isnan(Cd)) == 0


def test_sky_matmul_band_violation_and_dims():
    rows = [2, 2, 2]
    Ad = np.ones(F.sky_numentries(rows, rows, 1, 1))
    Cd = np.zeros(F.sky_numentries(rows, rows, 1, 1))
    with pytest.raises(IndexError):  # C (1,1) too narrow for a (2,2) product
        O.sky_matmul((rows, rows, 1, 1), (rows, rows, 1, 1), (rows, rows, 1, 1), 1.0, Ad, Ad, 0.0, Cd)
    with pytest.raises(ValueError):
        O.sky_matmul((rows, rows, 1, 1), ([2, 2], [2, 2], 1, 1), (rows, [2, 2], 2, 2), 1.0, Ad, Ad, 0.0, Cd)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("uplo,unit", [("U", False), ("U", True), ("L", False), ("L", True)])
def test_sky_trsm_vs_dense(dtype, uplo, unit, blas):
    rng = np.random.default_rng(5)
    for rows, l, u in [(list(range(1, 5)), 1, 1), ([8] * 6, 0, 3), ([8] * 6, 2, 0), ([3, 7, 1, 12, 5], 2, 1), ([6], 0, 0)]:
        m = sum(rows)
        data = random_sky(rng, rows, rows, l, u, dtype, scale=1.0 / np.sqrt(max(m, 1)))
        A = F.sky_to_dense(data, rows, rows, l, u)
        A += 10 * np.eye(m, dtype=dtype)  # well conditioned, cf. test/test_triblockbanded.jl:83
        data = F.sky_from_dense(A, rows, rows, l, u)
        T = np.triu(A) if uplo == "U" else np.tril(A)
        if unit:
            np.fill_diagonal(T, 1)
        for nrhs in (None, 3):
            B0 = np.asfortranarray(rng.standard_normal((m,) if nrhs is None else (m, nrhs)).astype(dtype))
            B = B0.copy(order="F")
            out = O.sky_trsm(rows, l, u, uplo, unit, data, B)
            assert out is B  # in place (test/test_linalg.jl:65,75)
            ref = np.linalg.solve(T.astype(np.float64), B0.astype(np.float64))
            assert relerr(B, ref) <= tol(dtype, m)


def test_sky_trsm_needs_diagonal_blocks():
    with pytest.raises(IndexError):
        O.sky_trsm([2, 2, 2], -1, 2, "U", False, np.ones(F.sky_numentries([2, 2, 2], [2, 2, 2], -1, 2)), np.ones(6))


@pytest.mark.parametrize("case", [
    (list(range(1, 11)), (1, 1), (1, 1), (1, 1), (1, 1)),      # test/test_linalg.jl:91-104
    (list(range(1, 6)), (-1, 1), (-1, 1), (1, -1), (1, -1)),   # test/test_linalg.jl:160-171
    ([7, 3, 12, 5, 9], (2, 1), (1, 2), (0, 2), (2, 0)),
])
def test_oracle_bbb_matmul_vs_dense(case):
    """Matrix(A*B) ~= Matrix(A)*Matrix(B) for the block-triple gbmm restatement."""
    import numpy as np
    from oracle import cpu as O
    from oracle import formats as F
    from tests.helpers import random_bbb, relerr, tol
    rows, luA, lmA, luB, lmB = case
    rng = np.random.default_rng(5)
    adata = random_bbb(rng, rows, rows, *luA, *lmA, poison=False)
    bdata = random_bbb(rng, rows, rows, *luB, *lmB, poison=False)
    luC, lmC = (luA[0] + luB[0], luA[1] + luB[1]), (lmA[0] + lmB[0], lmA[1] + lmB[1])
    c = np.zeros(F.bbb_data_shape(rows, *luC, *lmC), order="F")
    O.bbb_matmul((rows, rows, *luA, *lmA), (rows, rows, *luB, *lmB), (rows, rows, *luC, *lmC), 1.0, adata, bdata, 0.0, c)
    dense = F.bbb_to_dense(adata, rows, rows, *luA, *lmA).astype(np.longdouble) @ F.bbb_to_dense(bdata, rows, rows, *luB, *lmB).astype(np.longdouble)
    assert relerr(F.bbb_to_dense(c, rows, rows, *luC, *lmC), dense) <= tol(np.float64, 16)


@pytest.mark.parametrize("uplo,unit", [("U", False), ("U", True), ("L", False), ("L", True)])
def test_oracle_bbb_trsm_vs_dense(uplo, unit):
    """block substitution on the band storage == dense triangular solve (test/test_triblockbanded.jl:80-99)"""
    for rows, lu, lm in [(list(range(1, 11)), (1, 1), (1, 1)), ([7, 0, 3, 12, 1], (2, 1), (2, 3)), ([20] * 6, (0, 0), (1, 2))]:
        rng = np.random.default_rng(5)
        shape = F.bbb_data_shape(rows, *lu, *lm)
        data = np.asfortranarray(rng.standard_normal(shape) * 0.05)
        w = lm[0] + lm[1] + 1
        data[lu[1] * w + lm[1], :] = 2.0 + rng.random(shape[1])
        Dn = F.bbb_to_dense(data, rows, rows, *lu, *lm)
        Tm = np.triu(Dn) if uplo == "U" else np.tril(Dn)
        if unit:
            np.fill_diagonal(Tm, 1.0)
        Bm = np.asfortranarray(rng.standard_normal((sum(rows), 3)))
        ref = np.linalg.solve(Tm, Bm)
        got = Bm.copy(order="F")
        O.bbb_trsm(rows, *lu, *lm, uplo, unit, data, got)
        assert np.linalg.norm(got - ref) <= 1e-13 * np.linalg.norm(ref)


QR_CASES = [
    (list(range(1, 6)), list(range(1, 6)), 2, 1),     # test/test_blockskylineqr.jl:13-35 "Square QR"
    (list(range(1, 7)), list(range(1, 6)), 2, 1),     # :37-58 "Thin QR"
    ([4] * 6, [4] * 6, 1, 1),
    ([3, 5, 2, 4], [3, 5, 2, 4], 1, 2),
    ([7], [7], 0, 0),
]


@pytest.mark.parametrize("rows,cols,l,u", QR_CASES)
def test_oracle_block_banded_qr_matches_lapack(rows, cols, l, u):
    """test/test_blockskylineqr.jl:19-20,43-44: F.factors == qr(Matrix(A)).factors and F.tau == the dense
    unblocked Householder taus; lmul!(Q', b) and F \\ b against dense LAPACK (:29-34)."""
    rng = np.random.default_rng(17)
    m, n = sum(rows), sum(cols)
    Ad = F.sky_to_dense(F.sky_from_dense(rng.standard_normal((m, n)), rows, cols, l, u), rows, cols, l, u)
    data = F.sky_from_dense(Ad, rows, cols, l, l + u)            # `_qr` widens to (l, l+u), src/blockskylineqr.jl:75
    tau = O.sky_qr(rows, cols, l, l + u, data)
    h, t = np.linalg.qr(Ad, mode="raw")                          # LAPACK geqrf: same reflector convention
    fac = F.sky_to_dense(data, rows, cols, l, l + u)
    assert np.abs(fac - h.T).max() <= 1e-13 * np.abs(h).max()    # R and the Householder vectors, nothing outside the band
    assert np.abs(tau - t[:len(tau)]).max() <= 1e-13
    B0 = rng.standard_normal((m, 3))
    QtB = np.asfortranarray(B0.copy())
    O.sky_qr_lmul_adj(rows, cols, l, l + u, data, tau, QtB)
    Q, _ = np.linalg.qr(Ad, mode="complete")
    assert np.abs(QtB - Q.T @ B0).max() <= 1e-13 * np.abs(B0).max()
    if rows == cols:
        X = np.asfortranarray(B0.copy())
        O.sky_qr_ldiv(rows, l, l + u, data, tau, X)
        ref = np.linalg.solve(Ad, B0)
        assert np.linalg.norm(X - ref) <= 1e-10 * np.linalg.norm(ref)


def test_oracle_sky_mul_t_vs_dense():
    rng = np.random.default_rng(23)
    for rows, cols, l, u in [([5, 3, 7, 2, 6], [4, 6, 2, 5, 3, 7], 1, 2), ([3] * 7, [3] * 7, [2, 1, 2, 3, 1, 0, 0], [0, 1, 1, 2, 3, 1, 2])]:
        data = random_sky(rng, rows, cols, l, u)
        Dn = F.sky_to_dense(data, rows, cols, l, u)
        X = np.asfortranarray(rng.standard_normal((sum(rows), 3)))
        Y = np.asfortranarray(rng.standard_normal((sum(cols), 3)))
        ref = 2.0 * Dn.T @ X - 0.5 * Y
        O.sky_mul_t(rows, cols, l, u, 2.0, data, X, -0.5, Y)
        assert relerr(Y, ref) <= 1e-14


def test_oracle_sky_rmul_vs_dense():
    """`X*A` (test/test_linalg.jl:56-57): the restated block loop against dense `X @ Matrix(A)`, incl. beta = 0 with a NaN destination,
    rectangular and ragged structures and an empty block column."""
    rng = np.random.default_rng(29)
    for rows, cols, l, u in [([5, 3, 7, 2, 6], [4, 6, 2, 5, 3, 7], 1, 2), ([3] * 7, [3] * 7, [2, 1, 2, 3, 1, 0, 0], [0, 1, 1, 2, 3, 1, 2]),
                             (list(range(1, 5)), list(range(1, 5)), 1, 1), ([2, 2, 2], [2, 2, 2], [0, -1, 0], [0, -1, 0])]:
        data = random_sky(rng, rows, cols, l, u)
        Dn = F.sky_to_dense(data, rows, cols, l, u)
        X = np.asfortranarray(rng.standard_normal((4, sum(rows))))
        Y0 = np.asfortranarray(rng.standard_normal((4, sum(cols))))
        Y = Y0.copy(order="F")
        O.sky_rmul(rows, cols, l, u, 2.0, X, data, -0.5, Y)
        assert relerr(Y, 2.0 * X @ Dn - 0.5 * Y0) <= 1e-14
        Y = np.full_like(Y0, np.nan)
        O.sky_rmul(rows, cols, l, u, 1.0, X, data, 0.0, Y)
        assert relerr(Y, X @ Dn) <= 1e-14 if np.linalg.norm(X @ Dn) else not np.isnan(Y).any()


def test_oracle_bbb_sparse_matches_dense():
    import scipy.sparse as sp
    rng = np.random.default_rng(29)
    for rows, cols, lu, lm in [(list(range(1, 8)), list(range(1, 8)), (1, 1), (1, 1)), ([5, 0, 3, 7], [4, 2, 0, 6, 3], (1, 2), (2, 1))]:
        data = random_bbb(rng, rows, cols, *lu, *lm, poison=False)
        cp, rv, nz = F.bbb_sparse(data, rows, cols, *lu, *lm)
        S = sp.csc_matrix((nz, rv, cp), shape=(sum(rows), sum(cols)))
        assert S.has_sorted_indices or np.all(np.diff(rv[cp[0]:cp[1]]) > 0)
        assert np.array_equal(S.toarray(), F.bbb_to_dense(data, rows, cols, *lu, *lm))
        assert cp[-1] == F.bbb_valid_mask(rows, cols, *lu, *lm).sum()


@pytest.mark.parametrize("rows,l,u", [(list(range(1, 6)), 2, 1), ([3, 2, 1, 3], 1, 1), ([4] * 6, 1, 1), ([3, 5, 2, 4], 1, 2)])
def test_oracle_block_banded_ql_matches_lapack(rows, l, u):
    """test/test_blockskylineqr.jl:86-88: F.factors == ql(Matrix(A)).factors, F.tau == the dense taus.  scipy has no
    geqlf binding, so LAPACK enters through the identity geqlf(A) = reverse(geqrf(reverse(A)))."""
    rng = np.random.default_rng(19)
    m = sum(rows)
    Ad = F.sky_to_dense(F.sky_from_dense(rng.standard_normal((m, m)) + 3.0 * np.eye(m), rows, rows, l, u), rows, rows, l, u)
    data = F.sky_from_dense(Ad, rows, rows, l + u, u)
    tau = O.sky_ql(rows, rows, l + u, u, data)
    h, t = np.linalg.qr(Ad[::-1, ::-1], mode="raw")
    fac = F.sky_to_dense(data, rows, rows, l + u, u)
    assert np.abs(fac - h.T[::-1, ::-1]).max() <= 1e-13 * np.abs(h).max()
    assert np.abs(tau - t[::-1]).max() <= 1e-13
    Lm = np.tril(fac)
    assert np.abs(Lm.T @ Lm - Ad.T @ Ad).max() <= 1e-12 * np.abs(Ad.T @ Ad).max()
    B0 = rng.standard_normal((m, 3))
    X = np.asfortranarray(B0.copy())
    O.sky_ql_ldiv(rows, l + u, u, data, tau, X)
    ref = np.linalg.solve(Ad, B0)
    assert np.linalg.norm(X - ref) <= 1e-10 * np.linalg.norm(ref)
    with pytest.raises(ValueError):
        O.sky_ql([3, 3], [3, 3, 3], 2, 1, np.zeros(F.sky_numentries([3, 3], [3, 3, 3], 2, 1)))
