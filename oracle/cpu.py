"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the C oracle (oracle/bbm_oracle.c).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this.  See oracle/bbm_oracle.c for the reference
file:line each routine follows.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

i64 = C.c_int64
_p = C.c_void_p


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libbbm_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("bbm_oracle.c", "bbm_oracle_impl.inc")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libbbm_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_sky_numentries.restype = i64
    return _LIB


def find_openblas() -> str | None:
    """The image's OpenBLAS 0.3.30 (LP64, ``scipy_cblas_*`` symbols) bundled with scipy."""
    try:
        import scipy
    except Exception:  # pragma: no cover
        return None
    root = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    hits = sorted(glob.glob(os.path.join(root, "libscipy_openblas-*.so")))
    return hits[0] if hits else None


def use_openblas(threads: int | None = None) -> bool:
    """Route the per-block operations through OpenBLAS (what the reference's CPU path calls)."""
    path = find_openblas()
    if path is None:
        return False
    rc = lib().oracle_set_blas(path.encode())
    if rc != 0:
        return False
    if threads is not None:
        lib().oracle_set_blas_threads(int(threads))
    return True


def use_builtin() -> None:
    lib().oracle_set_blas(None)


def _ia(v):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.int64).reshape(-1))
    return a, a.ctypes.data_as(_p)


def _vecbw(b, Nc):
    b = np.asarray(b, dtype=np.int64)
    if b.ndim == 0:
        b = np.full(Nc, int(b), dtype=np.int64)
    return np.ascontiguousarray(b.reshape(-1))


def _suffix(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64", C.c_double
    if dtype == np.float32:
        return "f32", C.c_float
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def _check(rc, what):
    if rc == 2:
        raise ValueError(f"{what}: DimensionMismatch")
    if rc == 3:
        raise IndexError(f"{what}: BandError")
    if rc != 0:
        raise RuntimeError(f"{what}: oracle error {rc}")


def _mat(X, n):
    """View X (vector or F-ordered matrix) as (ptr, ld, nrhs)."""
    if X.ndim == 1:
        assert X.flags.c_contiguous
        return X.ctypes.data_as(_p), max(1, X.shape[0]), 1
    assert X.flags.f_contiguous, "matrices must be column-major"
    return X.ctypes.data_as(_p), max(1, X.shape[0]), X.shape[1]


def bbb_mul(rows, cols, l, u, lam, mu, alpha, data, X, beta, Y):
    """Y <- alpha*A*X + beta*Y in place; A given by its band storage ``data`` (P x n, F order)."""
    suf, ct = _suffix(data.dtype)
    r, rp = _ia(rows)
    c, cp = _ia(cols)
    assert X.dtype == data.dtype and Y.dtype == data.dtype
    if X.shape[0] != int(c.sum()) or Y.shape[0] != int(r.sum()):
        raise ValueError("DimensionMismatch")
    assert data.flags.f_contiguous or data.size == 0
    xp, ldx, nrhs = _mat(X, 0)
    yp, ldy, nrhs2 = _mat(Y, 0)
    assert nrhs == nrhs2
    f = getattr(lib(), f"oracle_bbb_mul_{suf}")
    rc = f(i64(len(r)), i64(len(c)), rp, cp, i64(l), i64(u), i64(lam), i64(mu), ct(alpha),
           data.ctypes.data_as(_p), xp, i64(ldx), i64(nrhs), ct(beta), yp, i64(ldy))
    _check(rc, "bbb_mul")
    return Y


def bbb_mul_t(rows, cols, l, u, lam, mu, alpha, data, X, beta, Y):
    """Y <- alpha*A^T*X + beta*Y in place (X has sum(rows) rows, Y has sum(cols) rows)."""
    suf, ct = _suffix(data.dtype)
    r, rp = _ia(rows)
    c, cp = _ia(cols)
    if X.shape[0] != int(r.sum()) or Y.shape[0] != int(c.sum()):
        raise ValueError("DimensionMismatch")
    xp, ldx, nrhs = _mat(X, 0)
    yp, ldy, _ = _mat(Y, 0)
    f = getattr(lib(), f"oracle_bbb_mul_t_{suf}")
    rc = f(i64(len(r)), i64(len(c)), rp, cp, i64(l), i64(u), i64(lam), i64(mu), ct(alpha),
           data.ctypes.data_as(_p), xp, i64(ldx), i64(nrhs), ct(beta), yp, i64(ldy))
    _check(rc, "bbb_mul_t")
    return Y


def sky_mul_t(rows, cols, l, u, alpha, data, X, beta, Y):
    """Y <- alpha*A^T*X + beta*Y in place for a block-skyline A (X has sum(rows) rows, Y has sum(cols) rows)."""
    suf, ct = _suffix(data.dtype)
    r, rp = _ia(rows)
    c, cp = _ia(cols)
    lv, uv = _vecbw(l, len(c)), _vecbw(u, len(c))
    if X.shape[0] != int(r.sum()) or Y.shape[0] != int(c.sum()):
        raise ValueError("DimensionMismatch")
    xp, ldx, nrhs = _mat(X, 0)
    yp, ldy, _ = _mat(Y, 0)
    f = getattr(lib(), f"oracle_sky_mul_t_{suf}")
    rc = f(i64(len(r)), i64(len(c)), rp, cp, lv.ctypes.data_as(_p), uv.ctypes.data_as(_p), ct(alpha),
           data.ctypes.data_as(_p), xp, i64(ldx), i64(nrhs), ct(beta), yp, i64(ldy))
    _check(rc, "sky_mul_t")
    return Y


def sky_mul(rows, cols, l, u, alpha, data, X, beta, Y):
    suf, ct = _suffix(data.dtype)
    r, rp = _ia(rows)
    c, cp = _ia(cols)
    lv, uv = _vecbw(l, len(c)), _vecbw(u, len(c))
    if X.shape[0] != int(c.sum()) or Y.shape[0] != int(r.sum()):
        raise ValueError("DimensionMismatch")
    xp, ldx, nrhs = _mat(X, 0)
    yp, ldy, _ = _mat(Y, 0)
    f = getattr(lib(), f"oracle_sky_mul_{suf}")
    rc = f(i64(len(r)), i64(len(c)), rp, cp, lv.ctypes.data_as(_p), uv.ctypes.data_as(_p), ct(alpha),
           data.ctypes.data_as(_p), xp, i64(ldx), i64(nrhs), ct(beta), yp, i64(ldy))
    _check(rc, "sky_mul")
    return Y


def sky_matmul(A, B, Cs, alpha, Adata, Bdata, beta, Cdata):
    """A, B, Cs are (rows, cols, l, u) structure tuples; Cdata is updated in place."""
    suf, ct = _suffix(Adata.dtype)
    args = []
    keep = []
    for rows, cols, l, u in (A, B, Cs):
        r, rp = _ia(rows)
        c, cp = _ia(cols)
        lv, uv = _vecbw(l, len(c)), _vecbw(u, len(c))
        keep += [r, c, lv, uv]
        args += [i64(len(r)), i64(len(c)), rp, cp, lv.ctypes.data_as(_p), uv.ctypes.data_as(_p)]
    f = getattr(lib(), f"oracle_sky_matmul_{suf}")
    rc = f(*args, ct(alpha), Adata.ctypes.data_as(_p), Bdata.ctypes.data_as(_p), ct(beta), Cdata.ctypes.data_as(_p))
    _check(rc, "sky_matmul")
    return Cdata


def bbb_matmul(A, B, Cs, alpha, Adata, Bdata, beta, Cdata):
    """A, B, Cs are (rows, cols, l, u, lam, mu) structure tuples; Cdata (band storage, F order) is updated in place."""
    suf, ct = _suffix(Adata.dtype)
    args, keep = [], []
    for rows, cols, l, u, lam, mu in (A, B, Cs):
        r, rp = _ia(rows)
        c, cp = _ia(cols)
        keep += [r, c]
        args += [i64(len(r)), i64(len(c)), rp, cp, i64(l), i64(u), i64(lam), i64(mu)]
    f = getattr(lib(), f"oracle_bbb_matmul_{suf}")
    rc = f(*args, ct(alpha), Adata.ctypes.data_as(_p), Bdata.ctypes.data_as(_p), ct(beta), Cdata.ctypes.data_as(_p))
    _check(rc, "bbb_matmul")
    return Cdata


def sky_trsm(rows, l, u, uplo, unit, data, B):
    """B <- T^-1 B in place, T = triangular part ('U'/'L', unit or not) of the skyline matrix."""
    suf, _ = _suffix(data.dtype)
    r, rp = _ia(rows)
    lv, uv = _vecbw(l, len(r)), _vecbw(u, len(r))
    if B.shape[0] != int(r.sum()):
        raise ValueError("DimensionMismatch")
    bp, ldb, nrhs = _mat(B, 0)
    f = getattr(lib(), f"oracle_sky_trsm_{suf}")
    rc = f(i64(len(r)), rp, lv.ctypes.data_as(_p), uv.ctypes.data_as(_p), C.c_int(1 if uplo.upper() == "U" else 0),
           C.c_int(1 if unit else 0), data.ctypes.data_as(_p), bp, i64(ldb), i64(nrhs))
    _check(rc, "sky_trsm")
    return B


def sky_qr(rows, cols, l, u, data):
    """qr!(A) in place on the skyline storage of the FACTORS structure (block bandwidths (l, u) with u = l_A+u_A
    already, src/blockskylineqr.jl:29-43,75); returns tau (blocked like the shorter axis)."""
    suf, _ = _suffix(data.dtype)
    r, rp = _ia(rows)
    c, cp = _ia(cols)
    lv, uv = _vecbw(l, len(c)), _vecbw(u, len(c))
    tau = np.zeros(int(c.sum()) if len(r) >= len(c) else int(r.sum()), dtype=data.dtype)
    f = getattr(lib(), f"oracle_sky_qr_{suf}")
    rc = f(i64(len(r)), i64(len(c)), rp, cp, lv.ctypes.data_as(_p), uv.ctypes.data_as(_p), data.ctypes.data_as(_p), tau.ctypes.data_as(_p))
    _check(rc, "sky_qr")
    return tau


def sky_qr_lmul_adj(rows, cols, l, u, data, tau, B):
    """B <- Q' B in place for the packed Q of sky_qr (lmul!(F.Q', B), src/blockskylineqr.jl:77-127)."""
    suf, _ = _suffix(data.dtype)
    r, rp = _ia(rows)
    c, cp = _ia(cols)
    lv, uv = _vecbw(l, len(c)), _vecbw(u, len(c))
    if B.shape[0] != int(r.sum()):
        raise ValueError("DimensionMismatch")
    bp, ldb, nrhs = _mat(B, 0)
    f = getattr(lib(), f"oracle_sky_qr_lmul_adj_{suf}")
    rc = f(i64(len(r)), i64(len(c)), rp, cp, lv.ctypes.data_as(_p), uv.ctypes.data_as(_p), data.ctypes.data_as(_p),
           tau.ctypes.data_as(_p), bp, i64(ldb), i64(nrhs))
    _check(rc, "sky_qr_lmul_adj")
    return B


def sky_qr_ldiv(rows, l, u, data, tau, B):
    """B <- F \\ B in place for a square block-banded QR (ldiv!(F, B), src/blockskylineqr.jl:131-145)."""
    suf, _ = _suffix(data.dtype)
    r, rp = _ia(rows)
    lv, uv = _vecbw(l, len(r)), _vecbw(u, len(r))
    if B.shape[0] != int(r.sum()):
        raise ValueError("DimensionMismatch")
    bp, ldb, nrhs = _mat(B, 0)
    f = getattr(lib(), f"oracle_sky_qr_ldiv_{suf}")
    rc = f(i64(len(r)), rp, lv.ctypes.data_as(_p), uv.ctypes.data_as(_p), data.ctypes.data_as(_p), tau.ctypes.data_as(_p), bp, i64(ldb), i64(nrhs))
    _check(rc, "sky_qr_ldiv")
    return B


def sky_ql(rows, cols, l, u, data):
    """ql!(A) in place on the skyline storage of the FACTORS structure (block bandwidths (l, u) with l = l_A+u_A
    already, src/blockskylineqr.jl:51-72,76); returns tau."""
    suf, _ = _suffix(data.dtype)
    r, rp = _ia(rows)
    c, cp = _ia(cols)
    lv, uv = _vecbw(l, len(c)), _vecbw(u, len(c))
    tau = np.zeros(int(c.sum()), dtype=data.dtype)
    f = getattr(lib(), f"oracle_sky_ql_{suf}")
    rc = f(i64(len(r)), i64(len(c)), rp, cp, lv.ctypes.data_as(_p), uv.ctypes.data_as(_p), data.ctypes.data_as(_p), tau.ctypes.data_as(_p))
    if rc == 1:
        raise ValueError("sky_ql: ArgumentError (wide block-QL / non-uniform bandwidths)")
    _check(rc, "sky_ql")
    return tau


def _sky_ql_call(name, rows, l, u, data, tau, B):
    suf, _ = _suffix(data.dtype)
    r, rp = _ia(rows)
    lv, uv = _vecbw(l, len(r)), _vecbw(u, len(r))
    if B.shape[0] != int(r.sum()):
        raise ValueError("DimensionMismatch")
    bp, ldb, nrhs = _mat(B, 0)
    f = getattr(lib(), f"{name}_{suf}")
    rc = f(i64(len(r)), rp, lv.ctypes.data_as(_p), uv.ctypes.data_as(_p), data.ctypes.data_as(_p), tau.ctypes.data_as(_p), bp, i64(ldb), i64(nrhs))
    _check(rc, name)
    return B


def sky_ql_lmul_adj(rows, l, u, data, tau, B):
    """B <- Q' B in place for the packed Q of a square block-banded QL (src/blockskylineqr.jl:96-111)."""
    return _sky_ql_call("oracle_sky_ql_lmul_adj", rows, l, u, data, tau, B)


def sky_ql_ldiv(rows, l, u, data, tau, B):
    """B <- F \\ B in place for a square block-banded QL (src/blockskylineqr.jl:149-157)."""
    return _sky_ql_call("oracle_sky_ql_ldiv", rows, l, u, data, tau, B)


def bbb_trsm(rows, l, u, lam, mu, uplo, unit, data, B):
    """B <- T^-1 B in place, T = triangular part ('U'/'L', unit or not) of the BandedBlockBandedMatrix."""
    suf, _ = _suffix(data.dtype)
    r, rp = _ia(rows)
    if B.shape[0] != int(r.sum()):
        raise ValueError("DimensionMismatch")
    bp, ldb, nrhs = _mat(B, 0)
    f = getattr(lib(), f"oracle_bbb_trsm_{suf}")
    rc = f(i64(len(r)), rp, i64(l), i64(u), i64(lam), i64(mu), C.c_int(1 if uplo.upper() == "U" else 0),
           C.c_int(1 if unit else 0), data.ctypes.data_as(_p), bp, i64(ldb), i64(nrhs))
    _check(rc, "bbb_trsm")
    return B


def sky_numentries(rows, cols, l, u) -> int:
    r, rp = _ia(rows)
    c, cp = _ia(cols)
    lv, uv = _vecbw(l, len(c)), _vecbw(u, len(c))
    return int(lib().oracle_sky_numentries(i64(len(r)), i64(len(c)), rp, cp, lv.ctypes.data_as(_p), uv.ctypes.data_as(_p)))


def sky_rmul(rows, cols, l, u, alpha, X, data, beta, Y):
    """Y <- alpha*X*A + beta*Y in place for a dense X (p x sum(rows)) and a block-skyline A: `X*A` / `MulAdd(X, A)`
    (test/test_linalg.jl:56-57).  [upstream] BlockArrays' block loop with X as a single block row: for every block column J
    and every K in blockcolsupport(A, J) (src/AbstractBlockBandedMatrix.jl:90-104), Y[:, J] += alpha * X[:, K] * A[K, J]
    (one gemm per block on the (ptr, ld) view of src/BlockSkylineMatrix.jl:459-465); beta == 0 overwrites Y without reading it.
    Plain numpy, per block -- small cases only."""
    from . import formats as F
    r, c = [int(v) for v in rows], [int(v) for v in cols]
    off, S, klo, khi = F.sky_layout(r, c, l, u)
    R = np.concatenate([[0], np.cumsum(r)]).astype(np.int64)
    Cc = np.concatenate([[0], np.cumsum(c)]).astype(np.int64)
    if X.shape[1] != R[-1] or Y.shape[1] != Cc[-1] or X.shape[0] != Y.shape[0]:
        raise ValueError("DimensionMismatch")
    if beta == 0:
        Y[...] = 0
    else:
        Y *= beta
    for J in range(len(c)):
        ld = int(S[J])
        if ld == 0 or c[J] == 0:
            continue
        pan = data[int(off[J]):int(off[J]) + ld * c[J]].reshape((ld, c[J]), order="F")
        for K in range(int(klo[J]), int(khi[J]) + 1):
            r0 = int(R[K] - R[int(klo[J])])
            Y[:, Cc[J]:Cc[J + 1]] += alpha * (X[:, R[K]:R[K + 1]] @ pan[r0:r0 + r[K], :])
    return Y
