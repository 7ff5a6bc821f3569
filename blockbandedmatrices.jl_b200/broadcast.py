"""Device-side structured ``+`` / ``-`` / ``alpha*A + B`` and ``A +- (I | Diagonal | Bidiagonal | Tridiagonal)``.

Host-side mirror of what the Julia extension keeps in Julia for these operations -- the result *structure*
(``similar``: maximum of the bandwidths, src/broadcast.jl:107-131, 391-411; ``copy_accommodating_diagonals``,
src/BlockSkylineMatrix.jl:502-531) -- followed by one or two C-ABI calls that do the arithmetic on the device
(``bbm_bbb_add_*``, ``bbm_sky_add_*``, ``bbm_*_add_diags_*``).  Semantics follow src/broadcast.jl:317-388 and
src/BlockSkylineMatrix.jl:535-638; pinned by test/test_broadcasting.jl:136-216 and test/test_blockskyline.jl:117-161.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .device import DeviceArray
from .matrices import BandedBlockBandedMatrix, BlockSkylineMatrix

_SUF = {np.dtype(np.float64): ("f64", C.c_double), np.dtype(np.float32): ("f32", C.c_float)}


# ---- the structured operands of LinearAlgebra the reference adds to a BlockSkylineMatrix ---------------------
class UniformScaling:
    """``lam*I``."""

    def __init__(self, lam=1.0):
        self.lam = lam


class Diagonal:
    def __init__(self, diag):
        self.diag = diag


class Bidiagonal:
    def __init__(self, dv, ev, uplo: str):
        if uplo not in ("U", "L"):
            raise ValueError("uplo must be 'U' or 'L'")
        self.dv, self.ev, self.uplo = dv, ev, uplo


class Tridiagonal:
    def __init__(self, dl, d, du):
        self.dl, self.d, self.du = dl, d, du


class SymTridiagonal:
    def __init__(self, dv, ev):
        self.dv, self.ev = dv, ev


_SPECIAL = (UniformScaling, Diagonal, Bidiagonal, Tridiagonal, SymTridiagonal)


def _parts(S):
    """(diagonals range, lambda, dl, d, du) of a structured operand."""
    if isinstance(S, UniformScaling):
        return (0, 0), S.lam, None, None, None
    if isinstance(S, Diagonal):
        return (0, 0), 0.0, None, S.diag, None
    if isinstance(S, Bidiagonal):
        return ((0, 1), 0.0, None, S.dv, S.ev) if S.uplo == "U" else ((-1, 0), 0.0, S.ev, S.dv, None)
    if isinstance(S, Tridiagonal):
        return (-1, 1), 0.0, S.dl, S.d, S.du
    if isinstance(S, SymTridiagonal):
        return (-1, 1), 0.0, S.ev, S.dv, S.ev
    raise TypeError(type(S).__name__)


# ---- result structures (host) --------------------------------------------------------------------------------
def _const_bw(A):
    """``blockbandwidths(A)``: the constant pair (maxima for a ragged skyline)."""
    if isinstance(A, BandedBlockBandedMatrix):
        return A.l, A.u
    if len(A.l) == 0:
        return 0, 0
    return int(np.max(A.l)), int(np.max(A.u))


def accommodating_bandwidths(rows, l, u, diagonals):
    """Per-column block bandwidths of ``copy_accommodating_diagonals(A, d0:d1)`` (src/BlockSkylineMatrix.jl:502-531):
    the smallest widening of (l, u) such that the diagonals d0..d1 lie inside stored blocks."""
    rows = np.asarray(rows, dtype=np.int64)
    l, u = np.array(l, dtype=np.int64, copy=True), np.array(u, dtype=np.int64, copy=True)
    d0, d1 = diagonals
    if d0 <= 0 <= d1:
        l, u = np.maximum(l, 0), np.maximum(u, 0)
    n = int(rows.sum())
    if n == 0:
        return l, u
    lasts = np.cumsum(rows)                         # blocklasts(ax), 1-based last index of every block
    for d in sorted({d0, d1}):
        if d == 0:
            continue
        v = u if d > 0 else l
        j = np.arange(1, n + 1)                     # 1-based column index
        i = np.clip(j - d, 1, n)                    # the row the diagonal d covers in column j
        md = np.searchsorted(lasts, j, side="left")     # searchsortedfirst(blocklasts, j) - 1
        db = np.searchsorted(lasts, i, side="left")
        np.maximum.at(v, md, np.abs(db - md))
    return l, u


def _require_device(A):
    if not A.on_device:
        raise TypeError("matrix data is host-resident: call A.to_device(handle) first (the device path has no CPU fallback)")
    return A.data.handle


def _descs(X, h):
    """(sky descriptor, bbb descriptor) of an operand, one of them None."""
    if X is None:
        return None, None
    if isinstance(X, BandedBlockBandedMatrix):
        return None, X.descriptor(h)
    return X.descriptor(h), None


def _same_axes(A, B):
    return len(A.rows) == len(B.rows) and len(A.cols) == len(B.cols) and (A.rows == B.rows).all() and (A.cols == B.cols).all()


def axpby_(Cm, alpha, A, beta, B):
    """``C .= alpha .* A .+ beta .* B`` on the device for block-banded operands of different bandwidths / formats
    (A or B may be None).  Cm may be B itself (``B .= 2.0 .* A .+ B``, test/test_broadcasting.jl:159)."""
    h = _require_device(Cm)
    for X in (A, B):
        if X is not None:
            _require_device(X)
            if not _same_axes(X, Cm):
                raise L.DimensionMismatch("block axes of the operands differ from the destination's")
            if X.dtype != Cm.dtype:
                raise TypeError("eltype mismatch")
    dt = Cm.dtype
    if dt not in _SUF:
        raise TypeError(f"device path takes Float32/Float64 only, got {dt}")
    suf, ct = _SUF[dt]
    pa = C.c_void_p(A.data.ptr) if A is not None else None
    pb = C.c_void_p(B.data.ptr) if B is not None else None
    if isinstance(Cm, BandedBlockBandedMatrix):
        for X in (A, B):
            if X is not None and not isinstance(X, BandedBlockBandedMatrix):
                raise TypeError("a BandedBlockBandedMatrix destination takes BandedBlockBandedMatrix operands "
                                "(a sum with a BlockBandedMatrix is a BlockBandedMatrix)")
        fn = getattr(h._lib, f"bbm_bbb_add_{suf}")
        rc = fn(h.h, A.descriptor(h) if A is not None else None, B.descriptor(h) if B is not None else None, Cm.descriptor(h),
                ct(alpha), pa, ct(beta), pb, C.c_void_p(Cm.data.ptr))
    else:
        (As, Ab), (Bs, Bb) = _descs(A, h), _descs(B, h)
        fn = getattr(h._lib, f"bbm_sky_add_{suf}")
        rc = fn(h.h, As, Ab, Bs, Bb, Cm.descriptor(h), ct(alpha), pa, ct(beta), pb, C.c_void_p(Cm.data.ptr))
    L.check(rc, h.h, "broadcast")
    return Cm


def similar_sum(A, B):
    """``similar(broadcasted(+, A, B))``: a BandedBlockBandedMatrix with the maxima of all four bandwidths when both
    operands are banded-block-banded, else a BlockBandedMatrix with the maxima of the (constant) block bandwidths
    (src/broadcast.jl:107-131, BroadcastStyle rules :16-23)."""
    h = _require_device(A)
    if not _same_axes(A, B):
        raise L.DimensionMismatch("A and B have different block axes")
    (lA, uA), (lB, uB) = _const_bw(A), _const_bw(B)
    lu = (max(lA, lB), max(uA, uB))
    if isinstance(A, BandedBlockBandedMatrix) and isinstance(B, BandedBlockBandedMatrix):
        lm = (max(A.lam, B.lam), max(A.mu, B.mu))
        return BandedBlockBandedMatrix(h.empty(BandedBlockBandedMatrix.data_shape(A.cols, lu, lm), A.dtype), A.rows, A.cols, lu, lm)
    n = BlockSkylineMatrix._numentries(A.rows, A.cols, lu[0], lu[1])
    return BlockSkylineMatrix(h.empty(n, A.dtype), A.rows, A.cols, lu[0], lu[1])


def _dev_vec(h, v, dt, n, what):
    if v is None:
        return None
    if not isinstance(v, DeviceArray):
        v = h.to_device(np.ascontiguousarray(np.asarray(v, dtype=dt)))
    if v.shape != (n,):
        raise L.DimensionMismatch(f"{what} must have {n} entries, got {v.shape}")
    return v


def add_diags_(A, alpha=1.0, lam=0.0, dl=None, d=None, du=None):
    """In place: ``A[i,i] += alpha*(lam + d[i])``, ``A[i,i+1] += alpha*du[i]``, ``A[i+1,i] += alpha*dl[i]``."""
    h = _require_device(A)
    dt = A.dtype
    suf, ct = _SUF[dt]
    m = A.shape[0]
    vd, vu, vl = _dev_vec(h, d, dt, m, "the diagonal"), _dev_vec(h, du, dt, max(m - 1, 0), "the super-diagonal"), _dev_vec(h, dl, dt, max(m - 1, 0), "the sub-diagonal")
    kind = "bbb" if isinstance(A, BandedBlockBandedMatrix) else "sky"
    fn = getattr(h._lib, f"bbm_{kind}_add_diags_{suf}")
    ptr = lambda v: C.c_void_p(v.ptr) if v is not None else None
    L.check(fn(h.h, A.descriptor(h), C.c_void_p(A.data.ptr), ct(alpha), ct(lam), ptr(vl), ptr(vd), ptr(vu)), h.h, "broadcast")
    return A   # (device copies of host diagonals made above are released by cudaFree, which waits for the kernel)


def _special(A, S, sign, reverse):
    """``A op S`` (reverse False) or ``S op A`` (reverse True), op = + / - (sign = +1 / -1)."""
    h = _require_device(A)
    diagonals, lam, dl, d, du = _parts(S)
    if isinstance(A, BlockSkylineMatrix):
        if len(A.rows) != len(A.cols) or (A.rows != A.cols).any():
            raise L.DimensionMismatch("blocks are not square")          # checksquareblocks
        l, u = accommodating_bandwidths(A.rows, A.l, A.u, diagonals)
        out = BlockSkylineMatrix(h.empty(BlockSkylineMatrix._numentries(A.rows, A.cols, l, u), A.dtype), A.rows, A.cols, l, u)
    else:
        out = BandedBlockBandedMatrix(h.empty(A.data.shape, A.dtype), A.rows, A.cols, (A.l, A.u), (A.lam, A.mu))
    # S op A starts from op(A) and ADDS S; A op S starts from A and applies op to S
    axpby_(out, float(sign) if reverse else 1.0, A, 0.0, None)
    return add_diags_(out, 1.0 if reverse else float(sign), lam, dl, d, du)


def add(A, B):
    """``A + B``."""
    if isinstance(B, _SPECIAL):
        return _special(A, B, +1, False)
    if isinstance(A, _SPECIAL):
        return _special(B, A, +1, True)
    return axpby_(similar_sum(A, B), 1.0, A, 1.0, B)


def sub(A, B):
    """``A - B``."""
    if isinstance(B, _SPECIAL):
        return _special(A, B, -1, False)
    if isinstance(A, _SPECIAL):
        return _special(B, A, -1, True)
    return axpby_(similar_sum(A, B), 1.0, A, -1.0, B)


def scaled_add(alpha, A, B):
    """``alpha*A + B`` (``2.0 .* A .+ B``, src/broadcast.jl:379-411)."""
    return axpby_(similar_sum(A, B), alpha, A, 1.0, B)
