"""``mul!`` / ``muladd!`` / ``*`` for the device-resident containers.

Mirrors the user-facing entry points of SURVEY.md 8(b): ``mul!(y, A, x[, alpha, beta])``,
``muladd!(alpha, A, B, beta, C)``, ``A*x`` -- each does its DimensionMismatch checks on the host
(as the Julia extension method would) and then makes exactly one C-ABI call.  A matrix whose
``data`` is still a host array is *not* taken: in the reference design those keep dispatching to
the reference's own CPU path, which this package deliberately does not contain.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .device import DeviceArray
from .matrices import Adjoint, BandedBlockBandedMatrix, BlockBandedQL, BlockBandedQR, BlockSkylineMatrix, _Triangular

_SUF = {np.dtype(np.float64): ("f64", C.c_double), np.dtype(np.float32): ("f32", C.c_float)}


def _require_device(A):
    if not A.on_device:
        raise TypeError("matrix data is host-resident: call A.to_device(handle) first "
                        "(the device path has no CPU fallback)")
    return A.data.handle


def _vec_dims(v, what):
    if len(v.shape) == 1:
        return v.shape[0], 1
    if len(v.shape) == 2:
        return v.shape[0], v.shape[1]
    raise L.DimensionMismatch(f"{what} must be a vector or a matrix")


def _ld(v):
    if isinstance(v, DeviceArray):
        return v.ld
    if v.ndim == 2 and v.shape[1] > 1:
        if not v.flags.f_contiguous:
            raise ValueError("host matrices must be column-major (Fortran order)")
        return max(1, v.shape[0])
    if not (v.flags.c_contiguous or v.flags.f_contiguous):
        raise ValueError("host vectors must be contiguous")
    return max(1, v.shape[0])


def mul_(y, A, x, alpha=1.0, beta=0.0):
    """``mul!(y, A, x, alpha, beta)``: y <- alpha*A*x + beta*y, returns y.

    x, y are both DeviceArrays (asynchronous, on the handle's stream) or both numpy arrays
    (host vectors: copied through the C ABI's *_host_* call, synchronous)."""
    if isinstance(A, BlockSkylineMatrix) and isinstance(x, BlockSkylineMatrix):
        return _sky_matmul_(y, A, x, alpha, beta)
    if isinstance(A, DeviceArray) and isinstance(x, BlockSkylineMatrix):
        return _sky_rmul_(y, A, x, alpha, beta)
    if isinstance(A, _Triangular):
        return _bbb_trimul_(y, A, x, alpha, beta)
    if isinstance(A, Adjoint):
        return _bbb_mul_t_(y, A, x, alpha, beta)
    if isinstance(A, BandedBlockBandedMatrix) and isinstance(x, BandedBlockBandedMatrix):
        return _bbb_matmul_(y, A, x, alpha, beta)
    if not isinstance(A, (BandedBlockBandedMatrix, BlockSkylineMatrix)):
        raise TypeError(f"unsupported matrix type {type(A).__name__}")
    kind = "bbb" if isinstance(A, BandedBlockBandedMatrix) else "sky"
    h = _require_device(A)
    m, n = A.shape
    xr, xc = _vec_dims(x, "x")
    yr, yc = _vec_dims(y, "y")
    if xr != n or yr != m or xc != yc:
        raise L.DimensionMismatch(f"A is {m}x{n}, x is {x.shape}, y is {y.shape}")
    dt = A.dtype
    if dt not in _SUF:
        raise TypeError(f"device path takes Float32/Float64 only, got {dt}")
    if np.dtype(x.dtype) != dt or np.dtype(y.dtype) != dt:
        raise TypeError("eltype mismatch between A, x and y")
    suf, ct = _SUF[dt]
    desc = A.descriptor(h)
    dev = isinstance(x, DeviceArray), isinstance(y, DeviceArray)
    if dev[0] != dev[1]:
        raise TypeError("x and y must both be device arrays or both be host arrays")
    if dev[0]:
        if y.ptr == x.ptr and m and n:
            raise ValueError("mul!: destination must not alias the source")
        fn = getattr(h._lib, f"bbm_{kind}_mul_{suf}")
        rc = fn(h.h, desc, ct(alpha), C.c_void_p(A.data.ptr), C.c_void_p(x.ptr), _ld(x), xc, ct(beta),
                C.c_void_p(y.ptr), _ld(y))
    else:
        if kind != "bbb":
            raise TypeError("host vectors are taken by the BandedBlockBandedMatrix product only; "
                            "move x and y to the device for a BlockSkylineMatrix")
        if not y.flags.writeable:
            raise ValueError("y is read-only")
        fn = getattr(h._lib, f"bbm_bbb_mul_host_{suf}")
        rc = fn(h.h, desc, ct(alpha), C.c_void_p(A.data.ptr), x.ctypes.data_as(C.c_void_p), _ld(x), xc, ct(beta),
                y.ctypes.data_as(C.c_void_p), _ld(y))
    L.check(rc, h.h, "mul!")
    return y


def _sky_rmul_(Y, X, A, alpha=1.0, beta=0.0):
    """``mul!(Y, X, A, alpha, beta)`` for a dense device matrix X (column-major) times a device-resident BlockSkylineMatrix
    (`X*A`, test/test_linalg.jl:56-57)."""
    h = _require_device(A)
    m, n = A.shape
    if not isinstance(Y, DeviceArray):
        raise TypeError("X and Y must both be device arrays")
    xr, xc = _vec_dims(X, "X")
    yr, yc = _vec_dims(Y, "Y")
    if xc != m or yc != n or xr != yr:
        raise L.DimensionMismatch(f"X is {X.shape}, A is {m}x{n}, Y is {Y.shape}")
    dt = A.dtype
    if dt not in _SUF or np.dtype(X.dtype) != dt or np.dtype(Y.dtype) != dt:
        raise TypeError("device path takes matching Float32/Float64 operands only")
    suf, ct = _SUF[dt]
    fn = getattr(h._lib, f"bbm_sky_rmul_{suf}")
    rc = fn(h.h, A.descriptor(h), ct(alpha), C.c_void_p(X.ptr), _ld(X), xr, C.c_void_p(A.data.ptr), ct(beta), C.c_void_p(Y.ptr), _ld(Y))
    L.check(rc, h.h, "mul!")
    return Y


def _bbb_matmul_(Cm, A, B, alpha=1.0, beta=0.0):
    """``mul!(C, A, B, alpha, beta)`` for three device-resident BandedBlockBandedMatrix operands."""
    if not isinstance(Cm, BandedBlockBandedMatrix):
        raise TypeError("the destination of a banded-block-banded product must be a BandedBlockBandedMatrix")
    h = _require_device(A)
    _require_device(B)
    _require_device(Cm)
    if A.shape[1] != B.shape[0] or Cm.shape != (A.shape[0], B.shape[1]):
        raise L.DimensionMismatch(f"A is {A.shape}, B is {B.shape}, C is {Cm.shape}")
    dt = A.dtype
    if dt not in _SUF or B.dtype != dt or Cm.dtype != dt:
        raise TypeError("device path takes matching Float32/Float64 operands only")
    suf, ct = _SUF[dt]
    fn = getattr(h._lib, f"bbm_bbb_matmul_{suf}")
    rc = fn(h.h, A.descriptor(h), B.descriptor(h), Cm.descriptor(h), ct(alpha), C.c_void_p(A.data.ptr), C.c_void_p(B.data.ptr),
            ct(beta), C.c_void_p(Cm.data.ptr))
    L.check(rc, h.h, "mul!")
    return Cm


def _bbb_mul_t_(y, At, x, alpha=1.0, beta=0.0):
    """``mul!(y, A', x, alpha, beta)`` for a device-resident real BandedBlockBandedMatrix."""
    A = At.parent
    if not isinstance(A, (BandedBlockBandedMatrix, BlockSkylineMatrix)):
        raise TypeError("adjoint products are taken for BandedBlockBandedMatrix / BlockSkylineMatrix parents only")
    h = _require_device(A)
    m, n = A.shape
    xr, xc = _vec_dims(x, "x")
    yr, yc = _vec_dims(y, "y")
    if xr != m or yr != n or xc != yc:
        raise L.DimensionMismatch(f"A' is {n}x{m}, x is {x.shape}, y is {y.shape}")
    if not (isinstance(x, DeviceArray) and isinstance(y, DeviceArray)):
        raise TypeError("adjoint products take device vectors")
    dt = A.dtype
    if dt not in _SUF or np.dtype(x.dtype) != dt or np.dtype(y.dtype) != dt:
        raise TypeError("device path takes matching Float32/Float64 operands only")
    suf, ct = _SUF[dt]
    fn = getattr(h._lib, f"bbm_bbb_mul_t_{suf}" if isinstance(A, BandedBlockBandedMatrix) else f"bbm_sky_mul_t_{suf}")
    rc = fn(h.h, A.descriptor(h), ct(alpha), C.c_void_p(A.data.ptr), C.c_void_p(x.ptr), _ld(x), xc, ct(beta),
            C.c_void_p(y.ptr), _ld(y))
    L.check(rc, h.h, "mul!")
    return y


def _bbb_trimul_(y, T, x, alpha=1.0, beta=0.0):
    """``mul!(y, UpperTriangular(A), x, alpha, beta)`` etc. for a device-resident BandedBlockBandedMatrix
    (the ``U*b`` / ``lmul(U,b)`` of test/test_triblockbanded.jl:17-53)."""
    A = T.parent
    if not isinstance(A, BandedBlockBandedMatrix):
        raise TypeError("triangular products are taken for BandedBlockBandedMatrix parents only")
    h = _require_device(A)
    m, n = A.shape
    xr, xc = _vec_dims(x, "x")
    yr, yc = _vec_dims(y, "y")
    if xr != n or yr != m or xc != yc:
        raise L.DimensionMismatch(f"A is {m}x{n}, x is {x.shape}, y is {y.shape}")
    if not (isinstance(x, DeviceArray) and isinstance(y, DeviceArray)):
        raise TypeError("triangular products take device vectors")
    dt = A.dtype
    if dt not in _SUF or np.dtype(x.dtype) != dt or np.dtype(y.dtype) != dt:
        raise TypeError("device path takes matching Float32/Float64 operands only")
    suf, ct = _SUF[dt]
    fn = getattr(h._lib, f"bbm_bbb_trimul_{suf}")
    rc = fn(h.h, A.descriptor(h), ord(T.uplo), 1 if T.unit else 0, ct(alpha), C.c_void_p(A.data.ptr), C.c_void_p(x.ptr), _ld(x), xc,
            ct(beta), C.c_void_p(y.ptr), _ld(y))
    L.check(rc, h.h, "mul!")
    return y


def _same_structure(A, B):
    if type(A) is not type(B) or A.shape != B.shape or len(A.rows) != len(B.rows) or len(A.cols) != len(B.cols):
        return False
    if (A.rows != B.rows).any() or (A.cols != B.cols).any():
        return False
    if isinstance(A, BandedBlockBandedMatrix):
        return (A.l, A.u, A.lam, A.mu) == (B.l, B.u, B.lam, B.mu)
    return (A.l == B.l).all() and (A.u == B.u).all()


def axpy_(a, X, Y):
    """``axpy!(a, X, Y)``: Y <- a*X + Y for device vectors or two device matrices of identical structure
    (src/BandedBlockBandedMatrix.jl:605-606, src/broadcast.jl:307-314); returns Y."""
    if isinstance(X, DeviceArray) and isinstance(Y, DeviceArray):
        dx, dy, h = X, Y, X.handle
        if X.shape != Y.shape:
            raise L.DimensionMismatch(f"{X.shape} vs {Y.shape}")
    else:
        if not _same_structure(X, Y):
            raise TypeError("device axpy! takes operands of identical structure (others keep the host path)")
        h = _require_device(X)
        _require_device(Y)
        dx, dy = X.data, Y.data
    dt = np.dtype(dx.dtype)
    if dt not in _SUF or np.dtype(dy.dtype) != dt:
        raise TypeError("device path takes matching Float32/Float64 operands only")
    suf, ct = _SUF[dt]
    n = int(np.prod(dx.shape, dtype=np.int64))
    L.check(getattr(h._lib, f"bbm_axpy_{suf}")(h.h, n, ct(a), C.c_void_p(dx.ptr), C.c_void_p(dy.ptr)), h.h, "axpy!")
    return Y


def lmul_(a, A):
    """``lmul!(a, A)`` / ``rmul!(A, a)``: scale a device vector or the storage of a device matrix in place."""
    d = A if isinstance(A, DeviceArray) else A.data
    h = d.handle
    dt = np.dtype(d.dtype)
    if dt not in _SUF:
        raise TypeError("device path takes Float32/Float64 only")
    suf, ct = _SUF[dt]
    n = int(np.prod(d.shape, dtype=np.int64))
    L.check(getattr(h._lib, f"bbm_scal_{suf}")(h.h, n, ct(a), C.c_void_p(d.ptr)), h.h, "lmul!")
    return A


def muladd_(alpha, A, B, beta, Cdest):
    """``ArrayLayouts.muladd!(alpha, A, B, beta, C)`` = ``materialize!(MulAdd(alpha,A,B,beta,C))``."""
    return mul_(Cdest, A, B, alpha, beta)


def add_bandwidths(A: BlockSkylineMatrix, B: BlockSkylineMatrix):
    """Column block-bandwidths of ``A*B`` (src/linalg.jl:8-30): ``Fill(lA+lB), Fill(uA+uB)`` for two
    BlockBandedMatrix operands, else per column ``v[i] = Bv[i] + maximum(Av[sel])`` with
    ``sel = max(i-Bu[i],1):min(i+Bl[i],length(Av))`` (skipped when empty)."""
    if A.constant_bandwidths and B.constant_bandwidths:
        lA, uA = (int(A.l[0]), int(A.u[0])) if len(A.l) else (0, 0)
        lB, uB = (int(B.l[0]), int(B.u[0])) if len(B.l) else (0, 0)
        return lA + lB, uA + uB
    l, u = B.l.copy(), B.u.copy()
    for v, Av in ((l, A.l), (u, A.u)):
        for i in range(len(v)):
            lo, hi = max(i - int(B.u[i]), 0), min(i + int(B.l[i]), len(Av) - 1)
            if hi >= lo:
                v[i] += int(np.max(Av[lo:hi + 1]))
    return l, u


def _sky_matmul_(Cm, A, B, alpha=1.0, beta=0.0):
    """``mul!(C, A, B, alpha, beta)`` for three BlockSkylineMatrix operands."""
    if not isinstance(Cm, BlockSkylineMatrix):
        raise TypeError("the destination of a block-banded product must be a BlockSkylineMatrix")
    h = _require_device(A)
    _require_device(B)
    _require_device(Cm)
    if A.shape[1] != B.shape[0] or Cm.shape != (A.shape[0], B.shape[1]):
        raise L.DimensionMismatch(f"A is {A.shape}, B is {B.shape}, C is {Cm.shape}")
    dt = A.dtype
    if dt not in _SUF or B.dtype != dt or Cm.dtype != dt:
        raise TypeError("device path takes matching Float32/Float64 operands only")
    suf, ct = _SUF[dt]
    fn = getattr(h._lib, f"bbm_sky_matmul_{suf}")
    rc = fn(h.h, A.descriptor(h), B.descriptor(h), Cm.descriptor(h), ct(alpha), C.c_void_p(A.data.ptr),
            C.c_void_p(B.data.ptr), ct(beta), C.c_void_p(Cm.data.ptr))
    L.check(rc, h.h, "mul!")
    return Cm


def _qr_call(name, F, B):
    A = F.factors
    h = _require_device(A)
    if not isinstance(B, DeviceArray) or not isinstance(F.tau, DeviceArray):
        raise TypeError("the right-hand side and tau must be device arrays")
    br, _bc = _vec_dims(B, "B")
    if br != A.shape[0]:
        raise L.DimensionMismatch(f"Q is {A.shape[0]} x {A.shape[0]}, B is {B.shape}")
    if A.dtype not in _SUF or np.dtype(B.dtype) != A.dtype or np.dtype(F.tau.dtype) != A.dtype:
        raise TypeError("the device QR solve takes matching Float32 / Float64 operands")
    name = name.replace("_f64", "_" + _SUF[A.dtype][0])
    rc = getattr(h._lib, name)(h.h, A.descriptor(h), C.c_void_p(A.data.ptr), C.c_void_p(F.tau.ptr), C.c_void_p(B.ptr), _ld(B), _bc)
    L.check(rc, h.h, name)
    return B


def sparse(A, index_base: int = 0):
    """``sparse(A)`` for a device-resident BandedBlockBandedMatrix (ext/BlockBandedMatricesSparseArraysExt.jl:10-27):
    returns the CSC arrays ``(colptr, rowval, nzval)`` as device arrays (``index_base=1`` for Julia's
    ``SparseMatrixCSC``); explicit zeros inside the bands are kept like the reference's."""
    if not isinstance(A, BandedBlockBandedMatrix):
        raise TypeError("sparse takes a BandedBlockBandedMatrix")
    h = _require_device(A)
    dt = A.dtype
    if dt not in _SUF:
        raise TypeError(f"device path takes Float32/Float64 only, got {dt}")
    n = A.shape[1]
    colptr = h.empty(n + 1, np.int64)
    nnz = C.c_int64(0)
    L.check(h._lib.bbm_bbb_sparse_colptr(h.h, A.descriptor(h), index_base, C.c_void_p(colptr.ptr), C.byref(nnz)), h.h, "sparse")
    rowval, nzval = h.empty(max(nnz.value, 1), np.int64), h.empty(max(nnz.value, 1), dt)
    fn = getattr(h._lib, f"bbm_bbb_sparse_fill_{_SUF[dt][0]}")
    L.check(fn(h.h, A.descriptor(h), C.c_void_p(A.data.ptr), C.c_void_p(colptr.ptr), index_base, C.c_void_p(rowval.ptr),
               C.c_void_p(nzval.ptr)), h.h, "sparse")
    if nnz.value == 0:
        rowval, nzval = DeviceArray(h, (0,), np.int64, ptr=rowval.ptr), DeviceArray(h, (0,), dt, ptr=nzval.ptr)
    return colptr, rowval, nzval


def qr_(A):
    """``qr!(A)`` on the device (src/blockskylineqr.jl:29-49): A must be a Float64 BlockBandedMatrix already in the
    factors' block bandwidths ``(l, l+u)`` (``A.with_bandwidths(l, l+u)`` on the host, then ``to_device``);
    its storage becomes ``F.factors``.  Returns the BlockBandedQR."""
    if not isinstance(A, BlockSkylineMatrix):
        raise TypeError("qr! takes a BlockBandedMatrix")
    h = _require_device(A)
    if A.dtype not in _SUF:
        raise TypeError("the device qr! takes Float32 / Float64")
    n_tau = A.shape[1] if len(A.rows) >= len(A.cols) else A.shape[0]
    tau = h.empty(max(n_tau, 1), A.dtype)
    rc = getattr(h._lib, f"bbm_sky_qr_{_SUF[A.dtype][0]}")(h.h, A.descriptor(h), C.c_void_p(A.data.ptr), C.c_void_p(tau.ptr))
    L.check(rc, h.h, "qr!")
    if n_tau == 0:
        tau = DeviceArray(h, (0,), A.dtype, ptr=tau.ptr)
    elif tau.shape[0] != n_tau:
        tau = DeviceArray(h, (n_tau,), A.dtype, ptr=tau.ptr)
    return BlockBandedQR(A, tau)


def ql_(A):
    """``ql!(A)`` on the device (src/blockskylineqr.jl:51-72): A must be a square Float64 BlockBandedMatrix already in
    the factors' block bandwidths ``(l+u, u)`` (``A.with_bandwidths(l+u, u)``); returns the BlockBandedQL."""
    if not isinstance(A, BlockSkylineMatrix):
        raise TypeError("ql! takes a BlockBandedMatrix")
    h = _require_device(A)
    if A.dtype not in _SUF:
        raise TypeError("the device ql! takes Float32 / Float64")
    if len(A.rows) < len(A.cols):
        raise ValueError("Wide block-QL not implemented")      # ArgumentError, src/blockskylineqr.jl:56
    n_tau = A.shape[1]
    tau = h.empty(max(n_tau, 1), A.dtype)
    rc = getattr(h._lib, f"bbm_sky_ql_{_SUF[A.dtype][0]}")(h.h, A.descriptor(h), C.c_void_p(A.data.ptr), C.c_void_p(tau.ptr))
    L.check(rc, h.h, "ql!")
    if tau.shape[0] != n_tau:
        tau = DeviceArray(h, (n_tau,), A.dtype, ptr=tau.ptr)
    return BlockBandedQL(A, tau)


def lmul_adj_(F, B):
    """``lmul!(F.Q', B)`` for a block-banded QR: B <- Q' B in place (src/blockskylineqr.jl:77-127)."""
    if not isinstance(F, BlockBandedQR):
        raise TypeError("lmul_adj_ takes a BlockBandedQR / BlockBandedQL")
    return _qr_call("bbm_sky_ql_lmul_adj_f64" if isinstance(F, BlockBandedQL) else "bbm_sky_qr_lmul_adj_f64", F, B)


def ldiv_(T, B):
    """``ldiv!(UpperTriangular(A), B)`` etc.: B <- T^-1 B in place on the device, returns B.
    ``ldiv!(F, B)`` for a square ``BlockBandedQR`` (src/blockskylineqr.jl:131-145): Q' B, then the triangular solve."""
    if isinstance(T, BlockBandedQR):
        if T.shape[0] != T.shape[1]:
            raise L.DimensionMismatch(f"F is {T.shape}: the device path solves square systems")
        return _qr_call("bbm_sky_ql_ldiv_f64" if isinstance(T, BlockBandedQL) else "bbm_sky_qr_ldiv_f64", T, B)
    if not isinstance(T, _Triangular) or not isinstance(T.parent, (BlockSkylineMatrix, BandedBlockBandedMatrix)):
        raise TypeError("ldiv! takes Upper/LowerTriangular wrappers of a BlockSkylineMatrix or BandedBlockBandedMatrix")
    A = T.parent
    h = _require_device(A)
    if not isinstance(B, DeviceArray):
        raise TypeError("the right-hand side must be a device array")
    br, bc = _vec_dims(B, "B")
    if br != A.shape[0] or A.shape[0] != A.shape[1]:
        raise L.DimensionMismatch(f"T is {A.shape}, B is {B.shape}")
    dt = A.dtype
    if dt not in _SUF or np.dtype(B.dtype) != dt:
        raise TypeError("device path takes matching Float32/Float64 operands only")
    suf, _ = _SUF[dt]
    fn = getattr(h._lib, f"bbm_sky_trsm_{suf}" if isinstance(A, BlockSkylineMatrix) else f"bbm_bbb_trsv_{suf}")
    rc = fn(h.h, A.descriptor(h), ord(T.uplo), 1 if T.unit else 0, C.c_void_p(A.data.ptr), C.c_void_p(B.ptr), _ld(B), bc)
    L.check(rc, h.h, "ldiv!")
    return B


class PreparedTriangular:
    """``ldiv!`` with one triangular BandedBlockBandedMatrix many times (the Gauss-Seidel loop of
    examples/finitedifference_2d.jl:17-47): ``bbm_bbb_trsv_prepare_*`` gathers the matrix into the solve kernel's
    wave-ordered planes once, ``ldiv_`` (``bbm_bbb_trsv_solve_*``) then only moves right-hand sides.  Valid while the
    matrix storage is unchanged."""

    def __init__(self, T):
        if not isinstance(T, _Triangular) or not isinstance(T.parent, BandedBlockBandedMatrix):
            raise TypeError("prepare_ldiv takes Upper/LowerTriangular wrappers of a BandedBlockBandedMatrix")
        A = T.parent
        self.T, self.h = T, _require_device(A)
        dt = A.dtype
        if dt not in _SUF:
            raise TypeError("device path takes Float32/Float64 only")
        self.suf = _SUF[dt][0]
        p = C.c_void_p()
        fn = getattr(self.h._lib, f"bbm_bbb_trsv_prepare_{self.suf}")
        L.check(fn(self.h.h, A.descriptor(self.h), ord(T.uplo), 1 if T.unit else 0, C.c_void_p(A.data.ptr), C.byref(p)), self.h.h, "prepare")
        self.plan = p

    def ldiv_(self, B):
        A = self.T.parent
        br, bc = _vec_dims(B, "B")
        if not isinstance(B, DeviceArray) or br != A.shape[0] or np.dtype(B.dtype) != A.dtype:
            raise L.DimensionMismatch(f"T is {A.shape}, B is {getattr(B, 'shape', None)}")
        fn = getattr(self.h._lib, f"bbm_bbb_trsv_solve_{self.suf}")
        L.check(fn(self.plan, C.c_void_p(B.ptr), _ld(B), bc), self.h.h, "ldiv!")
        return B

    def close(self):
        if getattr(self, "plan", None) and self.h.h:
            self.h._lib.bbm_bbb_tri_plan_destroy(self.plan)
            self.plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def prepare_ldiv(T):
    return PreparedTriangular(T)


def ldiv(T, B):
    """``T \\ B``: solves into a copy of B."""
    h = _require_device(T.factors if isinstance(T, BlockBandedQR) else T.parent)
    X = h.empty(B.shape, B.dtype)
    L.check(h._lib.bbm_memcpy_d2d(h.h, C.c_void_p(X.ptr), C.c_void_p(B.ptr), B.nbytes), h.h)
    return ldiv_(T, X)


def mul(A, x):
    """``A*x``: allocates the result where x lives (``similar`` + ``mul!``).  For two
    BlockSkylineMatrix operands the result structure is ``similar(::MulAdd)`` of src/linalg.jl:32-48."""
    if isinstance(A, DeviceArray) and isinstance(x, BlockSkylineMatrix):
        # X*A: dense times block-banded (test/test_linalg.jl:56-57); the result is dense like X
        h = _require_device(x)
        if len(A.shape) != 2:
            raise L.DimensionMismatch("X*A needs a matrix X")
        return _sky_rmul_(h.empty((A.shape[0], x.shape[1]), x.dtype), A, x, 1.0, 0.0)
    if isinstance(A, Adjoint):
        h = _require_device(A.parent)
        n = A.parent.shape[1]
        shape = (n,) if len(x.shape) == 1 else (n, x.shape[1])
        return mul_(h.empty(shape, A.parent.dtype), A, x, 1.0, 0.0)
    if isinstance(A, _Triangular):
        h = _require_device(A.parent)
        m = A.parent.shape[0]
        shape = (m,) if len(x.shape) == 1 else (m, x.shape[1])
        return mul_(h.empty(shape, A.parent.dtype), A, x, 1.0, 0.0)
    h = _require_device(A)
    if isinstance(A, BandedBlockBandedMatrix) and isinstance(x, BandedBlockBandedMatrix):
        # similar(::MulAdd{BBB,BBB}) of src/linalg.jl:50-71: summed block and sub-block bandwidths
        if len(A.cols) != len(x.rows) or (A.cols != x.rows).any():
            raise L.DimensionMismatch("*")
        lu, lm = (A.l + x.l, A.u + x.u), (A.lam + x.lam, A.mu + x.mu)
        shape = BandedBlockBandedMatrix.data_shape(x.cols, lu, lm)
        return _bbb_matmul_(BandedBlockBandedMatrix(h.empty(shape, A.dtype), A.rows, x.cols, lu, lm), A, x, 1.0, 0.0)
    if isinstance(A, BlockSkylineMatrix) and isinstance(x, BlockSkylineMatrix):
        if len(A.cols) != len(x.rows) or (A.cols != x.rows).any():
            raise L.DimensionMismatch("*")
        l, u = add_bandwidths(A, x)
        n = BlockSkylineMatrix._numentries(A.rows, x.cols, l, u)
        return _sky_matmul_(BlockSkylineMatrix(h.empty(n, A.dtype), A.rows, x.cols, l, u), A, x, 1.0, 0.0)
    m = A.shape[0]
    shape = (m,) if len(x.shape) == 1 else (m, x.shape[1])
    if isinstance(x, DeviceArray):
        y = h.empty(shape, A.dtype)
    else:
        y = np.empty(shape, dtype=A.dtype, order="F")
    return mul_(y, A, x, 1.0, 0.0)
