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
from .matrices import BandedBlockBandedMatrix

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
    if not isinstance(A, BandedBlockBandedMatrix):
        raise TypeError(f"unsupported matrix type {type(A).__name__}")
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
        fn = getattr(h._lib, f"bbm_bbb_mul_{suf}")
        rc = fn(h.h, desc, ct(alpha), C.c_void_p(A.data.ptr), C.c_void_p(x.ptr), _ld(x), xc, ct(beta),
                C.c_void_p(y.ptr), _ld(y))
    else:
        if not y.flags.writeable:
            raise ValueError("y is read-only")
        fn = getattr(h._lib, f"bbm_bbb_mul_host_{suf}")
        rc = fn(h.h, desc, ct(alpha), C.c_void_p(A.data.ptr), x.ctypes.data_as(C.c_void_p), _ld(x), xc, ct(beta),
                y.ctypes.data_as(C.c_void_p), _ld(y))
    L.check(rc, h.h, "mul!")
    return y


def muladd_(alpha, A, B, beta, Cdest):
    """``ArrayLayouts.muladd!(alpha, A, B, beta, C)`` = ``materialize!(MulAdd(alpha,A,B,beta,C))``."""
    return mul_(Cdest, A, B, alpha, beta)


def mul(A, x):
    """``A*x``: allocates the result where x lives (``similar`` + ``mul!``)."""
    h = _require_device(A)
    m = A.shape[0]
    shape = (m,) if len(x.shape) == 1 else (m, x.shape[1])
    if isinstance(x, DeviceArray):
        y = h.empty(shape, A.dtype)
    else:
        y = np.empty(shape, dtype=A.dtype, order="F")
    return mul_(y, A, x, 1.0, 0.0)
