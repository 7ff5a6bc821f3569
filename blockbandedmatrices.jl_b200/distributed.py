"""Row-block-partitioned BandedBlockBandedMatrix matvec over several GPUs (one process per GPU).

Host-side mirror of the ``bbm_bbb_partition_rows`` / ``bbm_bbb_dist_layout`` / ``bbm_comm_*`` /
``bbm_dist_bbb_*`` part of the C ABI.  ``torch.distributed`` is used only as the bootstrap channel
that ships the 128-byte NCCL unique id from rank 0 to the others (a Julia host would use MPI.jl or
Distributed for the same thing); the halo exchange itself is NCCL send/recv issued inside the
library, overlapped with the interior rows' product.

Reference precedent: examples/shared_array_backend.jl:92-122 (contiguous block-row ranges per
worker); sub-view structure src/BandedBlockBandedMatrix.jl:416-420, 450-456.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .matrices import _cum, _sizes


def partition_rows(rows, nranks: int) -> np.ndarray:
    """Balanced contiguous block-row ranges: ``Kbounds[p] .. Kbounds[p+1]`` is rank p's range."""
    r = _sizes(rows)
    kb = np.zeros(nranks + 1, dtype=np.int64)
    L.check(L.load().bbm_bbb_partition_rows(len(r), r.ctypes.data_as(C.c_void_p), nranks, kb.ctypes.data_as(C.c_void_p)),
            None, "bbm_bbb_partition_rows")
    return kb


def layout(rows, cols, l, u, nranks, rank, Kbounds=None):
    """Host-only layout of one rank: (dict of bbm_dist_layout fields, sends, recvs); sends / recvs
    are lists of (peer, offset into x_local, element count)."""
    r, c = _sizes(rows), _sizes(cols)
    kb = partition_rows(r, nranks) if Kbounds is None else np.ascontiguousarray(Kbounds, dtype=np.int64)
    lay = L.DistLayout()
    s = np.zeros(3 * max(1, nranks), dtype=np.int64)
    v = np.zeros(3 * max(1, nranks), dtype=np.int64)
    L.check(L.load().bbm_bbb_dist_layout(len(r), len(c), r.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), l, u,
                                         nranks, rank, kb.ctypes.data_as(C.c_void_p), C.byref(lay),
                                         s.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p)), None,
            "bbm_bbb_dist_layout")
    d = lay.asdict()
    sends = [tuple(int(t) for t in s[3 * i:3 * i + 3]) for i in range(d["nsend"])]
    recvs = [tuple(int(t) for t in v[3 * i:3 * i + 3]) for i in range(d["nrecv"])]
    return d, sends, recvs


def piece_structures(rows, cols, l, u, lay):
    """The (up to three) sub-views a rank multiplies: interior, left boundary, right boundary.
    Yields (K_lo, K_hi, J_lo, J_hi, l', u') with the shifted block bandwidths of the sub-view
    (src/BandedBlockBandedMatrix.jl:450-456)."""
    out = []
    for lo, hi in ((lay["Ka"], lay["Kb"]), (lay["K0"], lay["Ka"]), (lay["Kb"], lay["K1"])):
        if hi <= lo:
            continue
        Jl = min(max(lo - l, lay["Jlo"]), lay["Jhi"])
        Jh = min(max(hi + u, Jl), lay["Jhi"])
        s = lo - Jl
        out.append((lo, hi, Jl, Jh, l - s, u + s))
    return out


class Communicator:
    """``bbm_comm_t``: NCCL communicator + exchange stream, bootstrapped over torch.distributed."""

    def __init__(self, handle, dist=None, nranks: int = 1, rank: int = 0):
        self.handle = handle
        self._lib = handle._lib
        self.nranks, self.rank = nranks, rank
        idbuf = (C.c_uint8 * 128)()
        if nranks > 1 and dist is None:
            idbuf = None  # no NCCL: the caller fills the halo parts of x_local itself
        elif nranks > 1:
            import torch
            if rank == 0:
                L.check(self._lib.bbm_comm_unique_id(idbuf), handle.h, "bbm_comm_unique_id")
            t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda(handle.device)
            dist.broadcast(t, src=0)
            idbuf = (C.c_uint8 * 128)(*[int(b) for b in t.cpu().tolist()])
        c = C.c_void_p()
        L.check(self._lib.bbm_comm_create(handle.h, idbuf, nranks, rank, C.byref(c)), handle.h, "bbm_comm_create")
        self.c = c

    def close(self):
        if getattr(self, "c", None):
            self._lib.bbm_comm_destroy(self.c)
            self.c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DistBandedBlockBandedMatrix:
    """Rank-local part of a row-block-partitioned BandedBlockBandedMatrix.

    ``local_data(parent)`` cuts the rank's column slab out of a host parent array (or the caller
    generates only that slab); ``x_local`` / ``y_own`` are device buffers of the plan's sizes."""

    def __init__(self, comm: Communicator, rows, cols, lu, lm, dtype=np.float64, Kbounds=None):
        self.comm = comm
        self.rows, self.cols = _sizes(rows), _sizes(cols)
        self.l, self.u = int(lu[0]), int(lu[1])
        self.lam, self.mu = int(lm[0]), int(lm[1])
        self.dtype = np.dtype(dtype)
        h = comm.handle
        kb = None if Kbounds is None else np.ascontiguousarray(Kbounds, dtype=np.int64)
        p = C.c_void_p()
        L.check(h._lib.bbm_dist_bbb_plan_create(
            comm.c, len(self.rows), len(self.cols), self.rows.ctypes.data_as(C.c_void_p),
            self.cols.ctypes.data_as(C.c_void_p), self.l, self.u, self.lam, self.mu,
            None if kb is None else kb.ctypes.data_as(C.c_void_p), C.byref(p)), h.h, "bbm_dist_bbb_plan_create")
        self.plan = p
        lay = L.DistLayout()
        L.check(h._lib.bbm_dist_bbb_plan_layout(p, C.byref(lay)), h.h)
        self.layout = lay.asdict()
        self.data = None

    @property
    def P(self):
        return max(0, self.l + self.u + 1) * max(0, self.lam + self.mu + 1)

    def local_data(self, parent: np.ndarray) -> np.ndarray:
        c0 = self.layout["col_offset"]
        return np.asfortranarray(parent[:, c0:c0 + self.layout["n_local"]])

    def set_local_data(self, slab: np.ndarray):
        if tuple(slab.shape) != (self.P, self.layout["n_local"]):
            raise L.DimensionMismatch(f"local slab must be {self.P} x {self.layout['n_local']}, got {slab.shape}")
        h = self.comm.handle
        self.data = h.to_device(np.asfortranarray(slab.astype(self.dtype, copy=False))) if slab.size else h.empty(slab.shape, self.dtype)
        return self

    def mul_(self, y_own, x_local, alpha=1.0, beta=0.0, nrhs=1):
        """y_own <- alpha*A[K0:K1,:]*x + beta*y_own (collective; asynchronous on the handle's stream)."""
        h = self.comm.handle
        suf, ct = ("f64", C.c_double) if self.dtype == np.float64 else ("f32", C.c_float)
        fn = getattr(h._lib, f"bbm_dist_bbb_mul_{suf}")
        L.check(fn(self.plan, ct(alpha), C.c_void_p(self.data.ptr), C.c_void_p(x_local.ptr), nrhs, ct(beta),
                   C.c_void_p(y_own.ptr)), h.h, "bbm_dist_bbb_mul")
        return y_own

    def close(self):
        if getattr(self, "plan", None):
            self.comm.handle._lib.bbm_dist_bbb_plan_destroy(self.plan)
            self.plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
