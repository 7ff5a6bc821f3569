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

    def mul_host_(self, y_own: np.ndarray, x_own: np.ndarray, alpha=1.0, beta=0.0):
        """y_own <- alpha*A[K0:K1,:]*x + beta*y_own with the rank's vectors on the host (pinned arrays for full PCIe
        speed): ``bbm_dist_bbb_mul_host_*`` -- upload, halo exchange, product and download pipelined over block-row
        chunks.  Collective and synchronous."""
        h = self.comm.handle
        if x_own.shape != (self.layout["n_own"],) or y_own.shape != (self.layout["m_own"],):
            raise L.DimensionMismatch(f"x_own must have {self.layout['n_own']} entries and y_own {self.layout['m_own']}")
        if np.dtype(x_own.dtype) != self.dtype or np.dtype(y_own.dtype) != self.dtype:
            raise TypeError("eltype mismatch")
        suf, ct = ("f64", C.c_double) if self.dtype == np.float64 else ("f32", C.c_float)
        fn = getattr(h._lib, f"bbm_dist_bbb_mul_host_{suf}")
        L.check(fn(self.plan, ct(alpha), C.c_void_p(self.data.ptr), x_own.ctypes.data_as(C.c_void_p), ct(beta),
                   y_own.ctypes.data_as(C.c_void_p)), h.h, "bbm_dist_bbb_mul_host")
        return y_own

    def enable_peer(self) -> bool:
        """Collective.  Switch the halo exchange to the fused peer-memory push inside the band-walk
        kernel (``bbm_dist_bbb_enable_peer``); returns False if the partition does not qualify and
        the NCCL exchange stays in use."""
        h = self.comm.handle
        rc = h._lib.bbm_dist_bbb_enable_peer(self.plan, self.dtype.itemsize)
        if rc == 8:  # BBM_ERR_UNSUPPORTED
            return False
        L.check(rc, h.h, "bbm_dist_bbb_enable_peer")
        return True

    def peer_status(self):
        """(fused path active, a peer wait timed out) -- synchronises the stream."""
        h = self.comm.handle
        a, t = C.c_int32(0), C.c_int32(0)
        L.check(h._lib.bbm_dist_bbb_peer_status(self.plan, C.byref(a), C.byref(t)), h.h)
        return bool(a.value), bool(t.value)

    def close(self):
        if getattr(self, "plan", None):
            self.comm.handle._lib.bbm_dist_bbb_plan_destroy(self.plan)
            self.plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ==================================================================================================
# Row-block-partitioned BlockBandedMatrix (constant block bandwidths): matvec and A*B
# ==================================================================================================
# A contiguous block-row range [S0,S1) of a block-banded matrix is itself a BlockSkylineMatrix with
# per-column bandwidths (cf. the sub-view bandwidths of src/linalg.jl:78-87): block column J keeps the
# block rows colsupport(J) & [S0,S1).  Unlike the banded-block-banded format its storage is NOT a
# slice of the parent (every panel is cut by rows), so each rank builds / receives its own panels.
# Structure planning below is pure numpy (exercised on CPU by tests/test_dist_cpu.py); the products
# are the single-GPU kernels on the local structures; halos travel as packed blocks over
# torch.distributed point-to-point (NCCL on GPUs).

def sky_rowrange_structure(rows, cols, l, u, S0, S1):
    """Local structure of rows [S0,S1) of a BlockBandedMatrix(rows, cols, (l,u)).

    Returns (Jlo, Jhi, lloc, uloc): stored block columns [Jlo,Jhi) and the per-column local block
    bandwidths such that local block (K-S0, J-Jlo) is stored iff global block (K,J) is, K in [S0,S1)."""
    Nr, Nc = len(rows), len(cols)
    if S1 <= S0 or l + u + 1 <= 0:
        J0 = min(max(S0, 0), Nc)
        return J0, J0, np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    Jlo = min(max(S0 - l, 0), Nc)
    Jhi = min(max(S1 - 1 + u + 1, Jlo), Nc)
    lloc = np.zeros(Jhi - Jlo, dtype=np.int64)
    uloc = np.zeros(Jhi - Jlo, dtype=np.int64)
    for J in range(Jlo, Jhi):
        lo = max(J - u, 0, S0) - S0
        hi = min(J + l, Nr - 1, S1 - 1) - S0
        Jp = J - Jlo
        if hi < lo:
            lloc[Jp], uloc[Jp] = -1 - Jp, Jp   # empty column: max(Jp-u,0)=0 > min(Jp+l,..)=-1
        else:
            lloc[Jp], uloc[Jp] = hi - Jp, Jp - lo
    return Jlo, Jhi, lloc, uloc


def sky_cut_rows(parent, S0, S1):
    """Host helper: the local BlockSkylineMatrix holding rows [S0,S1) of a host BlockBandedMatrix."""
    from .matrices import BlockSkylineMatrix
    l, u = int(parent.l[0]) if len(parent.l) else 0, int(parent.u[0]) if len(parent.u) else 0
    Jlo, Jhi, ll, uu = sky_rowrange_structure(parent.rows, parent.cols, l, u, S0, S1)
    loc = BlockSkylineMatrix.zeros(parent.dtype, parent.rows[S0:S1], parent.cols[Jlo:Jhi], ll, uu)
    for Jp in range(Jhi - Jlo):
        for Kp in loc.blockcolsupport(Jp):
            blk = parent.block(S0 + Kp, Jlo + Jp)
            s, ld = loc.block_start(Kp, Jp), int(loc.block_strides[Jp])
            for j in range(blk.shape[1]):
                loc.data[s + j * ld:s + j * ld + blk.shape[0]] = blk[:, j]
    return loc, Jlo, Jhi


def owned_cols(Kbounds, Nc, p):
    """Block columns of x (= block rows of a right factor B) owned by rank p (same rule as bbm_bbb_dist_layout)."""
    nranks = len(Kbounds) - 1
    J0 = min(int(Kbounds[p]), Nc)
    J1 = Nc if p == nranks - 1 else min(int(Kbounds[p + 1]), Nc)
    return J0, max(J0, J1)


def matmul_plan(rowsA, colsA, lA, uA, colsB, lB, uB, Kbounds, rank):
    """Plan of rank `rank` for C[K0:K1,:] = A[K0:K1,:] * B with A, C partitioned by block rows and B's
    block rows owned like x.  Returns a dict with the stored ranges and the halo transfer lists:
    sends / recvs = {peer: [(N, J), ...]} blocks of B in a fixed (N, J) order both sides agree on."""
    nranks = len(Kbounds) - 1
    NrA, NcA, NcB = len(rowsA), len(colsA), len(colsB)

    def stored_B_rows(p):   # B block rows rank p multiplies with = column hull of its A rows
        K0, K1 = int(Kbounds[p]), int(Kbounds[p + 1])
        Jlo, Jhi, _, _ = sky_rowrange_structure(rowsA, colsA, lA, uA, K0, K1)
        return Jlo, Jhi

    def rowsupport_B(N):
        return range(max(N - lB, 0), min(N + uB, NcB - 1) + 1)

    K0, K1 = int(Kbounds[rank]), int(Kbounds[rank + 1])
    S0, S1 = stored_B_rows(rank)
    O0, O1 = owned_cols(Kbounds, NcA, rank)
    sends, recvs = {}, {}
    for q in range(nranks):
        if q == rank:
            continue
        q0, q1 = owned_cols(Kbounds, NcA, q)
        t0, t1 = stored_B_rows(q)
        need = [(N, J) for N in range(max(S0, q0), min(S1, q1)) for J in rowsupport_B(N) if colsA[N] and colsB[J]]
        give = [(N, J) for N in range(max(t0, O0), min(t1, O1)) for J in rowsupport_B(N) if colsA[N] and colsB[J]]
        if need:
            recvs[q] = need
        if give:
            sends[q] = give
    return {"K0": K0, "K1": K1, "S0": S0, "S1": S1, "O0": O0, "O1": O1, "sends": sends, "recvs": recvs}


class DistBlockBandedProduct:
    """C[K0:K1,:] <- alpha*A[K0:K1,:]*B + beta*C[K0:K1,:] on one rank of a row-block partition.

    A_loc, C_loc hold the rank's own block rows, B_loc holds the block rows [S0,S1) the rank multiplies
    with; rows of B_loc outside the owned range [O0,O1) are halo and are refreshed by `exchange()`
    (packed blocks over torch.distributed isend/irecv -- NCCL between GPUs).  Storage is torch tensors
    so that packing is ordinary strided copies; the product is the single-GPU DMMA kernel."""

    def __init__(self, handle, dist, torch, rows, cols, lA, uA, colsB, lB, uB, Kbounds, rank, dtype=np.float64):
        from .matrices import BlockSkylineMatrix
        from .linalg import add_bandwidths
        self.h, self.dist, self.torch, self.rank = handle, dist, torch, rank
        self.rows, self.cols, self.colsB = _sizes(rows), _sizes(cols), _sizes(colsB)
        self.plan = matmul_plan(self.rows, self.cols, lA, uA, self.colsB, lB, uB, Kbounds, rank)
        P = self.plan
        tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
        dev = torch.device("cuda", handle.device) if handle is not None else torch.device("cpu")
        self.dev = dev

        def make(rws, cls, l, u, S0, S1):
            Jlo, Jhi, ll, uu = sky_rowrange_structure(rws, cls, l, u, S0, S1)
            n = BlockSkylineMatrix.numentries(rws[S0:S1], cls[Jlo:Jhi], ll, uu)
            t = torch.zeros(max(n, 1), dtype=tdt, device=dev)
            if handle is not None:
                from .device import DeviceArray
                d = DeviceArray(handle, (n,), np.dtype(dtype), ptr=t.data_ptr())
            else:
                d = t.numpy()[:n]
            return BlockSkylineMatrix(d, rws[S0:S1], cls[Jlo:Jhi], ll, uu), t, Jlo

        self.A, self.tA, self.JloA = make(self.rows, self.cols, lA, uA, P["K0"], P["K1"])
        self.B, self.tB, self.JloB = make(self.cols, self.colsB, lB, uB, P["S0"], P["S1"])
        self.C, self.tC, self.JloC = make(self.rows, self.colsB, lA + lB, uA + uB, P["K0"], P["K1"])

    # ---- block views of the local B storage -------------------------------------------------
    def _b_block(self, N, J):
        Kp, Jp = N - self.plan["S0"], J - self.JloB
        s, ld = self.B.block_start(Kp, Jp), int(self.B.block_strides[Jp])
        rN, cJ = int(self.cols[N]), int(self.colsB[J])
        return self.tB.as_strided((cJ, rN), (ld, 1), s)     # [j][i] view of the column-major block

    def exchange(self):
        """Refresh the halo block rows of B_loc from their owners."""
        torch, dist = self.torch, self.dist
        ops, unpack = [], []
        for q, blocks in sorted(self.plan["recvs"].items()):
            n = sum(int(self.cols[N]) * int(self.colsB[J]) for N, J in blocks)
            buf = torch.empty(n, dtype=self.tB.dtype, device=self.dev)
            ops.append(dist.P2POp(dist.irecv, buf, q))
            unpack.append((buf, blocks))
        for q, blocks in sorted(self.plan["sends"].items()):
            buf = torch.cat([self._b_block(N, J).reshape(-1) for N, J in blocks])
            ops.append(dist.P2POp(dist.isend, buf, q))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for buf, blocks in unpack:
            o = 0
            for N, J in blocks:
                v = self._b_block(N, J)
                v.copy_(buf[o:o + v.numel()].view(v.shape))
                o += v.numel()

    def mul_(self, alpha=1.0, beta=0.0):
        from .linalg import mul_
        return mul_(self.C, self.A, self.B, alpha, beta)


class DistBlockBandedMatvec:
    """y[K0:K1) <- alpha*A[K0:K1,:]*x + beta*y[K0:K1) for a row-block-partitioned BlockBandedMatrix.

    A_loc holds the rank's own block rows (a ragged BlockSkylineMatrix, see sky_rowrange_structure);
    x_local = [halo | owned | halo] covers the block columns A_loc touches, the halo parts are refreshed
    by `exchange()` (torch.distributed isend/irecv -- NCCL between GPUs) with the same send/recv lists as
    the banded-block-banded plan (they depend on the block structure and (l,u) only)."""

    def __init__(self, handle, dist, torch, rows, cols, l, u, Kbounds, rank, dtype=np.float64):
        from .matrices import BlockSkylineMatrix
        self.h, self.dist, self.torch, self.rank = handle, dist, torch, rank
        self.rows, self.cols = _sizes(rows), _sizes(cols)
        world = len(Kbounds) - 1
        self.layout, self.sends, self.recvs = layout(self.rows, self.cols, l, u, world, rank, Kbounds)
        lay = self.layout
        K0, K1 = lay["K0"], lay["K1"]
        Jlo, Jhi, ll, uu = sky_rowrange_structure(self.rows, self.cols, l, u, K0, K1)
        # the skyline sub-structure only spans the touched columns; x_local spans the layout's hull
        self.Jlo, self.Jhi = Jlo, Jhi
        C = _cum(self.cols)
        self.x_off = int(C[Jlo] - lay["col_offset"]) if Jhi > Jlo else 0   # where A_loc's columns start in x_local
        tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
        self.dev = torch.device("cuda", handle.device) if handle is not None else torch.device("cpu")
        n = BlockSkylineMatrix.numentries(self.rows[K0:K1], self.cols[Jlo:Jhi], ll, uu)
        self.tA = torch.zeros(max(n, 1), dtype=tdt, device=self.dev)
        self.x_local = torch.zeros(max(lay["n_local"], 1), dtype=tdt, device=self.dev)
        self.y_own = torch.zeros(max(lay["m_own"], 1), dtype=tdt, device=self.dev)
        if handle is not None:
            from .device import DeviceArray
            data = DeviceArray(handle, (n,), np.dtype(dtype), ptr=self.tA.data_ptr())
        else:
            data = self.tA.numpy()[:n]
        self.A = BlockSkylineMatrix(data, self.rows[K0:K1], self.cols[Jlo:Jhi], ll, uu)
        self._call = None

    def exchange(self):
        dist, ops = self.dist, []
        for peer, off, cnt in self.recvs:
            ops.append(dist.P2POp(dist.irecv, self.x_local[off:off + cnt], peer))
        for peer, off, cnt in self.sends:
            ops.append(dist.P2POp(dist.isend, self.x_local[off:off + cnt], peer))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def mul_(self, alpha=1.0, beta=0.0):
        """Device path: y_own <- alpha*A_loc*x_local[cols of A_loc] + beta*y_own (after exchange()).
        One C-ABI call with cached arguments: at 8 GPUs the per-rank kernel is ~0.1 ms, shorter than the generic
        `mul_` wrapper's host-side checks."""
        import ctypes as C
        from . import _lib as L
        if self._call is None:
            n_loc, m = self.A.shape[1], self.A.shape[0]
            es = self.tA.element_size()
            f64 = np.dtype(self.A.dtype) == np.float64
            ct = C.c_double if f64 else C.c_float
            fn = getattr(self.h._lib, "bbm_sky_mul_f64" if f64 else "bbm_sky_mul_f32")
            args = (self.h.h, self.A.descriptor(self.h), C.c_void_p(self.A.data.ptr),
                    C.c_void_p(self.x_local.data_ptr() + self.x_off * es), max(1, n_loc), 1, C.c_void_p(self.y_own.data_ptr()), max(1, m))
            self._call = (fn, ct, args)
        fn, ct, a = self._call
        L.check(fn(a[0], a[1], ct(alpha), a[2], a[3], a[4], a[5], ct(beta), a[6], a[7]), self.h.h, "mul!")
        return self.y_own


# ==================================================================================================
# The same two sharded BlockBandedMatrix operations through the C ABI (bbm_dist_sky_*): planning, halo
# pack / exchange (ncclSend/ncclRecv inside the library) and the overlapped local kernels all live in
# csrc/distsky.cu.  The classes above stay as the torch/gloo mirror the CPU tests drive.
# ==================================================================================================
def rowrange_structure_c(Nr, Nc, l, u, S0, S1):
    """``bbm_sky_rowrange_structure``: (Jlo, Jhi, lvec, uvec) of block rows [S0,S1) -- host-only, no GPU."""
    lib = L.load()
    jlo, jhi = C.c_int64(), C.c_int64()
    L.check(lib.bbm_sky_rowrange_structure(Nr, Nc, l, u, S0, S1, C.byref(jlo), C.byref(jhi), None, None), None, "bbm_sky_rowrange_structure")
    n = jhi.value - jlo.value
    lv, uv = np.zeros(max(n, 1), dtype=np.int64), np.zeros(max(n, 1), dtype=np.int64)
    L.check(lib.bbm_sky_rowrange_structure(Nr, Nc, l, u, S0, S1, C.byref(jlo), C.byref(jhi), lv.ctypes.data_as(C.c_void_p),
                                           uv.ctypes.data_as(C.c_void_p)), None, "bbm_sky_rowrange_structure")
    return jlo.value, jhi.value, lv[:n], uv[:n]


def mm_layout_c(colsA, colsB, NrA, lA, uA, lB, uB, Kbounds, rank):
    """``bbm_dist_sky_mm_layout``: ({K0,K1,S0,S1,O0,O1}, blocks sent to each rank, blocks received from each rank)."""
    cA, cB = _sizes(colsA), _sizes(colsB)
    kb = np.ascontiguousarray(Kbounds, dtype=np.int64)
    nranks = len(kb) - 1
    rng6 = np.zeros(6, dtype=np.int64)
    sb, rb = np.zeros(nranks, dtype=np.int64), np.zeros(nranks, dtype=np.int64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    L.check(L.load().bbm_dist_sky_mm_layout(NrA, len(cA), len(cB), p(cA), p(cB), lA, uA, lB, uB, nranks, rank, p(kb), p(rng6), p(sb), p(rb)),
            None, "bbm_dist_sky_mm_layout")
    return dict(zip(("K0", "K1", "S0", "S1", "O0", "O1"), (int(v) for v in rng6))), sb, rb


class DistBlockBandedMatrix:
    """Rank-local part of a row-block-partitioned BlockBandedMatrix(rows, cols, (l,u)) for ``mul!`` (``bbm_dist_sky_plan_*``).

    ``local`` is the BlockSkylineMatrix structure of the rank's block rows (its ``data`` a device vector of
    ``numentries`` entries, filled with ``set_local_data``); ``x_local`` / ``y_own`` are device buffers of the layout's
    sizes like in the banded-block-banded case."""

    def __init__(self, comm: Communicator, rows, cols, l, u, dtype=np.float64, Kbounds=None):
        from .matrices import BlockSkylineMatrix
        self.comm = comm
        self.rows, self.cols = _sizes(rows), _sizes(cols)
        self.dtype = np.dtype(dtype)
        h = comm.handle
        kb = None if Kbounds is None else np.ascontiguousarray(Kbounds, dtype=np.int64)
        p = C.c_void_p()
        L.check(h._lib.bbm_dist_sky_plan_create(comm.c, len(self.rows), len(self.cols), self.rows.ctypes.data_as(C.c_void_p),
                                                self.cols.ctypes.data_as(C.c_void_p), int(l), int(u),
                                                None if kb is None else kb.ctypes.data_as(C.c_void_p), C.byref(p)), h.h, "bbm_dist_sky_plan_create")
        self.plan = p
        lay = L.DistLayout()
        jlo, jhi, ne, xo = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        L.check(h._lib.bbm_dist_sky_plan_layout(p, C.byref(lay), C.byref(jlo), C.byref(jhi), C.byref(ne), C.byref(xo)), h.h)
        self.layout = lay.asdict()
        self.Jlo, self.Jhi, self.numentries, self.x_off = jlo.value, jhi.value, ne.value, xo.value
        n = self.Jhi - self.Jlo
        lv, uv = np.zeros(max(n, 1), dtype=np.int64), np.zeros(max(n, 1), dtype=np.int64)
        L.check(h._lib.bbm_dist_sky_plan_bandwidths(p, lv.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p)), h.h)
        K0, K1 = self.layout["K0"], self.layout["K1"]
        self.local = BlockSkylineMatrix(h.empty(self.numentries, self.dtype), self.rows[K0:K1], self.cols[self.Jlo:self.Jhi], lv[:n], uv[:n])

    def set_local_data(self, vec: np.ndarray):
        if vec.shape != (self.numentries,):
            raise L.DimensionMismatch(f"local panels hold {self.numentries} entries, got {vec.shape}")
        if vec.size:
            self.local.data.copy_from_host(np.ascontiguousarray(vec, dtype=self.dtype))
        return self

    def mul_(self, y_own, x_local, alpha=1.0, beta=0.0, nrhs=1):
        """y_own <- alpha*A[K0:K1,:]*x + beta*y_own (collective; asynchronous on the handle's stream)."""
        h = self.comm.handle
        suf, ct = ("f64", C.c_double) if self.dtype == np.float64 else ("f32", C.c_float)
        fn = getattr(h._lib, f"bbm_dist_sky_mul_{suf}")
        L.check(fn(self.plan, ct(alpha), C.c_void_p(self.local.data.ptr), C.c_void_p(x_local.ptr), nrhs, ct(beta), C.c_void_p(y_own.ptr)),
                h.h, "bbm_dist_sky_mul")
        return y_own

    def close(self):
        if getattr(self, "plan", None):
            self.comm.handle._lib.bbm_dist_sky_plan_destroy(self.plan)
            self.plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DistBlockBandedMatmul:
    """C[K0:K1,:] <- alpha*A[K0:K1,:]*B + beta*C[K0:K1,:] on one rank of a row-block partition (``bbm_dist_sky_mm_plan_*``).

    ``A``, ``B``, ``C`` are the rank-local BlockSkylineMatrix structures with device storage: the rank's own block rows of
    A and C, and the block rows ``[S0,S1)`` of B it multiplies with (rows outside ``[O0,O1)`` are halo, refreshed by every
    ``mul_``)."""

    def __init__(self, comm: Communicator, rowsA, colsA, colsB, luA, luB, dtype=np.float64, Kbounds=None):
        from .matrices import BlockSkylineMatrix
        self.comm = comm
        self.rowsA, self.colsA, self.colsB = _sizes(rowsA), _sizes(colsA), _sizes(colsB)
        self.dtype = np.dtype(dtype)
        h = comm.handle
        kb = None if Kbounds is None else np.ascontiguousarray(Kbounds, dtype=np.int64)
        p = C.c_void_p()
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        L.check(h._lib.bbm_dist_sky_mm_plan_create(comm.c, len(self.rowsA), len(self.colsA), len(self.colsB), ptr(self.rowsA), ptr(self.colsA),
                                                   ptr(self.colsB), int(luA[0]), int(luA[1]), int(luB[0]), int(luB[1]),
                                                   None if kb is None else ptr(kb), C.byref(p)), h.h, "bbm_dist_sky_mm_plan_create")
        self.plan = p
        rng6, jlo, jhi, ne = (np.zeros(k, dtype=np.int64) for k in (6, 3, 3, 3))
        L.check(h._lib.bbm_dist_sky_mm_plan_info(p, ptr(rng6), ptr(jlo), ptr(jhi), ptr(ne)), h.h)
        self.ranges = dict(zip(("K0", "K1", "S0", "S1", "O0", "O1"), (int(v) for v in rng6)))
        R = self.ranges
        axes = ((self.rowsA[R["K0"]:R["K1"]], self.colsA), (self.colsA[R["S0"]:R["S1"]], self.colsB), (self.rowsA[R["K0"]:R["K1"]], self.colsB))
        mats = []
        for t in range(3):
            n = int(jhi[t] - jlo[t])
            lv, uv = np.zeros(max(n, 1), dtype=np.int64), np.zeros(max(n, 1), dtype=np.int64)
            L.check(h._lib.bbm_dist_sky_mm_plan_bandwidths(p, t, ptr(lv), ptr(uv)), h.h)
            mats.append(BlockSkylineMatrix(h.empty(int(ne[t]), self.dtype), axes[t][0], axes[t][1][int(jlo[t]):int(jhi[t])], lv[:n], uv[:n]))
        self.A, self.B, self.C = mats
        self.Jlo, self.Jhi = [int(v) for v in jlo], [int(v) for v in jhi]

    def mul_(self, alpha=1.0, beta=0.0):
        h = self.comm.handle
        suf, ct = ("f64", C.c_double) if self.dtype == np.float64 else ("f32", C.c_float)
        fn = getattr(h._lib, f"bbm_dist_sky_matmul_{suf}")
        L.check(fn(self.plan, ct(alpha), C.c_void_p(self.A.data.ptr), C.c_void_p(self.B.data.ptr), ct(beta), C.c_void_p(self.C.data.ptr)),
                h.h, "bbm_dist_sky_matmul")
        return self.C

    def close(self):
        if getattr(self, "plan", None):
            self.comm.handle._lib.bbm_dist_sky_mm_plan_destroy(self.plan)
            self.plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
