"""Host-side mirror of the reference's two containers.

The reference keeps constructors, block axes and bandwidth queries in host Julia and only the
``data`` buffer moves to the device (a package extension keyed on the buffer type).  These classes
play that role for the Python harness: same field names, same storage shapes, same queries
(0-based where Julia is 1-based).  ``data`` is either a numpy array (host matrix; only
construction / indexing / densification work) or a :class:`DeviceArray` (device-resident; the
products run through the C ABI).

Reference: src/BandedBlockBandedMatrix.jl:25-50 (struct, data axes), :277-304 (bandwidth getters),
:503-506 (block accessor); src/BlockSkylineMatrix.jl:47-53,158-173 (sizes + struct),
:6-43,93-104 (starts/strides/numentries); src/AbstractBlockBandedMatrix.jl:90-104 (supports).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .device import DeviceArray, Handle


def _sizes(v) -> np.ndarray:
    a = np.asarray(v, dtype=np.int64).reshape(-1)
    if (a < 0).any():
        raise ValueError("block sizes must be non-negative")
    return np.ascontiguousarray(a)


def _cum(sz: np.ndarray) -> np.ndarray:
    out = np.zeros(len(sz) + 1, dtype=np.int64)
    np.cumsum(sz, out=out[1:])
    return out


class _BlockAxes:
    """Block axes shared by both containers (``axes(A)`` / ``blocksize`` in the reference)."""

    def _init_axes(self, rows, cols):
        self.rows, self.cols = _sizes(rows), _sizes(cols)
        self.R, self.C = _cum(self.rows), _cum(self.cols)

    @property
    def shape(self):
        return (int(self.R[-1]), int(self.C[-1]))

    @property
    def blockshape(self):
        """``blocksize(A)``: number of row / column blocks."""
        return (len(self.rows), len(self.cols))

    def _find(self, prefix, i):
        if not (0 <= i < prefix[-1]):
            raise IndexError(f"index {i} out of bounds")
        return int(np.searchsorted(prefix, i, side="right") - 1)

    @property
    def on_device(self) -> bool:
        return isinstance(self.data, DeviceArray)


# ==================================================================================================
class BandedBlockBandedMatrix(_BlockAxes):
    """``_BandedBlockBandedMatrix(data, rows, cols, (l,u), (lam,mu))``.

    ``data`` is the dense column-major ``P x n`` band-of-bands array, ``P = (l+u+1)(lam+mu+1)``
    (0 rows when ``-l > u`` or ``-lam > mu``), exactly the parent array of the reference's
    ``BlockedMatrix``; entry ``A[R[K]+k, C[J]+j]`` is ``data[(u+K-J)*w + (mu+k-j), C[J]+j]``.
    """

    def __init__(self, data, rows, cols, lu, lm):
        self._init_axes(rows, cols)
        self.l, self.u = int(lu[0]), int(lu[1])
        self.lam, self.mu = int(lm[0]), int(lm[1])
        exp = self.data_shape(self.cols, lu, lm)
        if tuple(data.shape) != exp:  # check_data_sizes, src/BandedBlockBandedMatrix.jl:1-10
            raise ValueError(f"Data matrix must be {exp[0]} x {exp[1]} (number of block bands x sub-block bands), got {tuple(data.shape)}")
        if isinstance(data, np.ndarray) and data.size and not data.flags.f_contiguous:
            data = np.asfortranarray(data)
        self.data = data
        self._desc = None

    # ---- constructors ------------------------------------------------------------------------
    @staticmethod
    def data_shape(cols, lu, lm):
        nb, w = max(0, lu[0] + lu[1] + 1), max(0, lm[0] + lm[1] + 1)
        return (nb * w, int(np.sum(np.asarray(cols, dtype=np.int64))))

    @classmethod
    def undef(cls, dtype, rows, cols, lu, lm):
        """``BandedBlockBandedMatrix{T}(undef, rows, cols, (l,u), (lam,mu))``."""
        return cls(np.empty(cls.data_shape(cols, lu, lm), dtype=dtype, order="F"), rows, cols, lu, lm)

    @classmethod
    def zeros(cls, dtype, rows, cols, lu, lm):
        return cls(np.zeros(cls.data_shape(cols, lu, lm), dtype=dtype, order="F"), rows, cols, lu, lm)

    @classmethod
    def from_dense(cls, A, rows, cols, lu, lm, fill=0.0):
        """``BandedBlockBandedMatrix(A, rows, cols, (l,u), (lam,mu))``: keep the in-band entries of A.
        Unused storage slots get ``fill`` (the reference leaves them undefined)."""
        A = np.asarray(A)
        out = cls(np.full(cls.data_shape(cols, lu, lm), fill, dtype=A.dtype, order="F"), rows, cols, lu, lm)
        if A.shape != out.shape:
            raise L.DimensionMismatch(f"dense matrix {A.shape} vs block axes {out.shape}")
        for rho, gi, gj in out._slot_map():
            out.data[rho, gj] = A[gi, gj]
        return out

    @classmethod
    def finitedifference_2d(cls, n: int, dtype=np.float64, nblocks: int | None = None):
        """``finitedifference_2d(n)`` of examples/finitedifference_2d.jl:9-15: the 5-point Laplacian
        ``Kron(D2, I) + Kron(I, D2)`` with h = 1/n as a BandedBlockBandedMatrix, n blocks of size n
        (``nblocks`` block rows/columns of size n when given), (l,u) = (lam,mu) = (1,1).

        Band storage: every column holds n^2*[0,1,0, 1,-4,1, 0,1,0] (slab 0 = super-diagonal block
        = identity, slab 1 = diagonal block = tridiag(1,-4,1), slab 2 = sub-diagonal block).  Slots
        that fall outside a block are unused; like examples/cuarrays.jl:6-24 they keep the value."""
        nb = n if nblocks is None else int(nblocks)
        dtype = np.dtype(dtype)
        col = (np.array([0, 1, 0, 1, -4, 1, 0, 1, 0], dtype=np.float64) * float(n) ** 2).astype(dtype)
        data = np.empty((9, n * nb), dtype=dtype, order="F")
        data[:, :] = col[:, None]
        return cls(data, [n] * nb, [n] * nb, (1, 1), (1, 1))

    # ---- queries -----------------------------------------------------------------------------
    @property
    def dtype(self):
        return np.dtype(self.data.dtype)

    @property
    def blockbandwidths(self):
        return (self.l, self.u)

    @property
    def subblockbandwidths(self):
        return (self.lam, self.mu)

    def blockcolsupport(self, J: int) -> range:
        Nr = len(self.rows)
        return range(max(J - self.u, 0), max(min(J + self.l, Nr - 1), -1) + 1)

    def blockrowsupport(self, K: int) -> range:
        Nc = len(self.cols)
        return range(max(K - self.l, 0), max(min(K + self.u, Nc - 1), -1) + 1)

    def _slot_map(self):
        """Yield (data row, global rows, global cols) index arrays, one per row of ``data``, for
        the slots that encode a matrix entry."""
        w, nb = self.lam + self.mu + 1, self.l + self.u + 1
        if w <= 0 or nb <= 0 or self.shape[1] == 0:
            return
        Nr = len(self.rows)
        gcol = np.arange(self.shape[1], dtype=np.int64)
        Jc = np.repeat(np.arange(len(self.cols), dtype=np.int64), self.cols)
        jc = gcol - self.C[Jc]
        for s in range(nb):
            K = Jc + s - self.u
            okK = (K >= 0) & (K < Nr)
            Kc = np.clip(K, 0, max(Nr - 1, 0))
            rK = self.rows[Kc] if Nr else np.zeros_like(Kc)
            for i in range(w):
                k = jc + i - self.mu
                ok = okK & (k >= 0) & (k < rK)
                yield s * w + i, (self.R[Kc] + k)[ok], gcol[ok]

    def valid_mask(self) -> np.ndarray:
        mask = np.zeros(self.data_shape(self.cols, (self.l, self.u), (self.lam, self.mu)), dtype=bool, order="F")
        for rho, gi, gj in self._slot_map():
            mask[rho, gj] = True
        return mask

    def _host_data(self) -> np.ndarray:
        return self.data.to_host() if self.on_device else self.data

    def to_dense(self) -> np.ndarray:
        """``Matrix(A)``."""
        data = self._host_data()
        A = np.zeros(self.shape, dtype=self.dtype)
        for rho, gi, gj in self._slot_map():
            A[gi, gj] = data[rho, gj]
        return A

    def __getitem__(self, ij):
        i, j = ij
        K, J = self._find(self.R, i), self._find(self.C, j)
        k, jj = i - int(self.R[K]), j - int(self.C[J])
        if not (-self.l <= J - K <= self.u) or not (-self.lam <= jj - k <= self.mu):
            return self.dtype.type(0)
        w = self.lam + self.mu + 1
        return self._host_data()[(self.u + K - J) * w + (self.mu + k - jj), j]

    def block(self, K: int, J: int) -> np.ndarray:
        """``A[Block(K), Block(J)]`` as a dense array (zeros outside the block band)."""
        out = np.zeros((int(self.rows[K]), int(self.cols[J])), dtype=self.dtype)
        if not (-self.l <= J - K <= self.u) or self.lam + self.mu + 1 <= 0:
            return out
        w = self.lam + self.mu + 1
        band = self._host_data()[(self.u + K - J) * w:(self.u + K - J + 1) * w, self.C[J]:self.C[J + 1]]
        for jj in range(out.shape[1]):
            for k in range(max(0, jj - self.mu), min(out.shape[0] - 1, jj + self.lam) + 1):
                out[k, jj] = band[self.mu + k - jj, jj]
        return out

    def view(self, block_rows, block_cols) -> "BandedBlockBandedMatrix":
        """``view(A, Block.(K0:K1-1), Block.(J0:J1-1))`` for contiguous block ranges (0-based ``range``s or ints):
        again a BandedBlockBandedMatrix over the SAME storage -- the column slab ``C[J0]:C[J1]`` of ``data`` --
        with the block bandwidths shifted to ``(l-K0+J0, u+K0-J0)`` (test/test_triblockbanded.jl:55-78,
        layout BandedBlockBandedColumnMajor of block-range sub-views).  No copy, host or device."""
        def rng(b, n):
            if isinstance(b, (int, np.integer)):
                b = range(int(b), int(b) + 1)
            if not isinstance(b, range) or b.step != 1 or b.start < 0 or b.stop > n or b.stop < b.start:
                raise IndexError(f"block range {b!r} is not a contiguous range inside 0:{n}")
            return b
        br, bc = rng(block_rows, len(self.rows)), rng(block_cols, len(self.cols))
        K0, J0 = br.start, bc.start
        c0, c1 = int(self.C[bc.start]), int(self.C[bc.stop])
        P = self.data.shape[0]
        if isinstance(self.data, np.ndarray):
            d = self.data[:, c0:c1]
        else:
            d = DeviceArray(self.data.handle, (P, c1 - c0), self.dtype, ptr=self.data.ptr + P * c0 * self.dtype.itemsize)
            d._keep = self.data
        return BandedBlockBandedMatrix(d, self.rows[br.start:br.stop], self.cols[bc.start:bc.stop],
                                       (self.l - K0 + J0, self.u + K0 - J0), (self.lam, self.mu))

    # ---- device residency ----------------------------------------------------------------------
    def to_device(self, handle: Handle) -> "BandedBlockBandedMatrix":
        """Device-resident copy: one memcpy of the parent array, shape unchanged."""
        if self.on_device:
            return self
        d = handle.to_device(self.data) if self.data.size else handle.empty(self.data.shape, self.dtype)
        out = BandedBlockBandedMatrix(d, self.rows, self.cols, (self.l, self.u), (self.lam, self.mu))
        return out

    def descriptor(self, handle: Handle):
        if self._desc is None or self._desc[0] is not handle:
            d = C.c_void_p()
            L.check(handle._lib.bbm_bbb_desc_create(
                handle.h, len(self.rows), len(self.cols), self.rows.ctypes.data_as(C.c_void_p),
                self.cols.ctypes.data_as(C.c_void_p), self.l, self.u, self.lam, self.mu, C.byref(d)), handle.h,
                "bbm_bbb_desc_create")
            self._desc = (handle, d)
        return self._desc[1]

    def __del__(self):
        try:
            if self._desc is not None and self._desc[0].h:
                self._desc[0]._lib.bbm_bbb_desc_destroy(self._desc[1])
        except Exception:
            pass

    def __matmul__(self, x):
        from .linalg import mul
        return mul(self, x)


# ==================================================================================================
def _bwvec(b, Nc) -> np.ndarray:
    b = np.asarray(b, dtype=np.int64)
    if b.ndim == 0:
        b = np.full(Nc, int(b), dtype=np.int64)
    b = np.ascontiguousarray(b.reshape(-1))
    if len(b) != Nc:  # checkbandwidths, src/BlockSkylineMatrix.jl:1-3
        raise L.DimensionMismatch(f"{Nc} lower and upper column bandwidths are required, got {len(b)}")
    return b


class BlockSkylineMatrix(_BlockAxes):
    """``_BlockSkylineMatrix(data, BlockSkylineSizes(rows, cols, l, u))``; a ``BlockBandedMatrix`` when
    ``l``, ``u`` are scalars (``Fill`` bandwidths, src/BlockSkylineMatrix.jl:55-56, 173).

    ``data`` is the flat vector of column-block-contiguous panels: block column J stores block
    rows ``max(J-u[J],0) .. min(J+l[J],Nr-1)`` as one dense column-major panel of
    ``block_strides[J]`` rows (src/BlockSkylineMatrix.jl:6-43, 93-104)."""

    def __init__(self, data, rows, cols, l, u):
        self._init_axes(rows, cols)
        Nr, Nc = len(self.rows), len(self.cols)
        self.constant_bandwidths = np.ndim(l) == 0 and np.ndim(u) == 0
        self.l, self.u = _bwvec(l, Nc), _bwvec(u, Nc)
        J = np.arange(Nc, dtype=np.int64)
        self.klo = np.maximum(J - self.u, 0)
        hi = np.minimum(J + self.l, Nr - 1)
        self.khi = np.where(hi >= self.klo, hi, self.klo - 1)
        Rhi = self.R[np.clip(self.khi + 1, 0, Nr)]
        Rlo = self.R[np.clip(self.klo, 0, Nr)]
        self.block_strides = np.where(self.khi >= self.klo, Rhi - Rlo, 0).astype(np.int64)
        self.panel_offsets = np.zeros(Nc + 1, dtype=np.int64)
        np.cumsum(self.block_strides * self.cols, out=self.panel_offsets[1:])
        n = int(self.panel_offsets[-1])
        if int(np.prod(data.shape)) != n or len(data.shape) != 1:
            raise ValueError(f"data must be a vector of bb_numentries = {n} entries, got shape {tuple(data.shape)}")
        self.data = data
        self._desc = None

    # ---- constructors ------------------------------------------------------------------------
    @staticmethod
    def _numentries(rows, cols, l, u) -> int:
        """``bb_numentries`` (src/BlockSkylineMatrix.jl:93-104)."""
        r, c = _sizes(rows), _sizes(cols)
        R = _cum(r)
        Nr, Nc = len(r), len(c)
        lv, uv = _bwvec(l, Nc), _bwvec(u, Nc)
        J = np.arange(Nc, dtype=np.int64)
        lo = np.maximum(J - uv, 0)
        hi = np.minimum(J + lv, Nr - 1)
        S = np.where(hi >= lo, R[np.clip(hi + 1, 0, Nr)] - R[np.clip(lo, 0, Nr)], 0)
        return int(np.sum(S * c))

    numentries = _numentries

    @classmethod
    def undef(cls, dtype, rows, cols, l, u):
        """``BlockSkylineMatrix{T}(undef, rows, cols, (l,u))`` / ``BlockBandedMatrix{T}(undef, ...)``."""
        return cls(np.empty(cls._numentries(rows, cols, l, u), dtype=dtype), rows, cols, l, u)

    @classmethod
    def zeros(cls, dtype, rows, cols, l, u):
        return cls(np.zeros(cls._numentries(rows, cols, l, u), dtype=dtype), rows, cols, l, u)

    @classmethod
    def from_dense(cls, A, rows, cols, l, u):
        A = np.asarray(A)
        out = cls.zeros(A.dtype, rows, cols, l, u)
        if A.shape != out.shape:
            raise L.DimensionMismatch(f"dense matrix {A.shape} vs block axes {out.shape}")
        for J in range(len(out.cols)):
            S = int(out.block_strides[J])
            if S == 0 or out.cols[J] == 0:
                continue
            r0 = int(out.R[out.klo[J]])
            out.data[out.panel_offsets[J]:out.panel_offsets[J + 1]] = \
                A[r0:r0 + S, out.C[J]:out.C[J + 1]].reshape(-1, order="F")
        return out

    def with_bandwidths(self, l, u) -> "BlockSkylineMatrix":
        """``BlockBandedMatrix(A, (l, u))``: a host copy in wider (or equal) block bandwidths, new blocks zero --
        what the reference's ``_qr`` does before factorising (src/blockskylineqr.jl:75: ``(l, l+u)``)."""
        if self.on_device:
            raise TypeError("with_bandwidths re-lays the storage out on the host: call it before to_device")
        out = BlockSkylineMatrix(np.zeros(self._numentries(self.rows, self.cols, l, u), dtype=self.dtype), self.rows, self.cols, l, u)
        for J in range(len(self.cols)):
            S0, S1, c = int(self.block_strides[J]), int(out.block_strides[J]), int(self.cols[J])
            if S0 == 0 or c == 0:
                continue
            if out.klo[J] > self.klo[J] or out.khi[J] < self.khi[J]:
                raise L.BandError("the new block bandwidths do not contain the old ones")
            src = self.data[int(self.panel_offsets[J]):int(self.panel_offsets[J]) + S0 * c].reshape((S0, c), order="F")
            dst = out.data[int(out.panel_offsets[J]):int(out.panel_offsets[J]) + S1 * c].reshape((S1, c), order="F")
            r0 = int(self.R[self.klo[J]] - out.R[out.klo[J]])
            dst[r0:r0 + S0, :] = src
        return out

    # ---- queries -----------------------------------------------------------------------------
    @property
    def dtype(self):
        return np.dtype(self.data.dtype)

    @property
    def blockbandwidths(self):
        """(max l, max u) (src/BlockSkylineMatrix.jl:371-372 ``blockbandwidths`` of a skyline)."""
        if len(self.l) == 0:
            return (0, 0)
        return (int(self.l.max()), int(self.u.max()))

    @property
    def colblockbandwidths(self):
        return self.l.copy(), self.u.copy()

    def blockcolsupport(self, J: int) -> range:
        return range(int(self.klo[J]), int(self.khi[J]) + 1)

    def has_block(self, K: int, J: int) -> bool:
        return self.klo[J] <= K <= self.khi[J]

    def block_start(self, K: int, J: int) -> int:
        """0-based offset of block (K,J) in ``data`` (``block_starts[K,J]-1`` of the reference)."""
        if not self.has_block(K, J):
            raise L.BandError(f"block ({K},{J}) is outside the block band")
        return int(self.panel_offsets[J] + self.R[K] - self.R[self.klo[J]])

    def _host_data(self) -> np.ndarray:
        return self.data.to_host() if self.on_device else self.data

    def block(self, K: int, J: int) -> np.ndarray:
        out = np.zeros((int(self.rows[K]), int(self.cols[J])), dtype=self.dtype)
        if not self.has_block(K, J) or out.size == 0:
            return out
        S = int(self.block_strides[J])
        panel = self._host_data()[self.panel_offsets[J]:self.panel_offsets[J + 1]].reshape((S, int(self.cols[J])), order="F")
        k0 = int(self.R[K] - self.R[self.klo[J]])
        return panel[k0:k0 + out.shape[0], :].copy()

    def __getitem__(self, ij):
        i, j = ij
        K, J = self._find(self.R, i), self._find(self.C, j)
        if not self.has_block(K, J):
            return self.dtype.type(0)
        return self._host_data()[self.block_start(K, J) + (i - int(self.R[K])) + (j - int(self.C[J])) * int(self.block_strides[J])]

    def to_dense(self) -> np.ndarray:
        data = self._host_data()
        A = np.zeros(self.shape, dtype=self.dtype)
        for J in range(len(self.cols)):
            S = int(self.block_strides[J])
            if S == 0 or self.cols[J] == 0:
                continue
            r0 = int(self.R[self.klo[J]])
            A[r0:r0 + S, self.C[J]:self.C[J + 1]] = data[self.panel_offsets[J]:self.panel_offsets[J + 1]].reshape(
                (S, int(self.cols[J])), order="F")
        return A

    # ---- device residency ----------------------------------------------------------------------
    def to_device(self, handle: Handle) -> "BlockSkylineMatrix":
        if self.on_device:
            return self
        d = handle.to_device(self.data) if self.data.size else handle.empty(self.data.shape, self.dtype)
        return self._like(d)

    def _like(self, data):
        l = int(self.l[0]) if self.constant_bandwidths and len(self.l) else (0 if self.constant_bandwidths else self.l)
        u = int(self.u[0]) if self.constant_bandwidths and len(self.u) else (0 if self.constant_bandwidths else self.u)
        return BlockSkylineMatrix(data, self.rows, self.cols, l, u)

    def descriptor(self, handle: Handle):
        if self._desc is None or self._desc[0] is not handle:
            d = C.c_void_p()
            L.check(handle._lib.bbm_sky_desc_create(
                handle.h, len(self.rows), len(self.cols), self.rows.ctypes.data_as(C.c_void_p),
                self.cols.ctypes.data_as(C.c_void_p), self.l.ctypes.data_as(C.c_void_p),
                self.u.ctypes.data_as(C.c_void_p), C.byref(d)), handle.h, "bbm_sky_desc_create")
            self._desc = (handle, d)
        return self._desc[1]

    def __del__(self):
        try:
            if self._desc is not None and self._desc[0].h:
                self._desc[0]._lib.bbm_sky_desc_destroy(self._desc[1])
        except Exception:
            pass

    def __matmul__(self, x):
        from .linalg import mul
        return mul(self, x)


def BlockBandedMatrix(data, rows, cols, lu):
    """``_BlockBandedMatrix(data, rows, cols, (l,u))``: a BlockSkylineMatrix with constant bandwidths."""
    return BlockSkylineMatrix(data, rows, cols, int(lu[0]), int(lu[1]))


class _Triangular:
    """``UpperTriangular(A)`` / ``UnitUpperTriangular(A)`` / ``LowerTriangular(A)`` / ``UnitLowerTriangular(A)``."""

    def __init__(self, parent, uplo: str, unit: bool):
        self.parent, self.uplo, self.unit = parent, uplo, bool(unit)

    def __matmul__(self, x):
        from .linalg import mul
        return mul(self, x)

    @property
    def blockbandwidths(self):
        """src/triblockbanded.jl:10-25: clip to (min(0,l),u) / (l,min(0,u)) when the blocks match."""
        l, u = self.parent.blockbandwidths
        P = self.parent
        if len(P.rows) == len(P.cols) and (P.rows == P.cols).all():
            return (min(0, l), u) if self.uplo == "U" else (l, min(0, u))
        return (l, u)


class BlockBandedQR:
    """``F = qr(A)`` of a BlockBandedMatrix as the reference stores it (``MatrixFactorizations.QR{T,<:BlockSkylineMatrix}``,
    src/blockskylineqr.jl:45-49): ``factors`` is block-skyline with block bandwidths ``(l, l+u)`` -- R in and above
    the diagonal blocks, the Householder vectors of panel K below them -- and ``tau`` the reflector scalars.
    The factorisation is the reference's host-side ``qr!``; this wrapper only carries a device-resident copy of
    its result to ``lmul_adj_`` (``lmul!(F.Q', B)``) and ``ldiv_`` (``ldiv!(F, B)`` / ``F \\ B``)."""

    def __init__(self, factors: "BlockSkylineMatrix", tau):
        n_tau = int(factors.C[-1]) if len(factors.rows) >= len(factors.cols) else int(factors.R[-1])
        if int(np.prod(tau.shape)) != n_tau:
            raise L.DimensionMismatch(f"tau has {tau.shape} entries, the factors need {n_tau}")
        self.factors, self.tau = factors, tau

    @property
    def shape(self):
        return self.factors.shape

    def to_device(self, handle: Handle) -> "BlockBandedQR":
        t = self.tau if isinstance(self.tau, DeviceArray) else handle.to_device(np.ascontiguousarray(self.tau))
        return BlockBandedQR(self.factors.to_device(handle), t)


class BlockBandedQL(BlockBandedQR):
    """``F = ql(A)`` (``MatrixFactorizations.QL{T,<:BlockSkylineMatrix}``, src/blockskylineqr.jl:51-72): ``factors`` has
    block bandwidths ``(l+u, u)`` -- L in and below the diagonal blocks, Householder vectors above (geqlf packing)."""

    def to_device(self, handle: Handle) -> "BlockBandedQL":
        t = self.tau if isinstance(self.tau, DeviceArray) else handle.to_device(np.ascontiguousarray(self.tau))
        return BlockBandedQL(self.factors.to_device(handle), t)


class Adjoint:
    """``A'`` / ``transpose(A)`` of a real BandedBlockBandedMatrix: reversed bandwidths
    (src/adjtransblockbanded.jl:2-3), products through the transposed kernel."""

    def __init__(self, parent):
        self.parent = parent

    @property
    def shape(self):
        return (self.parent.shape[1], self.parent.shape[0])

    @property
    def blockbandwidths(self):
        l, u = self.parent.blockbandwidths
        return (u, l)

    @property
    def subblockbandwidths(self):
        lam, mu = self.parent.subblockbandwidths
        return (mu, lam)

    def __matmul__(self, x):
        from .linalg import mul
        return mul(self, x)


def UpperTriangular(A):
    return _Triangular(A, "U", False)


def UnitUpperTriangular(A):
    return _Triangular(A, "U", True)


def LowerTriangular(A):
    return _Triangular(A, "L", False)


def UnitLowerTriangular(A):
    return _Triangular(A, "L", True)
