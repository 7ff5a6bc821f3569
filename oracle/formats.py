"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's two storage formats.

This module is part of the *oracle*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
(``blockbandedmatrices.jl_b200``) never does.

Everything here is 0-based (as the C ABI sees it); the reference is 1-based Julia.  Each
function cites the reference file:line (relative to /root/reference) whose behaviour it restates.

Parity status: *layout pinned* -- every integer golden value the reference's tests hold for
these index maps is reproduced in ``tests/test_oracle_golden.py``.  The *arithmetic* of the
reference (BlockArrays / ArrayLayouts / BandedMatrices / OpenBLAS, un-vendored, no Manifest)
is only pinned to "structured ~= dense" by the reference's own tests, and that is what the
oracle's product routines are checked against.

Notation: block-row sizes r[K], block-col sizes c[J], prefix sums R, C; w = lam+mu+1,
P = (l+u+1)*w.
"""
from __future__ import annotations

import numpy as np


def prefix(sizes) -> np.ndarray:
    sizes = np.asarray(sizes, dtype=np.int64).reshape(-1)
    out = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(sizes, out=out[1:])
    return out


# ----------------------------------------------------------------------------------------------
# loop bounds
# ----------------------------------------------------------------------------------------------
def blockcolsupport(Nr: int, J: int, l: int, u: int) -> range:
    """In-band block rows of block column J: max(J-u,0) .. min(J+l,Nr-1), possibly empty.

    src/AbstractBlockBandedMatrix.jl:90-91,103 (1-based ``max(J-u,1):max(min(J+l,N),0)``).
    """
    lo = max(J - u, 0)
    hi = max(min(J + l, Nr - 1), -1)
    return range(lo, hi + 1)


def blockrowsupport(Nc: int, K: int, l: int, u: int) -> range:
    """In-band block columns of block row K (src/AbstractBlockBandedMatrix.jl:92-93,104)."""
    lo = max(K - l, 0)
    hi = max(min(K + u, Nc - 1), -1)
    return range(lo, hi + 1)


def tri_blockbandwidths(l: int, u: int, uplo: str) -> tuple[int, int]:
    """Block bandwidths of Upper/LowerTriangular(A) with matching diagonal blocks.

    src/triblockbanded.jl:10-25.
    """
    if uplo.upper() == "U":
        return (min(0, l), u)
    return (l, min(0, u))


# ----------------------------------------------------------------------------------------------
# BandedBlockBandedMatrix: banded storage of bands
# ----------------------------------------------------------------------------------------------
def bbb_data_shape(cols, l, u, lam, mu) -> tuple[int, int]:
    """Shape of ``data``: (max(0,l+u+1)*max(0,lam+mu+1), sum(cols)).

    src/BandedBlockBandedMatrix.jl:50 (_bbb_data_axes).
    """
    return (max(0, l + u + 1) * max(0, lam + mu + 1), int(np.sum(np.asarray(cols, dtype=np.int64))))


def bbb_check_data(data: np.ndarray, cols, l, u, lam, mu) -> None:
    """src/BandedBlockBandedMatrix.jl:1-10 (check_data_sizes) for the dense-backed case."""
    exp = bbb_data_shape(cols, l, u, lam, mu)
    if tuple(data.shape) != exp:
        raise ValueError(f"data must be {exp}, got {tuple(data.shape)}")


def bbb_entry_pos(R, C, l, u, lam, mu, i: int, j: int):
    """(row, col) inside ``data`` holding A[i,j], or None if (i,j) is outside the bands.

    A[R[K]+k, C[J]+j] = data[(u+K-J)*w + (mu+k-j), C[J]+j]
    src/BandedBlockBandedMatrix.jl:338-346 (getindex), :503-506 (block accessor),
    :538-545 (bandeddata of a block view) and BandedMatrices' band storage (row mu+k-j).
    """
    K = int(np.searchsorted(R, i, side="right") - 1)
    J = int(np.searchsorted(C, j, side="right") - 1)
    k = i - int(R[K])
    jj = j - int(C[J])
    if not (-l <= J - K <= u):
        return None
    if not (-lam <= jj - k <= mu):
        return None
    w = lam + mu + 1
    return ((u + K - J) * w + (mu + k - jj), j)


def bbb_getindex(data, rows, cols, l, u, lam, mu, i: int, j: int):
    R, C = prefix(rows), prefix(cols)
    pos = bbb_entry_pos(R, C, l, u, lam, mu, i, j)
    if pos is None:
        return data.dtype.type(0)
    return data[pos]


def bbb_to_dense(data, rows, cols, l, u, lam, mu) -> np.ndarray:
    """Dense copy; never reads an unused slot of ``data`` (they may hold NaN)."""
    R, C = prefix(rows), prefix(cols)
    m, n = int(R[-1]), int(C[-1])
    A = np.zeros((m, n), dtype=data.dtype)
    w = lam + mu + 1
    if l + u + 1 <= 0 or w <= 0:
        return A
    Nr, Nc = len(R) - 1, len(C) - 1
    for J in range(Nc):
        cJ = int(C[J + 1] - C[J])
        for K in blockcolsupport(Nr, J, l, u):
            rK = int(R[K + 1] - R[K])
            s = u + K - J
            for jj in range(cJ):
                klo = max(0, jj - mu)
                khi = min(rK - 1, jj + lam)
                for k in range(klo, khi + 1):
                    A[R[K] + k, C[J] + jj] = data[s * w + (mu + k - jj), C[J] + jj]
    return A


def bbb_from_dense(A, rows, cols, l, u, lam, mu, fill=0.0) -> np.ndarray:
    """``BandedBlockBandedMatrix(A, rows, cols, (l,u), (lam,mu))``: keep the in-band part of A.

    src/BandedBlockBandedMatrix.jl:198-220.  Unused slots are set to ``fill`` (the reference
    leaves them undefined) so tests can poison them with NaN.
    """
    R, C = prefix(rows), prefix(cols)
    shape = bbb_data_shape(cols, l, u, lam, mu)
    data = np.full(shape, fill, dtype=A.dtype, order="F")
    w = lam + mu + 1
    if shape[0] == 0:
        return data
    Nr, Nc = len(R) - 1, len(C) - 1
    for J in range(Nc):
        cJ = int(C[J + 1] - C[J])
        for K in blockcolsupport(Nr, J, l, u):
            rK = int(R[K + 1] - R[K])
            s = u + K - J
            for jj in range(cJ):
                for k in range(max(0, jj - mu), min(rK - 1, jj + lam) + 1):
                    data[s * w + (mu + k - jj), C[J] + jj] = A[R[K] + k, C[J] + jj]
    return data


def bbb_valid_mask(rows, cols, l, u, lam, mu) -> np.ndarray:
    """Boolean mask over ``data``: True where the slot encodes a matrix entry."""
    R, C = prefix(rows), prefix(cols)
    shape = bbb_data_shape(cols, l, u, lam, mu)
    mask = np.zeros(shape, dtype=bool, order="F")
    w = lam + mu + 1
    if shape[0] == 0:
        return mask
    Nr, Nc = len(R) - 1, len(C) - 1
    for J in range(Nc):
        cJ = int(C[J + 1] - C[J])
        if cJ == 0:
            continue
        jj = np.arange(cJ)
        for K in blockcolsupport(Nr, J, l, u):
            rK = int(R[K + 1] - R[K])
            s = u + K - J
            for i in range(w):
                k = jj + i - mu
                ok = (k >= 0) & (k < rK)
                mask[s * w + i, C[J] + jj[ok]] = True
    return mask


def bbb_sparse(data, rows, cols, l, u, lam, mu):
    """``sparse(A)`` of ext/BlockBandedMatricesSparseArraysExt.jl:10-27 as CSC arrays (0-based colptr, rowval, nzval):
    every in-band entry of every in-band block is stored, explicit zeros included; within a column the rows
    ascend (what ``sparse(i, j, z, m, n)`` produces from the per-block triplets)."""
    R, C = prefix(rows), prefix(cols)
    n = int(C[-1])
    w = lam + mu + 1
    colptr = np.zeros(n + 1, dtype=np.int64)
    rowval, nzval = [], []
    Nr, Nc = len(R) - 1, len(C) - 1
    for J in range(Nc):
        for j in range(int(C[J + 1] - C[J])):
            c = int(C[J]) + j
            cnt = 0
            if data.shape[0] > 0:
                for K in blockcolsupport(Nr, J, l, u):
                    rK = int(R[K + 1] - R[K])
                    for k in range(max(0, j - mu), min(rK - 1, j + lam) + 1):
                        rowval.append(int(R[K]) + k)
                        nzval.append(data[(u + K - J) * w + (mu + k - j), c])
                        cnt += 1
            colptr[c + 1] = colptr[c] + cnt
    return colptr, np.asarray(rowval, dtype=np.int64), np.asarray(nzval, dtype=data.dtype)


def bbb_nnz_per_row_max(l, u, lam, mu) -> int:
    return max(0, l + u + 1) * max(0, lam + mu + 1)


# ----------------------------------------------------------------------------------------------
# BlockSkylineMatrix / BlockBandedMatrix: column-block-contiguous skyline
# ----------------------------------------------------------------------------------------------
def _vec(b, Nc):
    b = np.asarray(b, dtype=np.int64)
    if b.ndim == 0:
        return np.full(Nc, int(b), dtype=np.int64)
    return b.reshape(-1)


def sky_colrange(Nr, J, lvec, uvec) -> range:
    """src/BlockSkylineMatrix.jl:91 (colrange): max(0,J-u[J]) .. min(Nr-1, J+l[J])."""
    lo = max(0, J - int(uvec[J]))
    hi = min(Nr - 1, J + int(lvec[J]))
    return range(lo, hi + 1)


def sky_layout(rows, cols, l, u):
    """Panel offsets, strides and total length of the skyline ``data`` vector.

    Returns (panel_off[Nc+1], strides[Nc], klo[Nc], khi[Nc]) with 0-based panel offsets;
    block (K,J) starts at panel_off[J] + R[K]-R[klo[J]] and element (k,j) of it sits
    at that start + k + j*strides[J].
    src/BlockSkylineMatrix.jl:6-27 (bb_blockstarts), :29-43 (bb_blockstrides),
    :93-104 (bb_numentries), :459-479 (pointer / in-block index).
    """
    R, C = prefix(rows), prefix(cols)
    Nr, Nc = len(R) - 1, len(C) - 1
    lvec, uvec = _vec(l, Nc), _vec(u, Nc)
    if Nc > 1 and (len(lvec) != Nc or len(uvec) != Nc):
        raise ValueError(f"{Nc} lower and upper column bandwidths are required")  # :1-3
    off = np.zeros(Nc + 1, dtype=np.int64)
    strides = np.zeros(Nc, dtype=np.int64)
    klo = np.zeros(Nc, dtype=np.int64)
    khi = np.full(Nc, -1, dtype=np.int64)
    for J in range(Nc):
        KR = sky_colrange(Nr, J, lvec, uvec)
        klo[J] = KR.start
        khi[J] = KR.stop - 1
        if len(KR) > 0:
            strides[J] = R[KR.stop] - R[KR.start]
        off[J + 1] = off[J] + strides[J] * (C[J + 1] - C[J])
    return off, strides, klo, khi


def sky_numentries(rows, cols, l, u) -> int:
    return int(sky_layout(rows, cols, l, u)[0][-1])


def sky_blockstart(rows, cols, l, u, K: int, J: int) -> int:
    """1-based ``block_starts[K+1,J+1]`` of the reference (0 if the block is not stored)."""
    R = prefix(rows)
    off, strides, klo, khi = sky_layout(rows, cols, l, u)
    if not (klo[J] <= K <= khi[J]):
        return 0
    return int(off[J] + R[K] - R[klo[J]] + 1)


def sky_block(data, rows, cols, l, u, K: int, J: int) -> np.ndarray:
    R, C = prefix(rows), prefix(cols)
    off, strides, klo, khi = sky_layout(rows, cols, l, u)
    rK, cJ = int(R[K + 1] - R[K]), int(C[J + 1] - C[J])
    out = np.zeros((rK, cJ), dtype=data.dtype)
    if klo[J] <= K <= khi[J]:
        st = int(off[J] + R[K] - R[klo[J]])
        S = int(strides[J])
        for jj in range(cJ):
            out[:, jj] = data[st + jj * S: st + jj * S + rK]
    return out


def sky_getindex(data, rows, cols, l, u, i: int, j: int):
    R, C = prefix(rows), prefix(cols)
    K = int(np.searchsorted(R, i, side="right") - 1)
    J = int(np.searchsorted(C, j, side="right") - 1)
    off, strides, klo, khi = sky_layout(rows, cols, l, u)
    if not (klo[J] <= K <= khi[J]):
        return data.dtype.type(0)
    return data[int(off[J] + (i - R[klo[J]]) + (j - C[J]) * strides[J])]


def sky_to_dense(data, rows, cols, l, u) -> np.ndarray:
    R, C = prefix(rows), prefix(cols)
    off, strides, klo, khi = sky_layout(rows, cols, l, u)
    A = np.zeros((int(R[-1]), int(C[-1])), dtype=data.dtype)
    for J in range(len(C) - 1):
        S, cJ = int(strides[J]), int(C[J + 1] - C[J])
        if S == 0 or cJ == 0:
            continue
        panel = data[off[J]: off[J + 1]].reshape((S, cJ), order="F")
        A[R[klo[J]]: R[klo[J]] + S, C[J]: C[J + 1]] = panel
    return A


def sky_from_dense(A, rows, cols, l, u) -> np.ndarray:
    R, C = prefix(rows), prefix(cols)
    off, strides, klo, khi = sky_layout(rows, cols, l, u)
    data = np.zeros(int(off[-1]), dtype=A.dtype)
    for J in range(len(C) - 1):
        S, cJ = int(strides[J]), int(C[J + 1] - C[J])
        if S == 0 or cJ == 0:
            continue
        data[off[J]: off[J + 1]] = A[R[klo[J]]: R[klo[J]] + S, C[J]: C[J + 1]].reshape(-1, order="F")
    return data


def add_bandwidths(Al, Au, Bl, Bu, constant: bool = False):
    """Column block-bandwidths of A*B.

    constant=True: BlockBandedMatrix x BlockBandedMatrix -> Fill(lA+lB), Fill(uA+uB)
    (src/linalg.jl:27-30).  Otherwise the ragged rule of src/linalg.jl:8-25:
    v[i] = Bv[i] + max(Av[sel]), sel = max(i-Bu[i],0)..min(i+Bl[i],len(Av)-1) (skip if empty).
    """
    Al, Au = np.asarray(Al, dtype=np.int64), np.asarray(Au, dtype=np.int64)
    Bl, Bu = np.asarray(Bl, dtype=np.int64), np.asarray(Bu, dtype=np.int64)
    if constant:
        return np.full(len(Bl), int(Al[0] + Bl[0]) if len(Al) else 0, dtype=np.int64), \
            np.full(len(Bu), int(Au[0] + Bu[0]) if len(Au) else 0, dtype=np.int64)
    l, u = Bl.copy(), Bu.copy()
    for v, Av in ((l, Al), (u, Au)):
        for i in range(len(v)):
            lo = max(i - int(Bu[i]), 0)
            hi = min(i + int(Bl[i]), len(Av) - 1)
            if hi < lo:
                continue
            v[i] += int(np.max(Av[lo: hi + 1]))
    return l, u


# ----------------------------------------------------------------------------------------------
# synthetic inputs of BASELINE.json's configs
# ----------------------------------------------------------------------------------------------
def laplacian_bbb_data(n: int, dtype=np.float64) -> np.ndarray:
    """2-D 5-point Laplacian on an n x n grid as BBB (1,1),(1,1) band storage, 9 x n^2.

    Every column is n^2*[0,1,0, 1,-4,1, 0,1,0] (examples/cuarrays.jl:6-24;
    examples/finitedifference_2d.jl:9-15 builds the same operator with h=1/n).  Slots that
    fall outside a block (first/last column of each block, first/last block column) are
    *unused*; they keep the stencil value here exactly like examples/cuarrays.jl does.
    """
    sc = float(n) ** 2
    col = np.array([0, 1, 0, 1, -4, 1, 0, 1, 0], dtype=dtype) * dtype(sc)
    return np.asfortranarray(np.repeat(col[:, None], n * n, axis=1))


def laplacian_apply(x: np.ndarray, n: int) -> np.ndarray:
    """Independent check: the 5-point stencil applied to x viewed as an n x n grid."""
    X = x.reshape(n, n)  # X[J, j]: block J, in-block j
    Y = -4.0 * X
    Y[1:, :] += X[:-1, :]
    Y[:-1, :] += X[1:, :]
    Y[:, 1:] += X[:, :-1]
    Y[:, :-1] += X[:, 1:]
    return (Y * float(n) ** 2).reshape(-1)
