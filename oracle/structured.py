"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's structured ``+`` / ``-`` / ``alpha*A + B`` and
``A +- (I | Diagonal | Bidiagonal | Tridiagonal | SymTridiagonal)``, block loop by block loop (pure numpy, small cases).

Operands are tuples ``("bbb", data, rows, cols, l, u, lam, mu)`` or ``("sky", data, rows, cols, lvec, uvec)`` over the
storage formats of oracle/formats.py.  Only tests/ may import this.
"""
from __future__ import annotations

import numpy as np

from . import formats as F


def _axes(op):
    return list(op[2]), list(op[3])


def blockbandwidths(op):
    """``blockbandwidths(A)``: constants; (maximum(l), maximum(u)) for a skyline (src/BlockSkylineMatrix.jl:373-374)."""
    if op[0] == "bbb":
        return int(op[4]), int(op[5])
    lv, uv = np.asarray(op[4]).reshape(-1), np.asarray(op[5]).reshape(-1)
    if len(lv) == 0:
        return 0, 0
    return int(lv.max()), int(uv.max())


def get_block(op, K, J):
    """``view(A, Block(K), Block(J))`` as a dense array -- zeros where nothing is stored."""
    rows, cols = _axes(op)
    out = np.zeros((rows[K], cols[J]), dtype=op[1].dtype)
    if op[0] == "bbb":
        _, data, _, _, l, u, lam, mu = op
        w = lam + mu + 1
        if not (-l <= J - K <= u) or w <= 0 or data.shape[0] == 0:
            return out
        C = F.prefix(cols)
        band = data[(u + K - J) * w:(u + K - J + 1) * w, C[J]:C[J + 1]]
        for jj in range(out.shape[1]):
            for k in range(max(0, jj - mu), min(out.shape[0] - 1, jj + lam) + 1):
                out[k, jj] = band[mu + k - jj, jj]
        return out
    _, data, _, _, lv, uv = op
    if K in F.sky_colrange(len(rows), J, F._vec(lv, len(cols)), F._vec(uv, len(cols))):
        out[:, :] = F.sky_block(data, rows, cols, lv, uv, K, J)
    return out


def set_block(op, K, J, blk):
    """``view(C, Block(K), Block(J)) .= blk``: stores the in-band part; a non-zero outside the structure is a BandError."""
    rows, cols = _axes(op)
    if op[0] == "bbb":
        _, data, _, _, l, u, lam, mu = op
        w = lam + mu + 1
        inband = (-l <= J - K <= u) and w > 0 and data.shape[0] > 0
        C = F.prefix(cols)
        for jj in range(blk.shape[1]):
            for k in range(blk.shape[0]):
                if inband and -lam <= jj - k <= mu:
                    data[(u + K - J) * w + (mu + k - jj), C[J] + jj] = blk[k, jj]
                elif blk[k, jj] != 0:
                    raise IndexError("BandError")
        return
    _, data, _, _, lv, uv = op
    lvv, uvv = F._vec(lv, len(cols)), F._vec(uv, len(cols))
    if K in F.sky_colrange(len(rows), J, lvv, uvv):
        off, S, klo, khi = F.sky_layout(rows, cols, lv, uv)
        R = F.prefix(rows)
        st = int(off[J] + R[K] - R[klo[J]])
        for jj in range(blk.shape[1]):
            data[st + jj * int(S[J]): st + jj * int(S[J]) + blk.shape[0]] = blk[:, jj]
    elif np.any(blk != 0):
        raise IndexError("BandError")


def add_into(Cop, sign, A, B):
    """``copyto!(C, broadcasted(op, A, B))`` for op = + (sign 1) / - (sign -1), src/broadcast.jl:317-372: per block
    column the joint block range, then the four one-operand ranges, then zeros in the rest of C's band."""
    rows, cols = _axes(Cop)
    N, M = len(rows), len(cols)
    A_l, A_u = blockbandwidths(A)
    B_l, B_u = blockbandwidths(B)
    C_l, C_u = blockbandwidths(Cop)
    op = (lambda a, b: a + b) if sign > 0 else (lambda a, b: a - b)
    uop = (lambda b: b) if sign > 0 else (lambda b: -b)
    r1 = lambda lo, hi: range(lo, hi + 1)      # 1-based inclusive Julia range
    for Jt in range(1, M + 1):
        J = Jt - 1
        for Kt in r1(max(1, Jt - min(A_u, B_u)), min(N, Jt + min(A_l, B_l))):
            set_block(Cop, Kt - 1, J, op(get_block(A, Kt - 1, J), get_block(B, Kt - 1, J)))
        for Kt in r1(max(1, Jt - A_u), min(N, Jt - B_u - 1)):
            set_block(Cop, Kt - 1, J, get_block(A, Kt - 1, J))
        for Kt in r1(max(1, Jt - B_u), min(N, Jt - A_u - 1)):
            set_block(Cop, Kt - 1, J, uop(get_block(B, Kt - 1, J)))
        for Kt in r1(max(1, Jt + B_l + 1), min(N, Jt + A_l)):
            set_block(Cop, Kt - 1, J, get_block(A, Kt - 1, J))
        for Kt in r1(max(1, Jt + A_l + 1), min(N, Jt + B_l)):
            set_block(Cop, Kt - 1, J, uop(get_block(B, Kt - 1, J)))
        for Kt in r1(max(Jt - C_u, 1), min(Jt - max(A_u, B_u) - 1, N)):
            set_block(Cop, Kt - 1, J, np.zeros((rows[Kt - 1], cols[J]), dtype=Cop[1].dtype))
        for Kt in r1(max(Jt + max(A_l, B_l) + 1, 1), min(Jt + C_l, N)):
            set_block(Cop, Kt - 1, J, np.zeros((rows[Kt - 1], cols[J]), dtype=Cop[1].dtype))
    return Cop


def axpy_into(dest, alpha, A, B):
    """``dest .= alpha .* A .+ B`` (src/broadcast.jl:379-388): blockbanded_copyto!(dest, B) unless dest is B, then
    blockbanded_axpy!(alpha, A, dest) = per stored block of A ``dest_KJ .= alpha .* A_KJ .+ dest_KJ`` (:307-314)."""
    rows, cols = _axes(dest)
    N, M = len(rows), len(cols)
    if dest is not B:
        for J in range(M):
            for K in range(N):
                blk = get_block(B, K, J)
                try:
                    set_block(dest, K, J, blk)
                except IndexError:
                    if np.any(blk != 0):
                        raise
    A_l, A_u = blockbandwidths(A)
    for J in range(M):
        for K in range(max(0, J - A_u), min(N - 1, J + A_l) + 1):
            a = get_block(A, K, J)
            set_block(dest, K, J, alpha * a + get_block(dest, K, J))
    return dest


def copy_accommodating_diagonals(rows, lvec, uvec, diagonals):
    """Bandwidths of ``copy_accommodating_diagonals(A, d0:d1)`` (src/BlockSkylineMatrix.jl:502-531), plain loops."""
    rows = [int(r) for r in rows]
    l, u = [int(v) for v in F._vec(lvec, len(rows))], [int(v) for v in F._vec(uvec, len(rows))]
    d0, d1 = diagonals
    if d0 <= 0 <= d1:
        l, u = [max(v, 0) for v in l], [max(v, 0) for v in u]
    n = sum(rows)
    lasts = np.cumsum(rows)
    for d in (d0, d1):
        if d == 0:
            continue
        v = u if d > 0 else l
        for j in range(1, n + 1):
            i = min(max(j - d, 1), n)
            md = int(np.searchsorted(lasts, j, side="left"))
            db = int(np.searchsorted(lasts, i, side="left"))
            v[md] = max(v[md], abs(db - md))
    return l, u


def special_op(A, sign, reverse, diagonals, lam=0.0, dl=None, d=None, du=None):
    """``A op S`` / ``S op A`` for a skyline A and S given by its diagonals (src/BlockSkylineMatrix.jl:535-638): returns
    (data, l, u) of the result."""
    _, data, rows, cols, lv, uv = A
    l, u = copy_accommodating_diagonals(rows, lv, uv, diagonals)
    dense = F.sky_to_dense(data, rows, cols, lv, uv)
    if reverse and sign < 0:
        dense = -dense
    n = dense.shape[0]
    s = 1.0 if reverse else float(sign)
    for i in range(n):
        dense[i, i] = dense[i, i] + s * (lam + (d[i] if d is not None else 0.0))
    for i in range(n - 1):
        if du is not None:
            dense[i, i + 1] = dense[i, i + 1] + s * du[i]
        if dl is not None:
            dense[i + 1, i] = dense[i + 1, i] + s * dl[i]
    return F.sky_from_dense(dense, rows, cols, l, u), l, u
