"""CPU side of the structured +/-: the oracle's block-loop restatement (oracle/structured.py, src/broadcast.jl:317-388,
src/BlockSkylineMatrix.jl:502-638) against dense arithmetic on the structures of the reference's own tests, and the
host-side structure rules of the package against the reference's goldens."""
import numpy as np
import pytest

import bbm_b200 as B
from oracle import formats as F
from oracle import structured as S
from tests.helpers import random_bbb, random_sky


def _dense(op):
    if op[0] == "bbb":
        return F.bbb_to_dense(np.nan_to_num(op[1]), *op[2:])
    return F.sky_to_dense(*op[1:])


def _mk(rng, kind, rows, cols, lu, lm=None):
    if kind == "bbb":
        d = random_bbb(rng, rows, cols, lu[0], lu[1], lm[0], lm[1], poison=False)
        return ("bbb", d, rows, cols, lu[0], lu[1], lm[0], lm[1])
    return ("sky", random_sky(rng, rows, cols, lu[0], lu[1]), rows, cols, lu[0], lu[1])


def _empty_like_sum(A, B):
    rows, cols = A[2], A[3]
    (lA, uA), (lB, uB) = S.blockbandwidths(A), S.blockbandwidths(B)
    l, u = max(lA, lB), max(uA, uB)
    if A[0] == "bbb" and B[0] == "bbb":
        lam, mu = max(A[6], B[6]), max(A[7], B[7])
        return ("bbb", np.full(F.bbb_data_shape(cols, l, u, lam, mu), np.nan, order="F"), rows, cols, l, u, lam, mu)
    return ("sky", np.full(F.sky_numentries(rows, cols, l, u), np.nan), rows, cols, l, u)


# (A kind, A bandwidths, B kind, B bandwidths) on Fill(4,4) x Fill(4,3): test/test_broadcasting.jl:191-216
DEGENERATE = [
    ("bbb", ((2, 0), (1, 1)), "bbb", ((1, -1), (1, 1))),
    ("bbb", ((0, 2), (1, 1)), "bbb", ((-1, 1), (1, 1))),
    ("sky", ((2, 0), None), "bbb", ((1, -1), (1, 1))),
    ("sky", ((2, 0), None), "sky", ((1, -1), None)),
]


@pytest.mark.parametrize("case", DEGENERATE)
@pytest.mark.parametrize("sign", [1, -1])
def test_oracle_add_degenerate_bands(case, sign):
    rng = np.random.default_rng(11)
    ka, ba, kb, bb = case
    rows, cols = [4] * 4, [4] * 3
    A = _mk(rng, ka, rows, cols, *ba)
    Bm = _mk(rng, kb, rows, cols, *bb)
    for X, Y in ((A, Bm), (Bm, A)):
        Cm = S.add_into(_empty_like_sum(X, Y), sign, X, Y)
        want = _dense(X) + sign * _dense(Y)
        assert np.array_equal(_dense(Cm), want)


def test_oracle_add_and_axpy_wider_destination():
    """test/test_broadcasting.jl:136-189: (1,1) + (2,2) into (3,3); 2A + B; in place into B."""
    rng = np.random.default_rng(12)
    N = 10
    rows = list(range(1, N + 1))
    for kind, lm in (("sky", None), ("bbb", True)):
        mk = lambda b: _mk(rng, kind, rows if kind == "sky" else [1] * N, rows if kind == "sky" else [1] * N, (b, b), (b, b) if lm else None)
        A, Bm, Cm = mk(1), mk(2), mk(3)
        S.add_into(Cm, 1, A, Bm)
        assert np.array_equal(_dense(Cm), _dense(A) + _dense(Bm))
        S.axpy_into(Cm, 2.0, A, Bm)
        assert np.array_equal(_dense(Cm), 2.0 * _dense(A) + _dense(Bm))
        want = 2.0 * _dense(A) + _dense(Bm)
        S.axpy_into(Bm, 2.0, A, Bm)
        assert np.array_equal(_dense(Bm), want)


SKY_LU = [([0, 0, 0, 0], [0, 0, 0, 0]), ([-1, -1, -1, -1], [1, 2, 2, 2]), ([-1, -2, -2, -2], [1, 1, 2, 2]),
          ([2, 2, 1, 1], [-1, -1, -1, -1]), ([2, 2, 1, 1], [-2, -2, -2, -1])]
# (diagonals, expected l, expected u) -- test/test_blockskyline.jl:146-152
SPECIAL = {"I": ((0, 0), [0, 0, 0, 0], [0, 0, 0, 0]), "D": ((0, 0), [0, 0, 0, 0], [0, 0, 0, 0]),
           "Bu": ((0, 1), [0, 0, 0, 0], [0, 1, 1, 1]), "Bl": ((-1, 0), [1, 1, 1, 0], [0, 0, 0, 0]),
           "T": ((-1, 1), [1, 1, 1, 0], [0, 1, 1, 1]), "ST": ((-1, 1), [1, 1, 1, 0], [0, 1, 1, 1])}


@pytest.mark.parametrize("lu", SKY_LU)
@pytest.mark.parametrize("name", list(SPECIAL))
def test_accommodating_bandwidths_goldens(lu, name):
    rows = [1, 2, 3, 4]
    diagonals, el, eu = SPECIAL[name]
    want = ([max(a, b) for a, b in zip(el, lu[0])], [max(a, b) for a, b in zip(eu, lu[1])])
    assert S.copy_accommodating_diagonals(rows, lu[0], lu[1], diagonals) == want
    l, u = B.accommodating_bandwidths(rows, lu[0], lu[1], diagonals)
    assert (list(l), list(u)) == want


def test_accommodating_bandwidths_empty_blocks():
    rows = [2, 0, 3, 0, 0, 1]
    for diagonals in ((0, 0), (0, 1), (-1, 0), (-1, 1)):
        l0, u0 = [-1] * 6, [-1] * 6
        want = S.copy_accommodating_diagonals(rows, l0, u0, diagonals)
        l, u = B.accommodating_bandwidths(rows, l0, u0, diagonals)
        assert (list(l), list(u)) == want
