"""Device-side structured +/-, alpha*A + B and A +- (I | Diagonal | Bidiagonal | Tridiagonal) against the oracle's
block-loop restatement (oracle/structured.py), bit for bit, on the structures of test/test_broadcasting.jl:136-216 and
test/test_blockskyline.jl:117-161.  Every call goes through the C ABI (bbm_bbb_add_*, bbm_sky_add_*, bbm_*_add_diags_*)."""
import numpy as np
import pytest

import bbm_b200 as B
from oracle import formats as F
from oracle import structured as S
from tests.helpers import random_bbb, random_sky
from tests.test_structured_add_cpu import DEGENERATE, SKY_LU, SPECIAL, _dense, _empty_like_sum, _mk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    h = B.Handle(0)
    yield h
    h.close()


def _to_dev(h, op):
    if op[0] == "bbb":
        return B.BandedBlockBandedMatrix(op[1], op[2], op[3], (op[4], op[5]), (op[6], op[7])).to_device(h)
    return B.BlockSkylineMatrix(op[1], op[2], op[3], op[4], op[5]).to_device(h)


def _host(M):
    d = M.data.to_host()
    if isinstance(M, B.BandedBlockBandedMatrix):
        return F.bbb_to_dense(d, M.rows, M.cols, M.l, M.u, M.lam, M.mu) if d.size else np.zeros(M.shape)
    return F.sky_to_dense(d, M.rows, M.cols, M.l, M.u)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", DEGENERATE)
def test_add_sub_degenerate_bands(handle, case, dtype):
    rng = np.random.default_rng(21)
    ka, ba, kb, bb = case
    rows, cols = [4] * 4, [4] * 3

    def mk(kind, bw):
        op = _mk(rng, kind, rows, cols, *bw)
        return (op[0], op[1].astype(dtype)) + op[2:]
    A, Bm = mk(ka, ba), mk(kb, bb)
    dA, dB = _to_dev(handle, A), _to_dev(handle, Bm)
    for (X, dX), (Y, dY) in (((A, dA), (Bm, dB)), ((Bm, dB), (A, dA))):
        for fn, sign in ((B.add, 1), (B.sub, -1)):
            got = fn(dX, dY)
            ref = S.add_into(_empty_like_sum(X, Y), sign, X, Y)
            assert type(got) is (B.BandedBlockBandedMatrix if ref[0] == "bbb" else B.BlockSkylineMatrix)
            assert np.array_equal(_host(got), _dense(ref))
            assert _host(got).dtype == np.dtype(dtype)


@pytest.mark.parametrize("kind", ["sky", "bbb"])
def test_add_axpy_wider_destination_and_in_place(handle, kind):
    """C .= A .+ B into (3,3); 2A + B; B .= 2.0 .* A .+ B (test/test_broadcasting.jl:136-189)."""
    rng = np.random.default_rng(22)
    N = 10
    rows = list(range(1, N + 1)) if kind == "sky" else [3] * N
    mk = lambda b: _mk(rng, kind, rows, rows, (b, b), (b, b) if kind == "bbb" else None)
    A, Bm, Cm = mk(1), mk(2), mk(3)
    dA, dB, dC = _to_dev(handle, A), _to_dev(handle, Bm), _to_dev(handle, Cm)
    B.axpby_(dC, 1.0, dA, 1.0, dB)
    assert np.array_equal(_host(dC), _dense(A) + _dense(Bm))
    S2 = B.add(dA, dB)
    assert (S2.blockbandwidths if kind == "bbb" else (int(S2.l[0]), int(S2.u[0]))) == (2, 2)
    if kind == "bbb":
        assert S2.subblockbandwidths == (2, 2)
    assert np.array_equal(_host(S2), _dense(A) + _dense(Bm))
    ref = S.axpy_into(Cm, 2.0, A, Bm)
    B.axpby_(dC, 2.0, dA, 1.0, dB)
    assert np.array_equal(_host(dC), _dense(ref))
    assert np.array_equal(_host(B.scaled_add(2.0, dA, dB)), _dense(ref))
    B.axpby_(dB, 2.0, dA, 1.0, dB)          # in place: the destination is the second operand
    assert np.array_equal(_host(dB), _dense(ref))
    # a destination that is too narrow is a BandError; different block axes a DimensionMismatch
    with pytest.raises(B.BandError):
        B.axpby_(dA, 1.0, dB, 0.0, None)
    other = _to_dev(handle, _mk(rng, kind, rows[:-1] + [rows[-1] + 1], rows, (1, 1), (1, 1) if kind == "bbb" else None))
    with pytest.raises(B.DimensionMismatch):
        B.axpby_(dC, 1.0, other, 1.0, dB)


def test_add_large_mixed_formats(handle):
    """BlockBandedMatrix + BandedBlockBandedMatrix with ragged blocks, several tiles per panel, NaN in the unused slots
    of the band storage (never read as entries)."""
    rng = np.random.default_rng(23)
    rows = [300, 17, 0, 513, 64, 129]
    A = ("sky", random_sky(rng, rows, rows, [1, 0, 2, 1, 0, 0], [0, 1, 1, 2, 1, 0]), rows, rows, [1, 0, 2, 1, 0, 0], [0, 1, 1, 2, 1, 0])
    Bb = ("bbb", random_bbb(rng, rows, rows, 1, 2, 3, 2, poison=True), rows, rows, 1, 2, 3, 2)
    dA, dB = _to_dev(handle, A), _to_dev(handle, Bb)
    for fn, sign in ((B.add, 1.0), (B.sub, -1.0)):
        got = _host(fn(dA, dB))
        assert np.array_equal(got, _dense(A) + sign * _dense(Bb))
        got = _host(fn(dB, dA))
        assert np.array_equal(got, _dense(Bb) + sign * _dense(A))


@pytest.mark.parametrize("lu", SKY_LU)
@pytest.mark.parametrize("name", list(SPECIAL))
def test_skyline_plus_minus_structured(handle, lu, name):
    """check_blockskylinematrix_arithmetic of test/test_blockskyline.jl:128-160 (Float64 instead of Int)."""
    rng = np.random.default_rng(24)
    rows = [1, 2, 3, 4]
    N = sum(rows)
    data = rng.standard_normal(F.sky_numentries(rows, rows, lu[0], lu[1]))
    A = ("sky", data, rows, rows, lu[0], lu[1])
    dA = _to_dev(handle, A)
    d, e1, e2 = rng.standard_normal(N), rng.standard_normal(N - 1), rng.standard_normal(N - 1)
    ops = {"I": (B.UniformScaling(3.0), dict(lam=3.0)), "D": (B.Diagonal(d), dict(d=d)),
           "Bu": (B.Bidiagonal(d, e1, "U"), dict(d=d, du=e1)), "Bl": (B.Bidiagonal(d, e1, "L"), dict(d=d, dl=e1)),
           "T": (B.Tridiagonal(e1, d, e2), dict(dl=e1, d=d, du=e2)), "ST": (B.SymTridiagonal(d, e1), dict(d=d, dl=e1, du=e1))}
    Sop, parts = ops[name]
    diagonals, el, eu = SPECIAL[name]
    for fn, sign in ((B.add, 1), (B.sub, -1)):
        for reverse in (False, True):
            got = fn(Sop, dA) if reverse else fn(dA, Sop)
            ref, l, u = S.special_op(A, sign, reverse, diagonals, **parts)
            assert list(got.l) == [max(a, b) for a, b in zip(el, lu[0])] and list(got.u) == [max(a, b) for a, b in zip(eu, lu[1])]
            assert np.array_equal(got.data.to_host(), ref)


def test_bbb_identity_minus_scaled_laplacian(handle):
    """A = I - dt*Delta of examples/finitedifference_2d.jl:50-51 assembled on the device: -dt*Delta, then the diagonal."""
    n = 40
    dt = (1.0 / n ** 2) / 4
    Lap = B.BandedBlockBandedMatrix.finitedifference_2d(n)
    dL = Lap.to_device(handle)
    scaled = B.BandedBlockBandedMatrix(handle.empty(Lap.data.shape, np.float64), Lap.rows, Lap.cols, (1, 1), (1, 1))
    B.axpby_(scaled, dt, dL, 0.0, None)
    Ad = B.sub(B.UniformScaling(1.0), scaled)
    want = np.eye(n * n) - dt * Lap.to_dense()
    assert np.array_equal(_host(Ad), want)
    # a diagonal that is not stored is a BandError (the reference's setindex! throws)
    off = B.BandedBlockBandedMatrix.zeros(np.float64, [3] * 4, [3] * 4, (1, 1), (-1, 2)).to_device(handle)
    with pytest.raises(B.BandError):
        B.add(off, B.UniformScaling(1.0))
