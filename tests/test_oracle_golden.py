"""Pins the oracle's index maps to every integer golden the reference's tests/docs hold
(tests/golden/reference_goldens.json, each entry cites the reference file:line)."""
import json
import os

import numpy as np
import pytest

from oracle import formats as F
from oracle import cpu as O

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")))


def iota(shape, dtype=np.float64):
    n = int(np.prod(shape))
    return np.arange(1, n + 1, dtype=dtype).reshape(shape, order="F")


@pytest.mark.parametrize("case", G["bbb_layout"], ids=lambda c: c["cite"])
def test_bbb_layout_goldens(case):
    rows, cols = case["rows"], case["cols"]
    (l, u), (lam, mu) = case["lu"], case["lm"]
    data = iota(F.bbb_data_shape(cols, l, u, lam, mu))
    for i, j, v in case["entries"]:
        assert F.bbb_getindex(data, rows, cols, l, u, lam, mu, i - 1, j - 1) == v
    D = F.bbb_to_dense(data, rows, cols, l, u, lam, mu)
    R, C = F.prefix(rows), F.prefix(cols)
    for b in case["blocks"]:
        K, J = b["K"] - 1, b["J"] - 1
        np.testing.assert_array_equal(D[R[K]:R[K + 1], C[J]:C[J + 1]], np.array(b["value"], dtype=float))
    for i, j, v in case["entries"]:
        assert D[i - 1, j - 1] == v


def test_bbb_bandeddata_pointer_golden():
    c = G["bbb_bandeddata_ptr"]
    (l, u), (lam, mu) = c["lu"], c["lm"]
    data = iota(F.bbb_data_shape(c["cols"], l, u, lam, mu))
    K, J = c["K"] - 1, c["J"] - 1
    w = lam + mu + 1
    Cp = F.prefix(c["cols"])
    # bandeddata(view(A,Block(K),Block(J))) = data[(u+K-J)*w ..., C[J] ...]  (src/BandedBlockBandedMatrix.jl:538-545)
    assert data[(u + K - J) * w, Cp[J]] == c["first_value"]


@pytest.mark.parametrize("case", G["bbb_from_dense"], ids=lambda c: c["cite"])
def test_bbb_from_dense_goldens(case):
    rows, cols = case["rows"], case["cols"]
    (l, u), (lam, mu) = case["lu"], case["lm"]
    A = np.ones((sum(rows), sum(cols))) if case["A"] == "ones" else np.array(case["A"], dtype=float)
    data = F.bbb_from_dense(A, rows, cols, l, u, lam, mu, fill=np.nan)
    assert data.shape == F.bbb_data_shape(cols, l, u, lam, mu)
    D = F.bbb_to_dense(data, rows, cols, l, u, lam, mu)
    np.testing.assert_array_equal(D, np.array(case["dense"], dtype=float))
    # the valid mask is exactly the set of slots from_dense wrote
    np.testing.assert_array_equal(~np.isnan(data), F.bbb_valid_mask(rows, cols, l, u, lam, mu))


@pytest.mark.parametrize("case", G["bbb_blocksupport"], ids=lambda c: c["cite"])
def test_blocksupport_goldens(case):
    l, u = case["lu"]
    for K1, (a, b) in case["rowsupport"].items():
        got = F.blockrowsupport(case["Nc"], int(K1) - 1, l, u)
        assert list(got) == list(range(a - 1, b))
    for J1, (a, b) in case["colsupport"].items():
        got = F.blockcolsupport(case["Nr"], int(J1) - 1, l, u)
        assert list(got) == list(range(a - 1, b))


def test_bbb_doc_layout():
    c = G["bbb_doc_layout"]
    (l, u), (lam, mu) = c["lu"], c["lm"]
    assert list(F.bbb_data_shape(c["cols"], l, u, lam, mu)) == c["data_shape"]
    R, C = F.prefix(c["rows"]), F.prefix(c["cols"])
    mask = F.bbb_valid_mask(c["rows"], c["cols"], l, u, lam, mu)
    for col1, slots in c["columns"].items():
        for row, ent in enumerate(slots):
            if ent is None:
                assert not mask[row, int(col1) - 1]
            else:
                assert F.bbb_entry_pos(R, C, l, u, lam, mu, ent[0] - 1, ent[1] - 1) == (row, int(col1) - 1)


@pytest.mark.parametrize("case", G["sky_layout"], ids=lambda c: c["cite"])
def test_sky_layout_goldens(case):
    rows, cols = case["rows"], case["cols"]
    l, u = case["lu"]
    n = F.sky_numentries(rows, cols, l, u)
    assert n == O.sky_numentries(rows, cols, l, u)  # C oracle agrees with the numpy restatement
    if case["numentries"] is not None:
        assert n == case["numentries"]
    data = np.arange(1, n + 1, dtype=np.float64)
    for i, j, v in case["entries"]:
        assert F.sky_getindex(data, rows, cols, l, u, i - 1, j - 1) == v
    for b in case["blocks"]:
        np.testing.assert_array_equal(F.sky_block(data, rows, cols, l, u, b["K"] - 1, b["J"] - 1), np.array(b["value"], dtype=float))
    for e in case.get("block_entries", []):
        blk = F.sky_block(data, rows, cols, l, u, e["K"] - 1, e["J"] - 1)
        assert blk[e["k"] - 1, e["j"] - 1] == e["value"]
    off, strides, klo, khi = F.sky_layout(rows, cols, l, u)
    for e in case.get("block_ptr", []):
        K, J = e["K"] - 1, e["J"] - 1
        assert strides[J] == e["stride"]
        st = F.sky_blockstart(rows, cols, l, u, K, J)  # 1-based like block_starts
        if e["first"] is not None:
            assert data[st - 1] == e["first"]
            assert data[st - 1 + strides[J]] == e["next_col"]
    D = F.sky_to_dense(data, rows, cols, l, u)
    for i, j, v in case["entries"]:
        assert D[i - 1, j - 1] == v
    np.testing.assert_array_equal(F.sky_from_dense(D, rows, cols, l, u), data)


def test_sky_panel_pointer_golden():
    c = G["sky_panel_ptr"]
    rows, cols = c["rows"], c["cols"]
    l, u = c["lu"]
    n = F.sky_numentries(rows, cols, l, u)
    data = np.random.default_rng(0).standard_normal(n)
    off, strides, klo, khi = F.sky_layout(rows, cols, l, u)
    R = F.prefix(rows)
    K0, K1, J = c["K_first"] - 1, c["K_last"] - 1, c["J"] - 1
    st = F.sky_blockstart(rows, cols, l, u, K0, J) - 1
    D = F.sky_to_dense(data, rows, cols, l, u)
    fi, fj = c["first_entry"]
    ni, nj = c["next_col_entry"]
    assert data[st] == D[fi - 1, fj - 1]
    assert data[st + strides[J]] == D[ni - 1, nj - 1]
    assert [int(R[K1 + 1] - R[K0]), cols[J]] == c["size"]
    # the panel view is one strided matrix: rows of blocks K0..K1 are contiguous in the panel
    panel = np.stack([data[st + jj * strides[J]: st + jj * strides[J] + c["size"][0]] for jj in range(cols[J])], axis=1)
    C = F.prefix(cols)
    np.testing.assert_array_equal(panel, D[R[K0]:R[K1 + 1], C[J]:C[J + 1]])


def test_add_bandwidths_golden():
    c = G["add_bandwidths"]
    l, u = F.add_bandwidths(c["l"], c["u"], c["l"], c["u"])
    assert list(l[0:7]) == c["MMl_1to7"]
    assert all(l[7:10] >= np.array(c["MMl_8to10_ge"]))
    assert all(u[1:4] >= np.array(c["MMu_2to4_ge"]))
    assert list(u[4:11]) == c["MMu_5to11"]
    # constant-bandwidth rule (src/linalg.jl:27-30): BlockBandedMatrix^2 stays BlockBanded
    l2, u2 = F.add_bandwidths([1] * 11, [1] * 11, [1] * 11, [1] * 11, constant=True)
    assert list(l2) == [2] * 11 and list(u2) == [2] * 11


@pytest.mark.parametrize("blas", ["builtin", "openblas"])
@pytest.mark.parametrize("case", G["sky_products_exact"], ids=lambda c: c["cite"][:40])
def test_sky_products_exact_goldens(case, blas):
    if blas == "openblas":
        if not O.use_openblas(threads=1):
            pytest.skip("OpenBLAS not found")
    else:
        O.use_builtin()
    try:
        A = case["A"]
        Al, Au = (A["lu"] if "lu" in A else (A["l"], A["u"]))
        if case.get("square"):
            Bs = A
            Bl, Bu = Al, Au
            Cl, Cu = F.add_bandwidths(Al, Au, Bl, Bu)
            Cs = {"rows": A["rows"], "cols": A["cols"]}
        else:
            Bs = case["B"]
            Bl, Bu = Bs["lu"]
            Cs = case["C"]
            Cl, Cu = Cs["lu"]
        Ad = np.ones(F.sky_numentries(A["rows"], A["cols"], Al, Au))
        Bd = np.ones(F.sky_numentries(Bs["rows"], Bs["cols"], Bl, Bu))
        if "sumA" in case:
            assert Ad.sum() == case["sumA"] and Bd.sum() == case["sumB"]
        Cd = np.full(F.sky_numentries(Cs["rows"], Cs["cols"], Cl, Cu), np.nan)
        O.sky_matmul((A["rows"], A["cols"], Al, Au), (Bs["rows"], Bs["cols"], Bl, Bu),
                     (Cs["rows"], Cs["cols"], Cl, Cu), 1.0, Ad, Bd, 0.0, Cd)
        dense = F.sky_to_dense(Ad, A["rows"], A["cols"], Al, Au) @ F.sky_to_dense(Bd, Bs["rows"], Bs["cols"], Bl, Bu)
        # integer-valued data: the reference asserts exact equality (==), so do we
        np.testing.assert_array_equal(F.sky_to_dense(Cd, Cs["rows"], Cs["cols"], Cl, Cu), dense)
    finally:
        O.use_builtin()
