"""No-GPU checks: the C-ABI library loads and exports every symbol include/bbm_b200.h declares,
fails loudly without a device, and the host-side containers agree with the oracle's index maps."""
import ctypes
import os
import re

import numpy as np
import pytest

import bbm_b200 as B
from bbm_b200 import _lib
from oracle import formats as F
from tests.helpers import BBB_CASES, random_bbb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "bbm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bbm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build_library()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/bbm_b200.h but not exported"
    # and the ctypes binding covers the same set
    assert sorted(_lib.SIGNATURES) == syms


def test_status_strings_and_version():
    lib = _lib.load()
    assert lib.bbm_version() >= 100
    assert lib.bbm_status_string(0) == b"ok"
    assert lib.bbm_status_string(2) == b"DimensionMismatch"
    assert lib.bbm_status_string(3) == b"BandError"


def test_invalid_handle_is_reported_not_crashed():
    lib = _lib.load()
    assert lib.bbm_synchronize(None) == 1
    assert lib.bbm_bbb_desc_destroy(None) == 1
    assert lib.bbm_last_error(None) == b"invalid handle"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(B.BBMError):
        B.Handle(0)
    A = B.BandedBlockBandedMatrix.zeros(np.float64, [2, 2], [2, 2], (1, 1), (1, 1))
    with pytest.raises(TypeError):  # host-resident data is never multiplied by this package
        B.mul(A, np.zeros(4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "blockbandedmatrices.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), f"{f} mentions the oracle"


@pytest.mark.parametrize("case", BBB_CASES, ids=lambda c: f"r{len(c[0])}c{len(c[1])}_l{c[2]}u{c[3]}_lam{c[4]}mu{c[5]}")
def test_host_container_matches_oracle_maps(case):
    rows, cols, l, u, lam, mu = case
    if sum(rows) * sum(cols) > 400 * 400:
        rows, cols = [min(r, 40) for r in rows], [min(c, 40) for c in cols]
    rng = np.random.default_rng(3)
    data = random_bbb(rng, rows, cols, l, u, lam, mu)
    A = B.BandedBlockBandedMatrix(data, rows, cols, (l, u), (lam, mu))
    dense = F.bbb_to_dense(data, rows, cols, l, u, lam, mu)
    np.testing.assert_array_equal(A.to_dense(), dense)
    np.testing.assert_array_equal(A.valid_mask(), F.bbb_valid_mask(rows, cols, l, u, lam, mu))
    A2 = B.BandedBlockBandedMatrix.from_dense(dense, rows, cols, (l, u), (lam, mu), fill=np.nan)
    np.testing.assert_array_equal(A2.data, F.bbb_from_dense(dense, rows, cols, l, u, lam, mu, fill=np.nan))
    m, n = A.shape
    for _ in range(min(50, m * n)):
        i, j = int(rng.integers(m)), int(rng.integers(n))
        assert A[i, j] == dense[i, j]
    for K in range(len(rows)):
        assert list(A.blockrowsupport(K)) == list(F.blockrowsupport(len(cols), K, l, u))
    for J in range(len(cols)):
        assert list(A.blockcolsupport(J)) == list(F.blockcolsupport(len(rows), J, l, u))
        for K in range(len(rows)):
            np.testing.assert_array_equal(A.block(K, J), dense[A.R[K]:A.R[K + 1], A.C[J]:A.C[J + 1]])


def test_data_shape_check():
    with pytest.raises(ValueError):
        B.BandedBlockBandedMatrix(np.zeros((8, 4)), [2, 2], [2, 2], (1, 1), (1, 1))
