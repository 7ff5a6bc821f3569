"""Host-side BlockSkylineMatrix mirror against the reference's golden layout values and the oracle's
index maps (no GPU)."""
import numpy as np
import pytest

import bbm_b200 as B
from oracle import formats as F
from tests.helpers import random_sky


def test_blockbanded_goldens():
    # test/test_blockbanded.jl:49-52
    A = B.BlockBandedMatrix(np.arange(1, 31, dtype=np.float64), [1, 2, 3], [1, 2, 3], (1, 1))
    assert A.block(0, 0).tolist() == [[1]]
    assert A.block(1, 1).tolist() == [[5, 11], [6, 12]]
    assert A.block(2, 2).tolist() == [[18, 23, 28], [19, 24, 29], [20, 25, 30]]
    # test/test_blockbanded.jl:59-90, test/test_linalg.jl:42-47
    A = B.BlockBandedMatrix(np.arange(1, 71, dtype=np.float64), [1, 2, 3, 4], [1, 2, 3, 4], (1, 1))
    assert [A[0, 0], A[1, 0], A[2, 0], A[3, 0], A[0, 1], A[0, 2]] == [1, 2, 3, 0, 4, 10]
    assert A.block(2, 3)[2, 0] == 45
    assert A.block_strides[3] == 7 and A.data[A.block_start(3, 3)] == 46 and A.data[A.block_start(3, 3) + 7] == 53
    # test/test_triblockbanded.jl:196: Block(5,5) of rows 1:5 has strides (1,9)
    A5 = B.BlockSkylineMatrix.zeros(np.float64, list(range(1, 6)), list(range(1, 6)), 1, 1)
    assert A5.block_strides[4] == 9


@pytest.mark.parametrize("case", [
    (list(range(1, 8)), list(range(1, 8)), 1, 2),
    ([3] * 7, [3] * 7, [2, 1, 2, 3, 1, 0, 0], [0, 1, 1, 2, 3, 1, 2]),
    ([3] * 7, [3] * 7, [-1, 1, 2, 0, 1, 0, 0], [1, 1, 0, 2, -1, 1, 2]),
    ([5, 3, 7, 2, 6], [4, 6, 2, 5, 3, 7], 1, 2),
    ([], [], 1, 1),
])
def test_skyline_mirror_matches_oracle_maps(case):
    rows, cols, l, u = case
    rng = np.random.default_rng(0)
    data = random_sky(rng, rows, cols, l, u)
    A = B.BlockSkylineMatrix(data, rows, cols, l, u)
    assert B.BlockSkylineMatrix.numentries(rows, cols, l, u) == F.sky_numentries(rows, cols, l, u) == data.size
    assert np.array_equal(A.to_dense(), F.sky_to_dense(data, rows, cols, l, u))
    back = B.BlockSkylineMatrix.from_dense(A.to_dense(), rows, cols, l, u)
    assert np.array_equal(back.data, data)
    for K in range(len(rows)):
        for J in range(len(cols)):
            if A.has_block(K, J):
                assert A.block_start(K, J) + 1 == F.sky_blockstart(rows, cols, l, u, K, J)  # oracle is 1-based like the reference


def test_product_bandwidths_golden():
    # test/test_blockskyline.jl:42-81: MMl[1:7] == [3,4,3,4,3,4,3] style goldens via the oracle rule
    rows = cols = [3] * 7
    l, u = [2, 1, 2, 3, 1, 0, 0], [0, 1, 1, 2, 3, 1, 2]
    M = B.BlockSkylineMatrix.zeros(np.float64, rows, cols, l, u)
    ml, mu = B.add_bandwidths(M, M)
    ol, ou = F.add_bandwidths(l, u, l, u)
    assert list(ml) == list(ol) and list(mu) == list(ou)
    A = B.BlockSkylineMatrix.zeros(np.float64, rows, cols, 1, 2)
    assert B.add_bandwidths(A, A) == (2, 4)      # src/linalg.jl:27-30


def test_triangular_bandwidths():
    A = B.BlockSkylineMatrix.zeros(np.float64, [2, 3], [2, 3], 1, 1)
    assert B.UpperTriangular(A).blockbandwidths == (0, 1)     # src/triblockbanded.jl:10-17
    assert B.LowerTriangular(A).blockbandwidths == (1, 0)
    A = B.BlockSkylineMatrix.zeros(np.float64, [2, 3], [3, 2], 1, 1)
    assert B.UpperTriangular(A).blockbandwidths == (1, 1)     # blocks do not match: unchanged


def test_bbb_block_range_views():
    """test/test_triblockbanded.jl:55-78: view(A, Block(3), Block.(3:5)) / view(A, Block.(2:4), Block(3)) stay
    banded-block-banded over the same storage with shifted block bandwidths."""
    rows = list(range(1, 11))
    rng = np.random.default_rng(3)
    A = B.BandedBlockBandedMatrix(rng.standard_normal(B.BandedBlockBandedMatrix.data_shape(rows, (1, 1), (1, 1))), rows, rows, (1, 1), (1, 1))
    D = A.to_dense()
    V = A.view(2, range(2, 5))
    assert V.blockbandwidths == (1, 1) and V.shape == (3, 3 + 4 + 5)
    assert np.array_equal(V.to_dense(), D[3:6, 3:15])
    V = A.view(range(1, 4), 2)
    assert V.blockbandwidths == (2, 0) and V.blockshape == (3, 1)      # reference: blockbandwidths(V) == (2,0)
    assert list(V.rows) == [2, 3, 4] and list(V.cols) == [3]
    assert np.array_equal(V.to_dense(), D[1:10, 3:6])
    assert np.shares_memory(V.data, A.data)
    V = A.view(range(0, 2), range(6, 9))                                 # entirely outside the block band
    assert not V.to_dense().any()
    with pytest.raises(IndexError):
        A.view(range(0, 11), 0)


def test_with_bandwidths_keeps_the_entries():
    """BlockBandedMatrix(A, (l, l+u)) of src/blockskylineqr.jl:75"""
    rows, cols = [3, 1, 4, 2], [2, 5, 1, 3]
    rng = np.random.default_rng(4)
    A = B.BlockSkylineMatrix(random_sky(rng, rows, cols, 1, 1), rows, cols, 1, 1)
    W = A.with_bandwidths(1, 2)
    assert W.blockbandwidths == (1, 2) and np.array_equal(W.to_dense(), A.to_dense())
    with pytest.raises(IndexError):
        A.with_bandwidths(0, 2)
