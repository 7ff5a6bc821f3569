"""BASELINE.json's configs at (or near) full size on the GPU, and the widened rows at the same scale.

Every config is checked against the CPU oracle at FULL size: cfg2 / cfg3 in one oracle call (< 1 s), cfg4
(2048 blocks of 256, all of C, block-column chunk by chunk) and cfg5 (1024 blocks of 512, all 64 right-hand
sides) against the oracle bound to threaded OpenBLAS (10-30 s each).  Reduced-size cfg4 / cfg5 cases add the
size-independent properties (associativity, solve-then-multiply round trip) on the GPU results themselves."""
import numpy as np
import pytest

import bbm_b200 as B
from oracle import cpu as O
from oracle import formats as F
from tests.helpers import relerr, tol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    h = B.Handle(0)
    yield h
    h.close()


def test_cfg2_full_size_laplacian_vs_oracle(handle):
    """2D Laplacian BandedBlockBandedMatrix, 4096 blocks of size 4096 (16.7M unknowns), Float64."""
    n = 4096
    A = B.BandedBlockBandedMatrix.finitedifference_2d(n)
    rng = np.random.default_rng(2)
    # the true reference leaves out-of-block slots undefined: poison them so that masking is exercised
    A.data[0:3, :n] = np.nan          # super-diagonal slab of the first block column (block row -1)
    A.data[6:9, -n:] = np.nan         # sub-diagonal slab of the last block column (block row Nr)
    x = rng.standard_normal(n * n)
    Ad = A.to_device(handle)
    dx, dy = handle.to_device(x), handle.empty(n * n, np.float64).fill_bytes(0xFF)
    B.mul_(dy, Ad, dx)
    y = dy.to_host()
    ref = np.zeros(n * n)
    O.use_openblas(threads=1)
    O.bbb_mul([n] * n, [n] * n, 1, 1, 1, 1, 1.0, A.data, x, 0.0, ref)
    O.use_builtin()
    assert not np.isnan(y).any()
    assert relerr(y, ref) <= tol(np.float64, 9)
    # linearity on the GPU result alone
    x2 = rng.standard_normal(n * n)
    y2 = (Ad @ handle.to_device(x2)).to_host()
    y12 = (Ad @ handle.to_device(2.0 * x - 3.0 * x2)).to_host()
    assert relerr(y12, 2.0 * y - 3.0 * y2) <= 4 * tol(np.float64, 9)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_cfg3_full_size_triangular_blocks_vs_oracle(handle, dtype):
    """BandedBlockBandedMatrix rows=cols=1:4000, (2,2),(2,2): 8.0M unknowns, P=25."""
    rows = list(range(1, 4001))
    rng = np.random.default_rng(3)
    n = sum(rows)
    data = rng.standard_normal((n, 25), dtype=np.float32).astype(dtype).T   # F-contiguous 25 x n, no copy
    x = rng.standard_normal(n).astype(dtype)
    Ad = B.BandedBlockBandedMatrix(data, rows, rows, (2, 2), (2, 2)).to_device(handle)
    y = (Ad @ handle.to_device(x)).to_host()
    ref = np.zeros(n, dtype=dtype)
    O.bbb_mul(rows, rows, 2, 2, 2, 2, 1.0, data, x, 0.0, ref)
    assert relerr(y, ref) <= tol(dtype, 25)


def _torch_handle():
    import torch
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    return torch, stream, B.Handle(0, stream=stream.cuda_stream)


def test_cfg4_full_size_vs_oracle():
    """cfg4 at FULL size: BlockBandedMatrix fill(256, 2048), (2,2) x (2,2) -> (4,4), Float64, 1.72 TFLOP on the FP64
    tensor cores; ALL of C (9.65 GB) against the oracle's per-block dgemm loop (test/test_linalg.jl:91-104 at scale)."""
    from benchmarks.configs import cfg4_oracle_check, wrap
    torch, stream, h = _torch_handle()
    rows = [256] * 2048
    gen = torch.Generator(device="cuda").manual_seed(1004)
    nA = B.BlockSkylineMatrix.numentries(rows, rows, 2, 2)
    nC = B.BlockSkylineMatrix.numentries(rows, rows, 4, 4)
    assert (nA, nC) == (670695424, 1206648832)
    with torch.cuda.stream(stream):
        a_t = torch.randn(nA, generator=gen, device="cuda", dtype=torch.float64) / 16
        b_t = torch.randn(nA, generator=gen, device="cuda", dtype=torch.float64) / 16
        c_t = torch.full((nC,), float("nan"), device="cuda", dtype=torch.float64)     # beta = 0 must never read C
        A = B.BlockSkylineMatrix(wrap(B, h, a_t), rows, rows, 2, 2)
        Bm = B.BlockSkylineMatrix(wrap(B, h, b_t), rows, rows, 2, 2)
        Cm = B.BlockSkylineMatrix(wrap(B, h, c_t), rows, rows, 4, 4)
        B.mul_(Cm, A, Bm)
        h.synchronize()
        assert not bool(torch.isnan(c_t).any().item())
        err, worst, secs, nchk = cfg4_oracle_check(torch, a_t, b_t, c_t, rows)
    assert nchk == 8
    assert err <= tol(np.float64, 1280) and worst <= tol(np.float64, 1280), (err, worst)
    h.close()


def test_cfg5_full_size_vs_oracle():
    """cfg5 at FULL size: UpperTriangular BlockBandedMatrix fill(512, 1024), (0,3), ldiv! with 64 right-hand sides;
    every column against the oracle's block back-substitution (test/test_triblockbanded.jl:80-99 at scale)."""
    from benchmarks.configs import cfg5_oracle_check, wrap
    torch, stream, h = _torch_handle()
    bs, nblk, nrhs = 512, 1024, 64
    rows = [bs] * nblk
    m = bs * nblk
    gen = torch.Generator(device="cuda").manual_seed(1005)
    nA = B.BlockSkylineMatrix.numentries(rows, rows, 0, 3)
    assert nA * 8 == 8577351680
    with torch.cuda.stream(stream):
        a_t = torch.randn(nA, generator=gen, device="cuda", dtype=torch.float64) / np.sqrt(bs)
        A = B.BlockSkylineMatrix(wrap(B, h, a_t), rows, rows, 0, 3)
        idx = torch.arange(bs, device="cuda", dtype=torch.int64)
        starts = torch.tensor([A.block_start(K, K) for K in range(nblk)], device="cuda", dtype=torch.int64)
        lds = torch.tensor([int(A.block_strides[K]) for K in range(nblk)], device="cuda", dtype=torch.int64)
        a_t[(starts[:, None] + idx[None, :] * (lds[:, None] + 1)).reshape(-1)] += 10.0      # +10 on the global diagonal
        b0 = torch.randn(nrhs, m, generator=gen, device="cuda", dtype=torch.float64)
        x_t = b0.clone()
        torch.cuda.synchronize()
        B.ldiv_(B.UpperTriangular(A), wrap(B, h, x_t))
        h.synchronize()
        err, secs = cfg5_oracle_check(torch, a_t, b0, x_t, rows)
    assert err <= tol(np.float64, 2048), err
    h.close()


def test_cfg4_block_size_256_vs_oracle(handle):
    """BlockBandedMatrix fill(256, N), (2,2) x (2,2) -> (4,4) on the FP64 tensor cores (N=24 of 2048)."""
    rows = [256] * 24
    rng = np.random.default_rng(4)
    na = B.BlockSkylineMatrix.numentries(rows, rows, 2, 2)
    adata, bdata = rng.standard_normal(na) / 16, rng.standard_normal(na) / 16
    Ad = B.BlockSkylineMatrix(adata, rows, rows, 2, 2).to_device(handle)
    Bd = B.BlockSkylineMatrix(bdata, rows, rows, 2, 2).to_device(handle)
    Cd = B.mul(Ad, Bd)
    assert Cd.blockbandwidths == (4, 4)
    got = Cd.data.to_host()
    ref = np.zeros_like(got)
    O.use_openblas(threads=8)
    O.sky_matmul((rows, rows, 2, 2), (rows, rows, 2, 2), (rows, rows, 4, 4), 1.0, adata, bdata, 0.0, ref)
    O.use_builtin()
    assert relerr(got, ref) <= tol(np.float64, 1280)
    # (A*B)*x == A*(B*x) on the GPU
    x = handle.to_device(rng.standard_normal(sum(rows)))
    assert relerr((Cd @ x).to_host(), (Ad @ (Bd @ x)).to_host()) <= 8 * tol(np.float64, 1280)


def test_cfg5_block_size_512_64rhs_vs_oracle(handle):
    """UpperTriangular BlockBandedMatrix fill(512, N), (0,3), 64 right-hand sides (N=12 of 1024)."""
    rows = [512] * 12
    rng = np.random.default_rng(5)
    m = sum(rows)
    data = rng.standard_normal(B.BlockSkylineMatrix.numentries(rows, rows, 0, 3)) / np.sqrt(512)
    A = B.BlockSkylineMatrix(data, rows, rows, 0, 3)
    for K in range(len(rows)):
        s, ld = A.block_start(K, K), int(A.block_strides[K])
        idx = np.arange(512)
        data[s + idx + idx * ld] += 10
    Ad = A.to_device(handle)
    B0 = np.asfortranarray(rng.standard_normal((m, 64)))
    dB = handle.to_device(B0)
    B.ldiv_(B.UpperTriangular(Ad), dB)
    got = dB.to_host()
    ref = B0.copy(order="F")
    O.use_openblas(threads=8)
    O.sky_trsm(rows, 0, 3, "U", False, data, ref)
    O.use_builtin()
    assert relerr(got, ref) <= tol(np.float64, 2048)
    # round trip: U * (U \ B) == B with U = upper triangle (zero the strictly lower part of the diagonal blocks)
    udata = data.copy()
    for K in range(len(rows)):
        s, ld = A.block_start(K, K), int(A.block_strides[K])
        for j in range(512):
            udata[s + j * ld + j + 1:s + j * ld + 512] = 0.0
    Ud = B.BlockSkylineMatrix(udata, rows, rows, 0, 3).to_device(handle)
    back = handle.empty((m, 64), np.float64)
    B.mul_(back, Ud, dB)
    assert relerr(back.to_host(), B0) <= 8 * tol(np.float64, 2048)


@pytest.mark.parametrize("uplo", ["L", "U"])
def test_gauss_seidel_half_step_full_size(handle, uplo):
    """ldiv!(LowerTriangular(A), b) / UpperTriangular on cfg2's 16.7M-unknown Laplacian (examples/finitedifference_2d.jl:21)
    against the oracle's block substitution, plus the round trip T * (T \\ b) == b on the GPU alone."""
    n = 4096
    A = B.BandedBlockBandedMatrix.finitedifference_2d(n)
    rng = np.random.default_rng(6)
    b = rng.standard_normal(n * n)
    Ad = A.to_device(handle)
    T = (B.LowerTriangular if uplo == "L" else B.UpperTriangular)(Ad)
    dx = B.ldiv(T, handle.to_device(b))
    x = dx.to_host()
    ref = b.copy()
    O.bbb_trsm([n] * n, 1, 1, 1, 1, uplo, False, A.data, ref)
    assert relerr(x, ref) <= tol(np.float64, 9)
    assert relerr((T @ dx).to_host(), b) <= 4 * tol(np.float64, 9)


def test_block_banded_qr_solve_moderate_size(handle):
    """F = qr(A); F \\ B on the device for a BlockBandedMatrix of 96 blocks of 128, (1,1), 16 right-hand sides:
    A * (F \\ B) == B, R'R == A'A on sampled columns (size-independent properties), factors vs the restated qr!."""
    rows, l, u = [128] * 96, 1, 1
    rng = np.random.default_rng(7)
    m = sum(rows)
    A = B.BlockSkylineMatrix(rng.standard_normal(B.BlockSkylineMatrix.numentries(rows, rows, l, u)) / np.sqrt(3 * 128), rows, rows, l, u)
    for K in range(len(rows)):
        s, ld = A.block_start(K, K), int(A.block_strides[K])
        i = np.arange(128)
        A.data[s + i + i * ld] += 2.0
    Aw = A.with_bandwidths(l, l + u)
    Fd = B.qr_(Aw.to_device(handle))
    B0 = np.asfortranarray(rng.standard_normal((m, 16)))
    dX = B.ldiv(Fd, handle.to_device(B0))
    back = handle.empty((m, 16), np.float64)
    B.mul_(back, A.to_device(handle), dX)
    assert relerr(back.to_host(), B0) <= tol(np.float64, 3 * 128)
    ref = Aw.data.copy()
    tau_ref = O.sky_qr(rows, rows, l, l + u, ref)
    got = Fd.factors.data.to_host()
    assert np.abs(got - ref).max() <= tol(np.float64, 2 * 128) * np.abs(ref).max()
    assert np.abs(Fd.tau.to_host() - tau_ref).max() <= tol(np.float64, 2 * 128)
    # Q' applied to the columns of A reproduces R: lmul!(Q', A e_j) == R e_j for sampled j
    Ad = A.to_dense()
    cols = rng.integers(0, m, 8)
    dC = handle.to_device(np.asfortranarray(Ad[:, cols]))
    B.lmul_adj_(Fd, dC)
    Rd = np.triu(F.sky_to_dense(got, rows, rows, l, l + u))
    assert relerr(dC.to_host(), Rd[:, cols]) <= tol(np.float64, 3 * 128)
