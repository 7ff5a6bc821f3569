#!/usr/bin/env python
"""Secondary measurements: every BASELINE.json config at FULL size on one B200, each with a parity
check (against the CPU oracle where the oracle finishes in seconds, through size-independent
properties otherwise) and a roofline line.  bench.py stays the headline (cfg2); this script feeds
DESIGN.md's per-kernel table.

  python benchmarks/configs.py [--configs 1,2,3,4,5] [--reps 10] [--no-oracle]

One JSON line per config on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

EPS = {8: 2.220446049250313e-16, 4: 1.1920929e-07}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


def relerr(a, b):
    den = float(np.linalg.norm(np.asarray(b, dtype=np.float64).ravel()))
    return float(np.linalg.norm((np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).ravel())) / max(den, 1e-300)


def time_gpu(torch, stream, fn, reps, warm=3):
    with torch.cuda.stream(stream):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record(stream)
        for i in range(reps):
            fn()
            ev[i + 1].record(stream)
        torch.cuda.synchronize()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    return float(np.mean(per)), float(np.min(per))


def dgemm_peak(torch, n=8192, reps=5):
    """cuBLAS FP64 GEMM throughput on this GPU = the tensor roofline's denominator for cfg4/cfg5."""
    a = torch.randn(n, n, device="cuda", dtype=torch.float64)
    b = torch.randn(n, n, device="cuda", dtype=torch.float64)
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def wrap(B, h, t):
    """DeviceArray view of a torch tensor's memory (column-major interpretation by the caller)."""
    d = B.DeviceArray(h, tuple(reversed(t.shape)) if t.dim() == 2 else tuple(t.shape), np.dtype(str(t.dtype).replace("torch.", "")), ptr=t.data_ptr())
    d._keep = t
    return d


def run_bbb(torch, B, h, stream, name, rows, cols, lu, lm, dtype, reps, oracle, laplacian=None):
    dt = np.dtype(dtype)
    P = max(0, lu[0] + lu[1] + 1) * max(0, lm[0] + lm[1] + 1)
    n, m = int(np.sum(cols)), int(np.sum(rows))
    tdt = torch.float64 if dt == np.float64 else torch.float32
    gen = torch.Generator(device="cuda").manual_seed(1003)
    if laplacian is not None:
        col = torch.tensor([0, 1, 0, 1, -4, 1, 0, 1, 0], dtype=tdt, device="cuda") * float(laplacian) ** 2
        data_t = col.repeat(n, 1).contiguous()            # n x P row-major == P x n column-major
    else:
        data_t = torch.randn(n, P, generator=gen, device="cuda", dtype=tdt)
    x_t = torch.randn(n, generator=gen, device="cuda", dtype=tdt)
    y_t = torch.full((m,), float("nan"), device="cuda", dtype=tdt)
    A = B.BandedBlockBandedMatrix(wrap(B, h, data_t), rows, cols, lu, lm)
    dx, dy = wrap(B, h, x_t), wrap(B, h, y_t)
    avg, best = time_gpu(torch, stream, lambda: B.mul_(dy, A, dx), reps)
    nbytes = dt.itemsize * (P * n + n + m)
    nnz = (lu[0] + lu[1] + 1) * (lm[0] + lm[1] + 1)
    out = {"config": name, "dtype": str(dt), "unknowns": m, "ms_avg": avg, "ms_min": best,
           "algorithmic_bytes": nbytes, "GBps": nbytes / (avg * 1e-3) / 1e9,
           "roofline": {"bound": "hbm", "peak_GBps": peaks(), "frac": nbytes / (avg * 1e-3) / 1e9 / peaks()},
           "tolerance": 64 * EPS[dt.itemsize] * nnz}
    y = y_t.cpu().numpy()
    if oracle:
        from oracle import cpu as O
        O.use_openblas(threads=1)
        data = np.asfortranarray(data_t.cpu().numpy().T)
        ref = np.zeros(m, dtype=dt)
        t0 = time.perf_counter()
        O.bbb_mul(rows, cols, lu[0], lu[1], lm[0], lm[1], 1.0, data, x_t.cpu().numpy(), 0.0, ref)
        out["cpu_seconds_full_size"] = time.perf_counter() - t0
        out["cpu_GBps"] = nbytes / out["cpu_seconds_full_size"] / 1e9
        out["rel_err_vs_oracle_full_size"] = relerr(y, ref)
        out["parity_ok"] = bool(out["rel_err_vs_oracle_full_size"] <= out["tolerance"])
    else:  # linearity
        x2 = torch.randn(n, generator=gen, device="cuda", dtype=tdt)
        y2 = torch.empty_like(y_t); y12 = torch.empty_like(y_t)
        B.mul_(wrap(B, h, y2), A, wrap(B, h, x2))
        B.mul_(wrap(B, h, y12), A, wrap(B, h, (2.0 * x_t - 3.0 * x2).contiguous()))
        h.synchronize()
        out["linearity_rel_err"] = relerr(y12.cpu().numpy(), (2.0 * y_t - 3.0 * y2).cpu().numpy())
        out["parity_ok"] = bool(out["linearity_rel_err"] <= 4 * out["tolerance"])
    return out


def sky_colrange(F, rows, l, u, Ja, Jb):
    """Block columns [Ja,Jb) of a BlockBandedMatrix(rows, rows, (l,u)) as a matrix of its own: a contiguous slice of
    the parent's `data` (panels are stored column block by column block) whose structure is block-skyline with the
    per-column bandwidths (l+Ja, u-Ja) -- negative ones included, which both the oracle and the library take.
    Returns (offset, length, lvec, uvec)."""
    off, S, klo, khi = F.sky_layout(rows, rows, l, u)
    n = Jb - Ja
    return int(off[Ja]), int(off[Jb] - off[Ja]), np.full(n, l + Ja, dtype=np.int64), np.full(n, u - Ja, dtype=np.int64)


def cfg4_oracle_check(torch, a_t, b_t, c_t, rows, chunks=None, nchunks=8, threads=None):
    """C = A*B of cfg4 against the CPU oracle (per-block OpenBLAS dgemm in the reference's loop order) at FULL size,
    block-column chunk by chunk so that the host never holds more than A and one chunk of B and C: chunk [Ja,Jb) of C
    is A times chunk [Ja,Jb) of B.  Returns (normwise rel. error over the checked chunks, worst chunk, seconds of
    oracle time, chunks checked)."""
    from oracle import cpu as O
    from oracle import formats as F
    O.use_openblas(threads=threads or (os.cpu_count() or 1))
    nblk = len(rows)
    a_h = a_t.cpu().numpy()
    bounds = [nblk * i // nchunks for i in range(nchunks + 1)]
    num2 = den2 = 0.0
    worst = 0.0
    secs = 0.0
    todo = list(range(nchunks)) if chunks is None else list(chunks)
    for ci in todo:
        Ja, Jb = bounds[ci], bounds[ci + 1]
        ob, nb_, lb, ub = sky_colrange(F, rows, 2, 2, Ja, Jb)
        oc, nc_, lc, uc = sky_colrange(F, rows, 4, 4, Ja, Jb)
        b_h = b_t[ob:ob + nb_].cpu().numpy()
        ref = np.zeros(nc_)
        t0 = time.perf_counter()
        O.sky_matmul((rows, rows, 2, 2), (rows, rows[Ja:Jb], lb, ub), (rows, rows[Ja:Jb], lc, uc), 1.0, a_h, b_h, 0.0, ref)
        secs += time.perf_counter() - t0
        ref_t = torch.from_numpy(ref).to(c_t.device)
        d = float(torch.linalg.vector_norm(c_t[oc:oc + nc_] - ref_t).item())
        r = float(torch.linalg.vector_norm(ref_t).item())
        if not np.isfinite(d):
            d = float("inf")
        num2 += d * d
        den2 += r * r
        worst = max(worst, d / max(r, 1e-300))
        del ref_t, ref, b_h
    O.use_builtin()
    return float(np.sqrt(num2) / max(np.sqrt(den2), 1e-300)), worst, secs, len(todo)


def cfg5_oracle_check(torch, a_t, b0, x_t, rows, threads=None):
    """U \\ B of cfg5 against the CPU oracle's block back-substitution (trsv + gemv per block and right-hand side
    column, the reference's loop) at FULL size, every right-hand side.  b0 / x_t are (nrhs, m) row-major tensors
    (= m x nrhs column-major).  Returns (normwise rel. error, seconds)."""
    from oracle import cpu as O
    O.use_openblas(threads=threads or (os.cpu_count() or 1))
    ref = np.asfortranarray(b0.cpu().numpy().T)
    data = a_t.cpu().numpy()
    t0 = time.perf_counter()
    O.sky_trsm(rows, 0, 3, "U", False, data, ref)
    secs = time.perf_counter() - t0
    O.use_builtin()
    ref_t = torch.from_numpy(np.ascontiguousarray(ref.T)).to(x_t.device)
    d = float(torch.linalg.vector_norm(x_t - ref_t).item())
    r = float(torch.linalg.vector_norm(ref_t).item())
    return (d / max(r, 1e-300)) if np.isfinite(d) else float("inf"), secs


def run_cfg4(torch, B, h, stream, reps, oracle, nblk=2048, bs=256, peak_tf=None, oracle_chunks=None):
    rows = [bs] * nblk
    gen = torch.Generator(device="cuda").manual_seed(1004)
    nA = B.BlockSkylineMatrix.numentries(rows, rows, 2, 2)
    nC = B.BlockSkylineMatrix.numentries(rows, rows, 4, 4)
    a_t = torch.randn(nA, generator=gen, device="cuda", dtype=torch.float64) / 16
    b_t = torch.randn(nA, generator=gen, device="cuda", dtype=torch.float64) / 16
    c_t = torch.full((nC,), float("nan"), device="cuda", dtype=torch.float64)
    A = B.BlockSkylineMatrix(wrap(B, h, a_t), rows, rows, 2, 2)
    Bm = B.BlockSkylineMatrix(wrap(B, h, b_t), rows, rows, 2, 2)
    Cm = B.BlockSkylineMatrix(wrap(B, h, c_t), rows, rows, 4, 4)
    assert B.add_bandwidths(A, Bm) == (4, 4)
    avg, best = time_gpu(torch, stream, lambda: B.mul_(Cm, A, Bm), reps, warm=2)
    # flops: sum over (J, N in colsupport(B,J), K in colsupport(A,N)) of 2*bs^3
    triples = 0
    for J in range(nblk):
        for N in range(max(0, J - 2), min(nblk - 1, J + 2) + 1):
            triples += min(nblk - 1, N + 2) - max(0, N - 2) + 1
    flops = 2.0 * bs ** 3 * triples
    nbytes = 8 * (2 * nA + nC)
    out = {"config": f"cfg4: BlockBandedMatrix {nblk} blocks of {bs}, (2,2) x (2,2) -> (4,4), Float64", "ms_avg": avg, "ms_min": best,
           "block_triples": triples, "flops": flops, "TFLOPs": flops / (avg * 1e-3) / 1e12, "algorithmic_bytes": nbytes,
           "GBps": nbytes / (avg * 1e-3) / 1e9, "tolerance": 64 * EPS[8] * 5 * bs}
    if peak_tf:
        out["roofline"] = {"bound": "tensor(fp64)", "peak_TFLOPs": peak_tf, "peak_source": "torch.matmul float64 8192^3 (cuBLAS DGEMM) measured in this run",
                           "frac": out["TFLOPs"] / peak_tf}
    # parity: a sample of C blocks against float64 block products computed by numpy on the host
    rng = np.random.default_rng(7)
    h.synchronize()
    worst = 0.0
    for _ in range(6):
        J = int(rng.integers(0, nblk))
        K = int(rng.integers(max(0, J - 4), min(nblk - 1, J + 4) + 1))
        ref = np.zeros((bs, bs))
        for N in range(max(0, J - 2), min(nblk - 1, J + 2) + 1):
            if abs(K - N) <= 2:
                sa, la = A.block_start(K, N), int(A.block_strides[N])
                sb, lb = Bm.block_start(N, J), int(Bm.block_strides[J])
                Ab = a_t[sa:sa + la * bs].cpu().numpy().reshape((la, bs), order="F")[:bs]
                Bb = b_t[sb:sb + lb * bs].cpu().numpy().reshape((lb, bs), order="F")[:bs]
                ref += Ab @ Bb
        sc, lc = Cm.block_start(K, J), int(Cm.block_strides[J])
        got = c_t[sc:sc + lc * bs].cpu().numpy().reshape((lc, bs), order="F")[:bs]
        worst = max(worst, relerr(got, ref) if np.linalg.norm(ref) else float(np.abs(got).max()))
    out["rel_err_sampled_blocks_vs_numpy"] = worst
    out["nan_in_C"] = bool(torch.isnan(c_t).any().item())
    out["parity_ok"] = bool(worst <= out["tolerance"] and not out["nan_in_C"])
    if oracle:
        err, worst_chunk, secs, nchk = cfg4_oracle_check(torch, a_t, b_t, c_t, rows, chunks=oracle_chunks)
        out["rel_err_vs_oracle"] = err
        out["rel_err_vs_oracle_worst_chunk"] = worst_chunk
        out["oracle_checked"] = f"{nchk} of 8 block-column chunks of C at full size (per-block OpenBLAS dgemm, reference loop order)"
        out["cpu_seconds_oracle_chunks"] = secs
        out["cpu_seconds_full_size_extrapolated"] = secs * 8 / max(nchk, 1)
        out["parity_ok"] = bool(out["parity_ok"] and err <= out["tolerance"])
    return out


def sgemm_simt_peak(torch, n=8192, reps=5):
    """cuBLAS FP32 GEMM with TF32 switched off (FFMA, exact single-precision products) = what a Float32 block product
    that has to agree with the reference's sgemm can be compared with."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        a = torch.randn(n, n, device="cuda", dtype=torch.float32)
        b = torch.randn(n, n, device="cuda", dtype=torch.float32)
        c = torch.empty_like(a)
        for _ in range(2):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def run_cfg4_f32(torch, B, h, stream, reps, nblk=2048, bs=256):
    """cfg4's structure in Float32 (`sky_gemm_simt_kernel<float>`: FFMA, no tensor cores -- single-precision products
    are exact only on the SIMT pipe or through a 3 x TF32 split): throughput against the same-run cuBLAS SGEMM (TF32 off)
    and the nominal FP32 FFMA peak; sampled blocks against float64 block products."""
    rows = [bs] * nblk
    gen = torch.Generator(device="cuda").manual_seed(1004)
    nA = B.BlockSkylineMatrix.numentries(rows, rows, 2, 2)
    nC = B.BlockSkylineMatrix.numentries(rows, rows, 4, 4)
    a_t = torch.randn(nA, generator=gen, device="cuda", dtype=torch.float32) / 16
    b_t = torch.randn(nA, generator=gen, device="cuda", dtype=torch.float32) / 16
    c_t = torch.full((nC,), float("nan"), device="cuda", dtype=torch.float32)
    A = B.BlockSkylineMatrix(wrap(B, h, a_t), rows, rows, 2, 2)
    Bm = B.BlockSkylineMatrix(wrap(B, h, b_t), rows, rows, 2, 2)
    Cm = B.BlockSkylineMatrix(wrap(B, h, c_t), rows, rows, 4, 4)
    avg, best = time_gpu(torch, stream, lambda: B.mul_(Cm, A, Bm), reps, warm=2)
    triples = 0
    for J in range(nblk):
        for N in range(max(0, J - 2), min(nblk - 1, J + 2) + 1):
            triples += min(nblk - 1, N + 2) - max(0, N - 2) + 1
    flops = 2.0 * bs ** 3 * triples
    peak = sgemm_simt_peak(torch)
    props = torch.cuda.get_device_properties(0)
    nominal = props.multi_processor_count * 128 * 2 * 1.965e9 / 1e12
    tf = flops / (avg * 1e-3) / 1e12
    out = {"config": f"cfg4 in Float32: BlockBandedMatrix {nblk} blocks of {bs}, (2,2) x (2,2) -> (4,4)", "ms_avg": avg, "ms_min": best,
           "TFLOPs": tf, "tolerance": 64 * EPS[4] * 5 * bs,
           "roofline": {"bound": "fp32 simt", "peak_TFLOPs": peak, "peak_source": "torch.matmul float32 8192^3 with allow_tf32 = False (cuBLAS SGEMM) measured in this run",
                        "frac": tf / peak, "nominal_ffma_TFLOPs": nominal, "frac_of_nominal": tf / nominal}}
    rng = np.random.default_rng(7)
    h.synchronize()
    worst = 0.0
    for _ in range(6):
        J = int(rng.integers(0, nblk))
        K = int(rng.integers(max(0, J - 4), min(nblk - 1, J + 4) + 1))
        ref = np.zeros((bs, bs))
        for N in range(max(0, J - 2), min(nblk - 1, J + 2) + 1):
            if abs(K - N) <= 2:
                sa, la = A.block_start(K, N), int(A.block_strides[N])
                sb, lb = Bm.block_start(N, J), int(Bm.block_strides[J])
                Ab = a_t[sa:sa + la * bs].cpu().numpy().astype(np.float64).reshape((la, bs), order="F")[:bs]
                Bb = b_t[sb:sb + lb * bs].cpu().numpy().astype(np.float64).reshape((lb, bs), order="F")[:bs]
                ref += Ab @ Bb
        sc, lc = Cm.block_start(K, J), int(Cm.block_strides[J])
        got = c_t[sc:sc + lc * bs].cpu().numpy().astype(np.float64).reshape((lc, bs), order="F")[:bs]
        worst = max(worst, relerr(got, ref) if np.linalg.norm(ref) else float(np.abs(got).max()))
    out["rel_err_sampled_blocks_vs_float64"] = worst
    out["nan_in_C"] = bool(torch.isnan(c_t).any().item())
    out["parity_ok"] = bool(worst <= out["tolerance"] and not out["nan_in_C"])
    return out


def run_bbb_mm(torch, B, h, stream, reps, oracle, n=2048):
    """A*A for the 2-D Laplacian BandedBlockBandedMatrix (n blocks of size n): (1,1),(1,1)^2 -> (2,2),(2,2)."""
    rows = [n] * n
    col = torch.tensor([0, 1, 0, 1, -4, 1, 0, 1, 0], dtype=torch.float64, device="cuda")
    a_t = col.repeat(n * n, 1).contiguous()
    c_t = torch.full((n * n, 25), float("nan"), device="cuda", dtype=torch.float64)
    A = B.BandedBlockBandedMatrix(wrap(B, h, a_t), rows, rows, (1, 1), (1, 1))
    Cm = B.BandedBlockBandedMatrix(wrap(B, h, c_t), rows, rows, (2, 2), (2, 2))
    avg, best = time_gpu(torch, stream, lambda: B.mul_(Cm, A, A), reps, warm=2)
    nbytes = 8 * (2 * 9 + 25) * n * n
    out = {"config": f"BBB x BBB: Laplacian^2, {n} blocks of {n}, (1,1),(1,1) x same -> (2,2),(2,2), Float64", "ms_avg": avg, "ms_min": best,
           "algorithmic_bytes": nbytes, "GBps": nbytes / (avg * 1e-3) / 1e9,
           "roofline": {"bound": "hbm", "peak_GBps": peaks(), "frac": nbytes / (avg * 1e-3) / 1e9 / peaks()}}
    # property: (A*A)*x == A*(A*x); interior stencil value check: centre of the 13-point biharmonic-like stencil = 20
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(n * n, generator=gen, device="cuda", dtype=torch.float64)
    y1 = torch.empty_like(x); t = torch.empty_like(x); y2 = torch.empty_like(x)
    B.mul_(wrap(B, h, y1), Cm, wrap(B, h, x))
    B.mul_(wrap(B, h, t), A, wrap(B, h, x)); B.mul_(wrap(B, h, y2), A, wrap(B, h, t))
    h.synchronize()
    out["assoc_rel_err"] = relerr(y1.cpu().numpy(), y2.cpu().numpy())
    out["parity_ok"] = bool(out["assoc_rel_err"] <= 64 * EPS[8] * 25 * 4)
    return out


def run_sky_mv(torch, B, h, stream, reps, oracle, nblk=2048, bs=256, nrhs=1):
    """Skyline matvec / multi-vector product on cfg4's operand (BlockBandedMatrix fill(256,2048), (2,2))."""
    rows = [bs] * nblk
    gen = torch.Generator(device="cuda").manual_seed(1006)
    nA = B.BlockSkylineMatrix.numentries(rows, rows, 2, 2)
    m = bs * nblk
    a_t = torch.randn(nA, generator=gen, device="cuda", dtype=torch.float64) / 16
    x_t = torch.randn(nrhs, m, generator=gen, device="cuda", dtype=torch.float64)
    y_t = torch.full((nrhs, m), float("nan"), device="cuda", dtype=torch.float64)
    A = B.BlockSkylineMatrix(wrap(B, h, a_t), rows, rows, 2, 2)
    dx = wrap(B, h, x_t if nrhs > 1 else x_t[0])
    dy = wrap(B, h, y_t if nrhs > 1 else y_t[0])
    avg, best = time_gpu(torch, stream, lambda: B.mul_(dy, A, dx), reps)
    nbytes = 8 * (nA + 2 * m * nrhs)
    out = {"config": f"skyline matvec: BlockBandedMatrix {nblk} blocks of {bs}, (2,2), Float64, nrhs={nrhs}", "ms_avg": avg, "ms_min": best,
           "algorithmic_bytes": nbytes, "GBps": nbytes / (avg * 1e-3) / 1e9,
           "roofline": {"bound": "hbm", "peak_GBps": peaks(), "frac": nbytes / (avg * 1e-3) / 1e9 / peaks()}, "tolerance": 64 * EPS[8] * 5 * bs}
    # parity on a sample of block rows against numpy
    h.synchronize()
    rng = np.random.default_rng(3)
    worst = 0.0
    xh = x_t.cpu().numpy()
    for _ in range(4):
        K = int(rng.integers(0, nblk))
        ref = np.zeros((bs, nrhs))
        for J in range(max(0, K - 2), min(nblk - 1, K + 2) + 1):
            s, ld = A.block_start(K, J), int(A.block_strides[J])
            blk = a_t[s:s + ld * bs].cpu().numpy().reshape((ld, bs), order="F")[:bs]
            ref += blk @ xh[:, J * bs:(J + 1) * bs].T
        got = y_t[:, K * bs:(K + 1) * bs].cpu().numpy().T
        worst = max(worst, relerr(got, ref))
    out["rel_err_sampled_rows_vs_numpy"] = worst
    out["parity_ok"] = bool(worst <= out["tolerance"])
    return out


def run_cfg5(torch, B, h, stream, reps, oracle, nblk=1024, bs=512, nrhs=64, peak_tf=None):
    rows = [bs] * nblk
    m = bs * nblk
    gen = torch.Generator(device="cuda").manual_seed(1005)
    nA = B.BlockSkylineMatrix.numentries(rows, rows, 0, 3)
    a_t = torch.randn(nA, generator=gen, device="cuda", dtype=torch.float64) / np.sqrt(bs)
    A = B.BlockSkylineMatrix(wrap(B, h, a_t), rows, rows, 0, 3)
    # +10 on the global diagonal (well conditioned)
    idx = torch.arange(bs, device="cuda", dtype=torch.int64)
    starts = torch.tensor([A.block_start(K, K) for K in range(nblk)], device="cuda", dtype=torch.int64)
    lds = torch.tensor([int(A.block_strides[K]) for K in range(nblk)], device="cuda", dtype=torch.int64)
    a_t[(starts[:, None] + idx[None, :] * (lds[:, None] + 1)).reshape(-1)] += 10.0
    b0 = torch.randn(nrhs, m, generator=gen, device="cuda", dtype=torch.float64)   # row-major (nrhs, m) == m x nrhs col-major
    b_t = b0.clone()
    dB = wrap(B, h, b_t)
    T = B.UpperTriangular(A)

    def step():
        b_t.copy_(b0)
        B.ldiv_(T, dB)

    # time the copy alone and subtract (the solve is in place)
    with torch.cuda.stream(stream):
        cavg, _ = time_gpu(torch, stream, lambda: b_t.copy_(b0), 5, warm=1)
    avg, best = time_gpu(torch, stream, step, reps, warm=1)
    avg -= cavg; best -= cavg
    trsm_flops = nblk * bs * bs * nrhs
    upd_blocks = sum(min(3, K) for K in range(nblk))
    flops = trsm_flops + 2.0 * upd_blocks * bs * bs * nrhs
    nbytes = 8 * (nA + 2 * m * nrhs)
    out = {"config": f"cfg5: UpperTriangular BlockBandedMatrix {nblk} blocks of {bs}, (0,3), {nrhs} RHS, Float64 ldiv!",
           "ms_avg": avg, "ms_min": best, "us_per_block_step": avg * 1e3 / nblk, "flops": flops, "TFLOPs": flops / (avg * 1e-3) / 1e12,
           "algorithmic_bytes": nbytes, "GBps": nbytes / (avg * 1e-3) / 1e9, "tolerance": 64 * EPS[8] * 4 * bs,
           "roofline": {"bound": "latency chain (1024 sequential block steps); stream 1x matrix + tensor flops as lower bounds",
                        "hbm_frac": nbytes / (avg * 1e-3) / 1e9 / peaks(), "fp64_frac": (flops / (avg * 1e-3) / 1e12 / peak_tf) if peak_tf else None}}
    h.synchronize()
    X = b_t.cpu().numpy()          # (nrhs, m): row q = solution of column q
    if oracle:
        err, secs = cfg5_oracle_check(torch, a_t, b0, b_t, rows)
        out["cpu_seconds_64_columns"] = secs
        out["cpu_seconds_per_rhs_column"] = secs / nrhs
        out["rel_err_vs_oracle"] = err
        out["oracle_checked"] = f"all {nrhs} right-hand sides at full size (OpenBLAS trsv + gemv per block, reference loop order)"
        out["parity_ok"] = bool(err <= out["tolerance"])
    else:
        out["parity_ok"] = bool(np.isfinite(X).all())
    return out


def host_block_banded_qr(F, rows, cols, l, u, data):
    """qr! of a block-banded matrix on the host with LAPACK (scipy dgeqrf per panel, dormqr on the trailing blocks):
    the loop of src/blockskylineqr.jl:29-43 on the skyline storage `data` of the (l, l+u) factors, in place."""
    from scipy.linalg import lapack
    as_strided = np.lib.stride_tricks.as_strided
    off, S, klo, khi = F.sky_layout(rows, cols, l, l + u)
    R = np.concatenate([[0], np.cumsum(rows)]); N = len(cols)
    tau = np.zeros(int(np.sum(cols)))
    C0 = np.concatenate([[0], np.cumsum(cols)])
    it = data.itemsize
    for K in range(N):
        K1 = min(K + l, len(rows) - 1)
        Sr, c = int(R[K1 + 1] - R[K]), int(cols[K])
        V = as_strided(data[int(off[K] + R[K] - R[klo[K]]):], shape=(Sr, c), strides=(it, it * int(S[K])))
        qr_, t_, _, info = lapack.dgeqrf(np.asfortranarray(V))
        assert info == 0
        V[:, :] = qr_
        tau[C0[K]:C0[K] + len(t_)] = t_
        for J in range(K + 1, min(K + l + u, N - 1) + 1):
            Bv = as_strided(data[int(off[J] + R[K] - R[klo[J]]):], shape=(Sr, int(cols[J])), strides=(it, it * int(S[J])))
            cq, _, info = lapack.dormqr("L", "T", qr_, t_, np.asfortranarray(Bv), max(64, 64 * int(cols[J])))
            assert info == 0
            Bv[:, :] = cq
    return tau


def host_qr_solve(F, rows, l, u, data, tau, Bm, O):
    """ldiv!(F, B) on the host: dormqr per panel (src/blockskylineqr.jl:77-127), then the block triangular solve"""
    from scipy.linalg import lapack
    as_strided = np.lib.stride_tricks.as_strided
    off, S, klo, khi = F.sky_layout(rows, rows, l, l + u)
    R = np.concatenate([[0], np.cumsum(rows)]); N = len(rows)
    it = data.itemsize
    for K in range(N):
        K1 = min(K + l, N - 1)
        Sr, c = int(R[K1 + 1] - R[K]), int(rows[K])
        V = np.asfortranarray(as_strided(data[int(off[K] + R[K] - R[klo[K]]):], shape=(Sr, c), strides=(it, it * int(S[K]))))
        cq, _, info = lapack.dormqr("L", "T", V, tau[R[K]:R[K] + min(Sr, c)], np.asfortranarray(Bm[R[K]:R[K1 + 1], :]), max(64, 64 * Bm.shape[1]))
        assert info == 0
        Bm[R[K]:R[K1 + 1], :] = cq
    O.sky_trsm(rows, l, l + u, "U", False, data, Bm)
    return Bm


def run_qr(torch, B, h, stream, reps, oracle, nblk=512, bs=256, nrhs=64, peak_tf=None):
    """F = qr(A) on the host (as the reference does), F \\ B on the device: BlockBandedMatrix nblk blocks of bs, (1,1)."""
    from oracle import formats as F
    from oracle import cpu as O
    rows = [bs] * nblk
    l, u = 1, 1
    rng = np.random.default_rng(1011)
    nA = F.sky_numentries(rows, rows, l, l + u)
    data = np.zeros(nA)
    off, S, klo, khi = F.sky_layout(rows, rows, l, l + u)
    Rr = np.concatenate([[0], np.cumsum(rows)])
    # A with bandwidths (1,1) inside the (1,2) factor structure: fill blocks K-1..K+1 of every block column
    for J in range(nblk):
        ld = int(S[J])
        pan = data[int(off[J]):int(off[J]) + ld * bs].reshape((ld, bs), order="F")
        lo = int(Rr[max(J - u, 0)] - Rr[klo[J]])
        pan[lo:, :] = rng.standard_normal((ld - lo, bs)) / np.sqrt(3 * bs)
        dg = int(Rr[J] - Rr[klo[J]])
        pan[dg + np.arange(bs), np.arange(bs)] += 2.0
    Adata = data.copy()
    t0 = time.perf_counter()
    tau = host_block_banded_qr(F, rows, rows, l, u, data)
    t_fact = time.perf_counter() - t0
    m = bs * nblk
    B0 = np.asfortranarray(rng.standard_normal((m, nrhs)))
    Fd = B.BlockBandedQR(B.BlockSkylineMatrix(data, rows, rows, l, l + u), tau).to_device(h)
    b0_t = torch.from_numpy(np.ascontiguousarray(B0.T)).cuda()        # (nrhs, m) row-major == m x nrhs column-major
    b_t = b0_t.clone()
    dB = wrap(B, h, b_t)

    def step():
        b_t.copy_(b0_t)
        B.ldiv_(Fd, dB)

    def step_q():
        b_t.copy_(b0_t)
        B.lmul_adj_(Fd, dB)

    with torch.cuda.stream(stream):
        cavg, _ = time_gpu(torch, stream, lambda: b_t.copy_(b0_t), 5, warm=1)
    qavg, _ = time_gpu(torch, stream, step_q, reps, warm=1)
    avg, best = time_gpu(torch, stream, step, reps, warm=1)
    h.synchronize()
    X = b_t.cpu().numpy().T
    out = {"config": f"qr: F \\ B for F = qr(BlockBandedMatrix {nblk} blocks of {bs}, (1,1)) computed on the host, {nrhs} RHS, Float64",
           "host_qr_seconds_lapack": t_fact, "ms_ldiv": avg - cavg, "ms_lmul_adj": qavg - cavg, "us_per_panel_lmul_adj": (qavg - cavg) * 1e3 / nblk,
           "tolerance": 64 * EPS[8] * 4 * bs}
    # qr! itself on the device, from the same matrix (the solve above used the host factors)
    a0_t = torch.from_numpy(Adata).cuda()
    a_t = a0_t.clone()
    Aq = B.BlockSkylineMatrix(wrap(B, h, a_t), rows, rows, l, l + u)

    def fact():
        a_t.copy_(a0_t)
        B.qr_(Aq)

    with torch.cuda.stream(stream):
        c2, _ = time_gpu(torch, stream, lambda: a_t.copy_(a0_t), 3, warm=1)
    favg, _ = time_gpu(torch, stream, fact, max(2, reps // 2), warm=1)
    out["ms_qr_device"] = favg - c2
    with torch.cuda.stream(stream):
        a_t.copy_(a0_t)
    Fq = B.qr_(Aq)
    h.synchronize()
    out["factors_rel_diff_device_vs_lapack"] = relerr(a_t.cpu().numpy(), data)
    out["tau_max_diff_device_vs_lapack"] = float(np.abs(Fq.tau.to_host() - tau).max())
    with torch.cuda.stream(stream):
        b_t.copy_(b0_t)
    B.ldiv_(Fq, dB)
    h.synchronize()
    out["rel_diff_solution_device_factors_vs_host_factors"] = relerr(b_t.cpu().numpy().T, X)
    with torch.cuda.stream(stream):
        b_t.copy_(torch.from_numpy(np.ascontiguousarray(X.T)).cuda())
    h.synchronize()
    # residual against the original matrix: A X = B
    Ad = B.BlockSkylineMatrix(Adata, rows, rows, l, l + u).to_device(h)
    back = h.empty((m, nrhs), np.float64)
    B.mul_(back, Ad, wrap(B, h, b_t))
    h.synchronize()
    out["residual_rel_AX_minus_B"] = relerr(back.to_host(), B0)
    if oracle:
        O.use_openblas(threads=os.cpu_count() or 1)
        ref = B0.copy(order="F")
        t0 = time.perf_counter()
        host_qr_solve(F, rows, l, u, data, tau, ref, O)
        out["cpu_seconds_ldiv_lapack"] = time.perf_counter() - t0
        O.use_builtin()
        out["rel_err_vs_host_lapack"] = relerr(X, ref)
        out["parity_ok"] = bool(out["rel_err_vs_host_lapack"] <= out["tolerance"])
    return out


def run_bbb_tri(torch, B, h, stream, reps, oracle, n=4096):
    """Gauss-Seidel sweep of examples/finitedifference_2d.jl: (D+L) \\ b and (D+U) \\ b on the n^2 Laplacian."""
    A = B.BandedBlockBandedMatrix.finitedifference_2d(n)
    hdata = A.data
    gen = torch.Generator(device="cuda").manual_seed(1007)
    a_t = torch.from_numpy(np.ascontiguousarray(hdata.T)).cuda()     # (n*n, 9) row-major == 9 x n*n column-major
    Ad = B.BandedBlockBandedMatrix(wrap(B, h, a_t), [n] * n, [n] * n, (1, 1), (1, 1))
    b0 = torch.randn(n * n, generator=gen, device="cuda", dtype=torch.float64)
    b_t = b0.clone()
    db = wrap(B, h, b_t)
    out = {"config": f"bbbtri: Lower/UpperTriangular of the 2-D Laplacian BBB, {n} blocks of {n}, Float64 ldiv!", "steps": 3 * n - 2,
           "tolerance": 64 * EPS[8] * 9}
    for uplo, T in (("L", B.LowerTriangular(Ad)), ("U", B.UpperTriangular(Ad))):
        def step():
            b_t.copy_(b0)
            B.ldiv_(T, db)
        with torch.cuda.stream(stream):
            cavg, _ = time_gpu(torch, stream, lambda: b_t.copy_(b0), 5, warm=1)
        avg, best = time_gpu(torch, stream, step, reps, warm=1)
        out[f"ms_avg_{uplo}"] = avg - cavg
        out[f"us_per_level_{uplo}"] = (avg - cavg) * 1e3 / (3 * n - 2)
        h.synchronize()
        if oracle:
            from oracle import cpu as O
            ref = b0.cpu().numpy().copy()
            t0 = time.perf_counter()
            O.bbb_trsm([n] * n, 1, 1, 1, 1, uplo, False, hdata, ref)
            out[f"cpu_seconds_{uplo}"] = time.perf_counter() - t0
            out[f"rel_err_{uplo}"] = relerr(b_t.cpu().numpy(), ref)
    if oracle:
        out["parity_ok"] = bool(max(out["rel_err_L"], out["rel_err_U"]) <= out["tolerance"])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,3,4,5")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (experiments)")
    args = ap.parse_args()
    import torch
    import bbm_b200 as B
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    h = B.Handle(0, stream=stream.cuda_stream)
    for kv in args.opt:
        k, v = kv.split("=")
        h.set_option(k, int(v))
    want = set(args.configs.split(","))
    oracle = not args.no_oracle
    peak_tf = dgemm_peak(torch) if want & {"4", "5"} else None
    if peak_tf:
        print(json.dumps({"measured": "cuBLAS DGEMM 8192^3 via torch.matmul", "TFLOPs": peak_tf}), flush=True)
    if "1" in want:
        r = [100] * 100
        print(json.dumps(run_bbb(torch, B, h, stream, "cfg1: Laplacian BBB 100x100 blocks of 100, F64 (L2-resident: latency, not a roofline)", r, r, (1, 1), (1, 1), np.float64, 200, oracle, laplacian=100)), flush=True)
    if "2" in want:
        r = [4096] * 4096
        print(json.dumps(run_bbb(torch, B, h, stream, "cfg2: Laplacian BBB 4096 blocks of 4096, F64", r, r, (1, 1), (1, 1), np.float64, args.reps, oracle, laplacian=4096)), flush=True)
    if "3" in want:
        r = list(range(1, 4001))
        for dt in (np.float64, np.float32):
            print(json.dumps(run_bbb(torch, B, h, stream, "cfg3: BBB rows=cols=1:4000, (2,2),(2,2)", r, r, (2, 2), (2, 2), dt, args.reps, oracle)), flush=True)
    if "4" in want:
        print(json.dumps(run_cfg4(torch, B, h, stream, max(2, args.reps // 3), oracle, peak_tf=peak_tf)), flush=True)
    if "4f32" in want:
        print(json.dumps(run_cfg4_f32(torch, B, h, stream, max(2, args.reps // 3))), flush=True)
    if "bbbmm" in want:
        print(json.dumps(run_bbb_mm(torch, B, h, stream, args.reps, oracle)), flush=True)
    if "qr" in want:
        print(json.dumps(run_qr(torch, B, h, stream, max(2, args.reps // 3), oracle, peak_tf=peak_tf)), flush=True)
    if "bbbtri" in want:
        print(json.dumps(run_bbb_tri(torch, B, h, stream, max(2, args.reps // 3), oracle)), flush=True)
    if "4mv" in want:
        for q in (1, 8):
            print(json.dumps(run_sky_mv(torch, B, h, stream, args.reps, oracle, nrhs=q)), flush=True)
    if "5" in want:
        print(json.dumps(run_cfg5(torch, B, h, stream, max(2, args.reps // 3), oracle, peak_tf=peak_tf)), flush=True)


if __name__ == "__main__":
    main()
