#!/usr/bin/env python
"""Benchmark of the hot path named in BASELINE.json: BandedBlockBandedMatrix matvec on the
16.7M-unknown 2-D Laplacian (4096 blocks of size 4096, (l,u)=(lam,mu)=(1,1), Float64).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours       one step = one y <- A*x through the C ABI (libbbm_b200.so).  `value` is algorithmic GB/s
           with A, x, y resident in HBM (CUDA events on the launch stream, max over ranks); `e2e`
           is the same metric through the host-vector call (bbm_bbb_mul_host_f64: pinned x -> HBM,
           kernel, y -> pinned host) timed by wall clock around the synchronous call.
reference  the reference's CPU path (per-block OpenBLAS dgbmv in the [upstream] block-loop order,
           restated in oracle/bbm_oracle.c because Julia is not in the image) on a bounded sample
           of the same workload, on the box's host cores.

Algorithmic bytes per matvec: sizeof(T)*(P*n + n + m) = 1 476 395 008 B for cfg2.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "bbb_matvec_algorithmic_GBps"
UNIT = "GB/s"
BLOCK = 4096  # cfg2: 4096 blocks of size 4096


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json; null if there is none for this kernel)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def algorithmic_bytes(P, n, m, itemsize=8, nrhs=1):
    return itemsize * (P * n + n * nrhs + m * nrhs)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for t, line in self.rows:
            if t0 is not None and not (t0 - 0.05 <= t <= t1 + 0.05):
                continue
            f = [s.strip() for s in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's per-block BLAS path (oracle restatement) on a bounded sample
# --------------------------------------------------------------------------------------------------
def cpu_reference_sample(nblocks, reps, threads_list):
    """Time y <- A*x on the first `nblocks` block rows/cols of the cfg2 Laplacian (same block size,
    same bandwidths, same values) through the oracle bound to OpenBLAS.  Returns the best config."""
    from oracle import cpu as O  # the one place bench.py may execute the oracle
    import bbm_b200 as B

    A = B.BandedBlockBandedMatrix.finitedifference_2d(BLOCK, nblocks=nblocks)
    rows = [BLOCK] * nblocks
    rng = np.random.default_rng(1001)
    x = rng.standard_normal(BLOCK * nblocks)
    y = np.zeros(BLOCK * nblocks)
    nbytes = algorithmic_bytes(9, BLOCK * nblocks, BLOCK * nblocks)
    best = None
    blas = O.use_openblas(threads=1)
    for th in threads_list if blas else [1]:
        if blas:
            O.use_openblas(threads=th)
        O.bbb_mul(rows, rows, 1, 1, 1, 1, 1.0, A.data, x, 0.0, y)  # warm-up
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            O.bbb_mul(rows, rows, 1, 1, 1, 1, 1.0, A.data, x, 0.0, y)
            ts.append(time.perf_counter() - t0)
        t = min(ts)
        cand = {"value": nbytes / t / 1e9, "unit": UNIT, "cores": int(th), "seconds": t,
                "blas": "OpenBLAS 0.3.30 cblas_dgbmv per block" if blas else "builtin reference-order loops"}
        if best is None or cand["value"] > best["value"]:
            best = cand
    best["kind"] = "port"
    best["sample"] = (f"first {nblocks} of 4096 block rows/cols of the cfg2 Laplacian (block size {BLOCK}, "
                      f"{3 * nblocks - 2} dgbmv calls, {nbytes} algorithmic bytes), best of {reps}")
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    nblocks = 256
    # one step = one pass over the sample; the driver's K/W are honoured
    from oracle import cpu as O
    import bbm_b200 as B
    A = B.BandedBlockBandedMatrix.finitedifference_2d(BLOCK, nblocks=nblocks)
    rows = [BLOCK] * nblocks
    x = np.random.default_rng(1001).standard_normal(BLOCK * nblocks)
    y = np.zeros(BLOCK * nblocks)
    nbytes = algorithmic_bytes(9, BLOCK * nblocks, BLOCK * nblocks)
    best = None
    blas = O.use_openblas(threads=1)
    for th in ([1, ncores] if blas and ncores > 1 else [1]):
        if blas:
            O.use_openblas(threads=th)
        for _ in range(max(1, args.warmup)):
            O.bbb_mul(rows, rows, 1, 1, 1, 1, 1.0, A.data, x, 0.0, y)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.bbb_mul(rows, rows, 1, 1, 1, 1, 1.0, A.data, x, 0.0, y)
        t = (time.perf_counter() - t0) / args.steps
        if best is None or t < best[0]:
            best = (t, th)
    t, th = best
    val = nbytes / t / 1e9
    sample = (f"first {nblocks} of 4096 block rows/cols of the cfg2 Laplacian (block size {BLOCK}, "
              f"{3 * nblocks - 2} dgbmv calls per step, {nbytes} algorithmic bytes per step)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2: 2D Laplacian BandedBlockBandedMatrix, 4096 blocks of size 4096, (1,1),(1,1), Float64 matvec",
                   "sample": sample, "host_threads_tried": [1, ncores]},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": th, "kind": "port", "sample": sample,
                         "blas": "OpenBLAS 0.3.30 cblas_dgbmv per block" if blas else "builtin loops"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    # Libraries (NCCL's version banner, for one) write to fd 1: keep stdout for the ONE JSON line by
    # pointing fd 1 at stderr until the result is printed.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import bbm_b200 as B

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # NCCL_DEBUG=VERSION/INFO banners go to stdout by default: keep stdout for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    nb = args.blocks
    n = BLOCK * nb
    stream = torch.cuda.Stream(device=local_rank)
    h = B.Handle(local_rank, stream=stream.cuda_stream)
    for kv in args.opt:
        k, v = kv.split("=")
        h.set_option(k, int(v))

    if world == 1:
        result = bench_single(args, torch, B, h, stream, nb, n, rank)
    else:
        result = bench_distributed(args, torch, dist, B, h, stream, nb, BLOCK, rank, world, local_rank)
    if rank == 0:
        peak, how = peaks()
        peak *= world  # aggregate achieved GB/s against N x one GPU's measured copy bandwidth
        result["roofline"]["peak"] = peak
        result["roofline"]["frac"] = result["roofline"]["achieved"] / peak
        result["roofline"]["peak_source"] = f"{world} x MEASURED_PEAKS.json hbm_gbs ({how})"
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(result), flush=True)
        os.dup2(2, 1)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def bench_single(args, torch, B, h, stream, nb, n, rank):
    K, W = args.steps, args.warmup
    A = B.BandedBlockBandedMatrix.finitedifference_2d(BLOCK, nblocks=nb)
    Ad = A.to_device(h)
    del A
    rng = np.random.default_rng(1001)
    hx = h.pinned(n, np.float64)
    hy = h.pinned(n, np.float64)
    hx.array[:] = rng.standard_normal(n)
    dx = h.to_device(hx.array)
    dy = h.empty(n, np.float64)
    nbytes = algorithmic_bytes(9, n, n)

    launches0 = h.launch_count
    with torch.cuda.stream(stream):
        for _ in range(W):
            B.mul_(dy, Ad, dx)
        torch.cuda.synchronize()
        sampler = ClockSampler(h.device).start()
        time.sleep(0.15)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        l0 = h.launch_count
        t0 = time.perf_counter()
        ev[0].record(stream)
        for i in range(K):
            B.mul_(dy, Ad, dx)
            ev[i + 1].record(stream)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        launches = h.launch_count - l0
        clocks = sampler.stop(t0, t1)
    total_ms = ev[0].elapsed_time(ev[K])
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    ms_per_step = total_ms / K
    value = nbytes / (ms_per_step * 1e-3) / 1e9
    kernel_ms = float(np.mean(per))

    # end to end through the reference-facing host-vector call: x from pinned host memory, y back to it
    for _ in range(max(1, min(W, 3))):
        B.mul_(hy.array, Ad, hx.array)
    h.synchronize()
    Ke = max(3, min(K, 10))
    t0 = time.perf_counter()
    for _ in range(Ke):
        B.mul_(hy.array, Ad, hx.array)
    h.synchronize()
    e2e_s = (time.perf_counter() - t0) / Ke
    # the e2e result is the same vector the device-resident path produced
    ydev = dy.to_host()
    if not np.array_equal(ydev, hy.array):
        raise SystemExit("bench: host-vector path and device path disagree")

    cpu = None
    if not args.no_cpu:
        cpu = cpu_reference_sample(256, 5, [1, os.cpu_count() or 1])

    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cfg2: 2D Laplacian BandedBlockBandedMatrix, {nb} blocks of size {BLOCK}, (1,1),(1,1), Float64 matvec"
                               + ("" if nb == BLOCK else " [REDUCED block count]"),
                   "unknowns": n, "algorithmic_bytes_per_step": nbytes, "l2": "inputs (1.48 GB) exceed the 126 MB L2; no flush",
                   "parallelism": "1 GPU"},
        "clocks": clocks,
        "e2e": {"value": nbytes / e2e_s / 1e9, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "steps": Ke,
                "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n, "call": "bbm_bbb_mul_host_f64"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "bbb_walk_kernel<double,3,3>", "achieved": nbytes / (kernel_ms * 1e-3) / 1e9,
                     "peak": None, "unit": UNIT, "frac": None,
                     "traffic": ncu_traffic("bbb_walk_kernel<double,3,3>") if nb == BLOCK else None,
                     "kernel_ms_avg": kernel_ms, "kernel_ms_min": float(min(per)), "algorithmic_bytes_per_launch": nbytes},
        "cpu_baseline": cpu,
    }


# --------------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1): strong scaling of the cfg2 Laplacian matvec
# --------------------------------------------------------------------------------------------------
def bench_distributed(args, torch, dist, B, h, stream, nb, block, rank, world, local_rank):
    from bbm_b200 import _lib as L
    from bbm_b200.distributed import Communicator, DistBandedBlockBandedMatrix
    K, W = args.steps, args.warmup
    n = block * nb
    rows = [block] * nb
    comm = Communicator(h, dist, world, rank)
    A = DistBandedBlockBandedMatrix(comm, rows, rows, (1, 1), (1, 1))
    lay = A.layout
    col = (np.array([0, 1, 0, 1, -4, 1, 0, 1, 0], dtype=np.float64) * float(block) ** 2)
    slab = np.empty((9, lay["n_local"]), order="F")
    slab[:, :] = col[:, None]
    A.set_local_data(slab)
    del slab
    # the same global x on every rank (seeded), each keeps its owned part; halos start as NaN
    xg = np.random.default_rng(1001).standard_normal(n)
    xl = np.full(lay["n_local"], np.nan)
    o = lay["own_offset"]
    xl[o:o + lay["n_own"]] = xg[lay["col_offset"] + o: lay["col_offset"] + o + lay["n_own"]]
    dx = h.to_device(xl)
    dy = h.empty(lay["m_own"], np.float64)
    nbytes = algorithmic_bytes(9, n, n)
    # fused halo push over NVLink peer memory (one kernel per step); falls back to the NCCL exchange
    no_peer = any(o.replace(" ", "") == "dist.peer=0" for o in args.opt)
    try:
        fused = (not no_peer) and A.enable_peer()   # collective: every rank takes the same branch
    except Exception as exc:  # a rank-local setup failure is reported by the library as "unsupported" on all ranks
        print(f"[rank {rank}] fused halo push unavailable: {exc}", file=sys.stderr)
        fused = False

    with torch.cuda.stream(stream):
        for _ in range(W):
            A.mul_(dy, dx)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(h.device).start() if rank == 0 else None
        time.sleep(0.15)
        dist.barrier()
        torch.cuda.synchronize()
        l0 = h.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(K):
            A.mul_(dy, dx)
        ev1.record(stream)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        launches = h.launch_count - l0
        dist.barrier()
        clocks = sampler.stop(t0, t1) if sampler else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=f"cuda:{local_rank}", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / K

    # end to end: every step uploads the owned x from pinned host memory and reads y_own back
    hx = h.pinned(lay["n_own"], np.float64)
    hy = h.pinned(lay["m_own"], np.float64)
    hx.array[:] = xl[o:o + lay["n_own"]]
    import ctypes as C_
    lib = h._lib
    xown_ptr = dx.ptr + 8 * o

    def e2e_step():
        L.check(lib.bbm_memcpy_h2d(h.h, C_.c_void_p(xown_ptr), C_.c_void_p(hx.ptr), 8 * lay["n_own"]), h.h)
        A.mul_(dy, dx)
        L.check(lib.bbm_memcpy_d2h(h.h, C_.c_void_p(hy.ptr), C_.c_void_p(dy.ptr), 8 * lay["m_own"]), h.h)
        h.synchronize()

    for _ in range(2):
        e2e_step()
    dist.barrier()
    Ke = max(3, min(K, 10))
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    e2e = torch.tensor([(time.perf_counter() - t0) / Ke], device=f"cuda:{local_rank}", dtype=torch.float64)
    dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e.item())

    # correctness of the distributed result: checksum of y against a single-rank recomputation of
    # the stencil on the owned rows (cheap: the 5-point stencil in numpy)
    y = dy.to_host()
    X = xg.reshape(nb, block)
    K0, K1 = lay["K0"], lay["K1"]
    Y = -4.0 * X[K0:K1]
    Y[:, 1:] += X[K0:K1, :-1]
    Y[:, :-1] += X[K0:K1, 1:]
    if K0 > 0:
        Y += X[K0 - 1:K1 - 1]
    else:
        Y[1:] += X[K0:K1 - 1]
    if K1 < nb:
        Y += X[K0 + 1:K1 + 1]
    else:
        Y[:-1] += X[K0 + 1:K1]
    Y *= float(block) ** 2
    err = float(np.linalg.norm(y - Y.reshape(-1)) / max(np.linalg.norm(Y), 1e-300))
    errt = torch.tensor([err], device=f"cuda:{local_rank}", dtype=torch.float64)
    dist.all_reduce(errt, op=dist.ReduceOp.MAX)
    if float(errt.item()) > 64 * 2.22e-16 * 9:
        raise SystemExit(f"bench: distributed matvec is wrong (rel err {float(errt.item()):.3e})")
    lt = torch.tensor([launches], device=f"cuda:{local_rank}", dtype=torch.float64)
    dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    active, timed_out = A.peer_status()
    if timed_out:
        raise SystemExit("bench: a peer-memory halo wait timed out")

    value = nbytes / (ms_per_step * 1e-3) / 1e9
    return {
        "metric": "bbb_matvec_algorithmic_GBps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cfg2: 2D Laplacian BandedBlockBandedMatrix, {nb} blocks of size {block}, (1,1),(1,1), Float64 matvec"
                               + ("" if nb == block else " [REDUCED block count]"),
                   "unknowns": n, "algorithmic_bytes_per_step": nbytes,
                   "l2": "per-rank inputs (%.0f MB) exceed the 126 MB L2; no flush" % (nbytes / world / 1e6),
                   "parallelism": (f"row-block partition over {world} GPUs, x halo (l+u block columns) pushed into the neighbours' "
                                   "x_local over NVLink peer memory by CTA 0 of the same band-walk kernel (1 launch per rank per step)"
                                   if active else
                                   f"row-block partition over {world} GPUs, x halo (l+u block columns) by NCCL send/recv overlapped with interior rows"),
                   "max_rel_err_vs_stencil": float(errt.item())},
        "clocks": clocks,
        "e2e": {"value": nbytes / e2e_s / 1e9, "unit": "GB/s", "ms_per_step": e2e_s * 1e3, "steps": Ke,
                "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n,
                "call": "bbm_memcpy_h2d(x_own) + bbm_dist_bbb_mul_f64 + bbm_memcpy_d2h(y_own), per rank"},
        "gpu_launches": int(lt.item()),
        "roofline": {"bound": "hbm", "kernel": "bbb_walk_kernel<double,3,3> (per rank, incl. halo push / exchange)",
                     "achieved": value, "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                     "note": "aggregate over ranks; peak below is one GPU's, so frac ~ N at perfect scaling"},
        "cpu_baseline": None,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--blocks", type=int, default=BLOCK, help="number of blocks (default 4096 = cfg2; smaller only for debugging)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (experiments), e.g. dist.graph=0")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
