#!/usr/bin/env python
"""Benchmark of the hot path named in BASELINE.json: BandedBlockBandedMatrix matvec on the
16.7M-unknown 2-D Laplacian (4096 blocks of size 4096, (l,u)=(lam,mu)=(1,1), Float64).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours       one step = one y <- A*x through the C ABI (libbbm_b200.so).  `value` is algorithmic GB/s
           with A, x, y resident in HBM (CUDA events on the launch stream, max over ranks); `e2e`
           is the same metric through the host-vector call (bbm_bbb_mul_host_f64: pinned x -> HBM,
           kernel, y -> pinned host) timed by wall clock around the synchronous call.
reference  the reference's CPU path (per-block OpenBLAS dgbmv in the [upstream] block-loop order,
           restated in oracle/bbm_oracle.c because Julia is not in the image) on the SAME full-size
           workload (one step = one full 4096-block matvec, ~0.3 s), on the box's host cores, timed
           with 1 BLAS thread and with all of them; the better one is the line's value.

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


def workload_config(nb, world):
    """`config` of the JSON line: a function of the workload alone, so that both arms print the same object."""
    n = BLOCK * nb
    nbytes = algorithmic_bytes(9, n, n)
    return {"workload": f"cfg2: 2D Laplacian BandedBlockBandedMatrix, {nb} blocks of size {BLOCK}, (1,1),(1,1), Float64 matvec"
                        + ("" if nb == BLOCK else " [REDUCED block count]"),
            "unknowns": n, "algorithmic_bytes_per_step": nbytes,
            "l2": "inputs (%.0f MB per GPU) exceed the 126 MB L2; no flush" % (nbytes / world / 1e6),
            "parallelism": "1 GPU" if world == 1 else f"block rows partitioned over {world} GPUs, x halo of l+u block columns exchanged"}


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's per-block BLAS path (oracle restatement) on the full workload
# --------------------------------------------------------------------------------------------------
def cpu_reference(nb, steps, warmup, threads_list, best_of=False):
    """Time y <- A*x for the cfg2 Laplacian with `nb` blocks through the oracle bound to OpenBLAS (the [upstream]
    block loop + one dgbmv per block), once per thread count.  Returns (per-thread-count results, best)."""
    from oracle import cpu as O  # the one place bench.py may execute the oracle
    import bbm_b200 as B

    A = B.BandedBlockBandedMatrix.finitedifference_2d(BLOCK, nblocks=nb)
    rows = [BLOCK] * nb
    x = np.random.default_rng(1001).standard_normal(BLOCK * nb)
    y = np.zeros(BLOCK * nb)
    nbytes = algorithmic_bytes(9, BLOCK * nb, BLOCK * nb)
    blas = O.use_openblas(threads=1)
    runs = []
    for th in (threads_list if blas else [1]):
        if blas:
            O.use_openblas(threads=th)
        for _ in range(max(1, warmup)):
            O.bbb_mul(rows, rows, 1, 1, 1, 1, 1.0, A.data, x, 0.0, y)
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            O.bbb_mul(rows, rows, 1, 1, 1, 1, 1.0, A.data, x, 0.0, y)
            ts.append(time.perf_counter() - t0)
        t = min(ts) if best_of else float(np.mean(ts))
        runs.append({"threads": int(th), "seconds_per_step": t, "value": nbytes / t / 1e9, "unit": UNIT})
    best = max(runs, key=lambda r: r["value"])
    what = "OpenBLAS 0.3.30 cblas_dgbmv per block" if blas else "builtin reference-order loops"
    return runs, best, what, nbytes


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    nb = args.blocks
    runs, best, what, nbytes = cpu_reference(nb, args.steps, args.warmup, [1, ncores] if ncores > 1 else [1])
    val, t = best["value"], best["seconds_per_step"]
    sample = (f"the full workload: {nb} block rows/cols of size {BLOCK}, {3 * nb - 2} dgbmv calls and {nbytes} algorithmic "
              f"bytes per step; mean of {args.steps} steps after {args.warmup} warm-up")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(nb, max(1, args.gpus)),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": best["threads"], "kind": "port", "sample": sample, "blas": what,
                         "by_threads": runs, "host_cores": ncores},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def bind_near_gpu(torch, local_rank):
    """N > 1: keep this rank's host threads -- and, through first-touch page placement, its pinned staging buffers -- on
    the CPUs NVML reports as nearest to the rank's GPU (what `numactl --cpunodebind --membind` per rank does on a
    multi-socket box: without it the pinned buffers of all ranks may sit on one socket and every copy of the other
    socket's GPUs crosses the inter-socket link).  Best effort: returns the CPU list, or None if anything is missing."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        hdl = pynvml.nvmlDeviceGetHandleByPciBusId(bus)
        ncpu = os.cpu_count() or 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(hdl, (ncpu + 63) // 64)
        near = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus = sorted(near & os.sched_getaffinity(0))
        if not cpus or len(cpus) == len(os.sched_getaffinity(0)):
            return None          # one NUMA node (or a cpuset that hides the topology): nothing to do
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


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
    args.near_cpus = bind_near_gpu(torch, local_rank) if world > 1 else None
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

    if args.workload == "cfg4":
        # secondary workload (not the driver's default): the sharded block-banded x block-banded product of cfg4
        from benchmarks import dist_cfg4
        result = dist_cfg4.run(args, torch, dist, B, h, stream, rank, world, local_rank)
        if rank == 0:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            print(json.dumps(result), flush=True)
            os.dup2(2, 1)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
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


def extra_configs(torch, B, h, stream):
    """BASELINE.json's other configs at full size on this GPU (the same code as benchmarks/configs.py): time, achieved
    rate, fraction of the bounding roofline, and the error against the CPU oracle at full size -- cfg3, cfg5 and
    cfg4 on all 8 block-column chunks of C."""
    from benchmarks import configs as CF
    out = {}
    hbm = CF.peaks()
    t0 = time.perf_counter()
    try:
        peak_tf = CF.dgemm_peak(torch)
        out["fp64_gemm_peak"] = {"TFLOPs": peak_tf, "how": "torch.matmul float64 8192^3 (cuBLAS DGEMM), best of 5, this run"}
        r3 = list(range(1, 4001))
        for dt, key in ((np.float64, "cfg3_f64"), (np.float32, "cfg3_f32")):
            r = CF.run_bbb(torch, B, h, stream, "cfg3", r3, r3, (2, 2), (2, 2), dt, 20, True)
            out[key] = {"workload": "BandedBlockBandedMatrix rows=cols=1:4000, (2,2),(2,2) matvec", "ms": r["ms_avg"], "ms_min": r["ms_min"],
                        "achieved": r["GBps"], "unit": "GB/s", "frac": r["GBps"] / hbm, "peak": hbm, "peak_source": "MEASURED_PEAKS.json hbm_gbs",
                        "rel_err_vs_oracle": r["rel_err_vs_oracle_full_size"], "tolerance": r["tolerance"], "cpu_seconds": r["cpu_seconds_full_size"]}
        torch.cuda.empty_cache()
        r = CF.run_cfg4(torch, B, h, stream, 3, True, peak_tf=peak_tf, oracle_chunks=None)
        out["cfg4"] = {"workload": "BlockBandedMatrix fill(256,2048), (2,2) x (2,2) -> (4,4), Float64 mul!", "ms": r["ms_avg"], "ms_min": r["ms_min"],
                       "achieved": r["TFLOPs"], "unit": "TFLOP/s", "frac": r["TFLOPs"] / peak_tf, "peak": peak_tf,
                       "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (fp64_gemm_peak)",
                       "rel_err_vs_oracle": r["rel_err_vs_oracle"], "oracle_checked": r["oracle_checked"], "tolerance": r["tolerance"],
                       "cpu_seconds_full_size_extrapolated": r["cpu_seconds_full_size_extrapolated"]}
        torch.cuda.empty_cache()
        r = CF.run_cfg5(torch, B, h, stream, 3, True, peak_tf=peak_tf)
        out["cfg5"] = {"workload": "UpperTriangular BlockBandedMatrix fill(512,1024), (0,3), Float64 ldiv!, 64 right-hand sides", "ms": r["ms_avg"],
                       "ms_min": r["ms_min"], "us_per_block_step": r["us_per_block_step"], "achieved": r["TFLOPs"], "unit": "TFLOP/s",
                       "frac": r["TFLOPs"] / peak_tf, "peak": peak_tf, "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (fp64_gemm_peak)",
                       "hbm_frac": r["roofline"]["hbm_frac"], "bound": "latency chain of 1024 dependent block steps; FP64 tensor flops and one pass over the matrix are lower bounds",
                       "rel_err_vs_oracle": r["rel_err_vs_oracle"], "oracle_checked": r["oracle_checked"], "tolerance": r["tolerance"],
                       "cpu_seconds": r["cpu_seconds_64_columns"]}
    except Exception as exc:  # the headline line must survive a failure of the secondary configs
        out["error"] = f"{type(exc).__name__}: {exc}"
    out["seconds"] = time.perf_counter() - t0
    return out


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

    with torch.cuda.stream(stream):
        for _ in range(W):
            B.mul_(dy, Ad, dx)
        torch.cuda.synchronize()
        sampler = ClockSampler(h.device).start()
        time.sleep(0.15)
        # the timed region: K back-to-back products between two events on the launch stream (an event between the steps
        # would break the kernel-to-kernel programmatic dependent launch the products chain with)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = h.launch_count
        t0 = time.perf_counter()
        ev0.record(stream)
        for i in range(K):
            B.mul_(dy, Ad, dx)
        ev1.record(stream)
        t_enq = time.perf_counter()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        launches = h.launch_count - l0
        clocks = sampler.stop(t0, t1)
        enqueue_us = (t_enq - t0) / K * 1e6
        total_ms = ev0.elapsed_time(ev1)
        # a second, separate pass with an event behind every step: min / median of a single launch (not the headline)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        ev[0].record(stream)
        for i in range(K):
            B.mul_(dy, Ad, dx)
            ev[i + 1].record(stream)
        torch.cuda.synchronize()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    ms_per_step = total_ms / K
    value = nbytes / (ms_per_step * 1e-3) / 1e9
    kernel_ms = ms_per_step          # the step is one launch of the band-walk kernel: its average duration over the timed region

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

    extra = None
    if not args.no_extra and nb == BLOCK:
        del Ad, dx, dy
        extra = extra_configs(torch, B, h, stream)

    cpu = None
    if not args.no_cpu:
        ncores = os.cpu_count() or 1
        runs, best, what, _ = cpu_reference(nb, 3, 1, [1, ncores] if ncores > 1 else [1], best_of=True)
        cpu = {"value": best["value"], "unit": UNIT, "cores": best["threads"], "kind": "port", "blas": what, "by_threads": runs,
               "host_cores": ncores,
               "sample": f"the full workload ({nb} block rows/cols of size {BLOCK}, {3 * nb - 2} dgbmv calls per pass), best of 3 passes per thread count"}

    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(nb, 1),
        "clocks": clocks,
        "e2e": {"value": nbytes / e2e_s / 1e9, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "steps": Ke,
                "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n, "call": "bbm_bbb_mul_host_f64"},
        "gpu_launches": int(launches),
        "host_enqueue_us_per_step": enqueue_us,
        "roofline": {"bound": "hbm", "kernel": "bbb_walk_kernel<double,3,3>", "achieved": nbytes / (kernel_ms * 1e-3) / 1e9,
                     "peak": None, "unit": UNIT, "frac": None,
                     "traffic": ncu_traffic("bbb_walk_kernel<double,3,3>") if nb == BLOCK else None,
                     "traffic_source": "profiles/ncu_traffic.json (one committed `ncu --set full` capture of this kernel on this workload; not measured in this run)",
                     "kernel_ms_avg": kernel_ms, "kernel_ms_min_single": float(min(per)), "kernel_ms_median_single": float(np.median(per)),
                     "kernel_ms_note": "avg = timed region / K launches; min/median from a separate pass with one event per launch (4 us event granularity, no dependent-launch overlap)",
                     "algorithmic_bytes_per_launch": nbytes},
        "cpu_baseline": cpu,
        "extra": extra,
    }


# --------------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1): strong scaling of the cfg2 Laplacian matvec
# --------------------------------------------------------------------------------------------------
def random_slab(lay, block, P=9, seed=1002):
    """Random band storage of this rank's column slab, a function of the GLOBAL block column alone (every rank that
    holds a block column generates the same numbers)."""
    Jlo = lay["col_offset"] // block
    nJ = lay["n_local"] // block
    slab = np.empty((P, lay["n_local"]), order="F")
    for j in range(nJ):
        slab[:, j * block:(j + 1) * block] = np.random.default_rng([seed, Jlo + j]).standard_normal((P, block))
    return slab


def bench_distributed(args, torch, dist, B, h, stream, nb, block, rank, world, local_rank):
    from bbm_b200 import _lib as L
    from bbm_b200.distributed import Communicator, DistBandedBlockBandedMatrix
    from oracle import cpu as O   # the checker of the distributed result (never the thing measured)
    K, W = args.steps, args.warmup
    n = block * nb
    rows = [block] * nb
    comm = Communicator(h, dist, world, rank)
    A = DistBandedBlockBandedMatrix(comm, rows, rows, (1, 1), (1, 1))
    lay = A.layout
    dev = f"cuda:{local_rank}"

    def allmax(v):
        t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # the same global x on every rank (seeded), each keeps its owned part; halos start as NaN
    xg = np.random.default_rng(1001).standard_normal(n)
    xl = np.full(lay["n_local"], np.nan)
    o = lay["own_offset"]
    xl[o:o + lay["n_own"]] = xg[lay["col_offset"] + o: lay["col_offset"] + o + lay["n_own"]]
    dx = h.to_device(xl)
    dy = h.empty(lay["m_own"], np.float64)
    nbytes = algorithmic_bytes(9, n, n)
    # fused halo push over NVLink peer memory (one kernel per step); falls back to the NCCL exchange
    no_peer = any(o_.replace(" ", "") == "dist.peer=0" for o_ in args.opt)
    try:
        fused = (not no_peer) and A.enable_peer()   # collective: every rank takes the same branch
    except Exception as exc:  # a rank-local setup failure is reported by the library as "unsupported" on all ranks
        print(f"[rank {rank}] fused halo push unavailable: {exc}", file=sys.stderr)
        fused = False

    # ---- parity BEFORE anything is timed: random coefficients (a constant-coefficient stencil hides index bugs),
    # each rank's owned rows against the CPU oracle on the same sub-view (block rows [K0,K1) over the local column
    # slab are banded-block-banded with bandwidths (l-s, u+s), src/BandedBlockBandedMatrix.jl:450-456); two products
    # so that both parities of the halo buffers are exercised
    slab = random_slab(lay, block)
    A.set_local_data(slab)
    K0, K1 = lay["K0"], lay["K1"]
    Jlo, Jhi = lay["col_offset"] // block, (lay["col_offset"] + lay["n_local"]) // block
    sft = K0 - Jlo
    xloc = xg[lay["col_offset"]: lay["col_offset"] + lay["n_local"]].copy()
    ref = np.zeros(lay["m_own"])
    O.bbb_mul(rows[K0:K1], rows[Jlo:Jhi], 1 - sft, 1 + sft, 1, 1, 1.0, slab, xloc, 0.0, ref)
    err_oracle = 0.0
    with torch.cuda.stream(stream):
        for _ in range(2):
            dy.fill_bytes(0xFF)
            A.mul_(dy, dx)
            y = dy.to_host()
            e = float(np.linalg.norm(y - ref) / max(np.linalg.norm(ref), 1e-300)) if np.isfinite(y).all() else float("inf")
            err_oracle = max(err_oracle, e)
    err_oracle = allmax(err_oracle)
    if not (err_oracle <= 64 * 2.22e-16 * 9):
        raise SystemExit(f"bench: distributed matvec differs from the oracle on random coefficients (rel err {err_oracle:.3e})")
    del slab, ref

    # ---- the workload of the metric: the Laplacian
    col = (np.array([0, 1, 0, 1, -4, 1, 0, 1, 0], dtype=np.float64) * float(block) ** 2)
    slab = np.empty((9, lay["n_local"]), order="F")
    slab[:, :] = col[:, None]
    A.set_local_data(slab)
    del slab

    LEAD = 8   # untimed lead-in products queued in front of ev0 (see below)
    with torch.cuda.stream(stream):
        for _ in range(W):
            A.mul_(dy, dx)
        torch.cuda.synchronize()
        sampler = ClockSampler(h.device).start() if rank == 0 else None
        time.sleep(0.15)
        dist.barrier()
        torch.cuda.synchronize()
        # The timed region is bracketed by barrier + synchronize on both sides.  Inside it, LEAD untimed products are
        # queued ahead of ev0: the ranks leave the host barrier tens of microseconds apart, and a first timed step
        # would wait for the slowest neighbour's halo -- a start skew of the HOST that 20 steps of ~35 us cannot
        # amortise.  Behind the lead-in the GPUs are in lock step through the halo handshake itself and ev0..ev1
        # measures K steady-state steps, which is what the metric names.
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        for _ in range(LEAD):
            A.mul_(dy, dx)
        l0 = h.launch_count               # gpu_launches counts the K timed products only
        ev0.record(stream)
        for _ in range(K):
            A.mul_(dy, dx)
        ev1.record(stream)
        t_enq = time.perf_counter()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        launches = h.launch_count - l0
        dist.barrier()
        clocks = sampler.stop(t0, t1) if sampler else None
        ms_per_step = allmax(ev0.elapsed_time(ev1)) / K
        enqueue_us = allmax((t_enq - t0) / (K + LEAD) * 1e6)
        # a second, separate pass with an event behind every step: min / median step time (not the headline)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        for _ in range(LEAD):
            A.mul_(dy, dx)
        evs[0].record(stream)
        for i in range(K):
            A.mul_(dy, dx)
            evs[i + 1].record(stream)
        torch.cuda.synchronize()
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(K)]
        step_min, step_med = allmax(min(per)), allmax(float(np.median(per)))

    # end to end: every step uploads the owned x from pinned host memory and reads y_own back
    hx = h.pinned(lay["n_own"], np.float64)
    hy = h.pinned(lay["m_own"], np.float64)
    hx.array[:] = xl[o:o + lay["n_own"]]
    import ctypes as C_
    lib = h._lib
    xown_ptr = dx.ptr + 8 * o
    host_call = hasattr(A, "mul_host_") and not args.e2e_serial

    def e2e_step():
        if host_call:
            A.mul_host_(hy.array, hx.array)
            return
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
    e2e_s = allmax((time.perf_counter() - t0) / Ke)

    # the Laplacian result itself: owned rows against the 5-point stencil in numpy (device path and host-vector path)
    X = xg.reshape(nb, block)
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
    errs = [float(np.linalg.norm(v - Y.reshape(-1)) / max(np.linalg.norm(Y), 1e-300)) for v in (dy.to_host(), hy.array)]
    err = allmax(max(errs))
    if not (err <= 64 * 2.22e-16 * 9):
        raise SystemExit(f"bench: distributed matvec is wrong (rel err {err:.3e})")
    lt = torch.tensor([launches], device=dev, dtype=torch.float64)
    dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    active, timed_out = A.peer_status()
    if timed_out:
        raise SystemExit("bench: a peer-memory halo wait timed out")

    value = nbytes / (ms_per_step * 1e-3) / 1e9
    cfg = workload_config(nb, world)
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "exchange": ("x halo pushed into the neighbours' halo buffers over NVLink peer memory by CTA 0 of the same band-walk kernel "
                     "(1 launch per rank per step)" if active else "x halo by NCCL send/recv overlapped with the interior block rows"),
        "max_rel_err_vs_oracle": err_oracle, "max_rel_err_vs_stencil": err,
        "oracle_check": "random-coefficient matrix (seeded per global block column), every rank's owned rows vs oracle bbb_mul on the same sub-view, 2 products, before the timed region",
        "timing": {"lead_in_steps": LEAD, "step_ms_min": step_min, "step_ms_median": step_med, "host_enqueue_us_per_step": enqueue_us,
                   "note": "barrier+sync | LEAD untimed products, ev0, K products, ev1 | sync+barrier; max over ranks of ev0..ev1; min/median from a separate pass with per-step events"},
        "clocks": clocks,
        "e2e": {"value": nbytes / e2e_s / 1e9, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "steps": Ke,
                "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n,
                "call": ("bbm_dist_bbb_mul_host_f64 (pinned x_own up, halo exchange, product and y_own down pipelined over block-row chunks), per rank"
                         if host_call else "bbm_memcpy_h2d(x_own) + bbm_dist_bbb_mul_f64 + bbm_memcpy_d2h(y_own), per rank"),
                "host_affinity": (f"rank 0 bound to the {len(args.near_cpus)} CPUs NVML reports nearest to its GPU (pinned buffers first-touched there); every rank does the same"
                                  if getattr(args, "near_cpus", None) else "none (one NUMA node, or the topology is not visible)")},
        "gpu_launches": int(lt.item()),
        "roofline": {"bound": "hbm", "kernel": "bbb_walk_kernel<double,3,3> (per rank, incl. halo push / exchange)",
                     "achieved": value, "peak": None, "unit": UNIT, "frac": None, "traffic": None,
                     "note": "aggregate over ranks against N x one GPU's measured copy bandwidth"},
        "cpu_baseline": None,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--blocks", type=int, default=BLOCK, help="number of blocks (default 4096 = cfg2; smaller only for debugging)")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4"], help="cfg2 (default, BASELINE.json's metric) or the sharded cfg4 product")
    ap.add_argument("--cfg4-blocks", type=int, default=2048, help="--workload cfg4: number of 256-row blocks (2048 = cfg4)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-extra", action="store_true", help="skip the full-size cfg3/cfg4/cfg5 lines of `extra` (profiling runs)")
    ap.add_argument("--e2e-serial", action="store_true", help="N > 1: time the serial copy / product / copy end-to-end step instead of the pipelined host call")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (experiments), e.g. dist.graph=0")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
