#!/usr/bin/env python
"""Fixed per-step cost of the band-walk matvec at the size one rank of the 8-GPU cfg2 run holds (512 block rows of
4096): K back-to-back products between two events (no event in between: programmatic dependent launch only chains
kernel to kernel), for a list of option sets.  One JSON line per variant.

  python benchmarks/step_floor.py [--blocks 512] [--steps 200] [--variants "bbb.stages=3;bbb.stages=3,bbb.pdl=2"]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT = ["", "bbb.stages=3", "bbb.pdl=1", "bbb.pdl=2", "bbb.stages=3,bbb.ctas_per_sm=2", "bbb.stages=3,bbb.ctas_per_sm=2,bbb.pdl=1",
           "bbb.stages=3,bbb.ctas_per_sm=2,bbb.pdl=2", "bbb.stages=3,bbb.pdl=1", "bbb.stages=3,bbb.pdl=2", "bbb.stages=2", "bbb.stages=2,bbb.ctas_per_sm=4,bbb.pdl=2",
           "bbb.stages=2,bbb.ctas_per_sm=3,bbb.pdl=2", "bbb.tile=128,bbb.stages=3", "bbb.tile=128,bbb.stages=3,bbb.pdl=2", "bbb.tile=128,bbb.stages=4,bbb.ctas_per_sm=4,bbb.pdl=2"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=512)
    ap.add_argument("--block-size", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--short", type=int, default=20)
    ap.add_argument("--variants", default=None)
    args = ap.parse_args()
    import torch
    import bbm_b200 as B
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    h = B.Handle(0, stream=stream.cuda_stream)
    bs, nb = args.block_size, args.blocks
    n = bs * nb
    A = B.BandedBlockBandedMatrix.finitedifference_2d(bs, nblocks=nb).to_device(h)
    x = np.random.default_rng(1).standard_normal(n)
    dx, dy = h.to_device(x), h.empty(n, np.float64)
    nbytes = 8 * (9 * n + 2 * n)
    variants = DEFAULT if args.variants is None else args.variants.split(";")
    keys = set()
    for v in variants:
        for kv in filter(None, v.split(",")):
            keys.add(kv.split("=")[0])
    ref = None
    for v in variants:
        for k in keys:
            h.set_option(k, 0)
        if "bbb.stages" in keys:
            h.set_option("bbb.stages", 4)
        for kv in filter(None, v.split(",")):
            k, val = kv.split("=")
            h.set_option(k, int(val))
        out = {"variant": v or "default", "blocks": nb}
        with torch.cuda.stream(stream):
            for K in (args.short, args.steps):
                for _ in range(5):
                    B.mul_(dy, A, dx)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record(stream)
                for _ in range(K):
                    B.mul_(dy, A, dx)
                e1.record(stream)
                t1 = time.perf_counter()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / K
                out[f"us_per_step_{K}"] = us
                out[f"enqueue_us_{K}"] = (t1 - t0) / K * 1e6
            out["GBps"] = nbytes / (out[f"us_per_step_{args.steps}"] * 1e-6) / 1e9
        y = dy.to_host()
        if ref is None:
            ref = y
        out["same_result_as_default"] = bool(np.array_equal(y, ref))
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
