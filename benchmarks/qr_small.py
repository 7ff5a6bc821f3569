"""Micro-benchmark of the device qr! (bbm_sky_qr_f64): 64 blocks of 256, (1,1) -> 512 sub-panel steps; prints the
mean time per step (panel kernel + two trailing-update launches).  Used with ncu to look at the panel kernel."""
import sys, json, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch, bbm_b200 as B
from oracle import formats as F
torch.cuda.set_device(0); stream = torch.cuda.Stream(); h = B.Handle(0, stream=stream.cuda_stream)
nblk, bs, l, u = 64, 256, 1, 1
rows = [bs] * nblk
rng = np.random.default_rng(1)
A = B.BlockSkylineMatrix(np.zeros(F.sky_numentries(rows, rows, l, u)), rows, rows, l, u)
A.data[:] = rng.standard_normal(A.data.shape) / 30
Aw = A.with_bandwidths(l, l + u)
for J in range(nblk):
    s = Aw.block_start(J, J); ld = int(Aw.block_strides[J]); i = np.arange(bs); Aw.data[s + i + i * ld] += 2
Ad = Aw.to_device(h)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    B.qr_(Ad); h.synchronize()
    Ad2 = Aw.to_device(h)
    e0.record(); B.qr_(Ad2); e1.record(); h.synchronize()
print(json.dumps({"nblk": nblk, "ms": e0.elapsed_time(e1), "us_per_step": e0.elapsed_time(e1) * 1e3 / (nblk * 8)}))
