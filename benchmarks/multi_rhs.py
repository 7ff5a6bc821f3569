import sys, json, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch, bbm_b200 as B
from benchmarks.configs import wrap, time_gpu
torch.cuda.set_device(0); stream = torch.cuda.Stream(); h = B.Handle(0, stream=stream.cuda_stream)
n = 4096; nb = 1024
A = B.BandedBlockBandedMatrix.finitedifference_2d(n, nblocks=nb).to_device(h)
m = n * nb
for q in (1, 2, 4, 8):
    x = torch.randn(q, m, device='cuda', dtype=torch.float64); y = torch.empty(q, m, device='cuda', dtype=torch.float64)
    dx, dy = wrap(B, h, x if q > 1 else x[0]), wrap(B, h, y if q > 1 else y[0])
    avg, best = time_gpu(torch, stream, lambda: B.mul_(dy, A, dx), 20)
    # check against column-by-column
    ok = True
    if q > 1:
        y1 = torch.empty(m, device='cuda', dtype=torch.float64)
        for c in range(q):
            B.mul_(wrap(B, h, y1), A, wrap(B, h, x[c].contiguous())); h.synchronize()
            ok = ok and bool(torch.equal(y1, y[c]))
    print(json.dumps({"nrhs": q, "ms": avg, "GBps_alg": 8 * (9 * m + 2 * q * m) / avg / 1e6, "bit_equal_to_single": ok}))
