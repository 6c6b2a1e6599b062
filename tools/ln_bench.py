"""Times the LayerNorm kernels at the backbone's size (R = 3*128*129 rows x 768) with a cold L2: achieved HBM GB/s."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from editor_b200 import lib

R, D = 3 * 128 * 129, 768
dev = "cuda"
x = torch.randn(R, D, device=dev)
y = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
mean, rstd = torch.empty(R, device=dev), torch.empty(R, device=dev)
gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
dy = torch.randn(R, D, device=dev).to(torch.bfloat16)
g = torch.randn(R, D, device=dev)
gb = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
dg, db, dc = (torch.zeros(D, device=dev) for _ in range(3))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
peak = 6546.9
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except OSError:
    pass


def t(fn, n=10):
    fn(); fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / n


fwd = t(lambda: lib.layernorm_fwd(x, gamma, beta, 1e-6, y, mean, rstd, R))
bwd = t(lambda: lib.layernorm_bwd(dy, x, mean, rstd, gamma, g, g, gb, dg, db, dc, R))
bf, bb = R * D * (4 + 2) / 1e6, R * D * (2 + 4 + 4 + 4 + 2) / 1e6
print("ln_fwd %.4f ms  %.0f GB/s (%.2f of %.0f)   ln_bwd %.4f ms  %.0f GB/s (%.2f)" % (fwd, bf / fwd, bf / fwd / peak, peak, bwd, bb / bwd, bb / bwd / peak))
