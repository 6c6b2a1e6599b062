"""Times the backbone attention kernels (tcgen05, 129 tokens) forward and backward at B=128 (S = 384 sequences)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from editor_b200 import lib

S, H = 384, 12
qkv = (torch.randn(S * 129, 2304, device="cuda") * 1.0).to(torch.bfloat16)
out = torch.empty(S * 129, 768, dtype=torch.bfloat16, device="cuda")
P = torch.zeros(S * H, 129, 136, dtype=torch.bfloat16, device="cuda")
d_out = torch.randn(S * 129, 768, device="cuda").to(torch.bfloat16)
d_qkv = torch.empty_like(qkv)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


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


fwd = t(lambda: lib.attention(qkv, out, P, S, H, 129, 0.125, fixed_len=129, p_rows=129, ldp=136))
bwd = t(lambda: lib.attention(qkv, out, P, S, H, 129, 0.125, fixed_len=129, p_rows=129, ldp=136, d_out=d_out, d_qkv=d_qkv, backward=True))
fb = (228 + 161 + 76) * 1.0
bb = (228 + 161 + 76 + 228) * 1.0
print("attn_tc fwd %.3f ms (%.0f GB/s algorithmic)   bwd %.3f ms (%.0f GB/s)" % (fwd, fb / fwd, bwd, bb / bwd))
