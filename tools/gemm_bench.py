"""Times every GEMM shape / epilogue of one backbone layer at B=128 (M = 3*128*129 rows) in isolation.
Usage (GPU box): python tools/gemm_bench.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from editor_b200 import lib

M = 3 * 128 * 129
dev = "cuda"
bf = torch.bfloat16


def t(fn, n=10):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / n


def main():
    x768 = torch.randn(M, 768, device=dev).to(bf)
    x3072 = torch.randn(M, 3072, device=dev).to(bf)
    x2304 = torch.randn(M, 2304, device=dev).to(bf)
    w_qkv = torch.randn(2304, 768, device=dev).to(bf)
    w_proj = torch.randn(768, 768, device=dev).to(bf)
    w_fc1 = torch.randn(3072, 768, device=dev).to(bf)
    w_fc2 = torch.randn(768, 3072, device=dev).to(bf)
    b768, b2304, b3072 = (torch.randn(n, device=dev) for n in (768, 2304, 3072))
    res = torch.randn(M, 768, device=dev)
    res2 = torch.empty(M, 768, device=dev)
    o768, o2304, o3072, o3072b = (torch.empty(M, n, device=dev, dtype=bf) for n in (768, 2304, 3072, 3072))
    g768 = torch.zeros(768, 768, device=dev); g2304 = torch.zeros(2304, 768, device=dev)
    g3072 = torch.zeros(3072, 768, device=dev); g768x = torch.zeros(768, 3072, device=dev)
    rows = []
    def add(name, flops, fn):
        res = []
        for mode in (1, 0):          # 1 = single-CTA 128x256 tiles, 0 = CTA pairs (cta_group::2, 256x256)
            lib.gemm_set_mode(mode)
            ms = t(fn)
            res += [ms, flops / ms / 1e9]
        lib.gemm_set_mode(0)
        rows.append((name, *res))
    add("qkv fwd  [M,768]x[2304,768] bias bf16", 2*M*2304*768, lambda: lib.gemm(x768, w_qkv, o2304, M, 2304, 768, bias=b2304))
    add("proj fwd [M,768]x[768,768] residual f32", 2*M*768*768, lambda: lib.gemm(x768, w_proj, res2, M, 768, 768, epilogue=lib.EPI_RESIDUAL, bias=b768, aux=res))
    add("fc1 fwd  [M,768]x[3072,768] gelu + gelu'", 2*M*3072*768, lambda: lib.gemm(x768, w_fc1, o3072, M, 3072, 768, epilogue=lib.EPI_GELU, bias=b3072, out2=o3072b))
    add("fc2 fwd  [M,3072]x[768,3072] residual", 2*M*3072*768, lambda: lib.gemm(x3072, w_fc2, res2, M, 768, 3072, epilogue=lib.EPI_RESIDUAL, bias=b768, aux=res))
    add("fc2 dgrad * saved gelu' -> [M,3072]", 2*M*3072*768, lambda: lib.gemm(x768, w_fc2, o3072, M, 3072, 768, b_mn=True, epilogue=lib.EPI_GELU_BWD, aux=x3072))
    add("fc1 dgrad -> [M,768]", 2*M*3072*768, lambda: lib.gemm(x3072, w_fc1, o768, M, 768, 3072, b_mn=True))
    add("proj dgrad -> [M,768]", 2*M*768*768, lambda: lib.gemm(x768, w_proj, o768, M, 768, 768, b_mn=True))
    add("qkv dgrad -> [M,768]", 2*M*2304*768, lambda: lib.gemm(x2304, w_qkv, o768, M, 768, 2304, b_mn=True))
    add("fc2 wgrad [768,3072] splitK2", 2*M*3072*768, lambda: lib.gemm(x768, x3072, g768x, 768, 3072, M, a_mn=True, b_mn=True, epilogue=lib.EPI_ATOMIC, split_k=2))
    add("fc1 wgrad [3072,768] splitK2", 2*M*3072*768, lambda: lib.gemm(x3072, x768, g3072, 3072, 768, M, a_mn=True, b_mn=True, epilogue=lib.EPI_ATOMIC, split_k=2))
    add("proj wgrad [768,768] splitK8", 2*M*768*768, lambda: lib.gemm(x768, x768, g768, 768, 768, M, a_mn=True, b_mn=True, epilogue=lib.EPI_ATOMIC, split_k=8))
    add("qkv wgrad [2304,768] splitK8", 2*M*2304*768, lambda: lib.gemm(x2304, x768, g2304, 2304, 768, M, a_mn=True, b_mn=True, epilogue=lib.EPI_ATOMIC, split_k=8))
    tot1 = tot2 = 0
    print("%-44s %22s   %22s" % ("", "single CTA 128x256", "CTA pair 256x256"))
    for name, ms1, tf1, ms2, tf2 in rows:
        print("%-44s %7.3f ms %7.1f TF/s   %7.3f ms %7.1f TF/s" % (name, ms1, tf1, ms2, tf2))
        tot1 += ms1
        tot2 += ms2
    print("sum per layer: %.3f ms (single)  %.3f ms (pair)" % (tot1, tot2))


if __name__ == "__main__":
    main()
