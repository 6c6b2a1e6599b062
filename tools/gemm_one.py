"""Launches ONE GEMM shape of the backbone layer a few times (for `ncu --set full -k regex:gemm_bf16 -s 2 -c 1`).
Usage: python tools/gemm_one.py {qkv|proj|fc1|fc2|fc2d|fc1d|projd|qkvd|fc2w|fc1w|projw|qkvw} [mode]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from editor_b200 import lib

M = 3 * 128 * 129
dev, bf = "cuda", torch.bfloat16


def main():
    which = sys.argv[1]
    lib.gemm_set_mode(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    r = lambda *s: torch.randn(*s, device=dev).to(bf)
    x768, x3072, x2304 = r(M, 768), r(M, 3072), r(M, 2304)
    w_qkv, w_proj, w_fc1, w_fc2 = r(2304, 768), r(768, 768), r(3072, 768), r(768, 3072)
    b768, b2304, b3072 = (torch.randn(n, device=dev) for n in (768, 2304, 3072))
    res, res2 = torch.randn(M, 768, device=dev), torch.empty(M, 768, device=dev)
    o768, o2304, o3072, o3072b = (torch.empty(M, n, device=dev, dtype=bf) for n in (768, 2304, 3072, 3072))
    g = {n: torch.zeros(*s, device=dev) for n, s in (("proj", (768, 768)), ("qkv", (2304, 768)), ("fc1", (3072, 768)),
                                                     ("fc2", (768, 3072)))}
    A = lib.EPI_ATOMIC
    fns = {
        "qkv": lambda: lib.gemm(x768, w_qkv, o2304, M, 2304, 768, bias=b2304),
        "proj": lambda: lib.gemm(x768, w_proj, res2, M, 768, 768, epilogue=lib.EPI_RESIDUAL, bias=b768, aux=res),
        "fc1": lambda: lib.gemm(x768, w_fc1, o3072, M, 3072, 768, epilogue=lib.EPI_GELU, bias=b3072, out2=o3072b),
        "fc2": lambda: lib.gemm(x3072, w_fc2, res2, M, 768, 3072, epilogue=lib.EPI_RESIDUAL, bias=b768, aux=res),
        "fc2d": lambda: lib.gemm(x768, w_fc2, o3072, M, 3072, 768, b_mn=True, epilogue=lib.EPI_GELU_BWD, aux=x3072),
        "fc1d": lambda: lib.gemm(x3072, w_fc1, o768, M, 768, 3072, b_mn=True),
        "projd": lambda: lib.gemm(x768, w_proj, o768, M, 768, 768, b_mn=True),
        "qkvd": lambda: lib.gemm(x2304, w_qkv, o768, M, 768, 2304, b_mn=True),
        "fc2w": lambda: lib.gemm(x768, x3072, g["fc2"], 768, 3072, M, a_mn=True, b_mn=True, epilogue=A, split_k=2),
        "fc1w": lambda: lib.gemm(x3072, x768, g["fc1"], 3072, 768, M, a_mn=True, b_mn=True, epilogue=A, split_k=2),
        "projw": lambda: lib.gemm(x768, x768, g["proj"], 768, 768, M, a_mn=True, b_mn=True, epilogue=A, split_k=8),
        "qkvw": lambda: lib.gemm(x2304, x768, g["qkv"], 2304, 768, M, a_mn=True, b_mn=True, epilogue=A, split_k=8),
    }
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(4):
        flush.zero_()
        fns[which]()
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
