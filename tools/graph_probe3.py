"""Third bisect: capture single C-ABI calls of the backward (no autograd involved) and report which one invalidates."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OPS = ["ln_fwd", "ln_bwd", "gemm_store", "gemm_gelu_bwd", "gemm_atomic_sk", "gemm_small_atomic", "attn_tc_bwd", "colsum",
       "embed_bwd", "sgd", "bn_bwd", "triplet", "engine_bb_bwd", "engine_hma_bwd"]


def op(name):
    import torch
    from editor_b200 import lib
    lib.load()
    bf = torch.bfloat16
    M = 129 * 12
    dev = "cuda"
    x768 = torch.randn(M, 768, device=dev)
    xb768 = x768.to(bf)
    xb3072 = torch.randn(M, 3072, device=dev).to(bf)
    w2 = torch.randn(768, 3072, device=dev).to(bf)
    o3072 = torch.empty(M, 3072, device=dev, dtype=bf)
    gw = torch.zeros(768, 3072, device=dev)
    mean, rstd = torch.zeros(M, device=dev), torch.ones(M, device=dev)
    gamma, beta = torch.ones(768, device=dev), torch.zeros(768, device=dev)
    fn = None
    if name == "ln_fwd":
        y = torch.empty(M, 768, device=dev, dtype=bf)
        fn = lambda: lib.layernorm_fwd(x768, gamma, beta, 1e-6, y, mean, rstd, M)
    elif name == "ln_bwd":
        g = torch.zeros(M, 768, device=dev)
        gb = torch.empty(M, 768, device=dev, dtype=bf)
        dg, db, dc = (torch.zeros(768, device=dev) for _ in range(3))
        fn = lambda: lib.layernorm_bwd(xb768, x768, mean, rstd, gamma, g, g, gb, dg, db, dc, M)
    elif name == "gemm_store":
        fn = lambda: lib.gemm(xb768, w2.t().contiguous(), o3072, M, 3072, 768)
    elif name == "gemm_gelu_bwd":
        cs = torch.zeros(3072, device=dev)
        fn = lambda: lib.gemm(xb768, w2, o3072, M, 3072, 768, b_mn=True, epilogue=lib.EPI_GELU_BWD, aux=xb3072, colsum=cs)
    elif name == "gemm_atomic_sk":
        fn = lambda: lib.gemm(xb768, xb3072, gw, 768, 3072, M, a_mn=True, b_mn=True, epilogue=lib.EPI_ATOMIC, split_k=4)
    elif name == "gemm_small_atomic":
        a = torch.randn(4, 176, device=dev).to(bf)
        b = torch.randn(4, 2304, device=dev).to(bf)
        d = torch.zeros(171, 2304, device=dev)
        fn = lambda: lib.gemm(a, b, d, 171, 2304, 4, a_mn=True, b_mn=True, epilogue=lib.EPI_ATOMIC)
    elif name == "attn_tc_bwd":
        S = 12
        qkv = torch.randn(S * 129, 2304, device=dev).to(bf)
        out = torch.empty(S * 129, 768, device=dev, dtype=bf)
        P = torch.zeros(S * 12, 129, 136, device=dev, dtype=bf)
        lib.attention(qkv, out, P, S, 12, 129, 0.125, fixed_len=129, p_rows=129, ldp=136)
        d_out = torch.randn(S * 129, 768, device=dev).to(bf)
        d_qkv = torch.empty_like(qkv)
        fn = lambda: lib.attention(qkv, out, P, S, 12, 129, 0.125, fixed_len=129, p_rows=129, ldp=136, d_out=d_out, d_qkv=d_qkv, backward=True)
    elif name == "colsum":
        o = torch.zeros(768, device=dev)
        fn = lambda: lib.colsum(xb768, o, M, 768)
    elif name in ("engine_bb_bwd", "engine_hma_bwd", "embed_bwd", "sgd", "bn_bwd", "triplet"):
        import __graft_entry__ as ge
        from editor_b200.train import editor_loss
        model, sd, x, label, cam, _ = ge._small_case(True, 4)
        model = model.cuda().train()
        x = {k: v.cuda() for k, v in x.items()}
        label, cam = label.cuda(), cam.cuda()
        eng = model.engine()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs = model(x, label=label, cam_label=cam, writer=None, epoch=1)
        # reach into the saved state of the two big autograd nodes
        node_hma = outs[1].grad_fn
        import gc
        ctxs = {}
        def walk(fn_, seen):
            if fn_ is None or fn_ in seen:
                return
            seen.add(fn_)
            nm = type(fn_).__name__
            if nm in ("_BackboneFnBackward", "_HMAFnBackward"):
                ctxs[nm] = fn_
            for nf, _ in fn_.next_functions:
                walk(nf, seen)
        walk(outs[1].grad_fn, set())
        walk(outs[2].grad_fn, set())
        bb, hma = ctxs["_BackboneFnBackward"], ctxs["_HMAFnBackward"]
        if name == "engine_bb_bwd":
            d_tok = torch.randn(12, 129, 768, device=dev) * 1e-3
            fn = lambda: eng.backbone_backward(bb.sv, d_tok.clone())
        elif name == "engine_hma_bwd":
            tokens = hma.saved_tensors[0]
            z = torch.randn(3, 4, 768, device=dev) * 1e-3
            one = torch.ones(1, device=dev)
            fn = lambda: eng.hma_backward(tokens, hma.sel, hma.sv, z, z.clone(), z.clone(), one)
        elif name == "embed_bwd":
            g = torch.randn(12 * 129, 768, device=dev)
            dpos, dsie = torch.zeros(129, 768, device=dev), torch.zeros(4, 768, device=dev)
            dpatch = torch.empty(12 * 128, 768, device=dev, dtype=bf)
            fn = lambda: lib.call("edb_embed_assemble_bwd", g.data_ptr(), 12, 4, 128, cam.data_ptr(), 3.0, dpos.data_ptr(),
                                  dsie.data_ptr(), dpatch.data_ptr(), 0, lib.stream_ptr())
        elif name == "sgd":
            a = eng.arena
            mom = torch.zeros_like(a.flat)
            flags = torch.zeros(a.total // 64, dtype=torch.uint8, device=dev)
            fn = lambda: lib.call("edb_sgd_step", a.flat.data_ptr(), a.grad.data_ptr(), mom.data_ptr(), a.flat16.data_ptr(),
                                  flags.data_ptr(), a.total, 0.001, 0.9, 1e-4, 1e-4, 2.0, 1.0, 0, lib.stream_ptr())
        elif name == "bn_bwd":
            def fn():
                l = outs[0].float().sum()
                torch.autograd.grad(l, [outs[1]], retain_graph=True)
        elif name == "triplet":
            feat = torch.randn(4, 2304, device=dev, requires_grad=True)
            def fn():
                l = editor_loss((torch.randn(4, 171, device=dev), feat, torch.zeros((), device=dev)), label)
                l.backward()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    print("OP %s OK" % name)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        op(sys.argv[1])
    else:
        for s in OPS:
            r = subprocess.run([sys.executable, __file__, s], capture_output=True, text=True, timeout=300)
            ok = [l for l in r.stdout.splitlines() if l.startswith("OP ")]
            print(ok[0] if ok else "OP %s FAILED: %s" % (s, " | ".join(r.stderr.strip().splitlines()[-3:])[:500]), flush=True)
