"""Each C-ABI op against a plain torch fp32 computation of the same op (tight tolerances: these pin the backward
formulas that the bf16 end-to-end test can only check to 1e-2)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from oracle import editor_oracle as orc

pytestmark = pytest.mark.gpu


def _g(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize("rows", [1, 129, 1000])
def test_layernorm_fwd_bwd(rows):
    from editor_b200 import lib
    x = (torch.randn(rows, 768, generator=_g(1)) * 2 + 0.3).cuda()
    gam, bet = (1 + 0.1 * torch.randn(768, generator=_g(2))).cuda(), (0.1 * torch.randn(768, generator=_g(3))).cuda()
    y = torch.empty(rows, 768, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    lib.layernorm_fwd(x, gam, bet, 1e-6, y, mean, rstd)
    xr = x.clone().requires_grad_(True)
    gr, br = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (768,), gr, br, 1e-6)
    assert _rel(y, ref.detach()) < 1e-5
    yb = torch.empty(rows, 768, dtype=torch.bfloat16, device="cuda")
    lib.layernorm_fwd(x, gam, bet, 1e-6, yb)
    assert _rel(yb.float(), ref.detach()) < 1e-2
    dy = torch.randn(rows, 768, generator=_g(4)).cuda()
    g_in = torch.randn(rows, 768, generator=_g(5)).cuda()
    scale = (torch.rand((rows + 128) // 129, generator=_g(6)) + 0.5).cuda()
    ref.backward(dy)
    g_out = torch.empty_like(x)
    gb16 = torch.empty(rows, 768, dtype=torch.bfloat16, device="cuda")
    dgam, dbet, dcol = (torch.zeros(768, device="cuda") for _ in range(3))
    lib.layernorm_bwd(dy, x, mean, rstd, gam, g_in, g_out, gb16, dgam, dbet, dcol, row_scale=scale, scale_group=129)
    want = g_in + xr.grad
    assert _rel(g_out, want) < 1e-4
    rs = scale.repeat_interleave(129)[:rows].unsqueeze(1)
    assert _rel(gb16.float(), want * rs) < 1e-2
    assert _rel(dgam, gr.grad) < 1e-4 and _rel(dbet, br.grad) < 1e-4
    assert _rel(dcol, (want * rs).sum(0)) < 1e-4
    # bf16 dy, in-place residual gradient
    dyb = dy.to(torch.bfloat16)
    g2 = g_in.clone()
    lib.layernorm_bwd(dyb, x, mean, rstd, gam, g2, g2, None, dgam, dbet, None)
    xr2 = x.clone().requires_grad_(True)
    F.layer_norm(xr2, (768,), gam, bet, 1e-6).backward(dyb.float())
    assert _rel(g2, g_in + xr2.grad) < 1e-4


def test_colsum_cast_split_gemm_fp32_faithful():
    from editor_b200 import lib
    src = torch.randn(1000, 3072, generator=_g(1)).cuda()
    out = torch.zeros(3072, device="cuda")
    lib.colsum(src.to(torch.bfloat16), out)
    assert _rel(out, src.to(torch.bfloat16).float().sum(0)) < 1e-4
    out.zero_()
    lib.colsum(src, out)
    assert _rel(out, src.sum(0)) < 1e-4
    # 3-piece bf16 split GEMM reproduces an fp32 matmul to ~1e-5 (EDB_PREC_FP32 path): the six cross products carry
    # 24 mantissa bits; what remains is the tensor core's fp32 accumulation over K = 6*768
    M, N, K = 300, 768, 768
    A, W = torch.randn(M, K, generator=_g(2)).cuda(), (0.05 * torch.randn(N, K, generator=_g(3))).cuda()
    As = torch.empty(M, 6 * K, dtype=torch.bfloat16, device="cuda")
    Ws = torch.empty(N, 6 * K, dtype=torch.bfloat16, device="cuda")
    lib.split3(A, As, 0)
    lib.split3(W, Ws, 1)
    D = torch.empty(M, N, device="cuda")
    lib.gemm(As, Ws, D, M, N, 6 * K)
    ref = (A.double() @ W.double().t()).float()
    assert _rel(D, ref) < 2e-5
    assert _rel(D, A.to(torch.bfloat16).float() @ W.to(torch.bfloat16).float().t()) > 1e-3   # plain bf16 is 100x worse


@pytest.mark.parametrize("H,W", [(256, 128), (128, 256)])
def test_patch_embed_matches_conv(H, W):
    from editor_b200 import lib
    B = 3
    imgs = [torch.randn(B, 3, H, W, generator=_g(i)).cuda() for i in range(3)]
    patches = torch.empty(3 * B * 128, 768, device="cuda")
    lib.call("edb_patch_im2col", imgs[0].data_ptr(), imgs[1].data_ptr(), imgs[2].data_ptr(), B, H, W, patches.data_ptr(),
             768, 1, lib.stream_ptr())
    wconv = (0.02 * torch.randn(768, 3, 16, 16, generator=_g(7))).cuda()
    ref = torch.cat([F.conv2d(i, wconv, stride=16).flatten(2).transpose(1, 2) for i in imgs], 0)   # [3B,128,768]
    got = (patches @ wconv.view(768, -1).t()).view(3 * B, 128, 768)
    assert _rel(got, ref) < 1e-4
    cls, pos = torch.randn(768, generator=_g(8)).cuda(), torch.randn(129, 768, generator=_g(9)).cuda()
    sie = torch.randn(4, 768, generator=_g(10)).cuda()
    cam = torch.tensor([0, 3, 1]).cuda()
    x = torch.empty(3 * B, 129, 768, device="cuda")
    lib.call("edb_embed_assemble", got.data_ptr(), cls.data_ptr(), pos.data_ptr(), sie.data_ptr(), cam.data_ptr(), 3.0,
             3 * B, B, 128, x.data_ptr(), lib.stream_ptr())
    want = torch.cat([cls.expand(3 * B, 1, 768), got], 1) + pos + 3.0 * sie[cam.repeat(3)].unsqueeze(1)
    assert _rel(x, want) < 1e-6
    g = torch.randn(3 * B, 129, 768, generator=_g(11)).cuda()
    dpos, dsie = torch.zeros(129, 768, device="cuda"), torch.zeros(4, 768, device="cuda")
    dpatch = torch.empty(3 * B * 128, 768, dtype=torch.bfloat16, device="cuda")
    lib.call("edb_embed_assemble_bwd", g.data_ptr(), 3 * B, B, 128, cam.data_ptr(), 3.0, dpos.data_ptr(), dsie.data_ptr(),
             dpatch.data_ptr(), 0, lib.stream_ptr())
    assert _rel(dpos, g.sum(0)) < 1e-5
    want_sie = torch.zeros(4, 768, device="cuda").index_add_(0, cam.repeat(3), 3.0 * g.sum(1))
    assert _rel(dsie, want_sie) < 1e-5
    assert _rel(dpatch.float().view(3 * B, 128, 768), g[:, 1:]) < 1e-2


def _torch_attention(qkv, lens, H=12):
    outs, maps, off = [], [], 0
    for L in lens:
        blk = qkv[off:off + L].view(L, 3, H, 64).permute(1, 2, 0, 3)
        q, k, v = blk[0], blk[1], blk[2]
        a = ((q @ k.transpose(-2, -1)) * 0.125).softmax(-1)
        outs.append((a @ v).transpose(0, 1).reshape(L, H * 64))
        maps.append(a)
        off += L
    return torch.cat(outs, 0), maps


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("lens", [[129, 129, 129], [11, 83, 1, 62], [249, 33]])
def test_attention_cuda_core_fwd_bwd(dtype, lens):
    from editor_b200 import lib
    T, H = sum(lens), 12
    qkv32 = torch.randn(T, 3 * H * 64, generator=_g(1)).cuda()
    qkv = qkv32.to(dtype)
    seq_off = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32).cuda()
    ml = max(lens)
    ldp = (ml + 7) // 8 * 8
    out = torch.empty(T, H * 64, dtype=dtype, device="cuda")
    P = torch.zeros(len(lens) * H, ml, ldp, dtype=dtype, device="cuda")
    lib.attention(qkv, out, P, len(lens), H, ml, 0.125, seq_off=seq_off, p_rows=ml, ldp=ldp, impl=1)
    qr = qkv.float().requires_grad_(True)
    ref, maps = _torch_attention(qr, lens)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert _rel(out.float(), ref.detach()) < tol
    for s, L in enumerate(lens):
        got = P.view(len(lens), H, ml, ldp)[s, :, :L, :L].float()
        assert _rel(got, maps[s].detach()) < tol
    d_out = torch.randn(T, H * 64, generator=_g(2)).cuda().to(dtype)
    ref.backward(d_out.float())
    d_qkv = torch.empty_like(qkv)
    lib.attention(qkv, None, P, len(lens), H, ml, 0.125, seq_off=seq_off, p_rows=ml, ldp=ldp, impl=1, d_out=d_out,
                  d_qkv=d_qkv, backward=True)
    assert _rel(d_qkv.float(), qr.grad) < (1e-4 if dtype == torch.float32 else 3e-2)


def test_attention_tensor_core_matches_cuda_core():
    """The tcgen05 kernel for 129-token sequences against the CUDA-core kernel and torch."""
    from editor_b200 import lib
    S, H = 7, 12
    qkv = (torch.randn(S * 129, 2304, generator=_g(1)) * 1.5).cuda().to(torch.bfloat16)
    outs, maps = [], []
    for impl in (0, 1):
        out = torch.zeros(S * 129, 768, dtype=torch.bfloat16, device="cuda")
        P = torch.full((S * H, 129, 136), 7.0, dtype=torch.bfloat16, device="cuda")
        lib.attention(qkv, out, P, S, H, 129, 0.125, fixed_len=129, p_rows=129, ldp=136, impl=impl)
        outs.append(out)
        maps.append(P)
    torch.cuda.synchronize()
    ref, rmaps = _torch_attention(qkv.float(), [129] * S)
    assert _rel(outs[0].float(), ref) < 2e-2
    assert _rel(outs[0].float(), outs[1].float()) < 2e-2
    got = maps[0].view(S, H, 129, 136)
    assert _rel(got[..., :129].float(), torch.stack(rmaps)) < 2e-2
    assert torch.all(got[..., 129:] == 0)
    d_out = torch.randn(S * 129, 768, generator=_g(2)).cuda().to(torch.bfloat16)
    grads = []
    for impl in (0, 1):
        d_qkv = torch.zeros_like(qkv)
        lib.attention(qkv, outs[1] if impl == 0 else None, maps[1], S, H, 129, 0.125, fixed_len=129, p_rows=129, ldp=136,
                      impl=impl, d_out=d_out, d_qkv=d_qkv, backward=True)
        grads.append(d_qkv)
    torch.cuda.synchronize()
    assert _rel(grads[0].float(), grads[1].float()) < 3e-2


@pytest.mark.parametrize("S", [1, 7, 64])
def test_attention_tensor_core_vs_torch_autograd(S):
    """attn_tc_fwd / attn_tc_bwd (the backbone's kernels) against torch autograd DIRECTLY (not through the CUDA-core kernel):
    out, the stored P maps, and d_qkv.  S = 64 -> 768 (sequence, head) items: several per CTA once the kernels loop."""
    from editor_b200 import lib
    H = 12
    qkv = (torch.randn(S * 129, 2304, generator=_g(3)) * 1.2).cuda().to(torch.bfloat16)
    out = torch.zeros(S * 129, 768, dtype=torch.bfloat16, device="cuda")
    P = torch.full((S * H, 129, 136), 7.0, dtype=torch.bfloat16, device="cuda")
    lib.attention(qkv, out, P, S, H, 129, 0.125, fixed_len=129, p_rows=129, ldp=136, impl=0)
    qr = qkv.float().requires_grad_(True)
    ref, rmaps = _torch_attention(qr, [129] * S)
    assert _rel(out.float(), ref.detach()) < 2e-2                    # tolerance: bf16 operands / bf16 P, fp32 accumulate
    got = P.view(S, H, 129, 136)
    assert _rel(got[..., :129].float(), torch.stack(rmaps).detach()) < 2e-2
    assert torch.all(got[..., 129:] == 0)
    d_out = torch.randn(S * 129, 768, generator=_g(4)).cuda().to(torch.bfloat16)
    ref.backward(d_out.float())
    d_qkv = torch.full_like(qkv, 3.0)
    lib.attention(qkv, out, P, S, H, 129, 0.125, fixed_len=129, p_rows=129, ldp=136, impl=0, d_out=d_out, d_qkv=d_qkv,
                  backward=True)
    torch.cuda.synchronize()
    err = ((d_qkv.float() - qr.grad).norm() / qr.grad.norm()).item()
    print("attn_tc_bwd vs torch autograd: L2 rel %.3e, max rel %.3e" % (err, _rel(d_qkv.float(), qr.grad)))
    assert err < 1e-2 and _rel(d_qkv.float(), qr.grad) < 3e-2        # tolerance: bf16 P / dS operands


def test_selection_kernels_bit_exact():
    from editor_b200 import lib, synth
    for (H, W) in ((256, 128), (128, 256)):
        x, _, _ = synth.synthetic_batch(6, H, W, seed=3)
        xg = {k: v.cuda() for k, v in x.items()}
        counts = torch.empty(6, 128, dtype=torch.int32, device="cuda")
        lib.call("edb_freq_counts", xg["RGB"].data_ptr(), xg["NI"].data_ptr(), xg["TI"].data_ptr(), 6, H, W,
                 counts.data_ptr(), lib.stream_ptr())
        assert torch.equal(counts.cpu(), orc.frequency_counts(x["RGB"], x["NI"], x["TI"], faithful=True))
    # top-k with heavy ties, every k, int and float rows -- against the rule measured on torch.topk (CUDA)
    g = _g(5)
    for k in (1, 2, 10, 16, 32, 64, 96, 128):
        vi = torch.randint(100, 140, (64, 128), generator=g, dtype=torch.int32)
        vi[0] = 256
        vf = (torch.randint(0, 20, (64, 128), generator=g).float() / 16.0)
        for vals, isf in ((vi, 0), (vf, 1)):
            mask = torch.zeros(64, 4, dtype=torch.int32, device="cuda")
            v = vals.cuda()
            lib.call("edb_topk_mask", v.data_ptr(), isf, 128, 64, 128, k, mask.data_ptr(), 0, lib.stream_ptr())
            bits = ((mask.cpu().view(-1, 4, 1) >> torch.arange(32).view(1, 1, 32)) & 1).bool().reshape(-1, 128)
            assert torch.equal(bits, orc.topk_mask(vals, k)), k
            idx = torch.topk(v, k, dim=1).indices
            tm = torch.zeros(64, 128, dtype=torch.bool, device="cuda").scatter_(1, idx, True)
            assert torch.equal(bits, tm.cpu()), ("torch.topk on this GPU", k)


def test_rollout_topk_matches_oracle():
    from editor_b200 import lib
    B, S, H, L = 2, 6, 12, 12
    maps = [torch.softmax(torch.randn(S, H, 129, 129, generator=_g(l)) * 2, -1) for l in range(L)]
    bufs = []
    for m in maps:
        b = torch.zeros(S * H, 129, 136, device="cuda")
        b[:, :, :129] = m.view(S * H, 129, 129).cuda()
        bufs.append(b)
    index = torch.zeros(B, 4, dtype=torch.int32, device="cuda")
    mod = torch.zeros(S, 4, dtype=torch.int32, device="cuda")
    rows = torch.empty(S * H, 128, device="cuda")
    arr = (ctypes.c_void_p * L)(*[b.data_ptr() for b in bufs])
    lib.call("edb_rollout_topk", arr, L, 1, S, B, H, 129, 136, 2, index.data_ptr(), mod.data_ptr(), rows.data_ptr(),
             lib.stream_ptr())
    want_rows = orc.rollout_full(maps)                       # SFTS.py:150-153 as written (full matrix product)
    assert _rel(rows.cpu().view(S, H, 128), want_rows) < 1e-4
    want = orc.part_attention_mask(maps, 2, full=True)
    bits = ((mod.cpu().view(-1, 4, 1) >> torch.arange(32).view(1, 1, 32)) & 1).bool().reshape(-1, 128)
    assert torch.equal(bits, want)
    ib = ((index.cpu().view(-1, 4, 1) >> torch.arange(32).view(1, 1, 32)) & 1).bool().reshape(-1, 128)
    assert torch.equal(ib, want[0:2] | want[2:4] | want[4:6])


def test_droppath_matches_oracle():
    """DropPath (vit_pytorch.py:52-69) as per-sequence row scales in the residual epilogues and LN backward."""
    import __graft_entry__ as ge
    model, sd, x, label, cam, al = ge._small_case(True, 4)
    model = model.cuda().train()
    eng = model.engine()
    g = _g(9)
    dp = [torch.floor(0.7 + torch.rand(12, generator=g)) / 0.7 for _ in range(24)]
    eng._droppath = lambda B, device: [d.to(device) for d in dp]
    xg = {k: v.cuda() for k, v in x.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = model(xg, label=label.cuda(), cam_label=cam.cuda(), writer=None, epoch=1)
    own_sel = ((eng.sel["index"].cpu().view(-1, 4, 1) >> torch.arange(32).view(1, 1, 32)) & 1).bool().reshape(-1, 128)
    dpo = [[d[m * 4:(m + 1) * 4] for d in dp] for m in range(3)]
    ref = orc.editor_forward(sd, x, cam, label=label, training=True, al=al, droppath=dpo, force_index=own_sel)
    for a, b in zip(outs, ref):
        assert _rel(a.float().cpu(), b.detach()) < 3e-2


def test_sgd_kernel_matches_reference_optimizer_arithmetic():
    from editor_b200 import lib
    from oracle import sgd_oracle
    n = 64 * 50
    g = _g(3)
    p0, gr = torch.randn(n, generator=g), torch.randn(n, generator=g)
    flags = torch.randint(0, 3, (n // 64,), generator=g, dtype=torch.uint8)
    pc, bc = p0.clone(), torch.zeros(n)
    pg, bg, p16 = p0.clone().cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, dtype=torch.bfloat16, device="cuda")
    for step in range(3):
        grs = gr * (step + 1)
        sgd_oracle.flat_sgd_step(pc, grs, bc, flags, 0.001, 0.9, 1e-4, 1e-4, 2.0, 0.5, step == 0)
        lib.call("edb_sgd_step", pg.data_ptr(), grs.cuda().data_ptr(), bg.data_ptr(), p16.data_ptr(), flags.cuda().data_ptr(),
                 n, 0.001, 0.9, 1e-4, 1e-4, 2.0, 0.5, int(step == 0), lib.stream_ptr())
    assert torch.allclose(pg.cpu(), pc, rtol=1e-6, atol=1e-7)
    live = (flags.repeat_interleave(64) & 2) == 0
    assert torch.equal(p16.cpu()[live], pc.to(torch.bfloat16)[live])


@pytest.mark.parametrize("lens", [[11, 83, 1, 62, 128], [249, 33, 130, 256, 7]])
def test_attention_tensor_core_varlen(lens):
    """tcgen05 var-len kernel (impl 2, HMA's packed AttentionMask) against torch and the CUDA-core kernel."""
    from editor_b200 import lib
    T, H = sum(lens), 12
    qkv = (torch.randn(T, 3 * H * 64, generator=_g(1)) * 1.2).cuda().to(torch.bfloat16)
    seq_off = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32).cuda()
    ml = max(lens)
    kp = 128 if ml <= 128 else 256
    pr = (ml + 127) // 128 * 128
    out = torch.zeros(T, H * 64, dtype=torch.bfloat16, device="cuda")
    P = torch.full((len(lens) * H, pr, kp), 3.0, dtype=torch.bfloat16, device="cuda")
    lib.attention(qkv, out, P, len(lens), H, ml, 0.125, seq_off=seq_off, p_rows=pr, ldp=kp, impl=2, total_rows=T)
    torch.cuda.synchronize()
    qr = qkv.float().requires_grad_(True)
    ref, maps = _torch_attention(qr, lens)
    assert _rel(out.float(), ref.detach()) < 2e-2
    Pv = P.view(len(lens), H, pr, kp)
    for s, L in enumerate(lens):
        assert _rel(Pv[s, :, :L, :L].float(), maps[s].detach()) < 2e-2
        nq = (L + 127) // 128 * 128
        assert torch.all(Pv[s, :, :nq, L:] == 0) and torch.all(Pv[s, :, L:nq, :] == 0)     # zero outside the sequence
    d_out = torch.randn(T, H * 64, generator=_g(2)).cuda().to(torch.bfloat16)
    ref.backward(d_out.float())
    d_qkv = torch.zeros_like(qkv)
    lib.attention(qkv, None, P, len(lens), H, ml, 0.125, seq_off=seq_off, p_rows=pr, ldp=kp, impl=2, d_out=d_out,
                  d_qkv=d_qkv, backward=True, total_rows=T)
    torch.cuda.synchronize()
    assert _rel(d_qkv.float(), qr.grad) < 3e-2


def test_rollout_topk_bf16_maps():
    from editor_b200 import lib
    B, S, H, L = 2, 6, 12, 12
    maps = [torch.softmax(torch.randn(S, H, 129, 129, generator=_g(20 + l)) * 2, -1).to(torch.bfloat16) for l in range(L)]
    bufs = []
    for m in maps:
        b = torch.zeros(S * H, 129, 136, dtype=torch.bfloat16, device="cuda")
        b[:, :, :129] = m.view(S * H, 129, 129).cuda()
        bufs.append(b)
    index = torch.zeros(B, 4, dtype=torch.int32, device="cuda")
    mod = torch.zeros(S, 4, dtype=torch.int32, device="cuda")
    rows = torch.empty(S * H, 128, device="cuda")
    arr = (ctypes.c_void_p * L)(*[b.data_ptr() for b in bufs])
    lib.call("edb_rollout_topk", arr, L, 0, S, B, H, 129, 136, 2, index.data_ptr(), mod.data_ptr(), rows.data_ptr(),
             lib.stream_ptr())
    mf = [m.float() for m in maps]
    want_rows = orc.rollout_cls_row(mf)
    assert _rel(rows.cpu().view(S, H, 128), want_rows) < 1e-4
    got = ((mod.cpu().view(-1, 4, 1) >> torch.arange(32).view(1, 1, 32)) & 1).bool().reshape(-1, 128)
    kth = torch.sort(want_rows, dim=-1, descending=True).values
    safe = ((kth[..., 1] - kth[..., 2]) / kth[..., 1] > 1e-3).all(1)     # rows whose top-2 is not a near-tie
    want = orc.part_attention_mask(mf, 2)
    assert safe.sum() >= S - 1 and torch.equal(got[safe], want[safe])


def test_tail_batchnorm_linear_ocfr_match_torch():
    from editor_b200 import lib, tail
    import __graft_entry__ as ge
    model, sd, x, label, cam, al = ge._small_case(True, 4)
    model = model.cuda().train()
    eng = model.engine()
    eng._ensure(torch.device("cuda", 0))
    eng.arena.refresh16()
    B = 16
    g = _g(4)
    # --- BatchNorm1d
    xin = torch.randn(B, 2304, generator=g).cuda().requires_grad_(True)
    bn_ref = torch.nn.BatchNorm1d(2304).cuda().train()
    bn_ref.load_state_dict(model.FUSE_BN.state_dict())
    y = tail.BatchNormFn.apply(eng, "FUSE_BN", model.FUSE_BN, xin)
    xr = xin.detach().clone().requires_grad_(True)
    yr = bn_ref(xr)
    assert _rel(y, yr.detach()) < 1e-5
    assert _rel(model.FUSE_BN.running_var, bn_ref.running_var) < 1e-6 and _rel(model.FUSE_BN.running_mean, bn_ref.running_mean) < 1e-5
    dy = torch.randn(B, 2304, generator=g).cuda()
    eng.arena.grad.zero_()
    y.backward(dy)
    yr.backward(dy)
    assert _rel(xin.grad, xr.grad) < 1e-4
    assert _rel(model.FUSE_BN.weight.grad, bn_ref.weight.grad) < 1e-4 and _rel(model.FUSE_BN.bias.grad, bn_ref.bias.grad) < 1e-4
    # --- Linear (REDUCE with bias, head without, N = 171)
    for name, K in (("RGB_REDUCE", 1536), ("FUSE_HEAD", 2304)):
        L = eng.tail_lin[name]
        mod = getattr(model, name)
        xi = (torch.randn(B, K, generator=g) * 0.5).cuda().requires_grad_(True)
        eng.arena.grad.zero_()
        mod.weight.grad = None
        yo = tail.LinearFn.apply(eng, L, "bf16", xi)
        ref = torch.nn.functional.linear(xi.detach().to(torch.bfloat16).float(), mod.weight.detach().to(torch.bfloat16).float(),
                                         None if mod.bias is None else mod.bias.detach())
        assert _rel(yo, ref) < 2e-3
        do = torch.randn_like(ref)
        yo.backward(do)
        dob = do.to(torch.bfloat16).float()
        assert _rel(xi.grad, dob @ mod.weight.detach().to(torch.bfloat16).float()) < 2e-3
        assert _rel(mod.weight.grad, dob.t() @ xi.detach().to(torch.bfloat16).float()) < 2e-3
        if mod.bias is not None:
            assert _rel(mod.bias.grad, do.sum(0)) < 1e-4
    # --- OCFR
    mem = model.FUSE_block.memory_cls
    C = mem.RGB_centers.shape[0]
    cen0 = [torch.randn(C, 768, generator=g) * 0.05 for _ in range(3)]
    for p, c in zip((mem.RGB_centers, mem.NIR_centers, mem.TIR_centers), cen0):
        p.data.copy_(c.cuda())
    lab = torch.tensor([3, 3, 3, 3, 9, 9, 9, 9, 0, 0, 0, 0, 170, 170, 170, 170])
    cm = torch.randn(3, B, 768, generator=g).cuda().requires_grad_(True)
    loss = tail.OcfrFn.apply(mem, cm, lab.cuda())
    cr = cm.detach().cpu().clone().requires_grad_(True)
    cen_ref = [c.clone() for c in cen0]
    lr = orc.ocfr([cr[0], cr[1], cr[2]], lab, cen_ref)
    assert abs(loss.item() - lr.item()) < 1e-5 * max(1.0, abs(lr.item()))
    for p, c in zip((mem.RGB_centers, mem.NIR_centers, mem.TIR_centers), cen_ref):
        assert _rel(p.data.cpu(), c) < 1e-5
    (loss * 3.0).backward()
    (lr * 3.0).backward()
    assert _rel(cm.grad.cpu(), cr.grad) < 1e-4


def test_loss_kernels_match_oracle_loss():
    from editor_b200 import tail
    B, C, F = 32, 171, 2304
    g = _g(6)
    label = torch.arange(8).repeat_interleave(4)
    logits = (torch.randn(B, C, generator=g) * 3).cuda().requires_grad_(True)
    feat = torch.randn(B, F, generator=g).cuda().requires_grad_(True)
    feat2 = (torch.randn(B, 768, generator=g) * 0.1).cuda().requires_grad_(True)
    aux = torch.tensor(0.7, device="cuda", requires_grad=True)
    outs = (logits, feat, logits * 0.5, feat2, aux)
    loss = tail.editor_loss(outs, label.cuda())
    ref_in = [t.detach().cpu().clone().requires_grad_(True) for t in (logits, feat, feat2, aux)]
    lr = orc.reference_loss((ref_in[0], ref_in[1], ref_in[0] * 0.5, ref_in[2], ref_in[3]), label)
    assert abs(loss.item() - lr.item()) < 1e-5 * abs(lr.item())
    (loss * 2.0).backward()
    (lr * 2.0).backward()
    assert _rel(logits.grad.cpu(), ref_in[0].grad) < 1e-4
    assert _rel(feat.grad.cpu(), ref_in[1].grad) < 1e-4
    assert _rel(feat2.grad.cpu(), ref_in[2].grad) < 1e-4
    assert abs(aux.grad.item() - 2.0) < 1e-6
