"""tcgen05 GEMM vs a plain torch fp32 matmul of the same bf16-rounded operands (all operand majors / epilogues)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(rows, cols, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(rows, cols, generator=g) * scale).to(torch.bfloat16).cuda()


def _ref(A, B, a_mn, b_mn):
    Af = A.float().t() if a_mn else A.float()
    Bf = B.float().t() if b_mn else B.float()
    return Af @ Bf.t()


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 768), (387, 2304, 768), (1000, 768, 3072),
                                   (258, 136, 200), (129 * 6, 768, 768), (64, 171, 2304)])
def test_gemm_majors(M, N, K, a_mn, b_mn):
    from editor_b200 import lib
    Mp, Np = (M + 7) // 8 * 8, (N + 7) // 8 * 8
    Kp = (K + 7) // 8 * 8
    A = _mk(K, Mp, 1)[:, :M] if a_mn else _mk(M, Kp, 1)[:, :K]
    B = _mk(K, Np, 2)[:, :N] if b_mn else _mk(N, Kp, 2)[:, :K]
    D = torch.full((M, Np), 7.0, dtype=torch.float32, device="cuda")
    lib.gemm(A, B, D, M, N, K, a_mn=a_mn, b_mn=b_mn)
    torch.cuda.synchronize()
    ref = _ref(A, B, a_mn, b_mn)
    err = (D[:, :N] - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err
    if Np > N:
        assert torch.all(D[:, N:] == 7.0)  # padding columns untouched


def test_gemm_bias_gelu_bf16_out():
    from editor_b200 import lib
    M, N, K = 516, 3072, 768
    A, B = _mk(M, K, 3, 0.5), _mk(N, K, 4, 0.05)
    bias = torch.randn(N, device="cuda") * 0.1
    D = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    pre = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    lib.gemm(A, B, D, M, N, K, epilogue=lib.EPI_GELU, bias=bias, out2=pre)      # bf16: out2 = gelu'(pre)
    D1 = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    lib.gemm(A, B, D1, M, N, K, epilogue=lib.EPI_GELU, bias=bias)                # eval: no second output
    torch.cuda.synchronize()
    ref_pre = (A.float() @ B.float().t() + bias).requires_grad_(True)
    ref = torch.nn.functional.gelu(ref_pre)
    ref.sum().backward()
    assert (pre.float() - ref_pre.grad).abs().max().item() < 6e-3       # bf16 rounding of values up to 1.13: 4e-3
    assert (D.float() - ref.detach()).abs().max().item() < 2e-2
    assert (D1.float() - ref.detach()).abs().max().item() < 2e-2
    # fp32 output (EDB_PREC_FP32): exact erf GELU and the pre-activation itself in out2
    Df, pref = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    lib.gemm(A, B, Df, M, N, K, epilogue=lib.EPI_GELU, bias=bias, out2=pref)
    torch.cuda.synchronize()
    assert (pref - ref_pre.detach()).abs().max().item() < 2e-3
    assert (Df - ref.detach()).abs().max().item() < 2e-3


def test_gelu_epilogue_accuracy_over_the_whole_range():
    """The one-tanh GELU / GELU' of the bf16 epilogues against exact erf GELU on a dense grid of pre-activations
    (identity weights, so acc = x exactly): errors stay below half a bf16 ulp of the outputs."""
    from editor_b200 import lib
    M, N = 256, 256
    x = torch.linspace(-12.0, 12.0, M * N, device="cuda").view(M, N).to(torch.bfloat16)
    eye = torch.eye(N, device="cuda").to(torch.bfloat16)
    g, d = (torch.empty(M, N, dtype=torch.bfloat16, device="cuda") for _ in range(2))
    lib.gemm(x, eye, g, M, N, N, epilogue=lib.EPI_GELU, out2=d)
    torch.cuda.synchronize()
    xr = x.double().requires_grad_(True)
    ref = torch.nn.functional.gelu(xr)
    ref.sum().backward()
    eg = (g.double() - ref.detach()).abs() / (ref.detach().abs() * 2.0 ** -8 + 1e-3)
    ed = (d.double() - xr.grad).abs() / (xr.grad.abs() * 2.0 ** -8 + 1e-3)
    assert eg.max().item() < 1.0, eg.max().item()       # < 1 bf16 ulp (+1e-3 absolute floor near zero)
    assert ed.max().item() < 1.0, ed.max().item()


def test_gemm_residual_inplace():
    from editor_b200 import lib
    M, N, K = 774, 768, 3072
    A, B = _mk(M, K, 5, 0.5), _mk(N, K, 6, 0.05)
    bias = torch.randn(N, device="cuda") * 0.1
    x = torch.randn(M, N, device="cuda")
    ref = x + A.float() @ B.float().t() + bias
    lib.gemm(A, B, x, M, N, K, epilogue=lib.EPI_RESIDUAL, bias=bias, aux=x)
    torch.cuda.synchronize()
    assert (x - ref).abs().max().item() < 2e-3 * ref.abs().max().item()


def test_gemm_gelu_bwd_and_splitk():
    from editor_b200 import lib
    M, N, K = 640, 3072, 768
    dY, W = _mk(M, K, 7, 0.5), _mk(K, N, 8, 0.05)           # dH = dY @ W  (W stored [K_in=768 rows(k), N cols])
    x = _mk(M, N, 9).float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    dact = x.grad.to(torch.bfloat16)                         # the factor the forward epilogue saves
    D = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    csum = torch.full((N,), 0.5, device="cuda")             # fused bias gradient: ACCUMULATES the column sums of D
    lib.gemm(dY, W, D, M, N, K, b_mn=True, epilogue=lib.EPI_GELU_BWD, aux=dact, colsum=csum)
    torch.cuda.synchronize()
    refD = (dY.float() @ W.float()) * dact.float()
    assert (D.float() - refD).abs().max().item() < 3e-2
    assert (csum - 0.5 - refD.sum(0)).abs().max().item() < 2e-3 * refD.abs().sum(0).max().item()
    # ragged rows (device-side row count): rows beyond it contribute nothing
    m_dev = torch.tensor([M - 77], dtype=torch.int32, device="cuda")
    csum2 = torch.zeros(N, device="cuda")
    lib.gemm(dY, W, D, M, N, K, b_mn=True, epilogue=lib.EPI_GELU_BWD, aux=dact, colsum=csum2, M_dev=m_dev.data_ptr())
    torch.cuda.synchronize()
    assert (csum2 - refD[:M - 77].sum(0)).abs().max().item() < 2e-3 * refD.abs().sum(0).max().item()
    # wgrad with split-K: dW[N', K'] = dY^T X, reduction over M rows
    Mr, Nw, Kw = 129 * 48, 768, 768
    dY2, X2 = _mk(Mr, Nw, 10, 0.1), _mk(Mr, Kw, 11, 0.1)
    dW = torch.zeros(Nw, Kw, dtype=torch.float32, device="cuda")
    lib.gemm(dY2, X2, dW, Nw, Kw, Mr, a_mn=True, b_mn=True, epilogue=lib.EPI_ATOMIC, split_k=8)
    torch.cuda.synchronize()
    ref = dY2.float().t() @ X2.float()
    assert (dW - ref).abs().max().item() < 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (129 * 6, 768, 768), (1000, 2304, 200), (49536 // 8, 3072, 768)])
def test_gemm_cta_pair_equals_single_cta(M, N, K, a_mn, b_mn):
    """The cta_group::2 path (256 x 256 MMA over a CTA pair) accumulates in the same order as the single-CTA path:
    bit-identical results, including the ragged last pair tile (M not a multiple of 256) and a device-side row count."""
    from editor_b200 import lib
    Kp, Mp, Np = (K + 7) // 8 * 8, (M + 7) // 8 * 8, (N + 7) // 8 * 8
    A = _mk(K, Mp, 21)[:, :M] if a_mn else _mk(M, Kp, 21)[:, :K]
    B = _mk(K, Np, 22)[:, :N] if b_mn else _mk(N, Kp, 22)[:, :K]
    bias = torch.randn(N, device="cuda")
    m_dev = torch.tensor([M - 131], dtype=torch.int32, device="cuda")
    outs = []
    try:
        for mode in (1, 0):
            lib.gemm_set_mode(mode)
            D = torch.full((M, N), 7.0, dtype=torch.float32, device="cuda")
            D2 = torch.full((M, N), 7.0, dtype=torch.float32, device="cuda")
            lib.gemm(A, B, D, M, N, K, a_mn=a_mn, b_mn=b_mn, bias=bias)
            lib.gemm(A, B, D2, M, N, K, a_mn=a_mn, b_mn=b_mn, M_dev=m_dev.data_ptr())
            torch.cuda.synchronize()
            outs.append((D, D2))
    finally:
        lib.gemm_set_mode(0)
    ref = _ref(A, B, a_mn, b_mn) + bias
    assert (outs[1][0] - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    assert torch.all(outs[1][1][M - 131:] == 7.0)      # rows beyond the device-side count stay untouched


def test_gemm_cta_pair_epilogues_and_splitk_equal_single_cta():
    from editor_b200 import lib
    M, N, K = 129 * 10, 3072, 768
    A, W1, W2 = _mk(M, K, 31, 0.5), _mk(N, K, 32, 0.05), _mk(K, N, 33, 0.05)
    H = _mk(M, N, 34, 0.5)
    bias, bias2 = torch.randn(N, device="cuda") * 0.1, torch.randn(K, device="cuda") * 0.1
    res = torch.randn(M, K, device="cuda")
    scale = torch.rand(10, device="cuda")
    outs = []
    try:
        for mode in (1, 0):
            lib.gemm_set_mode(mode)
            g, pre = (torch.empty(M, N, dtype=torch.bfloat16, device="cuda") for _ in range(2))
            lib.gemm(A, W1, g, M, N, K, epilogue=lib.EPI_GELU, bias=bias, out2=pre)
            r = torch.empty(M, K, device="cuda")
            lib.gemm(H, W2, r, M, K, N, epilogue=lib.EPI_RESIDUAL, bias=bias2, aux=res, row_scale=scale, scale_group=129)
            db = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
            lib.gemm(A, W2, db, M, N, K, b_mn=True, epilogue=lib.EPI_GELU_BWD, aux=H)
            dW = torch.zeros(N, K, device="cuda")
            lib.gemm(H, A, dW, N, K, M, a_mn=True, b_mn=True, epilogue=lib.EPI_ATOMIC, split_k=1)
            dW4 = torch.zeros(N, K, device="cuda")
            lib.gemm(H, A, dW4, N, K, M, a_mn=True, b_mn=True, epilogue=lib.EPI_ATOMIC, split_k=4)
            torch.cuda.synchronize()
            outs.append((g, pre, r, db, dW, dW4))
    finally:
        lib.gemm_set_mode(0)
    for a, b in list(zip(*outs))[:5]:
        assert torch.equal(a, b)
    ref = H.float().t() @ A.float()
    assert (outs[1][5] - ref).abs().max().item() < 2e-3 * ref.abs().max().item()   # split-K: atomic order is free
    ref_r = res + scale.repeat_interleave(129)[:, None] * (H.float() @ W2.float().t() + bias2)
    assert (outs[1][2] - ref_r).abs().max().item() < 2e-3 * ref_r.abs().max().item()


def test_gemm_throughput_report(capsys):
    """Not an assertion on speed: prints TFLOP/s of the fc1-shaped GEMM for the log."""
    from editor_b200 import lib
    M, N, K = 49536, 3072, 768
    A, B = _mk(M, K, 12, 0.5), _mk(N, K, 13, 0.05)
    D = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    try:
        for mode, name in ((1, "single CTA 128x256"), (0, "CTA pair 256x256")):
            lib.gemm_set_mode(mode)
            for _ in range(3):
                lib.gemm(A, B, D, M, N, K)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                lib.gemm(A, B, D, M, N, K)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 10
            with capsys.disabled():
                print("\n[gemm 49536x3072x768 bf16, %s] %.3f ms  %.1f TFLOP/s" % (name, ms, 2.0 * M * N * K / ms / 1e9))
    finally:
        lib.gemm_set_mode(0)
    ref = A[:256].float() @ B.float().t()
    assert (D[:256].float() - ref).abs().max().item() < 5e-2
