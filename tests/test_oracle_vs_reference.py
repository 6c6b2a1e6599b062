"""Optional live cross-check (build container only: the GPU box has no /root/reference): the torch oracle against the
UNMODIFIED reference model on configurations and seeds that are NOT in the committed goldens -- MSVR310's yml
(128x256, AL=0), another weight seed and another batch seed.  Skipped wherever the reference tree is absent."""
import pytest
import torch

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


@pytest.mark.parametrize("ds,C,cams,H,W,al", [("MSVR310", 155, 8, 128, 256, False), ("RGBNT201", 171, 4, 256, 128, True)])
def test_oracle_eval_forward_equals_live_reference(ds, C, cams, H, W, al):
    from editor_b200 import synth
    from oracle import editor_oracle as orc
    from oracle.editor_oracle import topk_mask
    torch.manual_seed(0)
    model, cfg = ref_import.load_reference(ds, C, cams, opts=("MODEL.DROP_PATH", 0.0))
    sd = synth.synthetic_state_dict(seed=2024, num_class=C, camera_num=cams, al=al)
    model.load_state_dict(sd, strict=True)
    x, label, cam = synth.synthetic_batch(2, H, W, seed=7, num_cams=cams, instances=2)
    native = torch.topk

    def cuda_rule_topk(t, k, dim=-1, largest=True, sorted=True):  # noqa: A002  (the tie rule measured on B200)
        m = topk_mask(t, k)
        idx = torch.nonzero(m)[:, 1].reshape(t.shape[0], k)
        return torch.gather(t, 1, idx), idx
    grabbed = {}
    orig = model.SFTS.forward

    def hook(*a, **k):
        r = orig(*a, **k)
        grabbed["index"] = r[3].detach().clone()
        return r
    model.SFTS.forward = hook
    torch.topk = cuda_rule_topk
    try:
        model.eval()
        with torch.no_grad():
            ref = model(x, cam_label=cam, view_label=None, mode=1, img_path=None)
    finally:
        torch.topk = native
    aux = {}
    with torch.no_grad():
        own = orc.editor_forward(sd, x, cam, training=False, al=al, aux=aux)
    assert torch.equal(aux["index"], grabbed["index"][..., 0].bool())
    assert ((own - ref).abs().max() / ref.abs().max()).item() < 2e-4


def test_oracle_training_forward_and_loss_equal_live_reference_msvr310():
    """MSVR310 yml (AL=0 -> 9-tuple, make_model.py:213), training mode: outputs, loss (processor.py:82-92 with the
    reference's own layers/), selection, BN running statistics after the step."""
    import sys
    from editor_b200 import synth
    from oracle import editor_oracle as orc
    from oracle.editor_oracle import topk_mask
    ds, C, cams, H, W, al = "MSVR310", 155, 8, 128, 256, False
    torch.manual_seed(0)
    model, cfg = ref_import.load_reference(ds, C, cams, opts=("MODEL.DROP_PATH", 0.0))
    sd = synth.synthetic_state_dict(seed=77, num_class=C, camera_num=cams, al=al)
    model.load_state_dict(sd, strict=True)
    x, label, cam = synth.synthetic_batch(4, H, W, seed=3, num_cams=cams, instances=2)
    native = torch.topk

    def cuda_rule_topk(t, k, dim=-1, largest=True, sorted=True):  # noqa: A002
        m = topk_mask(t, k)
        idx = torch.nonzero(m)[:, 1].reshape(t.shape[0], k)
        return torch.gather(t, 1, idx), idx
    torch.topk = cuda_rule_topk
    try:
        model.train()
        res = model(x, label=label, cam_label=cam, view_label=None, img_path=None, writer=ref_import.NullWriter(), epoch=1)
    finally:
        torch.topk = native
    sys.path.insert(0, ref_import.REF_ROOT)
    for n in list(sys.modules):
        if n.split(".")[0] == "layers":
            del sys.modules[n]
    try:
        from layers.softmax_loss import CrossEntropyLabelSmooth
        from layers.triplet_loss import TripletLoss
    finally:
        sys.path.remove(ref_import.REF_ROOT)
    xent, tri = CrossEntropyLabelSmooth(num_classes=C, use_gpu=False), TripletLoss()
    ref_loss = sum(xent(res[i], label) + tri(res[i + 1], label)[0] for i in range(0, len(res) - 1, 2)) + res[-1]
    state = {}
    own = orc.editor_forward(sd, x, cam, label=label, training=True, al=al, state_out=state)
    assert len(own) == len(res) == 9
    for a, b in zip(own, res):
        assert ((a - b.detach()).abs().max() / b.detach().abs().max().clamp_min(1e-6)).item() < 2e-4
    own_loss = orc.reference_loss(own, label)
    assert abs(own_loss.item() - ref_loss.item()) < 2e-4 * abs(ref_loss.item())
    st = model.state_dict()
    for k, v in state.items():
        if "running" in k:
            assert (v - st[k]).abs().max().item() < 1e-5, k
