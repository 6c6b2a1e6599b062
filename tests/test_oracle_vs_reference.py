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
