"""The CPU oracle (oracle/editor_oracle.py) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py, run in the build container where /root/reference exists)."""
import os

import pytest
import torch

from editor_b200 import synth
from oracle import editor_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    return torch.load(os.path.join(HERE, "golden", "ref_%s.pt" % name), weights_only=False)


def _setup(g):
    m = g["meta"]
    sd = synth.synthetic_state_dict(seed=m["weights_seed"], num_class=m["C"], camera_num=m["cams"], al=m["al"])
    x, label, cam = synth.synthetic_batch(m["B"], m["H"], m["W"], seed=m["batch_seed"], num_cams=m["cams"],
                                          instances=m["instances"])
    return m, sd, x, label, cam


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize("case", ["rgbnt201", "rgbnt100"])
def test_eval_forward_matches_reference(case):
    g = _load(case)
    m, sd, x, label, cam = _setup(g)
    aux = {}
    with torch.no_grad():
        out = orc.editor_forward(sd, x, cam, training=False, al=m["al"], aux=aux)
    assert torch.equal(aux["mask_fre"], g["eval_mask_fre"])
    assert torch.equal(aux["index"], g["eval_index"])
    assert out.shape == g["eval_cls4t"].shape == (m["B"], 2304)
    assert _rel(out, g["eval_cls4t"]) < 1e-4
    # the pixel-mean shortcut the CUDA kernel uses gives the same counts as the 4-level Haar round trip
    c0 = orc.frequency_counts(x["RGB"], x["NI"], x["TI"], faithful=True)
    c1 = orc.frequency_counts(x["RGB"], x["NI"], x["TI"], faithful=False)
    assert torch.equal(c0, c1)


@pytest.mark.parametrize("case", ["rgbnt201", "rgbnt100"])
def test_train_forward_backward_matches_reference(case):
    g = _load(case)
    m, sd, x, label, cam = _setup(g)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "centers" not in k and "running" not in k
              and not k.startswith("FREQ_INDEX") else v) for k, v in sd.items()}
    state, aux = {}, {}
    outs = orc.editor_forward(sd, x, cam, label=label, training=True, al=m["al"], state_out=state, aux=aux)
    assert len(outs) == len(g["train_outputs"])
    assert torch.equal(aux["index"], g["train_index"])
    for a, b in zip(outs, g["train_outputs"]):
        assert a.shape == b.shape
        assert _rel(a.detach(), b) < 2e-4
    assert abs(float(aux["num"].float().mean()) - g["num_count"]) < 1e-6
    loss = orc.reference_loss(outs, label)
    assert abs(loss.item() - g["loss"].item()) < 2e-4 * abs(g["loss"].item())
    loss.backward()
    for k, ref in g["grads"].items():
        gr = sd[k].grad
        assert gr is not None, k
        assert abs(gr.norm().item() - ref["norm"]) <= 2e-3 * ref["norm"] + 1e-7, k
    for k, ref in g["grads_full"].items():
        if ref.abs().max() < 1e-6:      # e.g. *_REDUCE.bias: cancelled exactly by the batch-stat BN / pairwise distances
            assert sd[k].grad.abs().max() < 1e-6, k
        else:
            assert _rel(sd[k].grad, ref) < 2e-3, k
    unused = [k for k, v in sd.items() if getattr(v, "requires_grad", False) and v.grad is None]
    assert all(k.startswith("BACKBONE.base.fc") or k.startswith("BACKBONE_") for k in unused), unused
    for k, ref in g["state_after"].items():
        got = state.get(k, sd[k])           # BN layers not called in this configuration keep their buffers
        if "centers" in k:
            got = got[label.unique()]
        assert _rel(got.float(), ref.float()) < 1e-4, k


def test_topk_rule_is_the_measured_cuda_rule():
    """tests/golden/topk_probe_b200.json: torch.topk on B200 matched the ascending-index tie rule on every row."""
    import json
    with open(os.path.join(HERE, "golden", "topk_probe_b200.json")) as f:
        probe = json.load(f)
    for c in probe["cases"]:
        assert c["cuda_matches_asc_rule"] == c["rows"], c
    ex = probe["example_k3"]
    m = orc.topk_mask(torch.tensor(ex["x"]), 3)
    assert sorted(torch.nonzero(m[0])[:, 0].tolist()) == sorted(ex["cuda_idx"][0])
    assert orc.topk_mask(torch.full((1, 128), 256, dtype=torch.int32), 10)[0].nonzero()[:, 0].tolist() == list(range(10))
