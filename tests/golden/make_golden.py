"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Weights come from ``editor_b200.synth.synthetic_state_dict`` (seeded, so they are not stored) loaded into the
reference model with ``load_state_dict(strict=True)``; inputs from ``synthetic_batch``.  The reference is run
twice: (a) with the native CPU ``torch.topk`` and (b) with ``torch.topk`` replaced by the CUDA tie rule measured
on B200 (tools/probe_topk.py) -- the reference's GPU behaviour.  (b) is what parity is gated on;
(a) is stored to show which rows depend on the tie order.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from editor_b200 import synth  # noqa: E402
from oracle import ref_import  # noqa: E402
from oracle.editor_oracle import topk_mask  # noqa: E402

_native_topk = torch.topk


def _cuda_rule_topk(x, k, dim=-1, largest=True, sorted=True):  # noqa: A002
    assert largest and dim in (1, -1) and x.dim() == 2
    m = topk_mask(x, k)
    idx = torch.nonzero(m)[:, 1].reshape(x.shape[0], k)
    return torch.gather(x, 1, idx), idx


CASES = {
    # name: (dataset yml, num_class, cams, H, W, batch, instances, AL)
    "rgbnt201": ("RGBNT201", 171, 4, 256, 128, 4, 2, True),
    "rgbnt100": ("RGBNT100", 50, 8, 128, 256, 4, 2, False),
}


def run_case(name, cuda_rule):
    ds, C, cams, H, W, B, inst, al = CASES[name]
    torch.manual_seed(0)
    model, cfg = ref_import.load_reference(ds, C, cams, opts=("MODEL.DROP_PATH", 0.0))
    sd = synth.synthetic_state_dict(seed=1111, num_class=C, camera_num=cams, al=al)
    missing = model.load_state_dict(sd, strict=True)
    x, label, cam = synth.synthetic_batch(B, H, W, seed=1, num_cams=cams, instances=inst)
    torch.topk = _cuda_rule_topk if cuda_rule else _native_topk
    out = {}
    try:
        # capture the selection mask the reference computes
        grabbed = {}
        orig = model.SFTS.forward

        def hook(*a, **k):
            r = orig(*a, **k)
            grabbed["index"] = r[3].detach().clone()
            return r
        model.SFTS.forward = hook
        fre = {}
        orig_f = model.FREQ_INDEX.forward

        def hook_f(*a, **k):
            r = orig_f(*a, **k)
            fre["mask"] = r.detach().clone()
            return r
        model.FREQ_INDEX.forward = hook_f

        model.eval()
        with torch.no_grad():
            feat = model(x, cam_label=cam, view_label=None, mode=1, img_path=None)
        out["eval_cls4t"] = feat.clone()
        out["eval_index"] = grabbed["index"][..., 0].clone()
        out["eval_mask_fre"] = fre["mask"].clone()

        model.train()
        w = ref_import.NullWriter()
        res = model(x, label=label, cam_label=cam, view_label=None, img_path=None, writer=w, epoch=1)
        out["train_outputs"] = [r.detach().clone() for r in res]
        out["train_index"] = grabbed["index"][..., 0].clone()
        out["num_count"] = w.scalars[0][1]
        # loss exactly as engine/processor.py:82-92 with layers/make_loss.py (imported from the reference)
        sys.path.insert(0, ref_import.REF_ROOT)
        for n in list(sys.modules):
            if n.split(".")[0] == "layers":
                del sys.modules[n]
        from layers.softmax_loss import CrossEntropyLabelSmooth
        from layers.triplet_loss import TripletLoss
        sys.path.remove(ref_import.REF_ROOT)
        xent = CrossEntropyLabelSmooth(num_classes=C, use_gpu=False)
        tri = TripletLoss()
        loss = 0
        for i in range(0, len(res) - 1, 2):
            loss = loss + xent(res[i], label) + tri(res[i + 1], label)[0]
        loss = loss + res[-1]
        out["loss"] = loss.detach().clone()
        loss.backward()
        grads = {}
        for k, p in model.named_parameters():
            if p.grad is not None:
                g = p.grad
                grads[k] = {"norm": g.norm().item(), "head": g.flatten()[:16].clone(),
                            "sum": g.double().sum().item()}
        out["grads"] = grads
        full = ("BACKBONE.base.cls_token", "BACKBONE.base.blocks.0.norm1.weight", "BACKBONE.base.blocks.11.mlp.fc2.bias",
                "FUSE_block.out_norm.weight", "FUSE_block.normR.bias", "RGB_REDUCE.bias", "BACKBONE.base.sie_embed")
        out["grads_full"] = {k: dict(model.named_parameters())[k].grad.clone() for k in full}
        st = model.state_dict()
        out["state_after"] = {k: st[k].clone() for k in st if ("running" in k or "centers" in k or "tracked" in k)}
        # keep only the rows of the centres that changed (C x 768 x 3 is too large to commit otherwise)
        for k in list(out["state_after"]):
            if "centers" in k:
                out["state_after"][k] = out["state_after"][k][label.unique()].clone()
    finally:
        torch.topk = _native_topk
    out["meta"] = {"case": name, "cuda_topk_rule": cuda_rule, "B": B, "H": H, "W": W, "C": C, "cams": cams,
                   "al": al, "instances": inst, "weights_seed": 1111, "batch_seed": 1, "torch": torch.__version__}
    return out


def main():
    torch.set_num_threads(8)
    for name in CASES:
        g = run_case(name, cuda_rule=True)
        native = run_case(name, cuda_rule=False)
        g["native_cpu_topk_index"] = native["train_index"]
        g["native_cpu_topk_eval_cls4t"] = native["eval_cls4t"]
        path = os.path.join(HERE, "ref_%s.pt" % name)
        torch.save(g, path)
        print(name, "saved", os.path.getsize(path), "bytes; loss", float(g["loss"]),
              "selected/sample", g["train_index"].sum(1).tolist(),
              "rows differing under CPU topk:", int((g["train_index"] != native["train_index"]).any(1).sum()))


if __name__ == "__main__":
    main()
