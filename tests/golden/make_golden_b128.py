"""Full-size eval goldens (BASELINE.json batch 128) from the UNMODIFIED reference on CPU, both orientations:

    python tests/golden/make_golden_b128.py        # build container only (needs /root/reference); ~1 min per case

Stores, per case, the reference's eval feature `cls4t` [128, 2304] as float16-free fp32 (1.2 MB), the selection index
[128, 128] bool and the frequency mask, for weights `synthetic_state_dict(seed=1111)` and `synthetic_batch(128, seed=1)`.
torch.topk is replaced by the CUDA tie rule measured on B200 (see make_golden.py); rows whose result depends on the tie
order under the native CPU rule are listed so that a reader can see how many there are."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from editor_b200 import synth  # noqa: E402
from oracle import ref_import  # noqa: E402
import make_golden as mg  # noqa: E402

CASES = {"rgbnt201": ("RGBNT201", 171, 4, 256, 128, True), "rgbnt100": ("RGBNT100", 50, 8, 128, 256, False)}


def run(name, cuda_rule):
    ds, C, cams, H, W, al = CASES[name]
    torch.manual_seed(0)
    model, cfg = ref_import.load_reference(ds, C, cams, opts=("MODEL.DROP_PATH", 0.0))
    model.load_state_dict(synth.synthetic_state_dict(seed=1111, num_class=C, camera_num=cams, al=al), strict=True)
    x, label, cam = synth.synthetic_batch(128, H, W, seed=1, num_cams=cams, instances=16)
    grabbed = {}
    orig = model.SFTS.forward

    def hook(*a, **k):
        r = orig(*a, **k)
        grabbed["index"] = r[3].detach().clone()
        return r
    model.SFTS.forward = hook
    torch.topk = mg._cuda_rule_topk if cuda_rule else mg._native_topk
    try:
        model.eval()
        with torch.no_grad():
            feat = model(x, cam_label=cam, view_label=None, mode=1, img_path=None)
    finally:
        torch.topk = mg._native_topk
    return feat.clone(), grabbed["index"][..., 0].clone()


def main():
    torch.set_num_threads(8)
    for name in CASES:
        feat, index = run(name, True)
        _, index_native = run(name, False)
        out = {"eval_cls4t": feat, "eval_index": index,
               "rows_depending_on_tie_order": torch.nonzero((index != index_native).any(1)).flatten(),
               "meta": {"case": name, "B": 128, "weights_seed": 1111, "batch_seed": 1, "torch": torch.__version__,
                        "cuda_topk_rule": True}}
        path = os.path.join(HERE, "ref_b128_%s.pt" % name)
        torch.save(out, path)
        print(name, os.path.getsize(path), "bytes; kept/sample mean", float(index.sum(1).float().mean()),
              "rows depending on tie order:", out["rows_depending_on_tie_order"].tolist())


if __name__ == "__main__":
    main()
