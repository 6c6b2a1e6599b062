"""Why do GradScaler + torch SGD and Trainer.step disagree at B=4?  Prints per-step loss, scale, found-inf, gradient norms."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402
from editor_b200.train import Trainer, editor_loss  # noqa: E402


def fresh():
    model, sd, x, label, cam, _ = ge._small_case(True, 4)
    return model.cuda().train(), {k: v.cuda() for k, v in x.items()}, label.cuda(), cam.cuda()


for scale in (65536.0, 1.0):
    m1, x, label, cam = fresh()
    groups = [{"params": [v], "lr": 0.001 * (2 if "bias" in k else 1), "weight_decay": 1e-4} for k, v in m1.named_parameters()
              if v.requires_grad]
    opt = torch.optim.SGD(groups, momentum=0.9)
    scaler = torch.amp.GradScaler("cuda", init_scale=scale)
    for it in range(2):
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.float16):
            outs = m1(x, label=label, cam_label=cam, writer=None, epoch=1)
            loss = editor_loss(outs, label)
        scaler.scale(loss).backward()
        gn = {k: (p.grad.float().norm().item(), bool(torch.isfinite(p.grad).all())) for k, p in m1.named_parameters() if p.grad is not None}
        bad = [k for k, (n, f) in gn.items() if not f]
        scaler.step(opt)
        scaler.update()
        print("scale %g it %d loss %.4f new_scale %g  nonfinite grads: %s   |g out_norm.bias|/scale %.4e  |g qkv0|/scale %.4e" % (
            scale, it, loss.item(), scaler.get_scale(), bad[:5], gn["FUSE_block.out_norm.bias"][0] / scale,
            gn["BACKBONE.base.blocks.0.attn.qkv.weight"][0] / scale))
m2, x, label, cam = fresh()
tr = Trainer(m2)
for it in range(2):
    loss, _ = tr.step(x, label, cam)
    g = m2.engine().arena
    print("trainer it %d loss %.4f |g out_norm.bias| %.4e |g qkv0| %.4e" % (
        it, loss.item(), g.gview("FUSE_block.out_norm.bias").norm().item(), g.gview("BACKBONE.base.blocks.0.attn.qkv.weight").norm().item()))
