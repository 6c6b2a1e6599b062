"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the optimizer step the reference performs:
solver/make_optimizer.py:6-22 (one SGD group per tensor; names containing "bias" get lr*BIAS_LR_FACTOR and
WEIGHT_DECAY_BIAS) over torch.optim.SGD(momentum) semantics, plus the data-parallel gradient average of DDP
(engine/processor.py:47-50).  Used by tests to check edb_sgd_step and the flat-arena flag logic."""
import torch


def reference_optimizer(named_params, lr=0.001, momentum=0.9, wd=1e-4, wd_bias=1e-4, bias_lr_factor=2.0):
    groups = []
    for name, p in named_params:
        if not p.requires_grad:
            continue
        g_lr, g_wd = lr, wd
        if "bias" in name:
            g_lr, g_wd = lr * bias_lr_factor, wd_bias
        groups.append({"params": [p], "lr": g_lr, "weight_decay": g_wd})
    return torch.optim.SGD(groups, momentum=momentum)


def flat_sgd_step(p, g, buf, flags, lr, momentum, wd, wd_bias, bias_lr_factor, gscale, first):
    """The arithmetic of edb_sgd_step on flat tensors; flags: uint8 per 64-element chunk (bit0 bias, bit1 skip)."""
    f = flags.repeat_interleave(64)[:p.numel()]
    skip = (f & 2) != 0
    isb = (f & 1) != 0
    l = torch.where(isb, torch.full_like(p, lr * bias_lr_factor), torch.full_like(p, lr))
    w = torch.where(isb, torch.full_like(p, wd_bias), torch.full_like(p, wd))
    gg = g * gscale + w * p
    nb = gg if first else momentum * buf + gg
    newp = p - l * nb
    p.copy_(torch.where(skip, p, newp))
    buf.copy_(torch.where(skip, buf, nb))
