"""The drop-in claim, executed: the reference's UNMODIFIED ``engine/processor.py::do_train`` (GradScaler, fp16
``amp.autocast``, one ``torch.optim.SGD`` group per tensor reading ``p.grad``, ``loss.item()``; :23-120) drives this repo's
``modeling.make_model`` on the GPU, next to the reference model driven by the same loop.  The reference sources come from
the git-ignored copy ``baseline/_ref`` (``baseline/install_ref.py``); where it is absent these tests skip.

Also covers the ``.grad`` contract the unchanged callers rely on: accumulation over micro-batches,
``zero_grad(set_to_none=False)``, foreign ``.grad`` tensors, shared heads called three times per step."""
import json
import os
import subprocess
import sys

import pytest
import torch

import __graft_entry__ as ge
from editor_b200.train import Trainer

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "engine"))


def _run_ref(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "run_ref.py"), *args], capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref (copy of the unmodified reference) not present")
@pytest.mark.parametrize("config", ["RGBNT201", "RGBNT100"])
def test_unmodified_do_train_drives_this_model(config, tmp_path):
    common = ["--config", config, "--batch", "16", "--steps", "2", "--warmup", "1", "--drop-path", "0.0"]
    ours = _run_ref("--model", "ours", "--amp", "fp16", *common)
    assert ours["imports"]["engine"].startswith("baseline/_ref") and not ours["imports"]["modeling"].startswith("baseline")
    assert ours["params_finite"] and ours["kernel_launches_c_abi"] > 1000
    ref = _run_ref("--model", "reference", "--amp", "bf16", *common)
    assert ref["imports"]["modeling"].startswith("baseline/_ref")
    print("do_train losses  ours:", ours["losses"], " reference(bf16 autocast):", ref["losses"])
    print("kept tokens      ours:", ours["num_count"], " reference:", ref["num_count"])
    assert len(ours["losses"]) == len(ref["losses"]) == 3
    a0, b0 = ours["losses"][0], ref["losses"][0]
    assert abs(a0 - b0) < 1e-2 * abs(b0), (ours["losses"], ref["losses"])      # tolerance: 1e-2 bf16 (same weights, same batch)
    # later iterations compare two bf16 TRAININGS: the first SGD steps change the loss by 60 %, so a gradient that differs
    # at the bf16 noise level (2-3 %, tests/test_bf16_spread_gpu.py) moves the next loss by a few percent of that change
    for a, b in zip(ours["losses"][1:], ref["losses"][1:]):
        assert abs(a - b) < 5e-2 * abs(b0 - b) + 1e-2 * abs(b), (ours["losses"], ref["losses"])
    for a, b in zip(ours["num_count"], ref["num_count"]):
        assert abs(a - b) < 1.0                                                 # mean kept tokens per sample


def _update_agreement(p1, p2, sd0):
    """Two runs of the same training from the same start: distance between the parameters relative to the distance
    travelled -- over the whole model, and the per-tensor median.  (Per-tensor maxima are meaningless here: tensors whose
    gradient is a near-cancelling sum, e.g. FUSE_block.out_norm.bias with |g| ~ 2e-3 next to |g| ~ 3e2 of the qkv weights,
    differ by 5 % between two runs of IDENTICAL arithmetic -- fp32 atomics reorder the sums; tools/scaler_probe.py.)"""
    num = den = 0.0
    per = []
    for k in p1:
        if p1[k].grad is None:
            continue
        start = sd0[k].to(p1[k].device)
        upd = (p1[k].detach() - start).double().norm().item()
        diff = (p1[k].detach() - p2[k].detach()).double().norm().item()
        num += diff ** 2
        den += upd ** 2
        if upd > 1e-12:
            per.append(diff / upd)
    per.sort()
    return (num / max(den, 1e-300)) ** 0.5, per[len(per) // 2], per[-1]


def _grads(model):
    return {k: p.grad.detach().float().clone() for k, p in model.named_parameters() if p.grad is not None}


def _fwd_bwd(model, x, label, cam, scale=1.0):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = model(x, label=label, cam_label=cam, writer=None, epoch=1)
        from editor_b200.train import editor_loss
        loss = editor_loss(outs, label) * scale
    loss.backward()
    return loss


@pytest.mark.parametrize("al", [True, False])
def test_grad_accumulates_like_torch(al):
    """fwd/bwd, fwd/bwd without zeroing == sum of the two gradients; zero_grad() in both flavours starts over; a foreign
    .grad tensor is honoured once (not once per contributor: BACKBONE_HEAD / BACKBONE_BN are called three times, AL=0)."""
    model, sd, x, label, cam, _ = ge._small_case(al, 4)
    model = model.cuda().train()
    xa = {k: v.cuda() for k, v in x.items()}
    xb = {k: v.cuda().flip(0).contiguous() for k, v in x.items()}
    lab, cg = label.cuda(), cam.cuda()
    bn = {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "centers" in k or "tracked" in k}

    def reset_state():
        model.load_state_dict(bn, strict=False)

    _fwd_bwd(model, xa, lab, cg)
    ga = _grads(model)
    model.zero_grad(set_to_none=True)
    reset_state()
    _fwd_bwd(model, xb, lab, cg)
    gb = _grads(model)
    # (1) accumulation: no zeroing between two backward passes
    model.zero_grad(set_to_none=False)              # in-place zeroing of the arena views
    assert all(float(p.grad.abs().max()) == 0.0 for p in model.parameters() if p.grad is not None)
    reset_state()
    _fwd_bwd(model, xa, lab, cg)
    reset_state()
    _fwd_bwd(model, xb, lab, cg)
    gab = _grads(model)
    assert set(gab) == set(ga) == set(gb)
    worst = 0.0
    for k in ga:
        want = ga[k] + gb[k]
        err = ((gab[k] - want).norm() / want.norm().clamp_min(1e-12)).item()
        if want.norm() > 1e-6:
            worst = max(worst, err)
    print("accumulated-gradient error vs g(a)+g(b):", worst)
    assert worst < 2e-3         # split-K atomics reorder fp32 sums; OCFR centres differ between (a) alone and (a) after (b)
    # (2) a foreign .grad (set by the caller) is adopted once
    model.zero_grad(set_to_none=True)
    name = "BACKBONE_HEAD.weight" if not al else "AL_HEAD.weight"
    p = dict(model.named_parameters())[name]
    p.grad = torch.ones_like(p)
    reset_state()
    _fwd_bwd(model, xa, lab, cg)
    got = p.grad.detach().float()
    want = ga[name] + 1.0
    assert ((got - want).norm() / want.norm()).item() < 2e-3


def test_gradscaler_torch_sgd_equals_fused_trainer():
    """Two iterations of the reference's loop body (GradScaler + per-tensor torch SGD with bias lr x2,
    solver/make_optimizer.py:6-22) == two Trainer.step() calls (fused arena SGD) from the same initial state."""
    def fresh():
        model, sd, x, label, cam, _ = ge._small_case(True, 4)
        return model.cuda().train(), {k: v.cuda() for k, v in x.items()}, label.cuda(), cam.cuda()

    m1, x, label, cam = fresh()
    groups = []
    for k, v in m1.named_parameters():
        if v.requires_grad:
            groups.append({"params": [v], "lr": 0.001 * (2 if "bias" in k else 1), "weight_decay": 1e-4})
    opt = torch.optim.SGD(groups, momentum=0.9)
    scaler = torch.amp.GradScaler("cuda")
    from editor_b200.train import editor_loss
    for _ in range(2):
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.float16):           # processor.py:79: fp16 autocast; the engine computes bf16
            outs = m1(x, label=label, cam_label=cam, writer=None, epoch=1)
            loss = editor_loss(outs, label)
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
    m2, x, label, cam = fresh()
    tr = Trainer(m2)
    for _ in range(2):
        tr.step(x, label, cam)
    torch.cuda.synchronize()
    p1, p2 = dict(m1.named_parameters()), dict(m2.named_parameters())
    sd0 = ge._small_case(True, 4)[1]
    glob, med, worst = _update_agreement(p1, p2, sd0)
    print("GradScaler + torch SGD vs fused Trainer after 2 steps: whole-model %.3e, per-tensor median %.3e, max %.3e" % (glob, med, worst))
    assert glob < 1e-2 and med < 2e-2, (glob, med, worst)     # tolerance: fp32 atomics + bf16 re-rounding of the 2nd step
    #                                                           (measured 2.1e-3 / 2.6e-3; one step alone agrees to 1e-6,
    #                                                           tools/scaler_probe.py)
    for k in ("FUSE_BN.running_mean", "FUSE_block.memory_cls.RGB_centers"):
        a, b = m1.state_dict()[k], m2.state_dict()[k]
        assert ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item() < 5e-3       # second step: bf16 re-rounding


def test_graphed_step_equals_eager_step():
    """Trainer.capture: the whole training step replayed as one CUDA graph == the same steps launched eagerly (parameters,
    BatchNorm statistics, OCFR centres after 3 steps on changing batches; DROP_PATH 0 so that no RNG stream is involved)."""
    def fresh(seed):
        model, sd, x, label, cam, _ = ge._small_case(False, 4, seed=seed)
        return model.cuda().train(), sd, {k: v.cuda() for k, v in x.items()}, label.cuda(), cam.cuda()

    batches = [fresh(s)[2:] for s in (1, 2, 3)]
    m1, sd0, *_ = fresh(1)
    t1 = Trainer(m1)
    m2, _, x, label, cam = fresh(1)
    t2 = Trainer(m2)
    assert t2.capture(x, label, cam, warmup=2), getattr(t2, "capture_error", None)
    m2.load_state_dict(sd0, strict=True)            # undo the warm-up / capture steps: same start as m1
    t2.mom.zero_()                                    # (captured with first=False: momentum buffer zero -> same arithmetic)
    m2.engine().arena.refresh16(force=True)
    p1, p2 = dict(m1.named_parameters()), dict(m2.named_parameters())
    for i, b in enumerate(batches):
        l1, _ = t1.step(*b)
        l2, _ = t2.step_graphed(*b)
        torch.cuda.synchronize()
        glob, med, worst = _update_agreement(p1, p2, sd0)
        print("graphed vs eager after %d step(s): loss %.5f / %.5f, whole-model %.3e, per-tensor median %.3e, max %.3e"
              % (i + 1, l1.item(), l2.item(), glob, med, worst))
        # one step: the same kernels on the same data, only the order of the fp32 atomics differs; later steps compare
        # two bf16 trainings whose weights already differ in the last bits (activations re-round differently)
        # (measured: step 2-3 whole-model 4.3e-3; the bar leaves room for another realisation of the atomics' order)
        assert glob < (2e-4 if i == 0 else 3e-2) and med < (1e-3 if i == 0 else 3e-2), (i, glob, med, worst)
        assert abs(l1.item() - l2.item()) < (1e-5 if i == 0 else 5e-3) * abs(l1.item())
    for k in ("FUSE_BN.running_var", "BACKBONE_BN.running_mean", "FUSE_block.memory_cls.TIR_centers"):
        a, b = m1.state_dict()[k], m2.state_dict()[k]
        assert ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item() < 5e-3, k
