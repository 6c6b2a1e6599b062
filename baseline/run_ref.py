#!/usr/bin/env python
"""Measured denominator of the north-star targets: the UNMODIFIED reference training loop on the GPU.

    python baseline/run_ref.py --model reference --amp fp16 ...   # the reference model, as engine/processor.py runs it
    python baseline/run_ref.py --model reference --amp bf16 ...   # same loop, `amp.autocast` of its namespace -> bf16
    python baseline/run_ref.py --model ours --amp fp16 ...        # THIS repo's make_model driven by that same loop

Everything on the timed path is the reference's own code from the git-ignored copy ``baseline/_ref`` (made by
``baseline/install_ref.py``): ``engine/processor.py::do_train`` (:23-215, unmodified: ``amp.autocast`` + ``GradScaler`` +
``loss.item()`` twice + ``torch.cuda.synchronize()`` per iteration), ``layers/make_loss.py::make_loss`` (:12-81) and
``solver/make_optimizer.py::make_optimizer`` (:4-28; one SGD parameter group per tensor).  What this script supplies is
only what ``train_net.py`` would: a cfg, a model from ``modeling.make_model`` (the reference's, or this repo's drop-in
shim with ``--model ours``), a loader of synthetic P x K batches held in pinned host memory, and a constant-lr scheduler.
``do_train`` synchronises the device at the end of every iteration (:107), so the wall-clock time between two
``next(loader)`` calls is the duration of one full training iteration including the host->device copy of the batch.

Stubs: ``yacs`` (absent from the image; config/defaults.py:1 only constructs and assigns),
``pywt`` (absent; Haar taps only), ``matplotlib`` / ``seaborn`` (dead visualisation imports); ``SummaryWriter`` is
replaced in ``engine.processor``'s namespace by an in-memory writer (do_train hard-codes a log directory under ``/``).

Prints ONE JSON object on stdout.
"""
import argparse
import json
import logging
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")

CASES = {   # yml -> (num_class, cameras): SURVEY.md section 8 preamble
    "RGBNT201": (171, 4), "RGBNT100": (50, 8), "MSVR310": (155, 8),
}


class MemWriter:
    last = None

    def __init__(self, *a, **k):
        self.scalars = []
        MemWriter.last = self

    def add_scalar(self, tag, value, step=None):
        self.scalars.append((tag, float(value), step))

    def close(self):
        pass


class ConstScheduler:
    def __init__(self, lr):
        self.lr = lr

    def step(self, epoch=None):
        pass

    def _get_lr(self, epoch=None):
        return [self.lr]


class SyntheticLoader:
    """Iterable over `n` identical-shape P x K batches in pinned host memory (what a DataLoader with pin_memory=True
    hands to do_train, data/datasets/make_dataloader.py:276-283); records a timestamp per batch request."""

    def __init__(self, batches, n, batch_size, resident_device=None):
        self.batches, self.n, self.batch_size = batches, n, batch_size
        self.stamps = []
        self.dev = resident_device

    def __len__(self):
        return self.n

    def __iter__(self):
        for i in range(self.n):
            self.stamps.append(time.perf_counter())
            x, label, cam = self.batches[i % len(self.batches)]
            yield x, label, cam, cam.clone().zero_(), tuple("synthetic_%d" % j for j in range(self.batch_size))
        self.stamps.append(time.perf_counter())


def setup_imports(which):
    """sys.path so that `engine`, `layers`, `solver`, `utils` are ALWAYS the reference's and `modeling` / `config` are the
    reference's (--model reference) or this repo's drop-in shims (--model ours)."""
    sys.path.insert(0, ROOT)
    from baseline import stubs               # yacs / pywt / matplotlib / seaborn stand-ins
    stubs.install_stubs()
    sys.path.remove(ROOT)
    if which == "reference":
        sys.path.insert(0, REF)
        sys.path.insert(1, ROOT)             # editor_b200.synth (seeded weights) only
    else:
        sys.path.insert(0, ROOT)             # modeling/, config/ -> editor_b200 shims
        sys.path.insert(1, REF)              # engine/, layers/, solver/, utils/ -> the reference


def run(args):
    import torch
    setup_imports(args.model)
    import config as config_pkg
    import modeling as modeling_pkg
    from engine import processor
    from layers.make_loss import make_loss
    from solver.make_optimizer import make_optimizer
    origin = {"modeling": os.path.relpath(modeling_pkg.__file__, ROOT), "config": os.path.relpath(config_pkg.__file__, ROOT),
              "engine": os.path.relpath(processor.__file__, ROOT)}
    assert origin["engine"].startswith("baseline/_ref"), origin
    assert origin["modeling"].startswith("baseline/_ref") == (args.model == "reference"), origin
    from editor_b200 import synth
    processor.SummaryWriter = MemWriter
    logging.getLogger("EDITOR.train").addHandler(logging.StreamHandler(sys.stderr))

    C, cams = CASES[args.config]
    cfg = config_pkg.cfg.clone()
    cfg_root = REF if args.model == "reference" else ROOT
    cfg.merge_from_file(os.path.join(cfg_root, "configs", args.config, "EDITOR.yml"))
    cfg.merge_from_list(["MODEL.PRETRAIN_CHOICE", "none", "SOLVER.MAX_EPOCHS", 1, "SOLVER.CHECKPOINT_PERIOD", 10 ** 9,
                         "SOLVER.EVAL_PERIOD", 10 ** 9, "SOLVER.LOG_PERIOD", 10 ** 9, "SOLVER.IMS_PER_BATCH", args.batch,
                         "MODEL.DIST_TRAIN", False])
    if args.drop_path is not None:
        cfg.merge_from_list(["MODEL.DROP_PATH", args.drop_path])
    H, W = cfg.INPUT.SIZE_TRAIN
    al = bool(cfg.MODEL.AL)
    torch.manual_seed(cfg.SOLVER.SEED)
    torch.backends.cudnn.allow_tf32 = not args.no_tf32
    torch.backends.cuda.matmul.allow_tf32 = False          # torch default, what the reference runs with
    if args.amp == "bf16":
        # do_train calls `amp.autocast(enabled=True)` (torch.cuda.amp: dtype defaults to float16 in its signature): give its
        # namespace an `amp` whose autocast is bf16 -- the only way to run the UNMODIFIED loop under bf16 autocast
        import types
        processor.amp = types.SimpleNamespace(
            GradScaler=torch.cuda.amp.GradScaler,
            autocast=lambda enabled=True: torch.autocast("cuda", dtype=torch.bfloat16, enabled=enabled))
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model = modeling_pkg.make_model(cfg, num_class=C, camera_num=cams)
        loss_fn, center_criterion = make_loss(cfg, num_classes=C)
        model.load_state_dict(synth.synthetic_state_dict(seed=1111, num_class=C, camera_num=cams, al=al), strict=True)
        optimizer, optimizer_center = make_optimizer(cfg, model, center_criterion)
    inst = 16 if (args.batch % 16 == 0 and args.batch >= 32) else 2      # >= 2 identities: batch-hard triplet needs negatives
    batches = []
    for s in range(args.distinct_batches):
        x, label, cam = synth.synthetic_batch(args.batch, H, W, seed=1 + s, num_cams=cams, instances=inst)
        if args.resident:
            batches.append(({k: v.cuda() for k, v in x.items()}, label.cuda(), cam.cuda()))
        else:
            batches.append(({k: v.pin_memory() for k, v in x.items()}, label.pin_memory(), cam.pin_memory()))
    n = args.warmup + args.steps
    loader = SyntheticLoader(batches, n, args.batch)
    torch.cuda.reset_peak_memory_stats()
    out_stream = io.StringIO()
    with contextlib.redirect_stdout(out_stream):
        processor.do_train(cfg, model, center_criterion, loader, None, optimizer, optimizer_center,
                           ConstScheduler(cfg.SOLVER.BASE_LR), loss_fn, 0, 0)
    torch.cuda.synchronize()
    st = loader.stamps
    per = [(b - a) * 1e3 for a, b in zip(st[:-1], st[1:])]
    timed = per[args.warmup:]
    ms = sum(timed) / len(timed)
    finite = all(torch.isfinite(p).all().item() for p in model.parameters())
    res = {"model": args.model, "loop": "engine/processor.py::do_train (unmodified, baseline/_ref)", "amp": args.amp,
           "config": args.config, "batch": args.batch, "steps": args.steps, "warmup": args.warmup,
           "inputs": "device-resident" if args.resident else "pinned host, H2D inside the iteration",
           "ms_per_step": ms, "ms_each": [round(t, 2) for t in timed], "images_per_sec": args.batch / (ms * 1e-3),
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "params_finite": finite,
           "imports": origin, "torch": torch.__version__, "drop_path": float(cfg.MODEL.DROP_PATH),
           "losses": [v for t, v, _ in MemWriter.last.scalars if t == "Loss"],
           "num_count": [v for t, v, _ in MemWriter.last.scalars if t == "num_count"]}
    if args.dump_state:
        sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
        torch.save(sd, args.dump_state)
    if args.model == "ours":
        from editor_b200 import lib
        res["kernel_launches_c_abi"] = lib.launch_count
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", choices=("reference", "ours"), default="reference")
    ap.add_argument("--amp", choices=("fp16", "bf16"), default="fp16")
    ap.add_argument("--config", default="RGBNT201", choices=sorted(CASES))
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--distinct-batches", type=int, default=2)
    ap.add_argument("--resident", action="store_true", help="keep the batches on the device (no H2D in the iteration)")
    ap.add_argument("--drop-path", type=float, default=None)
    ap.add_argument("--no-tf32", action="store_true", help="cudnn.allow_tf32=False (fp32 parity runs)")
    ap.add_argument("--dump-state", default=None, help="torch.save the state_dict after the last iteration here")
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, "engine")):
        print(json.dumps({"unavailable": "baseline/_ref absent (run baseline/install_ref.py where /root/reference exists)"}))
        return
    print(json.dumps(run(args)), flush=True)


if __name__ == "__main__":
    main()
