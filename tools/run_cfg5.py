#!/usr/bin/env python
"""BASELINE.json configs[4]: MSVR310 EDITOR.yml 3-modal train loop (triplet + ID loss), batch 128 per rank, fp32, N ranks,
selection mask compared with the reference every 50 iterations.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_cfg5.py --iters 200 --every 50 > gpurun_out/cfg5.json

Both sides run in the SAME process per rank, on the same per-rank shards, from the same seeded weights:

  reference   the UNMODIFIED reference (baseline/_ref): modeling.make_model + layers.make_loss + solver.make_optimizer,
              trained by its own engine/processor.py::do_train, which wraps the model in DistributedDataParallel
              (find_unused_parameters=True) because cfg.MODEL.DIST_TRAIN is set (:47-50).  do_train's `amp` namespace is
              replaced by a disabled autocast / GradScaler so that the loop runs in fp32 (TF32 off) as configs[4] asks.
  ours        editor_b200.modeling.make_model in its fp32-faithful mode driven by editor_b200.train.Trainer
              (bucketed NCCL allreduce over the gradient arena + fused SGD).

The loader handed to do_train is a generator: when do_train asks for batch i+1 the reference has finished iteration i, and
the generator then runs OUR iteration i on the same batch.  Two comparisons per checkpoint iteration k (0, 50, 100, ...):

  trajectory      index_k of our own run vs the reference's index_k: both trainings have gone through k SGD steps of fp32
                  arithmetic in different summation orders, so near-tied rollout scores may flip; the number of differing
                  bits (of 128 x 128 per rank) is REPORTED;
  teacher-forced  a third instance of our model loaded with the reference's state_dict as it was BEFORE its iteration k,
                  run on batch k: its index must equal the reference's index_k BIT FOR BIT, and its fused feature
                  cls4t (training output [1]) agrees to 1e-3 -- this is the kernel-parity statement "selection-mask bit-exact vs the
                  reference every 50 iters", free of trajectory drift.

DROP_PATH is 0 on both sides (the fp32-faithful backward does not implement DropPath, and seeded DropPath streams could
not agree anyway: the reference draws per modality call, engine._droppath).  Rank 0 prints one JSON report.
"""
import argparse
import contextlib
import io
import json
import logging
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "baseline", "_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="MSVR310")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--every", type=int, default=50)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--distinct-batches", type=int, default=8)
    ap.add_argument("--classes", type=int, default=155)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved = os.dup(1)
    os.dup2(2, 1)                      # NCCL banner -> stderr
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
    os.dup2(saved, 1)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    sys.path.insert(0, ROOT)
    from baseline import stubs
    stubs.install_stubs()
    from editor_b200 import synth
    from editor_b200.config import cfg as own_cfg
    from editor_b200.modeling import make_model as own_make_model
    from editor_b200.train import Trainer
    sys.path.insert(0, REF)            # `modeling`, `config`, `engine`, `layers`, `solver`, `utils` -> the reference
    for name in list(sys.modules):
        if name.split(".")[0] in ("modeling", "config"):
            del sys.modules[name]
    import config as ref_config_pkg
    import modeling as ref_modeling_pkg
    from engine import processor
    from layers.make_loss import make_loss
    from solver.make_optimizer import make_optimizer
    assert os.path.realpath(ref_modeling_pkg.__file__).startswith(os.path.realpath(REF))
    assert os.path.realpath(processor.__file__).startswith(os.path.realpath(REF))

    class MemWriter:
        def __init__(self, *a, **k):
            pass

        def add_scalar(self, *a, **k):
            pass

        def close(self):
            pass
    processor.SummaryWriter = MemWriter
    processor.amp = types.SimpleNamespace(GradScaler=lambda: torch.amp.GradScaler("cuda", enabled=False),
                                          autocast=lambda enabled=True: torch.autocast("cuda", enabled=False))
    logging.getLogger("EDITOR.train").addHandler(logging.StreamHandler(sys.stderr))

    C, cams = args.classes, 8
    opts = ["MODEL.PRETRAIN_CHOICE", "none", "MODEL.DROP_PATH", 0.0, "MODEL.DIST_TRAIN", True, "SOLVER.MAX_EPOCHS", 1,
            "SOLVER.CHECKPOINT_PERIOD", 10 ** 9, "SOLVER.EVAL_PERIOD", 10 ** 9, "SOLVER.LOG_PERIOD", 10 ** 9,
            "SOLVER.IMS_PER_BATCH", args.batch * world]
    rcfg = ref_config_pkg.cfg.clone()
    rcfg.merge_from_file(os.path.join(REF, "configs", args.config, "EDITOR.yml"))
    rcfg.merge_from_list(opts)
    ocfg = own_cfg.clone()
    ocfg.merge_from_file(os.path.join(ROOT, "configs", args.config, "EDITOR.yml"))
    ocfg.merge_from_list(opts)
    H, W = rcfg.INPUT.SIZE_TRAIN
    al = bool(rcfg.MODEL.AL)
    sd = synth.synthetic_state_dict(seed=1111, num_class=C, camera_num=cams, al=al)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = ref_modeling_pkg.make_model(rcfg, num_class=C, camera_num=cams)
        loss_fn, center_criterion = make_loss(rcfg, num_classes=C)
        ref.load_state_dict(sd, strict=True)
        optimizer, optimizer_center = make_optimizer(rcfg, ref, center_criterion)
        own = own_make_model(ocfg, C, cams)
        own.load_state_dict(sd, strict=True)
        probe = own_make_model(ocfg, C, cams)
    own = own.to(dev).train()
    own.precision = "fp32"
    # eval mode: no per-layer activations are kept (memory); the selection and cls4t are the same functions of weights and
    # inputs in both modes (DROP_PATH 0, no dropout; BatchNorm is applied after cls4t)
    probe = probe.to(dev).eval()
    probe.precision = "fp32"
    trainer = Trainer(own, lr=rcfg.SOLVER.BASE_LR, momentum=rcfg.SOLVER.MOMENTUM, weight_decay=rcfg.SOLVER.WEIGHT_DECAY,
                      weight_decay_bias=rcfg.SOLVER.WEIGHT_DECAY_BIAS, bias_lr_factor=rcfg.SOLVER.BIAS_LR_FACTOR)

    grabbed = {}
    orig = ref.SFTS.forward

    def hook(*a, **k):
        r = orig(*a, **k)
        grabbed["index"] = r[3].detach()[..., 0].clone()
        return r
    ref.SFTS.forward = hook
    ref_out = {}
    orig_fwd = ref.forward

    def fwd_hook(*a, **k):
        r = orig_fwd(*a, **k)
        ref_out["outs"] = [t.detach().float().clone() for t in r]
        return r
    ref.forward = fwd_hook

    # per-rank shards: `distinct` P x K batches (8 identities x 16), different identities per batch, cycled
    batches = []
    for j in range(args.distinct_batches):
        ids = [(rank * 8 * args.distinct_batches + 8 * j + t) % C for t in range(args.batch // 16)]
        x, label, cam = synth.synthetic_batch(args.batch, H, W, seed=10_000 * (rank + 1) + j, num_cams=cams, ids=ids, instances=16)
        batches.append(({k: v.to(dev) for k, v in x.items()}, label.to(dev), cam.to(dev)))

    def bits(idx):
        idx = idx.cpu()
        return ((idx.view(-1, 4, 1) >> torch.arange(32).view(1, 1, 32)) & 1).bool().reshape(-1, 128)

    def rel(a, b):
        return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()

    report = {"checkpoints": [], "config": args.config, "world": world, "batch_per_rank": args.batch, "iters": args.iters,
              "precision": "fp32 (reference: torch fp32, TF32 off; ours: 3-piece bf16 split GEMMs on tcgen05, fp32 attention)",
              "drop_path": 0.0}
    times = {"ref": [], "own": []}

    class Loader:
        batch_size = args.batch * world

        def __len__(self):
            return args.iters

        def __iter__(self):
            t_prev = None
            for i in range(args.iters):
                x, label, cam = batches[i % len(batches)]
                ck = i % args.every == 0
                state_before = {k: v.detach().clone() for k, v in ref.state_dict().items()} if ck else None
                t_prev = time.perf_counter()
                yield x, label, cam, torch.zeros_like(cam), tuple("s%d" % j for j in range(args.batch))
                # ---- the reference has finished iteration i (do_train synchronises every iteration)
                times["ref"].append(time.perf_counter() - t_prev)
                t0 = time.perf_counter()
                trainer.step(x, label, cam)
                torch.cuda.synchronize()
                times["own"].append(time.perf_counter() - t0)
                if not ck:
                    continue
                own_idx = bits(own.engine().sel["index"])
                ref_idx = grabbed["index"].cpu()
                probe.load_state_dict(state_before, strict=True)
                with torch.no_grad():
                    pout = probe(x, cam_label=cam)                  # cls4t [B, 2304] == training output [1]
                pidx = bits(probe.engine().sel["index"])
                outs_err = [rel(pout.float().cpu(), ref_out["outs"][1].cpu())]
                entry = {"iteration": i, "rank": rank,
                         "teacher_forced_bits_differing": int((pidx != ref_idx).sum()),
                         "teacher_forced_cls4t_rel_err": max(outs_err),
                         "trajectory_bits_differing": int((own_idx != ref_idx).sum()),
                         "trajectory_samples_differing": int((own_idx != ref_idx).any(1).sum()),
                         "kept_tokens_mean_ref": float(ref_idx.sum(1).float().mean())}
                report["checkpoints"].append(entry)

    loader = Loader()
    with contextlib.redirect_stdout(io.StringIO()):
        processor.do_train(rcfg, ref, center_criterion, loader, None, optimizer, optimizer_center,
                           types.SimpleNamespace(step=lambda e=None: None, _get_lr=lambda e=None: [rcfg.SOLVER.BASE_LR]),
                           loss_fn, 0, local)
    torch.cuda.synchronize()
    # final parameters: relative distance between the two trainings, in units of the distance travelled
    p_ref = dict(ref.named_parameters())
    drift = []
    for k, p in own.named_parameters():
        if p.grad is None or k not in p_ref:
            continue
        moved = (p_ref[k].detach() - sd[k].to(dev)).norm().item()
        if moved > 1e-9:
            drift.append(((p.detach() - p_ref[k].detach()).norm().item() / moved, k))
    drift.sort()
    report["final_param_distance_over_distance_travelled"] = {"median": drift[len(drift) // 2][0], "max": drift[-1][0],
                                                              "argmax": drift[-1][1]}
    report["ms_per_iter"] = {"reference_ddp_fp32": 1e3 * sum(times["ref"][5:]) / max(len(times["ref"][5:]), 1),
                             "ours_trainer_fp32": 1e3 * sum(times["own"][5:]) / max(len(times["own"][5:]), 1)}
    gathered = [None] * world
    dist.all_gather_object(gathered, report["checkpoints"])
    if rank == 0:
        report["checkpoints"] = sorted([e for g in gathered for e in g], key=lambda e: (e["iteration"], e["rank"]))
        tf_bad = sum(e["teacher_forced_bits_differing"] for e in report["checkpoints"])
        report["teacher_forced_bit_exact"] = "%d/%d (rank, checkpoint) pairs" % (
            sum(e["teacher_forced_bits_differing"] == 0 for e in report["checkpoints"]), len(report["checkpoints"]))
        report["teacher_forced_total_bits_differing"] = tf_bad
        print(json.dumps(report), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
