"""Worker of tests/test_dist_gpu.py (one process per GPU, NCCL).  Checks, from identical initial weights and per-rank batches:

1. after Trainer.step the flat gradient arena of EVERY bucket holds the SUM over ranks of the per-rank gradients (the
   fused SGD applies 1/world) -- compared with per-rank gradients of a second, untouched model instance that are
   all-reduced here with one plain NCCL call;
2. torch's DistributedDataParallel(find_unused_parameters=True) wrapped around this repo's model (what
   engine/processor.py:47-50 does) + the reference's per-tensor SGD  ==  Trainer (bucketed overlapped allreduce + fused
   SGD) after two steps: parameters, BN running statistics.
Prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402
from editor_b200.train import Trainer, editor_loss  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved = os.dup(1)
    os.dup2(2, 1)                     # NCCL banner -> stderr
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
    os.dup2(saved, 1)
    al = os.environ.get("EDB_TEST_AL", "1") == "1"

    def fresh():
        model, sd, x, label, cam, _ = ge._small_case(al, 4, seed=1 + rank)
        return model.to(dev).train(), sd, {k: v.to(dev) for k, v in x.items()}, label.to(dev), cam.to(dev)

    report = {"world": world, "al": al}
    # ---- 1. bucketed allreduce == sum of per-rank gradients
    m0, sd, x, label, cam = fresh()
    m0.param_grads = "arena"
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = editor_loss(m0(x, label=label, cam_label=cam, writer=None, epoch=1), label)
    loss.backward()
    local_grad = m0.engine().arena.grad.clone()
    dist.all_reduce(local_grad)                         # expected: plain sum over ranks
    m1, _, _, _, _ = fresh()
    tr = Trainer(m1)
    tr.step(x, label, cam)
    got = m1.engine().arena.grad
    worst = {}
    for name, (a, b) in tr.buckets.items():
        ref = local_grad[a:b]
        worst[name] = ((got[a:b] - ref).norm() / ref.norm().clamp_min(1e-20)).item()
    report["bucket_rel_err"] = worst
    # parameters identical on every rank after the step
    flat = m1.engine().arena.flat
    mx, mn = flat.clone(), flat.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    report["params_rank_spread"] = (mx - mn).abs().max().item()
    # ---- 2. DDP + torch SGD == Trainer
    m2, _, _, _, _ = fresh()
    ddp = torch.nn.parallel.DistributedDataParallel(m2, device_ids=[local], find_unused_parameters=True)
    groups = [{"params": [v], "lr": 0.001 * (2 if "bias" in k else 1), "weight_decay": 1e-4}
              for k, v in m2.named_parameters() if v.requires_grad]
    opt = torch.optim.SGD(groups, momentum=0.9)
    m3, _, _, _, _ = fresh()
    tr3 = Trainer(m3)
    def agreement():
        p2, p3 = dict(m2.named_parameters()), dict(m3.named_parameters())
        num = den = 0.0
        per = []
        for k in p2:
            if p2[k].grad is None:
                continue
            upd = (p3[k].detach().cpu() - sd[k]).double().norm().item()
            diff = (p2[k].detach() - p3[k].detach()).double().norm().item()
            num, den = num + diff ** 2, den + upd ** 2
            if upd > 1e-12:
                per.append(diff / upd)
        per.sort()
        return {"whole_model": (num / max(den, 1e-300)) ** 0.5, "per_tensor_median": per[len(per) // 2], "per_tensor_max": per[-1]}

    report["ddp_vs_trainer_param_err_rel_update"] = []
    for _ in range(2):
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            l2 = editor_loss(ddp(x, label=label, cam_label=cam, writer=None, epoch=1), label)
        l2.backward()
        opt.step()
        l3, _ = tr3.step(x, label, cam)
        torch.cuda.synchronize()
        report["ddp_vs_trainer_param_err_rel_update"].append(agreement())
    report["ddp_loss"], report["trainer_loss"] = l2.item(), l3.item()
    a, b = m2.state_dict()["FUSE_BN.running_mean"], m3.state_dict()["FUSE_BN.running_mean"]
    report["bn_running_mean_err"] = ((a - b).abs().max() / b.abs().max()).item()
    if rank == 0:
        print(json.dumps(report), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
