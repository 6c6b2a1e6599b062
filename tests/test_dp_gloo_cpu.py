"""Data-parallel host logic on CPU (gloo, world_size 2): flat-arena gradient allreduce + the optimizer arithmetic of
edb_sgd_step must equal the reference recipe -- per-rank backward, gradient average (DDP), torch.optim.SGD with the
reference's parameter groups (solver/make_optimizer.py:6-22)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sgd_oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [("blocks.0.attn.qkv.weight", (96, 32)), ("blocks.0.attn.qkv.bias", (96,)), ("norm.weight", (32,)),
              ("norm.bias", (32,)), ("base.fc.weight", (10, 32)), ("head.weight", (7, 32))]
    return [(n, torch.randn(s, generator=g)) for n, s in shapes]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = _build(0)                                  # replicated parameters
    grads = [[g for _, g in _build(100 + rank + 10 * step)] for step in range(3)]     # per-rank gradients, 3 steps
    # ---- flat arena exactly as engine.Arena / train.Trainer lay it out
    offs, off = {}, 0
    for n, p in params:
        offs[n] = (off, p.numel())
        off += (p.numel() + 63) // 64 * 64
    flat, buf = torch.zeros(off), torch.zeros(off)
    flags = torch.full((off // 64,), 2, dtype=torch.uint8)
    for n, p in params:
        o, k = offs[n]
        flat[o:o + k] = p.flatten()
        flags[o // 64:(o + k + 63) // 64] = 2 if n.startswith("base.fc.") else (1 if "bias" in n else 0)
    for step in range(3):
        garena = torch.zeros(off)
        for (n, _), g in zip(params, grads[step]):
            o, k = offs[n]
            garena[o:o + k] = g.flatten()
        dist.all_reduce(garena, op=dist.ReduceOp.SUM)      # the one collective of the path
        sgd_oracle.flat_sgd_step(flat, garena, buf, flags, 0.001, 0.9, 1e-4, 1e-4, 2.0, 1.0 / world, step == 0)
    if rank == 0:
        torch.save({"flat": flat, "offs": offs}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_sgd_matches_reference_recipe(tmp_path):
    out = str(tmp_path / "r0.pt")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = torch.load(out)
    # reference recipe on one process: average the per-rank gradients, reference optimizer groups
    named = [(n, torch.nn.Parameter(p.clone(), requires_grad=not n.startswith("base.fc."))) for n, p in _build(0)]
    opt = sgd_oracle.reference_optimizer(named)
    for step in range(3):
        per_rank = [[g for _, g in _build(100 + r + 10 * step)] for r in range(world)]
        for i, (n, p) in enumerate(named):
            if p.requires_grad:
                p.grad = sum(per_rank[r][i] for r in range(world)) / world
        opt.step()
    for n, p in named:
        o, k = got["offs"][n]
        assert torch.allclose(got["flat"][o:o + k].view(p.shape), p.detach(), rtol=1e-6, atol=1e-7), n


def test_trainer_flags_and_gradient_buckets():
    """train.Trainer on the real parameter arena (built on CPU): bias chunks get lr x2, the never-used
    BACKBONE.base.fc.* and alignment padding are skipped, and the allreduce buckets partition the arena in the order the
    backward completes them."""
    import __graft_entry__ as ge
    from editor_b200.engine import Arena
    from editor_b200.train import Trainer
    model, *_ = ge._small_case(True, 2)
    arena = Arena(model, torch.device("cpu"))
    t = Trainer(model)
    t._setup(arena)
    flags = t.flags

    def chunk_flags(name):
        o, n, _ = arena.offsets[name]
        return set(flags[o // 64:(o + n + 63) // 64].tolist())
    assert chunk_flags("BACKBONE.base.blocks.3.attn.qkv.weight") == {0}
    assert chunk_flags("BACKBONE.base.blocks.3.attn.qkv.bias") == {1}
    assert chunk_flags("FUSE_BN.bias") == {1} and chunk_flags("FUSE_BN.weight") == {0}
    assert chunk_flags("BACKBONE.base.fc.weight") == {2} and chunk_flags("BACKBONE.base.fc.bias") == {2}
    spans = sorted(t.buckets.values())
    assert spans[0][0] == 0 and spans[-1][1] == arena.total
    assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))          # contiguous, no overlap, no hole
    off = lambda n: arena.offsets[n][0]                                        # noqa: E731
    lo, hi = t.buckets["after_backbone"]
    assert all(lo <= off(n) < hi for n in arena.names if not n.startswith("BACKBONE.base."))
    from editor_b200.engine import GRAD_STAGE_BLOCKS
    assert list(t.buckets) == ["after_backbone"] + ["blocks_from_%d" % l for l in GRAD_STAGE_BLOCKS] + ["rest"]
    top = 12
    for l in GRAD_STAGE_BLOCKS:              # bucket "blocks_from_l" = blocks l .. (previous stage - 1); the first one also
        lo, hi = t.buckets["blocks_from_%d" % l]                               # holds the final norm and the unused fc
        assert all(lo <= off(n) < hi for n in arena.names
                   if any(n.startswith("BACKBONE.base.blocks.%d." % i) for i in range(l, top)))
        top = l
    lo, hi = t.buckets["blocks_from_%d" % GRAD_STAGE_BLOCKS[0]]
    assert lo <= off("BACKBONE.base.norm.weight") < hi
    lo, hi = t.buckets["rest"]
    assert all(lo <= off(n) < hi for n in ("BACKBONE.base.cls_token", "BACKBONE.base.pos_embed",
                                           "BACKBONE.base.patch_embed.proj.weight", "BACKBONE.base.blocks.0.mlp.fc2.bias"))
    # the model still owns its parameters after the arena re-binds their storage
    sd = model.state_dict()
    assert sd["BACKBONE.base.pos_embed"].data_ptr() == arena.view("BACKBONE.base.pos_embed").data_ptr()


def test_arena_grad_contract_accumulates_like_torch():
    """engine.Arena.prepare_grads / finish_grads (host logic, CPU tensors): `.grad` views of the flat arena accumulate over
    backward passes until the caller zeroes them; a foreign `.grad` is adopted once; parameters without a `.grad` carry
    nothing.  (The kernels of one backward WRITE most gradients, so each backward starts from a zeroed arena.)"""
    import __graft_entry__ as ge
    from editor_b200.engine import Arena
    model, *_ = ge._small_case(False, 2)
    arena = Arena(model, torch.device("cpu"))
    names = ["BACKBONE.base.blocks.3.attn.qkv.bias", "FUSE_BN.weight", "BACKBONE_HEAD.weight"]
    params = dict(model.named_parameters())

    def fake_backward(value):
        """what a step's kernels do: overwrite / fill their slices of the (zeroed) arena, then expose them"""
        for n in names:
            arena.gview(n).fill_(value)
        arena.attach_grads(set(names))
        arena.finish_grads()

    # 1. zero_grad(set_to_none=True): every .grad None -> one memset, plain gradients
    arena.grad.fill_(7.0)
    arena.prepare_grads()
    assert float(arena.grad.abs().max()) == 0.0 and not arena.carry_live
    fake_backward(1.0)
    assert all(torch.all(params[n].grad == 1.0) for n in names)
    assert params[names[0]].grad.data_ptr() == arena.gview(names[0]).data_ptr()
    # 2. second backward without zeroing: old + new
    arena.prepare_grads()
    assert arena.carry_live and float(arena.grad.abs().max()) == 0.0
    fake_backward(2.0)
    assert all(torch.all(params[n].grad == 3.0) for n in names)
    # 3. zero_grad(set_to_none=False) zeroes the views in place -> the next backward starts over
    for n in names:
        params[n].grad.zero_()
    arena.prepare_grads()
    fake_backward(5.0)
    assert all(torch.all(params[n].grad == 5.0) for n in names)
    # 4. a foreign .grad tensor is adopted once and re-pointed at the arena; a parameter whose .grad is None carries nothing
    params[names[1]].grad = None
    params[names[2]].grad = torch.full_like(params[names[2]], 10.0)
    arena.prepare_grads()
    fake_backward(1.0)
    assert torch.all(params[names[0]].grad == 6.0)          # 5 carried + 1
    assert torch.all(params[names[1]].grad == 1.0)          # nothing carried
    assert torch.all(params[names[2]].grad == 11.0)         # foreign 10 adopted + 1
    assert params[names[2]].grad.data_ptr() == arena.gview(names[2]).data_ptr()


def _trainer_bucket_worker(rank, world, port, out):
    """The REAL train.Trainer bucket code (Trainer._setup + Trainer._on_grad_stage) on a CPU arena over gloo: the stages
    are reported in the order the engine's backward reports them; afterwards every element of the gradient arena must
    hold the sum over ranks -- covered exactly once, no hole, no double reduction."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as ge
    from editor_b200.engine import Arena, GRAD_STAGE_BLOCKS
    from editor_b200.train import Trainer
    model, *_ = ge._small_case(False, 2)
    arena = Arena(model, torch.device("cpu"))
    eng = model.engine()
    eng.arena = arena                                    # (the engine would build it on the first CUDA forward)
    tr = Trainer(model)
    assert tr.world == world
    tr._setup(arena)
    g = torch.Generator().manual_seed(1000 + rank)
    arena.grad.copy_(torch.randn(arena.total, generator=g))
    local = arena.grad.clone()
    # engine.backbone_backward: "after_backbone" first, then the block stages in descending order, then "rest"
    for stage in ["after_backbone"] + ["blocks_from_%d" % l for l in GRAD_STAGE_BLOCKS] + ["rest"]:
        tr._on_grad_stage(stage)
    assert len(tr.pending) == 2 + len(GRAD_STAGE_BLOCKS)
    for w in tr.pending:
        w.wait()
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    want = sum(gathered)
    if rank == 0:
        torch.save({"max_err": float((arena.grad - want).abs().max()), "total": arena.total}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_trainer_buckets_allreduce_every_element_once(tmp_path):
    out = str(tmp_path / "buckets.pt")
    mp.spawn(_trainer_bucket_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    assert got["total"] > 100_000_000 and got["max_err"] == 0.0        # fp32 sums of two ranks: exact, whatever the bucket
