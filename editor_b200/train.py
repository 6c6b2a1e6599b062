"""Training step around the hot path: loss (layers/make_loss.py:36-56 as called by engine/processor.py:82-92), one
gradient allreduce over the flat arena (the one collective of the path, processor.py:47-50) and the fused SGD-momentum
update (solver/make_optimizer.py:6-22).  Rows f-1 / f-2 of SURVEY.md section 8."""
import torch
import torch.distributed as dist

from . import lib
from .engine import GRAD_STAGE_BLOCKS


from .tail import editor_loss  # noqa: E402,F401  (CUDA kernels: label-smoothed CE + batch-hard soft-margin triplet)


class Trainer:
    """forward -> loss -> backward -> allreduce(arena) -> fused SGD, one process per GPU."""

    def __init__(self, model, lr=0.001, momentum=0.9, weight_decay=1e-4, weight_decay_bias=1e-4, bias_lr_factor=2.0):
        self.model = model
        self.lr, self.momentum, self.wd, self.wd_bias, self.blf = lr, momentum, weight_decay, weight_decay_bias, bias_lr_factor
        self.mom = None
        self.flags = None
        self.first = True
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.graph = None

    def _setup(self, arena):
        if self.mom is not None and self.mom.numel() == arena.total and self.mom.device == arena.flat.device:
            return
        self.mom = torch.zeros_like(arena.flat)
        flags = torch.full((arena.total // 64,), 2, dtype=torch.uint8)        # padding chunks are skipped
        for name in arena.names:
            o, n, _ = arena.offsets[name]
            c0, c1 = o // 64, (o + n + 63) // 64
            unused = name.startswith("BACKBONE.base.fc.")                       # never used by EDITOR (vit_pytorch.py:522)
            flags[c0:c1] = 2 if unused else (1 if "bias" in name else 0)        # make_optimizer.py:12-15
        self.flags = flags.to(arena.flat.device)
        self.first = True
        self.tail = [(n, p) for n, p in zip(arena.names, arena.params)
                     if not (n.startswith("BACKBONE.base.") or n.startswith("FUSE_block."))]
        # gradient buckets = contiguous arena slices in the order the backward completes them (the arena follows
        # named_parameters(): backbone embeddings, blocks 0..11, norm, fc, then FUSE_block and the heads)
        off = lambda n: arena.offsets[n][0]                                    # noqa: E731
        after = off("FUSE_block.normR.weight")
        self.buckets = {"after_backbone": (after, arena.total)}
        hi = after
        for l in GRAD_STAGE_BLOCKS:                                            # descending block indices
            lo = off("BACKBONE.base.blocks.%d.norm1.weight" % l)
            self.buckets["blocks_from_%d" % l] = (lo, hi)
            hi = lo
        self.buckets["rest"] = (0, hi)
        self.pending = []

    def _on_grad_stage(self, stage):
        """Called by the engine during backward: allreduce the finished slice on NCCL's stream while the remaining
        dgrad/wgrad GEMMs keep running (the one collective of the path, overlapped)."""
        if self.world == 1:
            return
        a, b = self.buckets[stage]
        arena = self.model.engine().arena
        self.pending.append(dist.all_reduce(arena.grad[a:b], op=dist.ReduceOp.SUM, async_op=True))

    def step(self, x, label, cam, writer=None, epoch=1):
        model = self.model
        eng = model.engine()
        eng.stats["grad_hook"] = self._on_grad_stage            # also tells the engine that the Trainer owns the exchange
        eng.stats["trainer_owns_grads"] = True                  # ... and the gradient arena: zeroed here, once per step
        eng.stats["pending_backward"] = False
        if eng.arena is not None:
            eng.arena.grad.zero_()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outputs = model(x, label=label, cam_label=cam, view_label=None, img_path=None, writer=writer, epoch=epoch)
            loss = editor_loss(outputs, label)
        arena = model.engine().arena
        self._setup(arena)
        loss.backward()
        for w in self.pending:
            w.wait()
        self.pending = []
        lib.call("edb_sgd_step", arena.flat.data_ptr(), arena.grad.data_ptr(), self.mom.data_ptr(),
                 arena.flat16.data_ptr(), self.flags.data_ptr(), arena.total, self.lr, self.momentum, self.wd,
                 self.wd_bias, self.blf, 1.0 / self.world, int(self.first), lib.stream_ptr())
        arena.generation += 1
        arena._versions = [p._version for p in arena.params]   # the bf16 shadow was rewritten by the optimizer kernel
        self.first = False
        return loss, outputs

    # ------------------------------------------------------------------ whole step as ONE CUDA graph
    def capture(self, x, label, cam, warmup=3):
        """Capture forward + loss + backward + bucketed allreduce + fused SGD into one CUDA graph (static input buffers):
        ~700 kernel launches, the autograd bookkeeping and the ctypes calls of a step leave the critical path, the
        launch gaps between the small tail / loss kernels close, and host jitter can no longer stall the GPU.  The bf16
        step has no host synchronisation and no data-dependent launch configuration (packed row counts stay on the
        device), which is what makes it capturable.  `warmup` eager steps run first on a side stream (workspace
        allocation, first-step flags, NCCL communicator).  Returns False (and stays eager) if capture is impossible."""
        model = self.model
        if getattr(model, "precision", "auto") == "fp32":
            return False                    # the fp32 parity mode reads row counts on the host
        self.static_in = ({k: v.clone() for k, v in x.items()}, label.clone(), cam.clone())
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(*self.static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        # no autograd graph of an earlier step may survive into the capture: the AccumulateGrad nodes of the parameters are
        # cached while any graph references them and carry the stream they were created on; the autograd engine would
        # sync the capture stream with that uncaptured stream (cudaErrorStreamCaptureIsolation)
        import gc
        gc.collect()
        n0 = lib.launch_count
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss, outs = self.step(*self.static_in)
                self.static_loss, self.static_out = loss.detach(), tuple(o.detach() for o in outs)
                del loss, outs
        except Exception as e:          # noqa: BLE001 - any capture failure leaves the trainer in eager mode
            self.graph, self.capture_error = None, repr(e)[:300]
            torch.cuda.synchronize()
            return False
        self.graph = g
        self.launches_per_replay = lib.launch_count - n0
        return True

    def step_graphed(self, x, label, cam):
        """Replay the captured step on new inputs (device tensors; copied into the static buffers on the current stream)."""
        if self.graph is None:
            return self.step(x, label, cam)
        sx, sl, sc = self.static_in
        for k in sx:
            if sx[k].data_ptr() != x[k].data_ptr():
                sx[k].copy_(x[k], non_blocking=True)
        if sl.data_ptr() != label.data_ptr():
            sl.copy_(label, non_blocking=True)
        if sc.data_ptr() != cam.data_ptr():
            sc.copy_(cam, non_blocking=True)
        self.graph.replay()
        lib.launch_count += self.launches_per_replay
        return self.static_loss, self.static_out
