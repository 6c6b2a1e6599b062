"""Host-side orchestration of the EDITOR hot path over the C ABI (include/editor_b200.h).

Python here is plumbing only: it owns device memory (a flat parameter / gradient arena and named workspace buffers),
sequences kernel launches on the current CUDA stream and wires the result into autograd.  All arithmetic of the path --
3-stream ViT-B/16 backbone, SFTS selection, HMA fusion, forward and backward -- runs in hand-written sm_100a kernels.
The small [B, 768..2304] tail (REDUCE linears, BNNeck heads, OCFR) is listed in DESIGN.md as the next rows to move.

Reference call sites: modeling/make_model.py:150-258 (EDITOR.forward), modeling/backbones/vit_pytorch.py:623-644
(Trans.forward), :215-220 (Block), :309-352 (BlockMask), modeling/fusion_part/SFTS.py:145-230,
modeling/fusion_part/Frequency.py:42-84, modeling/fusion_part/OCFR.py:22-84.
"""
import ctypes

import torch

from . import lib
from . import tail as _tail

DIM, HEADS, HID, NTOK, NPATCH = 768, 12, 3072, 129, 128
P_LD = 136                      # pitch of the stored attention maps (129 padded to a 16-byte multiple)
MAX_SEL = 82                    # <= 24 per modality x 3 + FREQUENCY_KEEP 10 with the default config
SCALE = 64 ** -0.5
BF16, FP32 = "bf16", "fp32"
# the backward reports a finished slice of the gradient arena after these backbone blocks (bucketed allreduce, train.py):
# every two blocks, so that the last bucket ("rest" = blocks 1-0 + embeddings, 60 MB) is all that cannot overlap
GRAD_STAGE_BLOCKS = (10, 8, 6, 4, 2)


def _align(n, a=64):
    return (n + a - 1) // a * a


class _Lin:
    """One nn.Linear (or the patch conv) as GEMM operands: W is [out, in] row-major, i.e. K-major B operand."""

    def __init__(self, eng, wname, bname):
        self.eng, self.wname, self.bname = eng, wname, bname
        a = eng.arena
        w = a.view(wname)
        self.out_f, self.in_f = w.shape[0], w.numel() // w.shape[0]
        self.w32 = w.view(self.out_f, self.in_f)
        self.w16 = a.view16(wname).view(self.out_f, self.in_f)
        self.gw = a.gview(wname).view(self.out_f, self.in_f)
        self.b = a.view(bname) if bname is not None else None
        self.gb = a.gview(bname) if bname is not None else None
        self._split = None
        self._split_version = -1

    def wsplit(self):
        """[out, 6*in] bf16 three-piece split of the fp32 weight (fp32-faithful path), cached per parameter version."""
        ver = (self.eng.arena.version(self.wname), self.eng.arena.generation)
        if self._split is None or ver != self._split_version:
            if self._split is None:
                self._split = torch.empty(self.out_f, 6 * self.in_f, dtype=torch.bfloat16, device=self.w32.device)
            lib.split3(self.w32, self._split, 1)
            self._split_version = ver
        return self._split


    def wsplit_rows(self, pad_rows=None):
        """[6*out(_padded), in] bf16: the split pieces concatenated along rows -- B operand of the fp32 dgrad."""
        ver = (self.eng.arena.version(self.wname), self.eng.arena.generation)
        rows = pad_rows or self.out_f
        if getattr(self, "_split_r", None) is None or ver != self._split_r_version:
            src = self.w32
            if rows != self.out_f:
                src = torch.zeros(rows, self.in_f, dtype=torch.float32, device=self.w32.device)
                src[:self.out_f] = self.w32
            self._split_r = torch.empty(6 * rows, self.in_f, dtype=torch.bfloat16, device=self.w32.device)
            lib.split3(src, self._split_r, 3)
            self._split_r_version = ver
        return self._split_r


class _Norm:
    def __init__(self, eng, prefix, eps):
        a = eng.arena
        self.g, self.b = a.view(prefix + ".weight"), a.view(prefix + ".bias")
        self.gg, self.gb = a.gview(prefix + ".weight"), a.gview(prefix + ".bias")
        self.eps = eps


class _BlockP:
    def __init__(self, eng, n1, attn, n2, mlp, eps, bias):
        b = (lambda s: s) if bias else (lambda s: None)
        self.ln1, self.ln2 = _Norm(eng, n1, eps), _Norm(eng, n2, eps)
        self.qkv = _Lin(eng, attn + ".qkv.weight", b(attn + ".qkv.bias"))
        self.proj = _Lin(eng, attn + ".proj.weight", b(attn + ".proj.bias"))
        self.fc1 = _Lin(eng, mlp + ".fc1.weight", b(mlp + ".fc1.bias"))
        self.fc2 = _Lin(eng, mlp + ".fc2.weight", b(mlp + ".fc2.bias"))


class Arena:
    """All trainable parameters as views of ONE flat fp32 buffer (+ a bf16 shadow and a flat gradient buffer): one
    cast kernel per step, one allreduce, one optimizer kernel.  ``state_dict`` / ``load_state_dict`` / ``.grad`` keep
    working because every ``nn.Parameter`` stays a normal parameter whose storage is a slice of the arena."""

    def __init__(self, model, device):
        self.model = model
        self.device = device
        self.names, self.params, self.offsets = [], [], {}
        off = 0
        for name, p in model.named_parameters():
            if not p.requires_grad:
                continue
            self.names.append(name)
            self.params.append(p)
            self.offsets[name] = (off, p.numel(), tuple(p.shape))
            off += _align(p.numel())
        self.total = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.flat16 = torch.zeros(off, dtype=torch.bfloat16, device=device)
        self.grad = torch.zeros(off, dtype=torch.float32, device=device)
        for name, p in zip(self.names, self.params):
            v = self.view(name)
            v.copy_(p.data)
            p.data = v
        self._ptrs = [p.data_ptr() for p in self.params]
        self._versions = None
        self.generation = 0          # bumped by the fused optimizer kernel (it updates parameters without torch knowing)
        self.carry, self.carry_live = None, False

    def view(self, name):
        o, n, shape = self.offsets[name]
        return self.flat[o:o + n].view(shape)

    def view16(self, name):
        o, n, shape = self.offsets[name]
        return self.flat16[o:o + n].view(shape)

    def gview(self, name):
        o, n, shape = self.offsets[name]
        return self.grad[o:o + n].view(shape)

    def version(self, name):
        return self.params[self.names.index(name)]._version

    def intact(self):
        return all(p.data_ptr() == q for p, q in zip(self.params, self._ptrs))

    def refresh16(self, force=False):
        vers = [p._version for p in self.params]
        if force or vers != self._versions:
            lib.cast_bf16(self.flat, self.flat16, self.total)
            self._versions = vers

    def prepare_grads(self):
        """Called at the start of every training forward.  torch semantics: ``.grad`` ACCUMULATES until the caller zeroes
        it (engine/processor.py:72 ``optimizer.zero_grad()``; gradient accumulation over micro-batches must keep working),
        while the kernels of one backward WRITE most gradients (bias / LayerNorm column sums) and only accumulate the
        split-K weight gradients -- they need a zeroed arena per backward.
        * every ``p.grad`` is None (``zero_grad(set_to_none=True)``, the torch default) -> one memset of the arena;
        * some ``p.grad`` is live (accumulation, or ``zero_grad(set_to_none=False)``) -> what the caller holds is saved to
          a second flat buffer (``carry``), the arena is zeroed, and ``finish_grads`` -- called by the last backward node
          of the step -- adds the carry back: ``p.grad`` ends up as old + new, as with torch's AccumulateGrad.  Foreign
          ``.grad`` tensors (assigned by the caller) are adopted into the carry once and re-pointed at the arena view."""
        self.carry_live = False
        if not any(p.grad is not None for p in self.params):
            self.grad.zero_()
            return
        if self.carry is None:
            self.carry = torch.empty_like(self.grad)
        # live arena views carry their old value through one flat copy; parameters without a .grad carry nothing, foreign
        # tensors carry their own value
        self.carry.copy_(self.grad)
        for name, p in zip(self.names, self.params):
            o, n, shape = self.offsets[name]
            if p.grad is None:
                self.carry[o:o + n].zero_()
            elif p.grad.data_ptr() != self.grad[o:o + n].data_ptr():
                self.carry[o:o + n].view(shape).copy_(p.grad)
                p.grad = self.grad[o:o + n].view(shape)
        self.grad.zero_()
        self.carry_live = True

    def finish_grads(self):
        """End of the step's backward (the backbone node runs last): old + new for callers that accumulate."""
        if self.carry_live:
            self.grad.add_(self.carry)
            self.carry_live = False

    def attach_grads(self, names=None):
        """Expose the gradient arena through ``p.grad`` (what GradScaler / torch optimizers of the unchanged caller
        read, engine/processor.py:94-96).  Idempotent: a parameter with several contributors per step (BACKBONE_HEAD /
        BACKBONE_BN are called three times when AL = 0) is attached by the first one."""
        for name, p in zip(self.names, self.params):
            if names is not None and name not in names:
                continue
            if p.grad is None:
                p.grad = self.gview(name)


class Workspace:
    def __init__(self, device):
        self.device = device
        self.bufs = {}

    def get(self, name, shape, dtype):
        n = 1
        for s in shape:
            n *= int(s)
        t = self.bufs.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            # zero-filled: rows past the (device-side) row count of a packed matrix must stay finite
            t = torch.zeros(max(n, 1), dtype=dtype, device=self.device)
            self.bufs[name] = t
        return t[:n].view(shape)

    def bytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values())


def _pick_split(tiles, kb_total, sms=148):
    best, best_eff = 1, 0.0
    for s in range(1, 17):
        if s > kb_total:
            break
        items = tiles * s
        waves = (items + sms - 1) // sms
        eff = items / (waves * sms)
        if eff > best_eff + 0.02:
            best, best_eff = s, eff
    return best


class EditorEngine:
    def __init__(self, model):
        self.model = model
        self.arena = None
        self.ws = None
        self.sel = None
        self.stats = {}

    # ------------------------------------------------------------------ setup
    def _ensure(self, device):
        if not torch.cuda.is_available() or device.type != "cuda":
            raise lib.EdbError("editor_b200 runs on a CUDA device (sm_100a) only; there is no CPU fallback")
        lib.load()
        if self.arena is not None and self.arena.device == device and self.arena.intact():
            return
        m = self.model
        frozen = [n for n, p in m.named_parameters() if not p.requires_grad and ".memory_cls." not in n]
        if frozen:      # the arena holds the trainable parameters only; OCFR's centre banks are the reference's own frozen ones
            raise lib.EdbError("parameters with requires_grad=False are not supported by the flat parameter arena "
                               "(%s ...): keep them trainable and give them lr 0 / exclude them from the optimizer" % frozen[0])
        self.arena = Arena(m, device)
        self.ws = Workspace(device)
        base = "BACKBONE.base."
        self.patch = _Lin(self, base + "patch_embed.proj.weight", base + "patch_embed.proj.bias")
        self.bb_blocks = [_BlockP(self, base + "blocks.%d.norm1" % i, base + "blocks.%d.attn" % i,
                                  base + "blocks.%d.norm2" % i, base + "blocks.%d.mlp" % i, 1e-6, True) for i in range(12)]
        self.bb_norm = _Norm(self, base + "norm", 1e-6)
        f = "FUSE_block."
        self.hma_blocks = [_BlockP(self, f + n1, f + at, f + n2, f + ml, 1e-5, False)
                           for n1, at, n2, ml in (("normR", "attnR", "normR_", "mlpR"), ("normN", "attnN", "normN_", "mlpN"),
                                                  ("normT", "attnT", "normT_", "mlpT"), ("norm1", "attn1", "norm2", "mlp"))]
        self.hma_out = _Norm(self, f + "out_norm", 1e-5)
        self.tail_lin = {n: _Lin(self, n + ".weight", (n + ".bias") if (n + ".bias") in self.arena.offsets else None)
                         for n in ("RGB_REDUCE", "NIR_REDUCE", "TIR_REDUCE", "FUSE_HEAD", "BACKBONE_HEAD", "AL_HEAD")
                         if (n + ".weight") in self.arena.offsets}
        self.bb_names = [n for n in self.arena.names if n.startswith("BACKBONE.base.") and ".fc." not in n]
        self.hma_names = [n for n in self.arena.names if n.startswith("FUSE_block.")]
        self.bb_plist = [self.arena.params[self.arena.names.index(n)] for n in self.bb_names]
        self.hma_plist = [self.arena.params[self.arena.names.index(n)] for n in self.hma_names]

    def mark_params_dirty(self):
        """Call after changing parameters in a way torch's version counters do not see (`p.data.mul_()`, EMA, clipping
        through `.data`, a custom optimizer writing raw storage): the bf16 shadow and the fp32 split caches are refreshed
        when `p._version` or the arena generation changes, and `.data` updates bump neither."""
        if self.arena is not None:
            self.arena.generation += 1
            self.arena.refresh16(force=True)

    def _mark(self, name):
        ev = self.stats.get("events")
        if ev is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append((name, e))

    # ------------------------------------------------------------------ building blocks
    def _linear(self, x, L, out, rows, prec, epi=lib.EPI_STORE, aux=None, out2=None, row_scale=None, group=1, rd=None):
        """rows = launch bound; rd = optional device pointer to the actual row count (packed HMA rows)."""
        if prec == BF16:
            lib.gemm(x, L.w16, out, rows, L.out_f, L.in_f, epilogue=epi, bias=L.b, aux=aux, out2=out2,
                     row_scale=row_scale, scale_group=group, M_dev=rd)
        else:
            xs = self.ws.get("split_a", (x.shape[0], 6 * L.in_f), torch.bfloat16)
            lib.split3(x, xs, 0, rows)
            lib.gemm(xs, L.wsplit(), out, rows, L.out_f, 6 * L.in_f, epilogue=epi, bias=L.b, aux=aux, out2=out2,
                     row_scale=row_scale, scale_group=group, M_dev=rd)
        return out

    def _wgrad(self, dy, x, L, rows, rd=None):
        """dW[out,in] += dy[rows,out]^T x[rows,in]   (split-K, fp32 atomic accumulate into the gradient arena)."""
        tiles = ((L.out_f + 127) // 128) * ((L.in_f + 255) // 256 if L.in_f > 128 else 1)
        split = _pick_split(tiles, (rows + 63) // 64)
        lib.gemm(dy, x, L.gw, L.out_f, L.in_f, rows, a_mn=True, b_mn=True, epilogue=lib.EPI_ATOMIC, split_k=split, K_dev=rd)

    def _dgrad32(self, dy, L, out, rows):
        """fp32-faithful dgrad: out[rows, in] = dy[rows, out] @ W  (3-piece bf16 split of both operands)."""
        ds = self.ws.get("split_a", (dy.shape[0], 6 * L.out_f), torch.bfloat16)
        lib.split3(dy, ds, 0, rows)
        lib.gemm(ds, L.wsplit_rows(), out, rows, L.in_f, 6 * L.out_f, b_mn=True)

    def _wgrad32(self, dy, x, L, rows):
        """fp32-faithful wgrad: dW[out, in] += dy^T x, both operands split along the reduction (row) dimension."""
        da = self.ws.get("split_ra", (6 * dy.shape[0] * L.out_f,), torch.bfloat16)[:6 * rows * L.out_f].view(6 * rows, L.out_f)
        xb = self.ws.get("split_rb", (6 * x.shape[0] * L.in_f,), torch.bfloat16)[:6 * rows * L.in_f].view(6 * rows, L.in_f)
        lib.split3(dy, da, 2, rows)
        lib.split3(x, xb, 3, rows)
        tiles = ((L.out_f + 127) // 128) * ((L.in_f + 255) // 256 if L.in_f > 128 else 1)
        split = _pick_split(tiles, (6 * rows + 63) // 64)
        lib.gemm(da, xb, L.gw, L.out_f, L.in_f, 6 * rows, a_mn=True, b_mn=True, epilogue=lib.EPI_ATOMIC, split_k=split)

    def _block_bwd32(self, g, rows, bp, sv, attn_bwd, dcol_prev):
        """fp32-faithful backward of one block (EDB_PREC_FP32; DROP_PATH must be 0).  g: fp32 gradient w.r.t. the block
        output, updated in place to the gradient w.r.t. the block input."""
        ws, cap = self.ws, g.shape[0]
        dh = ws.get("dh32", (cap, HID), torch.float32)
        self._dgrad32(g, bp.fc2, dh, rows)
        dpre = ws.get("dpre32", (cap, HID), torch.float32)
        lib.call("edb_gelu_bwd_f32", dh.data_ptr(), sv["pre"].data_ptr(), dpre.data_ptr(), rows * HID, lib.stream_ptr())
        self._wgrad32(g, sv["h"], bp.fc2, rows)
        if bp.fc1.gb is not None:
            lib.colsum(dpre, bp.fc1.gb, rows, HID)
        dln = ws.get("dln32", (cap, DIM), torch.float32)
        self._dgrad32(dpre, bp.fc1, dln, rows)
        self._wgrad32(dpre, sv["ln2"], bp.fc1, rows)
        lib.layernorm_bwd(dln, sv["x1"], sv["m2"], sv["r2"], bp.ln2.g, g, g, None, bp.ln2.gg, bp.ln2.gb, bp.proj.gb, rows)
        datt = ws.get("datt32", (cap, DIM), torch.float32)
        self._dgrad32(g, bp.proj, datt, rows)
        self._wgrad32(g, sv["att"], bp.proj, rows)
        dqkv = ws.get("dqkv32", (cap, 3 * DIM), torch.float32)
        attn_bwd(sv["qkv"], sv["P"], datt, dqkv, sv["att"])
        if bp.qkv.gb is not None:
            lib.colsum(dqkv, bp.qkv.gb, rows, 3 * DIM)
        self._dgrad32(dqkv, bp.qkv, dln, rows)
        self._wgrad32(dqkv, sv["ln1"], bp.qkv, rows)
        lib.layernorm_bwd(dln, sv["x"], sv["m1"], sv["r1"], bp.ln1.g, g, g, None, bp.ln1.gg, bp.ln1.gb, dcol_prev, rows)

    def _block_fwd(self, x, x1, x2, rows, bp, attn, tag, prec, rs_attn=None, rs_mlp=None, group=1, rd=None, keep=True):
        """One transformer block (vit_pytorch.py:215-220 / :311-317,328-329) on `rows` packed token rows.
        x, x1, x2: fp32 residual stream before / after attention / after MLP.  Returns what the backward needs."""
        ws, cap = self.ws, x.shape[0]
        adt = torch.bfloat16 if prec == BF16 else torch.float32
        ln1 = ws.get(tag + "ln1", (cap, DIM), adt)
        m1, r1 = ws.get(tag + "m1", (cap,), torch.float32), ws.get(tag + "r1", (cap,), torch.float32)
        lib.layernorm_fwd(x, bp.ln1.g, bp.ln1.b, bp.ln1.eps, ln1, m1, r1, rows, rows_dev=rd)
        qkv = ws.get(tag + "qkv", (cap, 3 * DIM), adt)
        self._linear(ln1, bp.qkv, qkv, rows, prec, rd=rd)
        att = ws.get(tag + "att", (cap, DIM), adt)
        P = attn(qkv, att, tag)
        self._linear(att, bp.proj, x1, rows, prec, epi=lib.EPI_RESIDUAL, aux=x, row_scale=rs_attn, group=group, rd=rd)
        ln2 = ws.get(tag + "ln2", (cap, DIM), adt)
        m2, r2 = ws.get(tag + "m2", (cap,), torch.float32), ws.get(tag + "r2", (cap,), torch.float32)
        lib.layernorm_fwd(x1, bp.ln2.g, bp.ln2.b, bp.ln2.eps, ln2, m2, r2, rows, rows_dev=rd)
        # "pre": bf16 mode -> gelu'(pre-activation), written by the fc1 epilogue for the backward (skipped when nothing
        # is kept); fp32-faithful mode -> the pre-activation itself (edb_gelu_bwd_f32 differentiates it exactly)
        pre = ws.get(tag + "pre", (cap, HID), adt) if (keep or prec != BF16) else None
        h = ws.get(tag + "h", (cap, HID), adt)
        self._linear(ln2, bp.fc1, h, rows, prec, epi=lib.EPI_GELU, out2=pre, rd=rd)
        self._linear(h, bp.fc2, x2, rows, prec, epi=lib.EPI_RESIDUAL, aux=x1, row_scale=rs_mlp, group=group, rd=rd)
        return dict(x=x, x1=x1, ln1=ln1, m1=m1, r1=r1, qkv=qkv, att=att, P=P, ln2=ln2, m2=m2, r2=r2, pre=pre, h=h)

    def _block_bwd(self, g, gb, rows, bp, sv, attn_bwd, dcol_prev, rs_attn=None, rs_prev=None, group=1, rd=None):
        """Backward of `_block_fwd`.  g (fp32) / gb (bf16, already DropPath-scaled): gradient w.r.t. the block output;
        on return they hold the gradient w.r.t. the block input (gb scaled by `rs_prev`).  With a device-side row count
        (rd) the 64 rows after the last valid one of every wgrad A-operand are zeroed: they pad the split-K reduction."""
        ws, cap = self.ws, g.shape[0]
        dpre = ws.get("dpre", (cap + 64, HID), torch.bfloat16)
        dqkv = ws.get("dqkv", (cap + 64, 3 * DIM), torch.bfloat16)
        if rd is not None:
            for buf, width in ((gb, DIM), (dpre, HID), (dqkv, 3 * DIM)):
                lib.call("edb_zero_rows", buf.data_ptr(), width * 2, rd, 64, lib.stream_ptr())
        # dpre = (gb @ W2) * gelu'(pre); the fc1 bias gradient (column sums of dpre) comes out of the same epilogue
        lib.gemm(gb, bp.fc2.w16, dpre, rows, HID, DIM, b_mn=True, epilogue=lib.EPI_GELU_BWD, aux=sv["pre"], M_dev=rd,
                 colsum=bp.fc1.gb)
        self._wgrad(gb, sv["h"], bp.fc2, rows, rd)
        dln = ws.get("dln", (cap, DIM), torch.bfloat16)
        lib.gemm(dpre, bp.fc1.w16, dln, rows, DIM, HID, b_mn=True, M_dev=rd)
        self._wgrad(dpre, sv["ln2"], bp.fc1, rows, rd)
        lib.layernorm_bwd(dln, sv["x1"], sv["m2"], sv["r2"], bp.ln2.g, g, g, gb, bp.ln2.gg, bp.ln2.gb, bp.proj.gb, rows,
                          row_scale=rs_attn, scale_group=group, rows_dev=rd)
        datt = ws.get("datt", (cap, DIM), torch.bfloat16)
        # qkv bias gradient without a pass over dqkv [rows, 2304]:
        #   V part: sum_keys dV = sum_q dO_q (sum_k P_qk) = column sums of datt (softmax rows sum to 1) -> fused into the
        #           epilogue of the proj dgrad GEMM that produces datt;
        #   K part: sum_keys dK = sum_q Q_q (sum_k dS_qk) = 0 exactly (softmax is invariant to a per-query shift of the
        #           scores, which is all a key bias does) -- autograd in the reference returns fp32 rounding noise here;
        #   Q part: column sums of dqkv[:, :768] (a third of the matrix).
        lib.gemm(gb, bp.proj.w16, datt, rows, DIM, DIM, b_mn=True, M_dev=rd,
                 colsum=bp.qkv.gb[2 * DIM:] if bp.qkv.gb is not None else None)
        self._wgrad(gb, sv["att"], bp.proj, rows, rd)
        attn_bwd(sv["qkv"], sv["P"], datt, dqkv, sv["att"])
        if bp.qkv.gb is not None:
            lib.colsum(dqkv, bp.qkv.gb, rows, DIM)
        lib.gemm(dqkv, bp.qkv.w16, dln, rows, DIM, 3 * DIM, b_mn=True, M_dev=rd)
        self._wgrad(dqkv, sv["ln1"], bp.qkv, rows, rd)
        lib.layernorm_bwd(dln, sv["x"], sv["m1"], sv["r1"], bp.ln1.g, g, g, gb, bp.ln1.gg, bp.ln1.gb, dcol_prev, rows,
                          row_scale=rs_prev, scale_group=group, rows_dev=rd)

    # ------------------------------------------------------------------ backbone (3 modalities batched: S = 3B sequences)
    def backbone_forward(self, rgb, ni, ti, cam, prec, keep, droppath=None):
        """Trans.forward for the three modality batches at once (shared weights, make_model.py:158-160).
        keep=True stores per-layer activations for the backward.  Returns (tokens [3B,129,768] fp32, saved dict)."""
        ws = self.ws
        B, _, H, W = rgb.shape
        S, R = 3 * B, 3 * B * NTOK
        adt = torch.bfloat16 if prec == BF16 else torch.float32
        patches = ws.get("patches", (S * NPATCH, DIM), adt)
        lib.call("edb_patch_im2col", rgb.data_ptr(), ni.data_ptr(), ti.data_ptr(), B, H, W, patches.data_ptr(), DIM,
                 int(prec == FP32), lib.stream_ptr())
        pe = ws.get("pe_out", (S * NPATCH, DIM), torch.float32)
        self._linear(patches, self.patch, pe, S * NPATCH, prec)
        a = self.arena
        base = "BACKBONE.base."
        sie = a.view(base + "sie_embed") if (base + "sie_embed") in a.offsets else None
        x = ws.get("x_0", (R, DIM), torch.float32)
        lib.call("edb_embed_assemble", pe.data_ptr(), a.view(base + "cls_token").data_ptr(),
                 a.view(base + "pos_embed").data_ptr(), lib.ptr(sie), cam.data_ptr(), self.model.sie_coe, S, B, NPATCH,
                 x.data_ptr(), lib.stream_ptr())
        pdt = torch.bfloat16 if prec == BF16 else torch.float32
        saved, maps = [], []
        for l, bp in enumerate(self.bb_blocks):
            tag = "bb%d_" % (l if keep else 0)
            Pm = ws.get("bbP%d" % l, (S * HEADS, NTOK, P_LD), pdt)
            maps.append(Pm)

            def attn(qkv, out, _tag, Pm=Pm):
                lib.attention(qkv, out, Pm, S, HEADS, NTOK, SCALE, fixed_len=NTOK, p_rows=NTOK, ldp=P_LD,
                              impl=self.stats.get("attn_impl", 0))
                return Pm
            x1 = ws.get("x_%d" % (2 * l + 1 if keep else 1), (R, DIM), torch.float32)
            x2 = ws.get("x_%d" % (2 * l + 2 if keep else 2 - (l & 1) * 2), (R, DIM), torch.float32)
            rs_a = rs_m = None
            if droppath is not None:
                rs_a, rs_m = droppath[2 * l], droppath[2 * l + 1]
            sv = self._block_fwd(x, x1, x2, R, bp, attn, tag, prec, rs_a, rs_m, NTOK, keep=keep)
            if keep:
                saved.append(sv)
            x = x2
        tokens = torch.empty(S, NTOK, DIM, dtype=torch.float32, device=rgb.device)
        mf, rf = ws.get("bb_mf", (R,), torch.float32), ws.get("bb_rf", (R,), torch.float32)
        lib.layernorm_fwd(x, self.bb_norm.g, self.bb_norm.b, self.bb_norm.eps, tokens.view(R, DIM), mf, rf, R)
        return tokens, dict(blocks=saved, x_last=x, mf=mf, rf=rf, maps=maps, patches=patches, B=B, cam=cam,
                            droppath=droppath, prec=prec)

    def _grad_stage(self, stage):
        """Tell the trainer that a contiguous slice of the gradient arena is final (bucketed allreduce, train.py)."""
        hook = self.stats.get("grad_hook")
        if hook is not None:
            hook(stage)

    def backbone_backward(self, sv, d_tokens):
        ws, a = self.ws, self.arena
        self._grad_stage("after_backbone")      # BackboneFn runs last: every tail / HMA gradient is complete
        B = sv["B"]
        S, R = 3 * B, 3 * B * NTOK
        dp = sv["droppath"]
        if sv["prec"] == FP32:
            return self._backbone_backward32(sv, d_tokens)
        g = ws.get("g", (R, DIM), torch.float32)
        gb = ws.get("gb", (R, DIM), torch.bfloat16)
        last = self.bb_blocks[-1]
        lib.layernorm_bwd(d_tokens.reshape(R, DIM), sv["x_last"], sv["mf"], sv["rf"], self.bb_norm.g, None, g, gb,
                          self.bb_norm.gg, self.bb_norm.gb, last.fc2.gb, R,
                          row_scale=None if dp is None else dp[23], scale_group=NTOK)

        def attn_bwd(qkv, P, datt, dqkv, att):
            # att = the forward output O: the tensor-core backward takes delta_i = dO_i . O_i from it
            lib.attention(qkv, att, P, S, HEADS, NTOK, SCALE, fixed_len=NTOK, p_rows=NTOK, ldp=P_LD,
                          impl=self.stats.get("attn_impl", 0), d_out=datt, d_qkv=dqkv, backward=True)
        for l in range(11, -1, -1):
            bp = self.bb_blocks[l]
            prev_gb = self.bb_blocks[l - 1].fc2.gb if l > 0 else None
            rs_a = None if dp is None else dp[2 * l]
            rs_prev = None if (dp is None or l == 0) else dp[2 * l - 1]
            self._block_bwd(g, gb, R, bp, sv["blocks"][l], attn_bwd, prev_gb, rs_a, rs_prev, NTOK)
            if l in GRAD_STAGE_BLOCKS:
                self._grad_stage("blocks_from_%d" % l)
        base = "BACKBONE.base."
        dpatch = ws.get("dpatch", (S * NPATCH, DIM), torch.bfloat16)
        dpos = a.gview(base + "pos_embed").view(NTOK, DIM)
        dsie = a.gview(base + "sie_embed") if (base + "sie_embed") in a.offsets else None
        lib.call("edb_embed_assemble_bwd", g.data_ptr(), S, B, NPATCH, sv["cam"].data_ptr(), self.model.sie_coe,
                 dpos.data_ptr(), lib.ptr(dsie), dpatch.data_ptr(), 0, lib.stream_ptr())
        a.gview(base + "cls_token").view(DIM).add_(dpos[0])
        lib.colsum(dpos[1:], self.patch.gb, NPATCH, DIM)
        self._wgrad(dpatch, sv["patches"], self.patch, S * NPATCH)
        self._grad_stage("rest")

    def _backbone_backward32(self, sv, d_tokens):
        ws, a = self.ws, self.arena
        B = sv["B"]
        S, R = 3 * B, 3 * B * NTOK
        if sv["droppath"] is not None:
            raise lib.EdbError("the fp32-faithful backward supports MODEL.DROP_PATH 0.0 only")
        g = ws.get("g", (R, DIM), torch.float32)
        lib.layernorm_bwd(d_tokens.reshape(R, DIM), sv["x_last"], sv["mf"], sv["rf"], self.bb_norm.g, None, g, None,
                          self.bb_norm.gg, self.bb_norm.gb, self.bb_blocks[-1].fc2.gb, R)

        def attn_bwd(qkv, P, datt, dqkv, att):
            lib.attention(qkv, None, P, S, HEADS, NTOK, SCALE, fixed_len=NTOK, p_rows=NTOK, ldp=P_LD, impl=1, d_out=datt,
                          d_qkv=dqkv, backward=True)
        for l in range(11, -1, -1):
            prev_gb = self.bb_blocks[l - 1].fc2.gb if l > 0 else None
            self._block_bwd32(g, R, self.bb_blocks[l], sv["blocks"][l], attn_bwd, prev_gb)
            if l in GRAD_STAGE_BLOCKS:
                self._grad_stage("blocks_from_%d" % l)
        base = "BACKBONE.base."
        dpatch = ws.get("dpatch32", (S * NPATCH, DIM), torch.float32)
        dpos = a.gview(base + "pos_embed").view(NTOK, DIM)
        dsie = a.gview(base + "sie_embed") if (base + "sie_embed") in a.offsets else None
        lib.call("edb_embed_assemble_bwd", g.data_ptr(), S, B, NPATCH, sv["cam"].data_ptr(), self.model.sie_coe,
                 dpos.data_ptr(), lib.ptr(dsie), dpatch.data_ptr(), 1, lib.stream_ptr())
        a.gview(base + "cls_token").view(DIM).add_(dpos[0])
        lib.colsum(dpos[1:], self.patch.gb, NPATCH, DIM)
        self._wgrad32(dpatch, sv["patches"], self.patch, S * NPATCH)
        self._grad_stage("rest")

    # ------------------------------------------------------------------ SFTS selection (no gradient)
    def select(self, rgb, ni, ti, maps, prec, want_debug=False):
        """mask_fre | RGB_index | NIR_index | TIR_index as 128-bit sets + the packed-row offsets (SFTS.py:183-190)."""
        ws, m = self.ws, self.model
        B, _, H, W = rgb.shape
        S = 3 * B
        counts = ws.get("freq_counts", (B, NPATCH), torch.int32)
        lib.call("edb_freq_counts", rgb.data_ptr(), ni.data_ptr(), ti.data_ptr(), B, H, W, counts.data_ptr(),
                 lib.stream_ptr())
        index = torch.empty(B, 4, dtype=torch.int32, device=rgb.device)
        lib.call("edb_topk_mask", counts.data_ptr(), 0, NPATCH, B, NPATCH, int(m.FREQ_INDEX.keep), index.data_ptr(), 0,
                 lib.stream_ptr())
        dbg = {}
        if want_debug:
            dbg["counts"] = counts.clone()
            dbg["mask_fre"] = index.clone()
            dbg["mod_mask"] = torch.zeros(S, 4, dtype=torch.int32, device=rgb.device)
            dbg["rows"] = torch.empty(S * HEADS, NPATCH, dtype=torch.float32, device=rgb.device)
        arr = (ctypes.c_void_p * len(maps))(*[t.data_ptr() for t in maps])
        lib.call("edb_rollout_topk", arr, len(maps), int(prec == FP32), S, B, HEADS, NTOK, P_LD, int(m.head_keep),
                 index.data_ptr(), lib.ptr(dbg.get("mod_mask")), lib.ptr(dbg.get("rows")), lib.stream_ptr())
        seq_off = torch.empty(B + 1, dtype=torch.int32, device=rgb.device)
        seq_off3 = torch.empty(B + 1, dtype=torch.int32, device=rgb.device)
        lib.call("edb_index_finalize", index.data_ptr(), B, seq_off.data_ptr(), seq_off3.data_ptr(), lib.stream_ptr())
        # static bounds from the configuration: <= HEAD_KEEP tokens per head, 12 heads, 3 modalities, + FREQUENCY_KEEP
        ml_cap = min(NTOK, 1 + 3 * HEADS * int(m.head_keep) + int(m.FREQ_INDEX.keep))
        sel = dict(index=index, seq_off=seq_off, seq_off3=seq_off3, B=B, ml_cap=ml_cap, T_cap=B * ml_cap,
                   T_dev=seq_off.data_ptr() + 4 * B, T3_dev=seq_off3.data_ptr() + 4 * B, debug=dbg)
        if prec == FP32:
            # parity mode: exact row counts on the host (one sync, like make_model.py:200); the split operands of the
            # fp32-faithful GEMMs are laid out by the actual row count
            off_host = seq_off.cpu()
            sel.update(T_cap=int(off_host[-1]), ml_cap=int((off_host[1:] - off_host[:-1]).max()), T_dev=None, T3_dev=None)
        return sel

    # ------------------------------------------------------------------ HMA on packed kept tokens
    def _varlen_attn(self, seq_off, nseq, max_len, tagP, prec, total_rows):
        """AttentionMask on packed kept tokens: tensor-core var-len kernel (bf16, <= 256 tokens) or the CUDA-core kernel
        (fp32-faithful mode, or sequences whose padded key count does not fit the tensor-core tiles)."""
        ws = self.ws
        use_tc = prec == BF16 and max_len <= 256 and self.stats.get("hma_attn_impl", 2) == 2
        if use_tc:
            kp = 128 if max_len <= 128 else 256
            p_rows, ldp, impl = (max_len + 127) // 128 * 128, kp, 2
        else:
            if max_len > 256:
                raise lib.EdbError("HMA sequences longer than 256 kept tokens are not supported (got %d)" % max_len)
            p_rows, ldp, impl = max_len, _align(max_len, 8), 1
        pdt = torch.bfloat16 if prec == BF16 else torch.float32

        def fwd(qkv, out, tag):
            Pm = ws.get(tagP + tag, (nseq * HEADS, p_rows, ldp), pdt)
            lib.attention(qkv, out, Pm, nseq, HEADS, max_len, SCALE, seq_off=seq_off, p_rows=p_rows, ldp=ldp, impl=impl,
                          total_rows=total_rows)
            return Pm

        def bwd(qkv, P, datt, dqkv, att):
            lib.attention(qkv, None, P, nseq, HEADS, max_len, SCALE, seq_off=seq_off, p_rows=p_rows, ldp=ldp, impl=impl,
                          d_out=datt, d_qkv=dqkv, backward=True, total_rows=total_rows)
        return fwd, bwd

    def hma_forward(self, tokens, sel, prec, training):
        """SFTS masking (SFTS.py:208-222) + BlockMask.forward (vit_pytorch.py:309-352) + pooling
        (make_model.py:186-203) on the packed kept rows.  tokens: [3B,129,768] fp32."""
        ws = self.ws
        B, T, ml = sel["B"], sel["T_cap"], sel["ml_cap"]          # launch bounds; actual counts stay on the device
        td, t3d = sel["T_dev"], sel["T3_dev"]
        cap = T
        dev = tokens.device
        xp = ws.get("hma_xp", (3, cap, DIM), torch.float32)
        loss_bcc = torch.zeros(1, dtype=torch.float32, device=dev) if training else None
        lib.call("edb_sfts_pack_fwd", tokens.data_ptr(), sel["index"].data_ptr(), sel["seq_off"].data_ptr(), B, cap,
                 xp.data_ptr(), lib.ptr(loss_bcc), lib.stream_ptr())
        afwd, abwd = self._varlen_attn(sel["seq_off"], B, ml, "hmaP", prec, T)
        x2all = ws.get("hma_x2", (3, cap, DIM), torch.float32)
        x1all = ws.get("hma_x1", (3, cap, DIM), torch.float32)
        saved = []
        for m in range(3):
            saved.append(self._block_fwd(xp[m], x1all[m], x2all[m], T, self.hma_blocks[m], afwd, "hma%d_" % m, prec, rd=td,
                                          keep=training))
        cls_mid = None
        if training:
            cls_mid = torch.empty(3, B, DIM, dtype=torch.float32, device=dev)
            lib.call("edb_cls_rows", x2all.data_ptr(), cap, sel["seq_off"].data_ptr(), B, cls_mid.data_ptr(), 0,
                     lib.stream_ptr())
        xj = ws.get("hma_xj", (3 * cap, DIM), torch.float32)
        lib.call("edb_joint_gather", x2all.data_ptr(), cap, xj.data_ptr(), sel["seq_off"].data_ptr(), B, ml, 0,
                 lib.stream_ptr())
        jfwd, jbwd = self._varlen_attn(sel["seq_off3"], B, 3 * ml, "hmaPj", prec, 3 * T)
        xj1 = ws.get("hma_xj1", (3 * cap, DIM), torch.float32)
        xj2 = ws.get("hma_xj2", (3 * cap, DIM), torch.float32)
        svj = self._block_fwd(xj, xj1, xj2, 3 * T, self.hma_blocks[3], jfwd, "hmaJ_", prec, rd=t3d, keep=training)
        xo = ws.get("hma_xo", (3 * cap, DIM), torch.float32)
        mo, ro = ws.get("hma_mo", (3 * cap,), torch.float32), ws.get("hma_ro", (3 * cap,), torch.float32)
        lib.layernorm_fwd(xj2, self.hma_out.g, self.hma_out.b, self.hma_out.eps, xo, mo, ro, 3 * T, rows_dev=t3d)
        cls_out = torch.empty(3, B, DIM, dtype=torch.float32, device=dev)
        patch_mean = torch.empty(3, B, DIM, dtype=torch.float32, device=dev)
        num = torch.empty(B, dtype=torch.int32, device=dev)
        lib.call("edb_pool_fwd", xo.data_ptr(), sel["seq_off"].data_ptr(), B, cls_out.data_ptr(), patch_mean.data_ptr(),
                 num.data_ptr(), lib.stream_ptr())
        sv = dict(mods=saved, joint=svj, xj2=xj2, mo=mo, ro=ro, num=num, abwd=abwd, jbwd=jbwd, cap=cap, prec=prec)
        return cls_out, patch_mean, cls_mid, loss_bcc, num, sv

    def hma_backward(self, tokens, sel, sv, d_cls, d_patch, d_mid, d_bcc):
        ws = self.ws
        B, T, ml, cap = sel["B"], sel["T_cap"], sel["ml_cap"], sv["cap"]
        td, t3d = sel["T_dev"], sel["T3_dev"]
        dxo = ws.get("hma_dxo", (3 * cap, DIM), torch.float32)
        lib.call("edb_pool_bwd", d_cls.data_ptr(), d_patch.data_ptr(), sel["seq_off"].data_ptr(), sv["num"].data_ptr(), B,
                 ml, dxo.data_ptr(), lib.stream_ptr())
        gj = ws.get("hma_gj", (3 * cap, DIM), torch.float32)
        fp32 = sv["prec"] == FP32
        gbj = None if fp32 else ws.get("hma_gbj", (3 * cap + 64, DIM), torch.bfloat16)
        lib.layernorm_bwd(dxo, sv["xj2"], sv["mo"], sv["ro"], self.hma_out.g, None, gj, gbj, self.hma_out.gg,
                          self.hma_out.gb, None, 3 * T, rows_dev=t3d)
        if fp32:
            self._block_bwd32(gj, 3 * T, self.hma_blocks[3], sv["joint"], sv["jbwd"], None)
        else:
            self._block_bwd(gj, gbj, 3 * T, self.hma_blocks[3], sv["joint"], sv["jbwd"], None, rd=t3d)
        gm = ws.get("hma_gm", (3, cap, DIM), torch.float32)
        lib.call("edb_joint_gather", gm.data_ptr(), cap, gj.data_ptr(), sel["seq_off"].data_ptr(), B, ml, 1,
                 lib.stream_ptr())
        if d_mid is not None:
            lib.call("edb_cls_rows", gm.data_ptr(), cap, sel["seq_off"].data_ptr(), B, d_mid.data_ptr(), 1,
                     lib.stream_ptr())
        gbm = None if fp32 else ws.get("hma_gbm", (cap + 64, DIM), torch.bfloat16)
        for m in range(3):
            if fp32:
                self._block_bwd32(gm[m], T, self.hma_blocks[m], sv["mods"][m], sv["abwd"], None)
                continue
            lib.call("edb_cast_rows_f32_bf16", gm[m].data_ptr(), gbm.data_ptr(), T, DIM, td, lib.stream_ptr())
            self._block_bwd(gm[m], gbm, T, self.hma_blocks[m], sv["mods"][m], sv["abwd"], None, rd=td)
        d_tokens = torch.empty_like(tokens)
        lib.call("edb_sfts_pack_bwd", tokens.data_ptr(), sel["index"].data_ptr(), sel["seq_off"].data_ptr(), B, cap,
                 gm.data_ptr(), lib.ptr(d_bcc), d_tokens.data_ptr(), lib.stream_ptr())
        return d_tokens

    # ------------------------------------------------------------------ model-level forward
    def _precision(self, training):
        p = self.model.precision
        if p == "auto":
            return BF16 if (training or torch.is_autocast_enabled("cuda")) else FP32
        return p

    def _warn_pending(self):
        """The gradient arena holds ONE backward at a time: fwd/bwd, fwd/bwd accumulates correctly; fwd, fwd, bwd, bwd does
        not (the second forward re-zeroes the arena the first backward is going to write).  A training forward that is
        never back-propagated is harmless, so this only warns, once."""
        if self.stats.get("pending_backward") and not self.stats.get("warned_pending"):
            import warnings
            warnings.warn("editor_b200: a training forward started before the backward of the previous one ran; gradients "
                          "of overlapping forward passes are not supported (run fwd/bwd, fwd/bwd)")
            self.stats["warned_pending"] = True

    def _check_cam(self, cam, B, training):
        """`sie_embed[cam[b]]` is gathered by the embed kernel: range-check on the device (no host sync) on the first
        forward of every (batch size, mode), or on every one with ``model.validate_inputs = True``."""
        m = self.model
        if cam.shape != (B,):
            raise lib.EdbError("cam_label must have shape [%d]" % B)
        seen = self.stats.setdefault("validated", set())
        check = getattr(m, "validate_inputs", False) or (B, training) not in seen
        seen.add((B, training))
        cams = int(getattr(m.BACKBONE.base, "cam_num", 0))
        if check and cams > 1:
            torch._assert_async(((cam >= 0) & (cam < cams)).all(), "cam_label out of range [0, camera_num)")
        return check

    def _check_labels(self, label, B, device, check):
        """The tail kernels index with these (`centers[label[b]]`, `logits[b][label[b]]`): int64, on the device,
        contiguous, in range -- where the reference raises an index error we must not read out of bounds.  OCFR needs
        P x K contiguous identities (OCFR.py:31-42 takes ``label_[::chunk]``; data/datasets/sampler.py delivers exactly
        that)."""
        if label is None:
            raise lib.EdbError("the training forward needs `label` (make_model.py:150, OCFR.py:44)")
        label = label.to(device=device, dtype=torch.int64).contiguous()
        if label.shape != (B,):
            raise lib.EdbError("label must have shape [%d]" % B)
        if check:
            C = self.model.FUSE_block.memory_cls.RGB_centers.shape[0]
            torch._assert_async(((label >= 0) & (label < C)).all(), "label out of range [0, num_class)")
            # OCFR.py:31-42 pairs row j with the centre of label_[(j // chunk) * chunk], chunk = B // #identities; the
            # kernel pairs row j with the centre of its OWN label -- identical exactly when the following holds
            srt = torch.sort(label).values
            n_ids = (srt[1:] != srt[:-1]).sum() + 1
            chunk = torch.div(B, n_ids, rounding_mode="floor")
            j = torch.arange(B, device=device)
            anchor = torch.div(j, chunk, rounding_mode="floor") * chunk
            ok = (label[anchor] == label).all() & (B % chunk == 0)
            torch._assert_async(ok, "labels are not P x K contiguous (OCFR.py:31-42: K instances per identity in a row)")
        return label

    def _droppath(self, B, device):
        """Per-sample keep/keep_prob factors, one vector of length 3B (sequence s = m*B+b) per residual branch: 24 vectors
        from ONE torch.rand((24, 3B)) per step.  Same distribution as the reference (floor(keep_prob + U[0,1)) / keep_prob
        per sample, per branch, per modality call; block 0 has rate 0 -> nn.Identity, vit_pytorch.py:52-69,210,215-220),
        but NOT the same random stream: the reference draws torch.rand((B,1,1)) 23 x 3 times in call order, so seeded runs
        agree in distribution only.  Parity tests inject the masks (test_droppath_matches_oracle) or set DROP_PATH 0."""
        rates = self.model.BACKBONE.base.drop_path_rates
        if max(rates) <= 0.0:
            return None
        keep = self.stats.get("droppath_keep")          # cached: a host->device copy is not allowed under graph capture
        if keep is None or keep.device != device:
            keep = 1.0 - torch.tensor(rates, dtype=torch.float32, device=device).repeat_interleave(2).unsqueeze(1)  # [24,1]
            self.stats["droppath_keep"] = keep
        rnd = torch.rand((24, 3 * B), dtype=torch.float32, device=device)     # one draw for the whole step
        scales = torch.floor(keep + rnd) / keep                               # keep_prob + rand, floor, / keep_prob
        return list(scales.unbind(0))

    def forward(self, x, cam_label, label, writer, epoch):
        m = self.model
        rgb, ni, ti = x["RGB"], x["NI"], x["TI"]
        self._ensure(rgb.device)
        for t in (rgb, ni, ti):
            if t.dtype != torch.float32 or not t.is_contiguous() or t.shape != rgb.shape:
                raise lib.EdbError("inputs must be contiguous float32 [B,3,H,W] CUDA tensors of one shape")
        if tuple(rgb.shape[2:]) != m.image_size:
            raise AssertionError("Input image size (%d*%d) doesn't match model (%d*%d)." %
                                 (rgb.shape[2], rgb.shape[3], m.image_size[0], m.image_size[1]))   # vit_pytorch.py:453-454
        B = rgb.shape[0]
        if cam_label is None:
            cam_label = torch.zeros(B, dtype=torch.int64, device=rgb.device)
        cam = cam_label.to(device=rgb.device, dtype=torch.int64).contiguous()
        training = m.training
        check = self._check_cam(cam, B, training)
        prec = self._precision(training)
        # model.precision = "fp32" trains fp32-faithfully (every GEMM as a 3-piece bf16 split, fp32 attention): the parity
        # mode of BASELINE.json configs[4]; ~6x the tensor work of the bf16 mode
        self.arena.refresh16()
        if not training:
            with torch.no_grad():
                tokens, sv = self.backbone_forward(rgb, ni, ti, cam, prec, keep=False)
                sel = self.select(rgb, ni, ti, sv["maps"], prec, want_debug=self.stats.get("debug", False))
                self.sel = sel
                cls_out, patch_mean, _, _, num, _ = self.hma_forward(tokens, sel, prec, False)
                self.last = dict(tokens=tokens, num=num)
                return self._reduce(cls_out, patch_mean, prec)
        ag = self.autograd_params()
        if ag:
            # autograd / DistributedDataParallel mode: the arena is scratch for ONE forward+backward, parameter gradients
            # leave through the autograd graph (AccumulateGrad -> DDP reducer hooks), p.grad is torch's business
            self._warn_pending()
            self.arena.grad.zero_()
        elif self.stats.get("trainer_owns_grads"):
            pass                            # editor_b200.train.Trainer zeroes the arena itself, once per step
        else:
            self._warn_pending()
            self.arena.prepare_grads()
        self.stats["pending_backward"] = True
        P = (lambda *names: tuple(self.arena.params[self.arena.names.index(n)] for n in names)) if ag else (lambda *n: ())
        label = self._check_labels(label, B, rgb.device, check)
        dp = self._droppath(B, rgb.device)
        # the cls tokens leave the backbone as a second output: slicing `tokens` instead would make autograd materialise
        # (zeros + scatter + add) a full [3B,129,768] gradient per slice -- 0.4 ms of fills and adds per step
        tokens, cls3 = _BackboneFn.apply(self, rgb, ni, ti, cam, prec, dp, ag, *self.bb_plist)
        cls_bb = [cls3[i] for i in range(3)]
        lin, BN, LIN = self.tail_lin, _tail.BatchNormFn.apply, _tail.LinearFn.apply
        if m.AL:
            ori = torch.cat(cls_bb, dim=-1)
            ori_score = LIN(self, lin["AL_HEAD"], prec, BN(self, "AL_BN", m.AL_BN, ori, *P("AL_BN.weight", "AL_BN.bias")),
                            *P("AL_HEAD.weight"))
        else:   # three separate BN calls, each with its own batch statistics
            scores = [LIN(self, lin["BACKBONE_HEAD"], prec,
                          BN(self, "BACKBONE_BN", m.BACKBONE_BN, c, *P("BACKBONE_BN.weight", "BACKBONE_BN.bias")),
                          *P("BACKBONE_HEAD.weight")) for c in cls_bb]
        cls_out, patch_mean, cls_mid, loss_bcc = _HMAFn.apply(self, tokens, prec, ag, *self.hma_plist)
        loss_ocfr = _tail.OcfrFn.apply(m.FUSE_block.memory_cls, cls_mid, label)        # label: checked int64 above
        if writer is not None:
            writer.add_scalar("num_count", self.last["num"].float().mean(), epoch)      # make_model.py:199-200
        cls4t = self._reduce(cls_out, patch_mean, prec, P)
        score = LIN(self, lin["FUSE_HEAD"], prec, BN(self, "FUSE_BN", m.FUSE_BN, cls4t, *P("FUSE_BN.weight", "FUSE_BN.bias")),
                    *P("FUSE_HEAD.weight"))
        aux = loss_bcc.reshape(()) + loss_ocfr.reshape(())
        if m.AL:
            return score, cls4t, ori_score, ori, aux
        return score, cls4t, scores[0], cls_bb[0], scores[1], cls_bb[1], scores[2], cls_bb[2], aux

    def _reduce(self, cls_out, patch_mean, prec, P=lambda *n: ()):
        """*_REDUCE(cat(cls, patch mean)) and the concatenation to cls4t [B, 2304] (make_model.py:205-208)."""
        outs = [_tail.LinearFn.apply(self, self.tail_lin[n], prec, torch.cat([cls_out[i], patch_mean[i]], dim=-1),
                                     *P(n + ".weight", n + ".bias"))
                for i, n in enumerate(("RGB_REDUCE", "NIR_REDUCE", "TIR_REDUCE"))]
        return torch.cat(outs, dim=-1)

    def autograd_params(self):
        """True -> parameter gradients are RETURNED by the autograd Functions instead of being exposed as views of the
        gradient arena.  Needed by ``torch.nn.parallel.DistributedDataParallel(find_unused_parameters=True)``
        (engine/processor.py:47-50): its reducer listens on every parameter's AccumulateGrad node.  ``model.param_grads``:
        "arena" (default single-process fast path, also what ``editor_b200.train.Trainer`` drives), "autograd", or
        "auto" = autograd exactly when torch.distributed is initialised with more than one rank and no Trainer hook is
        installed (i.e. somebody else -- DDP -- is responsible for the gradient exchange)."""
        mode = getattr(self.model, "param_grads", "auto")
        if mode == "auto":
            import torch.distributed as dist
            return bool(dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
                        and self.stats.get("grad_hook") is None)
        return mode == "autograd"


class _BackboneFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, rgb, ni, ti, cam, prec, dp, ag, *params):
        ctx.ag = ag
        eng._mark("bb_fwd_start")
        tokens, sv = eng.backbone_forward(rgb, ni, ti, cam, prec, keep=True, droppath=dp)
        eng._mark("bb_fwd_end")
        eng.sel = eng.select(rgb, ni, ti, sv["maps"], prec, want_debug=eng.stats.get("debug", False))
        eng._mark("select_end")
        ctx.eng, ctx.sv, ctx.nparams, ctx.prec = eng, sv, len(params), prec
        B = rgb.shape[0]
        cls3 = tokens.view(3, B, NTOK, DIM)[:, :, 0].contiguous()       # [3, B, 768]
        return tokens, cls3

    @staticmethod
    def backward(ctx, d_tokens, d_cls3):
        eng = ctx.eng
        eng._mark("bb_bwd_start")
        if d_tokens is None:
            B = d_cls3.shape[1]
            d_tokens = torch.zeros(3 * B, NTOK, DIM, dtype=torch.float32, device=d_cls3.device)
        else:
            d_tokens = d_tokens.contiguous().float()
        if d_cls3 is not None:
            # d_tokens is the tensor our own HMA backward allocated (or the zeros above): add the cls rows in place
            d_tokens.view(3, -1, NTOK, DIM)[:, :, 0] += d_cls3.float()
        eng.backbone_backward(ctx.sv, d_tokens)
        eng._mark("bb_bwd_end")
        eng.stats["pending_backward"] = False
        eng.arena.finish_grads()
        if ctx.ag:      # clones: AccumulateGrad may keep the tensor it is handed as p.grad, the arena is reused next step
            return (None,) * 8 + tuple(eng.arena.gview(n).clone() for n in eng.bb_names)
        eng.arena.attach_grads(set(eng.bb_names))
        return (None,) * (8 + ctx.nparams)


class _HMAFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, tokens, prec, ag, *params):
        ctx.ag = ag
        eng._mark("hma_fwd_start")
        cls_out, patch_mean, cls_mid, loss_bcc, num, sv = eng.hma_forward(tokens, eng.sel, prec, True)
        eng._mark("hma_fwd_end")
        # detached: a graph-attached tensor kept here would keep the whole previous step's autograd graph -- and with it the
        # AccumulateGrad node of every parameter, bound to the stream it was created on -- alive into the next step; under
        # CUDA-graph capture the engine then syncs the capture stream with that foreign "leaf stream"
        # (cudaErrorStreamCaptureIsolation)
        eng.last = dict(num=num, tokens=tokens.detach())
        ctx.eng, ctx.sv, ctx.sel, ctx.nparams, ctx.prec = eng, sv, eng.sel, len(params), prec
        ctx.save_for_backward(tokens)
        return cls_out, patch_mean, cls_mid, loss_bcc

    @staticmethod
    def backward(ctx, d_cls, d_patch, d_mid, d_bcc):
        eng = ctx.eng
        (tokens,) = ctx.saved_tensors
        z = lambda t, ref: torch.zeros_like(ref) if t is None else t.contiguous().float()   # noqa: E731
        shape_ref = torch.empty(3, ctx.sel["B"], DIM, device=tokens.device)
        eng._mark("hma_bwd_start")
        d_tokens = eng.hma_backward(tokens, ctx.sel, ctx.sv, z(d_cls, shape_ref), z(d_patch, shape_ref),
                                    None if d_mid is None else d_mid.contiguous().float(),
                                    None if d_bcc is None else d_bcc.contiguous().float())
        eng._mark("hma_bwd_end")
        if ctx.ag:
            return (None, d_tokens, None, None) + tuple(eng.arena.gview(n).clone() for n in eng.hma_names)
        eng.arena.attach_grads(set(eng.hma_names))
        return (None, d_tokens, None, None) + (None,) * ctx.nparams
