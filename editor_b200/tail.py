"""autograd wiring of the tail kernels (csrc/tail.cu): BNNeck BatchNorm1d, the REDUCE / classifier linears on the tcgen05
GEMM, OCFR, and the loss (label-smoothed CE + batch-hard soft-margin triplet).  Reference: make_model.py:162-171,
205-209; fusion_part/OCFR.py:44-84; layers/softmax_loss.py:23-34; layers/triplet_loss.py:122-136."""
import torch

from . import lib

BF16, FP32 = "bf16", "fp32"


def _pad8(n):
    return (n + 7) // 8 * 8


class LinearFn(torch.autograd.Function):
    """y = x W^T (+ b) for one of the small tail linears; W, b live in the parameter arena (L is an engine._Lin)."""

    @staticmethod
    def forward(ctx, eng, L, prec, x, *params):
        """params: () in the arena mode (gradients are written into the flat gradient arena and exposed through
        ``p.grad``), or (W[, b]) in the autograd mode (``engine.EditorEngine.autograd_params``: DistributedDataParallel
        needs every parameter gradient to come out of the autograd graph) -- then the backward RETURNS them."""
        ctx.nparams = len(params)
        B, K, N = x.shape[0], L.in_f, L.out_f
        x = x.contiguous().float()
        y = torch.empty(B, _pad8(N), dtype=torch.float32, device=x.device)
        if prec == BF16:
            xb = torch.empty(B, K, dtype=torch.bfloat16, device=x.device)
            lib.cast_bf16(x, xb)
            lib.gemm(xb, L.w16, y, B, N, K, bias=L.b)
        else:
            xb = torch.empty(B, 6 * K, dtype=torch.bfloat16, device=x.device)
            lib.split3(x, xb, 0)
            lib.gemm(xb, L.wsplit(), y, B, N, 6 * K, bias=L.b)
        ctx.eng, ctx.L, ctx.xb, ctx.prec, ctx.x32 = eng, L, xb, prec, (x if prec != BF16 else None)
        ctx.wshape = params[0].shape if params else None
        return y[:, :N]

    @staticmethod
    def backward(ctx, dy):
        eng, L, xb = ctx.eng, ctx.L, ctx.xb
        B, K, N = dy.shape[0], L.in_f, L.out_f
        Np = _pad8(N)
        dyf = torch.zeros(B, Np, dtype=torch.float32, device=dy.device)
        dyf[:, :N] = dy
        dx = torch.empty(B, K, dtype=torch.float32, device=dy.device)
        # autograd mode: this call's own contribution goes to fresh tensors (a shared head is called three times per step)
        gw = L.gw if ctx.nparams == 0 else torch.zeros_like(L.gw)
        gb = L.gb if (ctx.nparams == 0 or L.gb is None) else torch.zeros_like(L.gb)
        if ctx.prec == BF16:
            dyb = torch.empty(B, Np, dtype=torch.bfloat16, device=dy.device)
            lib.cast_bf16(dyf, dyb)
            lib.gemm(dyb, L.w16, dx, B, K, N, b_mn=True)
            lib.gemm(dyb, xb, gw, N, K, B, a_mn=True, b_mn=True, epilogue=lib.EPI_ATOMIC)
        else:       # fp32-faithful: 3-piece splits along the reduction dimension of each product
            ds = torch.empty(B, 6 * Np, dtype=torch.bfloat16, device=dy.device)
            lib.split3(dyf, ds, 0)
            lib.gemm(ds, L.wsplit_rows(Np), dx, B, K, 6 * Np, b_mn=True)
            da = torch.empty(6 * B, Np, dtype=torch.bfloat16, device=dy.device)
            xr = torch.empty(6 * B, K, dtype=torch.bfloat16, device=dy.device)
            lib.split3(dyf, da, 2)
            lib.split3(ctx.x32, xr, 3)
            lib.gemm(da, xr, gw, N, K, 6 * B, a_mn=True, b_mn=True, epilogue=lib.EPI_ATOMIC)
        names = {L.wname}
        if gb is not None:
            lib.colsum(dyf, gb, B, N)
            names.add(L.bname)
        if ctx.nparams:
            return (None, None, None, dx, gw.view(ctx.wshape)) + ((gb,) if ctx.nparams > 1 else ())
        eng.arena.attach_grads(names)
        return None, None, None, dx


class BatchNormFn(torch.autograd.Function):
    """nn.BatchNorm1d in training mode on [B, F]; `bn` is the nn.BatchNorm1d module holding the buffers."""

    @staticmethod
    def forward(ctx, eng, name, bn, x, *params):
        ctx.nparams = len(params)          # (weight, bias) in the autograd mode, see LinearFn
        a = eng.arena
        x = x.contiguous().float()
        B, F = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(F, dtype=torch.float32, device=x.device)
        invstd = torch.empty(F, dtype=torch.float32, device=x.device)
        lib.call("edb_bn1d_fwd", x.data_ptr(), F, B, F, a.view(name + ".weight").data_ptr(), a.view(name + ".bias").data_ptr(),
                 bn.running_mean.data_ptr(), bn.running_var.data_ptr(), float(bn.momentum), float(bn.eps), y.data_ptr(), F,
                 mean.data_ptr(), invstd.data_ptr(), lib.stream_ptr())
        bn.num_batches_tracked.add_(1)
        ctx.eng, ctx.name, ctx.saved = eng, name, (x, mean, invstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        eng, name = ctx.eng, ctx.name
        a = eng.arena
        x, mean, invstd = ctx.saved
        B, F = x.shape
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        if ctx.nparams:
            dg, db = torch.zeros(F, dtype=torch.float32, device=x.device), torch.zeros(F, dtype=torch.float32, device=x.device)
        else:
            dg, db = a.gview(name + ".weight"), a.gview(name + ".bias")
        lib.call("edb_bn1d_bwd", dy.data_ptr(), F, x.data_ptr(), F, B, F, a.view(name + ".weight").data_ptr(), mean.data_ptr(),
                 invstd.data_ptr(), dx.data_ptr(), F, dg.data_ptr(), db.data_ptr(), lib.stream_ptr())
        if ctx.nparams:
            return None, None, None, dx, dg, db
        a.attach_grads({name + ".weight", name + ".bias"})
        return None, None, None, dx


class OcfrFn(torch.autograd.Function):
    """OCFR.forward on the HMA cls tokens [3, B, 768]; updates the three centre banks in place, returns the loss [1]."""

    @staticmethod
    def forward(ctx, mem, cls_mid, label):
        x = cls_mid.contiguous().float()
        B = x.shape[1]
        C = mem.RGB_centers.shape[0]
        fn = torch.empty_like(x)
        inv = torch.empty(3, B, dtype=torch.float32, device=x.device)
        loss = torch.zeros(1, dtype=torch.float32, device=x.device)
        label = label.contiguous()
        lib.call("edb_ocfr_fwd", x.data_ptr(), label.data_ptr(), B, C, mem.RGB_centers.data_ptr(), mem.NIR_centers.data_ptr(),
                 mem.TIR_centers.data_ptr(), float(mem.momentum), fn.data_ptr(), inv.data_ptr(), loss.data_ptr(),
                 lib.stream_ptr())
        ctx.mem, ctx.saved = mem, (fn, inv, label)
        return loss

    @staticmethod
    def backward(ctx, g):
        mem = ctx.mem
        fn, inv, label = ctx.saved
        dx = torch.empty_like(fn)
        g = g.contiguous().float()
        lib.call("edb_ocfr_bwd", fn.data_ptr(), inv.data_ptr(), label.data_ptr(), fn.shape[1], mem.RGB_centers.data_ptr(),
                 mem.NIR_centers.data_ptr(), mem.TIR_centers.data_ptr(), g.data_ptr(), dx.data_ptr(), lib.stream_ptr())
        return None, dx, None


class CeSmoothFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, label, eps):
        z = logits.float()
        if z.stride(1) != 1:
            z = z.contiguous()
        B, C = z.shape
        loss = torch.zeros(1, dtype=torch.float32, device=z.device)
        dz = torch.empty(B, C, dtype=torch.float32, device=z.device)
        lib.call("edb_ce_smooth", z.data_ptr(), z.stride(0), label.data_ptr(), B, C, float(eps), loss.data_ptr(), dz.data_ptr(),
                 C, lib.stream_ptr())
        ctx.dz = dz
        return loss

    @staticmethod
    def backward(ctx, g):
        dz = ctx.dz
        out = torch.empty_like(dz)
        lib.call("edb_scale_by", dz.data_ptr(), g.contiguous().float().data_ptr(), out.data_ptr(), dz.numel(), lib.stream_ptr())
        return out, None, None


class TripletFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, label):
        x = feat.contiguous().float()
        B, F = x.shape
        nbytes = lib.load().edb_triplet_workspace_bytes(B)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        loss = torch.zeros(1, dtype=torch.float32, device=x.device)
        lib.call("edb_triplet_fwd", x.data_ptr(), F, label.data_ptr(), B, F, loss.data_ptr(), ws.data_ptr(), nbytes,
                 lib.stream_ptr())
        ctx.saved = (x, ws)
        return loss

    @staticmethod
    def backward(ctx, g):
        x, ws = ctx.saved
        B, F = x.shape
        dx = torch.empty_like(x)
        lib.call("edb_triplet_bwd", x.data_ptr(), F, B, F, ws.data_ptr(), g.contiguous().float().data_ptr(), dx.data_ptr(), F, 0,
                 lib.stream_ptr())
        return dx, None


def editor_loss(outputs, label, id_w=1.0, tri_w=1.0, eps=0.1):
    """engine/processor.py:82-92 over layers/make_loss.py:36-56 (softmax_triplet sampler, label smoothing on)."""
    label = label.to(torch.int64).contiguous()
    total = outputs[-1].float().reshape(())
    for i in range(0, len(outputs) - 1, 2):
        total = total + id_w * CeSmoothFn.apply(outputs[i], label, eps).reshape(()) \
            + tri_w * TripletFn.apply(outputs[i + 1], label).reshape(())
    return total
