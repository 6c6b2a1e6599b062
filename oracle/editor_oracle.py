"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain torch fp32) of the reference hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module, and only as the checker / the timed CPU baseline.  Nothing under ``editor_b200/`` imports it:
the product path fails loudly when ``libeditor_b200.so`` is missing.

Every function cites the reference file:line it follows (paths relative to the reference root, commit 473cab62).
The restatement is *pinned*: ``tests/golden/make_golden.py`` runs the UNMODIFIED reference (imported from
/root/reference in the build container) on seeded inputs/weights and commits its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those vectors.  ``torch.topk`` tie order is the one
third-party behaviour the results depend on: the rule used here -- keep every element strictly above
the k-th value, then fill with elements equal to it in ASCENDING index order -- was measured on a B200 with
torch 2.11 (tools/probe_topk.py, 4096/4096 rows in every dtype; summary in tests/golden/topk_probe_b200.json).

The state is a plain ``dict`` with exactly the reference's ``state_dict`` keys.
"""
import math

import torch
import torch.nn.functional as F

HEADS = 12
SIE_COE = 3.0          # config/defaults.py:54
EPS_BACKBONE = 1e-6    # vit_pytorch.py:699
EPS_HMA = 1e-5         # vit_pytorch.py:265 (nn.LayerNorm default)
OCFR_MOMENTUM = 0.8    # make_model.py:98


# --------------------------------------------------------------------------------------- selection primitives
def topk_mask(x, k):
    """bool mask of torch.topk(x, k, dim=1) -> sort -> scatter_ (SFTS.py:155-158, Frequency.py:58-62) with the
    CUDA tie rule (module docstring).  Only the *set* matters."""
    xs = x.double() if x.is_floating_point() else x.long()
    kth = torch.sort(xs, dim=1, descending=True).values[:, k - 1:k]
    gt = xs > kth
    eq = xs == kth
    need = k - gt.sum(1, keepdim=True)
    rank_eq = torch.cumsum(eq.long(), 1)
    return gt | (eq & (rank_eq <= need))


def haar_dwt_level(x):
    """One level of DWTForward(wave='haar', mode='zero') on even sizes (pytorch_wavelets/dwt/lowlevel.py:91-172,
    336-347; taps 1/sqrt2 from pywt 1.4.1 'haar').  Returns (ll, (lh, hl, hh))."""
    s = 1.0 / math.sqrt(2.0)
    a = x[..., 0::2, :]
    b = x[..., 1::2, :]
    lo_r = (a + b) * s          # column (height) low-pass
    hi_r = (a - b) * s
    def rows(t):
        c = t[..., :, 0::2]
        d = t[..., :, 1::2]
        return (c + d) * s, (c - d) * s
    ll, lh = rows(lo_r)
    hl, hh = rows(hi_r)
    return ll, (lh, hl, hh)


def haar_idwt_level(ll, highs):
    """Inverse of :func:`haar_dwt_level` (lowlevel.py:226-271, 671-680)."""
    s = 1.0 / math.sqrt(2.0)
    lh, hl, hh = highs
    def rows(lo, hi):
        out = torch.empty(lo.shape[:-1] + (lo.shape[-1] * 2,), dtype=lo.dtype)
        out[..., 0::2] = (lo + hi) * s
        out[..., 1::2] = (lo - hi) * s
        return out
    lo_r = rows(ll, lh)
    hi_r = rows(hl, hh)
    out = torch.empty(lo_r.shape[:-2] + (lo_r.shape[-2] * 2, lo_r.shape[-1]), dtype=ll.dtype)
    out[..., 0::2, :] = (lo_r + hi_r) * s
    out[..., 1::2, :] = (lo_r - hi_r) * s
    return out


def frequency_counts(rgb, ni, ti, stride=16, faithful=True, levels=4):
    """int32 [B, H/16 * W/16] -- positive-pixel count of each 16x16 window of IDWT(mean_m DWT(x_m)) averaged over
    channels (Frequency.py:42-56, 65-81).  ``faithful`` runs the 4-level Haar round trip; otherwise the
    algebraically identical pixel mean that the CUDA kernel uses."""
    if faithful:
        pyr = []
        for x in (rgb, ni, ti):
            ll, hs = x, []
            for _ in range(levels):
                ll, h = haar_dwt_level(ll)
                hs.append(h)
            pyr.append((ll, hs))
        low = (pyr[0][0] + pyr[1][0] + pyr[2][0]) / 3
        high = [tuple((pyr[0][1][j][c] + pyr[1][1][j][c] + pyr[2][1][j][c]) / 3 for c in range(3))
                for j in range(levels)]
        inv = low
        for j in reversed(range(levels)):
            inv = haar_idwt_level(inv, high[j])
    else:
        inv = (rgb + ni + ti) / 3
    img = inv.mean(dim=1)                                   # Frequency.py:44
    B, H, W = img.shape
    pos = (img > 0).to(torch.int32)                         # Frequency.py:52
    cnt = pos.reshape(B, H // stride, stride, W // stride, stride).sum(dim=(2, 4))
    return cnt.reshape(B, -1).to(torch.int32)


def frequency_mask(rgb, ni, ti, keep=10, stride=16, faithful=True):
    """bool [B,128] (Frequency.py:58-63)."""
    return topk_mask(frequency_counts(rgb, ni, ti, stride, faithful), int(keep))


def rollout_cls_row(attn_list):
    """[B,12,128] -- row 0 / columns 1: of A_11 @ ... @ A_0 (SFTS.py:148-153), via the row-vector chain of
    SURVEY.md App. A-2 (identical top-k sets; the faithful full product is :func:`rollout_full`)."""
    r = attn_list[-1][:, :, 0:1, :]
    for a in reversed(attn_list[:-1]):
        r = r @ a
    return r[:, :, 0, 1:]


def rollout_full(attn_list):
    last = attn_list[0]
    for a in attn_list[1:]:
        last = a @ last                                     # SFTS.py:150-152
    return last[:, :, 0, 1:]


def part_attention_mask(attn_list, head_keep=2, full=False):
    """bool [B,128]: per-head top-k of the rollout row, OR over heads (SFTS.py:153-162)."""
    row = rollout_full(attn_list) if full else rollout_cls_row(attn_list)
    B, Hh, N = row.shape
    k = int(N * ((1.0 / N) * int(head_keep)))               # make_model.py:93, SFTS.py:155
    out = torch.zeros(B, N, dtype=torch.bool)
    for h in range(Hh):
        out |= topk_mask(row[:, h, :], k)
    return out


# --------------------------------------------------------------------------------------- backbone
def _ln(x, sd, name, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def attention(x, sd, pre, bias=True):
    """Attention.forward (vit_pytorch.py:184-198); returns (out, post-softmax map)."""
    B, N, C = x.shape
    qkv = F.linear(x, sd[pre + "qkv.weight"], sd.get(pre + "qkv.bias") if bias else None)
    qkv = qkv.reshape(B, N, 3, HEADS, C // HEADS).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * ((C // HEADS) ** -0.5)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B, N, C)
    out = F.linear(out, sd[pre + "proj.weight"], sd.get(pre + "proj.bias") if bias else None)
    return out, attn


def mlp(x, sd, pre, bias=True):
    """Mlp.forward (vit_pytorch.py:139-145); exact-erf GELU."""
    h = F.gelu(F.linear(x, sd[pre + "fc1.weight"], sd.get(pre + "fc1.bias") if bias else None))
    return F.linear(h, sd[pre + "fc2.weight"], sd.get(pre + "fc2.bias") if bias else None)


def backbone(img, cam, sd, droppath=None, depth=12):
    """Trans.forward (vit_pytorch.py:623-644) with PatchEmbed_overlap (:455-457), camera SIE (:632-633).
    ``droppath``: optional list of 2*depth per-sample scale vectors [B] (keep/keep_prob of vit_pytorch.py:52-69)."""
    b = "BACKBONE.base."
    x = F.conv2d(img, sd[b + "patch_embed.proj.weight"], sd[b + "patch_embed.proj.bias"], stride=16)
    x = x.flatten(2).transpose(1, 2)
    B = x.shape[0]
    x = torch.cat([sd[b + "cls_token"].expand(B, -1, -1), x], dim=1)
    x = x + sd[b + "pos_embed"]
    if (b + "sie_embed") in sd:
        x = x + SIE_COE * sd[b + "sie_embed"][cam]
    maps = []
    for i in range(depth):
        p = b + "blocks.%d." % i
        a, m = attention(_ln(x, sd, p + "norm1", EPS_BACKBONE), sd, p + "attn.")
        if droppath is not None:
            a = a * droppath[2 * i].view(-1, 1, 1)
        x = x + a
        h = mlp(_ln(x, sd, p + "norm2", EPS_BACKBONE), sd, p + "mlp.")
        if droppath is not None:
            h = h * droppath[2 * i + 1].view(-1, 1, 1)
        x = x + h
        maps.append(m)
    return _ln(x, sd, b + "norm", EPS_BACKBONE), maps


# --------------------------------------------------------------------------------------- SFTS
def sfts(feats, maps, mask_fre, head_keep=2, training=False, full_rollout=False, force_index=None):
    """SFTS.forward (SFTS.py:181-230).  feats/maps: 3-lists (RGB, NIR, TIR).  Returns (3 masked feats, index, bcc).
    ``force_index`` (tests only): use this bool [B,128] selection instead of the computed one, so that a reduced-
    precision run whose top-k flipped on a near-tie can still be compared arithmetically."""
    idx = mask_fre.clone()
    for m in maps:
        idx |= part_attention_mask(m, head_keep, full_rollout)
    if force_index is not None:
        idx = force_index.clone()
    index = idx.unsqueeze(-1)
    out = [torch.cat([f[:, :1], f[:, 1:] * index], dim=1) for f in feats]
    loss = None
    if training:
        bg = [f[:, 1:] * (~index) for f in feats]
        loss = F.mse_loss(bg[0], bg[1]) + F.mse_loss(bg[0], bg[2]) + F.mse_loss(bg[1], bg[2])
    return out, index, loss


# --------------------------------------------------------------------------------------- HMA
def attention_mask(x, mask, sd, pre):
    """AttentionMask.forward (vit_pytorch.py:240-258); bias-free linears (make_model.py:98)."""
    B, N, C = x.shape
    if N != mask.shape[1]:
        mask = mask.repeat(1, N // mask.shape[1], 1)
    x = x * mask
    qkv = F.linear(x, sd[pre + "qkv.weight"]).reshape(B, N, 3, HEADS, C // HEADS).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * ((C // HEADS) ** -0.5)
    m4 = mask.unsqueeze(1).repeat(1, HEADS, 1, 1)
    attn = attn.masked_fill((m4 @ m4.transpose(-2, -1)) == 0, -65504.0)
    attn = attn.softmax(dim=-1) * m4
    out = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(out, sd[pre + "proj.weight"])


def mlp_masked(x, mask, sd, pre):
    """MlpMasked.forward (vit_pytorch.py:158-168)."""
    if x.shape[1] != mask.shape[1]:
        mask = mask.repeat(1, x.shape[1] // mask.shape[1], 1)
    x = x * mask
    return F.linear(F.gelu(F.linear(x, sd[pre + "fc1.weight"])), sd[pre + "fc2.weight"])


def ocfr(cls3, label, centers, momentum=OCFR_MOMENTUM):
    """OCFR.forward/update/compute_center/compute_intra_loss (OCFR.py:22-84).  ``centers``: list of 3 [C,768]
    tensors updated IN PLACE (EMA before the loss, :53).  Labels must be P x K contiguous (OCFR.py:33-36)."""
    feats = [F.normalize(c, dim=1) for c in cls3]
    uniq = label.unique()
    for f, cen in zip(feats, centers):
        bc = torch.stack([f[label == u].mean(dim=0) for u in uniq], dim=0).detach()
        cen[uniq] = momentum * bc + (1 - momentum) * cen[uniq]
    loss = 0
    chunk = label.shape[0] // uniq.shape[0]
    lab = label[::chunk]
    for f, cen in zip(feats, centers):
        cm = cen[uniq]
        rows = torch.stack([cm[uniq == lab[i]].repeat(chunk, 1) for i in range(lab.shape[0])], 0)
        loss = loss + F.mse_loss(rows.reshape(-1, f.shape[1]), f)
    return loss


def hma(feats, index, sd, label=None, training=False, centers=None):
    """BlockMask.forward (vit_pytorch.py:309-352).  Returns (x [B,387,768], ocfr loss or None)."""
    f = "FUSE_block."
    B = feats[0].shape[0]
    mask = torch.cat([torch.ones(B, 1, 1), index.float()], dim=1)
    xs = []
    for x, (n1, at, n2, ml) in zip(feats, (("normR", "attnR", "normR_", "mlpR"), ("normN", "attnN", "normN_", "mlpN"),
                                           ("normT", "attnT", "normT_", "mlpT"))):
        x = x + attention_mask(_ln(x, sd, f + n1, EPS_HMA), mask, sd, f + at + ".")
        x = x + mlp_masked(_ln(x, sd, f + n2, EPS_HMA), mask, sd, f + ml + ".")
        xs.append(x)
    loss = None
    if training:
        loss = ocfr([x[:, 0] for x in xs], label, centers)
    x = torch.cat(xs, dim=1)
    x = x + attention_mask(_ln(x, sd, f + "norm1", EPS_HMA), mask, sd, f + "attn1.")
    x = x + mlp_masked(_ln(x, sd, f + "norm2", EPS_HMA), mask, sd, f + "mlp.")
    x = _ln(x, sd, f + "out_norm", EPS_HMA) * mask.repeat(1, 3, 1)
    return x, loss


# --------------------------------------------------------------------------------------- heads
def _bn(x, sd, name, training, state_out):
    """nn.BatchNorm1d (make_model.py:114-141): train = batch stats + running-stat update (momentum .1, unbiased var)."""
    if training:
        y = F.batch_norm(x, None, None, sd[name + ".weight"], sd[name + ".bias"], True, 0.1, 1e-5)
        if state_out is not None:
            n = x.shape[0]
            rm = state_out.get(name + ".running_mean", sd[name + ".running_mean"])
            rv = state_out.get(name + ".running_var", sd[name + ".running_var"])
            state_out[name + ".running_mean"] = 0.9 * rm + 0.1 * x.detach().mean(0)
            state_out[name + ".running_var"] = 0.9 * rv + 0.1 * x.detach().var(0, unbiased=True) if n > 1 else rv
            nb = state_out.get(name + ".num_batches_tracked", sd[name + ".num_batches_tracked"])
            state_out[name + ".num_batches_tracked"] = nb + 1
        return y
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], False, 0.1, 1e-5)


def pool_reduce(x, sd):
    """make_model.py:186-208 / 237-257: split, cls + patch-sum / count(RGB rows != 0), *_REDUCE, concat."""
    N = x.shape[1] // 3
    parts = [x[:, i * N:(i + 1) * N] for i in range(3)]
    num = (parts[0][:, 1:].sum(dim=2) != 0).sum(dim=1).unsqueeze(-1)
    outs = []
    for p, name in zip(parts, ("RGB", "NIR", "TIR")):
        v = torch.cat([p[:, 0], p[:, 1:].sum(dim=1) / num], dim=-1)
        outs.append(F.linear(v, sd[name + "_REDUCE.weight"], sd[name + "_REDUCE.bias"]))
    return torch.cat(outs, dim=-1), num


def editor_forward(sd, x, cam_label, label=None, training=False, al=True, head_keep=2, freq_keep=10,
                   faithful=True, droppath=None, state_out=None, aux=None, force_index=None):
    """EDITOR.forward (make_model.py:150-258).  ``x`` = {'RGB','NI','TI'} float32 [B,3,H,W] CPU tensors.

    eval  -> cls4t [B,2304]
    train -> (score, cls4t, ori_score, ori, loss) if ``al`` else
             (score, cls4t, s_R, cls_R, s_N, cls_N, s_T, cls_T, loss); BN running stats / OCFR centres that the
             reference updates in place are written to ``state_out``.
    ``aux`` (dict) receives intermediates for the parity tests: index, mask_fre, tokens, hma output, num.
    ``droppath``: None or 3 lists (one per modality call, RGB/NI/TI order) of 24 per-sample scale vectors.
    """
    imgs = [x["RGB"], x["NI"], x["TI"]]
    mask_fre = frequency_mask(imgs[0], imgs[1], imgs[2], freq_keep, 16, faithful)
    feats, maps = [], []
    for m, img in enumerate(imgs):
        t, a = backbone(img, cam_label, sd, None if droppath is None else droppath[m])
        feats.append(t)
        maps.append(a)
    cls_bb = [t[:, 0] for t in feats]
    if training:
        if al:
            ori = torch.cat(cls_bb, dim=-1)
            ori_score = F.linear(_bn(ori, sd, "AL_BN", True, state_out), sd["AL_HEAD.weight"])
        else:
            bb_scores = []
            for c in cls_bb:                                  # three separate BN calls
                cur = dict(sd)
                if state_out is not None:
                    cur.update({k: v for k, v in state_out.items() if k.startswith("BACKBONE_BN")})
                bb_scores.append(F.linear(_bn(c, cur, "BACKBONE_BN", True, state_out), sd["BACKBONE_HEAD.weight"]))
    feats_s, index, loss_bcc = sfts(feats, maps, mask_fre, head_keep, training, force_index=force_index)
    centers = None
    if training:
        centers = [sd["FUSE_block.memory_cls.%s_centers" % m].clone() for m in ("RGB", "NIR", "TIR")]
    xh, loss_ocfr = hma(feats_s, index, sd, label, training, centers)
    cls4t, num = pool_reduce(xh, sd)
    if aux is not None:
        aux.update(index=index[..., 0], mask_fre=mask_fre, tokens=feats, hma=xh, num=num[:, 0],
                   loss_bcc=loss_bcc, loss_ocfr=loss_ocfr)
    if not training:
        return cls4t
    if state_out is not None:
        for m, c in zip(("RGB", "NIR", "TIR"), centers):
            state_out["FUSE_block.memory_cls.%s_centers" % m] = c
    score = F.linear(_bn(cls4t, sd, "FUSE_BN", True, state_out), sd["FUSE_HEAD.weight"])
    loss = loss_bcc + loss_ocfr
    if al:
        return score, cls4t, ori_score, ori, loss
    return (score, cls4t, bb_scores[0], cls_bb[0], bb_scores[1], cls_bb[1], bb_scores[2], cls_bb[2], loss)


# --------------------------------------------------------------------------------------- loss (row f-1 of SURVEY 8f)
def label_smooth_ce(logits, target, eps=0.1):
    """CrossEntropyLabelSmooth (layers/softmax_loss.py:23-34)."""
    C = logits.shape[1]
    logp = F.log_softmax(logits, dim=1)
    t = torch.zeros_like(logp).scatter_(1, target.unsqueeze(1), 1)
    t = (1 - eps) * t + eps / C
    return (-t * logp).mean(0).sum()


def triplet_soft_margin(feat, label):
    """TripletLoss(margin=None): batch-hard mining + SoftMarginLoss (layers/triplet_loss.py:16-31,51-105,122-136)."""
    xx = feat.pow(2).sum(1, keepdim=True)
    dist = xx + xx.t() - 2 * feat @ feat.t()
    dist = dist.clamp(min=1e-12).sqrt()
    same = label.unsqueeze(0) == label.unsqueeze(1)
    ap = torch.where(same, dist, dist.new_full((), -1e30)).max(1).values
    an = torch.where(~same, dist, dist.new_full((), 1e30)).min(1).values
    return F.soft_margin_loss(an - ap, torch.ones_like(an))


def reference_loss(outputs, label, id_w=1.0, tri_w=1.0):
    """engine/processor.py:82-92 over layers/make_loss.py:36-56 (sampler softmax_triplet, label smoothing on,
    ID_LOSS_WEIGHT = TRIPLET_LOSS_WEIGHT = 1.0, config/defaults.py:29-30): sum over (score, feat) pairs of
    CE_smooth + soft-margin triplet, plus the trailing aux loss."""
    total = 0
    for i in range(0, len(outputs) - 1, 2):
        total = total + id_w * label_smooth_ce(outputs[i], label) + tri_w * triplet_soft_margin(outputs[i + 1], label)
    return total + outputs[-1]
