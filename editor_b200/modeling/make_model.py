"""Drop-in model surface of the reference's ``modeling/make_model.py``: ``make_model(cfg, num_class, camera_num)``
(:371-374; ``build_model`` is an alias) returning an ``EDITOR`` ``nn.Module`` with

* the reference's exact ``state_dict`` schema (222 keys for RGBNT201) -- the sub-modules below are
  parameter holders with the reference's names, shapes and initialisers (make_model.py:10-31,86-141;
  vit_pytorch.py:261-307,420-531; OCFR.py:14-16; pytorch_wavelets/dwt/transform2d.py:36-40,101-105);
* the reference's ``forward`` signature and return tuples (make_model.py:150-258);
* all arithmetic executed by hand-written sm_100a CUDA behind the C ABI (``editor_b200.engine``).  There is no CPU or
  eager fallback: on a machine without the built library / a CUDA device ``forward`` raises.
"""
import math

import torch
import torch.nn as nn

from .. import engine as _engine


def _trunc_normal_(t, std=0.02):
    return nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0)


class _Attention(nn.Module):
    def __init__(self, dim, bias):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=bias)
        self.proj = nn.Linear(dim, dim, bias=bias)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden, bias):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden, bias=bias)
        self.fc2 = nn.Linear(hidden, dim, bias=bias)


class _Block(nn.Module):
    """Parameter holder of vit_pytorch.py:201-213 (Block)."""

    def __init__(self, dim, eps):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = _Attention(dim, True)
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = _Mlp(dim, 4 * dim, True)


class _PatchEmbed(nn.Module):
    def __init__(self, img_size, stride, dim):
        super().__init__()
        self.num_x = (img_size[1] - 16) // stride + 1
        self.num_y = (img_size[0] - 16) // stride + 1
        self.num_patches = self.num_x * self.num_y
        self.proj = nn.Conv2d(3, dim, kernel_size=16, stride=stride)
        n = 16 * 16 * dim
        self.proj.weight.data.normal_(0, math.sqrt(2.0 / n))          # vit_pytorch.py:438-441


class _Trans(nn.Module):
    """Parameter holder of vit_pytorch.py:463-531 (Trans, ViT-B/16)."""

    def __init__(self, img_size, stride, camera, drop_path_rate, sie_coe, dim=768, depth=12):
        super().__init__()
        self.patch_embed = _PatchEmbed(img_size, stride, dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, dim))
        self.cam_num = camera
        self.sie_xishu = sie_coe
        if camera > 1:
            self.sie_embed = nn.Parameter(torch.zeros(camera, 1, dim))
            _trunc_normal_(self.sie_embed)
        self.drop_path_rates = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]   # :511
        self.blocks = nn.ModuleList([_Block(dim, 1e-6) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.fc = nn.Linear(dim, 1000)                               # never used by EDITOR (:522), kept for the schema
        _trunc_normal_(self.cls_token)
        _trunc_normal_(self.pos_embed)
        self.apply(_init_vit)

    def load_param(self, model_path):
        """ImageNet checkpoint loading with pos-embed resize (vit_pytorch.py:646-690)."""
        param_dict = torch.load(model_path, map_location="cpu")
        for key in ("model", "state_dict"):
            if key in param_dict:
                param_dict = param_dict[key]
        own = self.state_dict()
        for k, v in param_dict.items():
            if "head" in k or "dist" in k:
                continue
            if "patch_embed.proj.weight" in k and v.dim() < 4:
                O, I, H, W = self.patch_embed.proj.weight.shape
                v = v.reshape(O, -1, H, W)
            elif k == "pos_embed" and v.shape != self.pos_embed.shape:
                if "distilled" in model_path:
                    v = torch.cat([v[:, 0:1], v[:, 2:]], dim=1)
                v = resize_pos_embed(v, self.pos_embed, self.patch_embed.num_y, self.patch_embed.num_x)
            if k in own and own[k].shape == v.shape:
                own[k].copy_(v)
            else:
                print("shape do not match in k :{}".format(k))


def resize_pos_embed(posemb, posemb_new, height, width):
    """vit_pytorch.py:674-690: bilinear resize of the checkpoint's square grid to (height, width)."""
    tok, grid = posemb[:, :1], posemb[0, 1:]
    gs = int(math.sqrt(len(grid)))
    grid = grid.reshape(1, gs, gs, -1).permute(0, 3, 1, 2)
    grid = nn.functional.interpolate(grid, size=(height, width), mode="bilinear")
    grid = grid.permute(0, 2, 3, 1).reshape(1, height * width, -1)
    return torch.cat([tok, grid], dim=1)


def _init_vit(m):
    if isinstance(m, nn.Linear):
        _trunc_normal_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)


class build_transformer(nn.Module):
    """make_model.py:34-83."""

    def __init__(self, num_classes, cfg, camera_num):
        super().__init__()
        if cfg.MODEL.TRANSFORMER_TYPE not in ("vit_base_patch16_224", "deit_base_patch16_224"):
            raise NotImplementedError("editor_b200 implements the ViT-B/16 backbone only (got %s)"
                                      % cfg.MODEL.TRANSFORMER_TYPE)
        self.token_dim = 768
        cams = camera_num if cfg.MODEL.SIE_CAMERA else 0
        stride = cfg.MODEL.STRIDE_SIZE[0] if isinstance(cfg.MODEL.STRIDE_SIZE, (list, tuple)) else cfg.MODEL.STRIDE_SIZE
        if stride != 16:
            raise NotImplementedError("only STRIDE_SIZE 16 (the value of every shipped config) is implemented")
        self.base = _Trans(tuple(cfg.INPUT.SIZE_TRAIN), stride, cams, cfg.MODEL.DROP_PATH, cfg.MODEL.SIE_COE)
        if cfg.MODEL.PRETRAIN_CHOICE == "imagenet":
            self.base.load_param(cfg.MODEL.PRETRAIN_PATH_T)
            print("Loading pretrained ImageNet model......from {}".format(cfg.MODEL.PRETRAIN_PATH_T))


class _HaarBank(nn.Module):
    """Buffers of pytorch_wavelets DWTForward/DWTInverse('haar') kept for state_dict compatibility
    (transform2d.py:36-40,101-105); the kernels use the closed form."""

    def __init__(self, prefix):
        super().__init__()
        s = 1.0 / math.sqrt(2.0)
        lo, hi = torch.tensor([s, s]), torch.tensor([s, -s])
        self.register_buffer(prefix + "0_col", lo.reshape(1, 1, 2, 1).clone())
        self.register_buffer(prefix + "1_col", hi.reshape(1, 1, 2, 1).clone())
        self.register_buffer(prefix + "0_row", lo.reshape(1, 1, 1, 2).clone())
        self.register_buffer(prefix + "1_row", hi.reshape(1, 1, 1, 2).clone())


class _FreqIndex(nn.Module):
    def __init__(self, keep, stride):
        super().__init__()
        self.DWT = _HaarBank("h")
        self.IDWT = _HaarBank("g")
        self.keep = keep
        self.stride = stride


class _OCFR(nn.Module):
    """State of OCFR.py:9-20."""

    def __init__(self, dim, num_class, momentum):
        super().__init__()
        self.RGB_centers = nn.Parameter(torch.zeros(num_class, dim), requires_grad=False)
        self.NIR_centers = nn.Parameter(torch.zeros(num_class, dim), requires_grad=False)
        self.TIR_centers = nn.Parameter(torch.zeros(num_class, dim), requires_grad=False)
        self.momentum = momentum


class _BlockMask(nn.Module):
    """Parameter holder of HMA (vit_pytorch.py:261-307); all linears bias-free (make_model.py:98)."""

    def __init__(self, dim, num_class, momentum):
        super().__init__()
        for n1, at, n2, ml in (("normR", "attnR", "normR_", "mlpR"), ("normN", "attnN", "normN_", "mlpN"),
                               ("normT", "attnT", "normT_", "mlpT"), ("norm1", "attn1", "norm2", "mlp")):
            setattr(self, n1, nn.LayerNorm(dim))
            setattr(self, at, _Attention(dim, False))
            setattr(self, n2, nn.LayerNorm(dim))
            setattr(self, ml, _Mlp(dim, 4 * dim, False))
        self.out_norm = nn.LayerNorm(dim)
        self.memory_cls = _OCFR(dim, num_class, momentum)
        self.apply(_init_vit)


def _kaiming_linear(m):
    nn.init.kaiming_normal_(m.weight, a=0, mode="fan_out")     # make_model.py:10-14
    nn.init.constant_(m.bias, 0.0)


class EDITOR(nn.Module):
    """make_model.py:86-258."""

    def __init__(self, num_classes, cfg, camera_num):
        super().__init__()
        self.BACKBONE = build_transformer(num_classes, cfg, camera_num)
        size, stride = cfg.INPUT.SIZE_TRAIN, cfg.MODEL.STRIDE_SIZE
        self.num_patches = int(size[0] // stride[0]) * int(size[1] // stride[1])
        self.ratio = (1 / self.num_patches) * int(cfg.MODEL.HEAD_KEEP)
        self.head_keep = int(self.num_patches * self.ratio)        # SFTS.py:155 int(N * ratio)
        self.SFTS = nn.Module()
        self.FREQ_INDEX = _FreqIndex(cfg.MODEL.FREQUENCY_KEEP, stride[0])
        self.FUSE_block = _BlockMask(768, num_classes, 0.8)
        for name in ("RGB_REDUCE", "NIR_REDUCE", "TIR_REDUCE"):
            lin = nn.Linear(2 * 768, 768)
            _kaiming_linear(lin)
            setattr(self, name, lin)
        self.FUSE_HEAD = nn.Linear(3 * 768, num_classes, bias=False)
        self.FUSE_BN = nn.BatchNorm1d(3 * 768)
        nn.init.normal_(self.FUSE_HEAD.weight, std=0.001)          # make_model.py:26-31
        self.BACKBONE_HEAD = nn.Linear(768, num_classes, bias=False)
        self.BACKBONE_BN = nn.BatchNorm1d(768)
        nn.init.normal_(self.BACKBONE_HEAD.weight, std=0.001)
        self.AL = cfg.MODEL.AL
        if self.AL:
            self.AL_HEAD = nn.Linear(3 * 768, num_classes, bias=False)
            self.AL_BN = nn.BatchNorm1d(3 * 768)
            nn.init.normal_(self.AL_HEAD.weight, std=0.001)
        self.image_size = (int(size[0]), int(size[1]))
        self.sie_coe = float(cfg.MODEL.SIE_COE)
        self.precision = "auto"     # "auto": bf16 under autocast / in training, fp32-faithful otherwise
        self._engine = None

    # ------------------------------------------------------------------ reference API
    def load_param(self, trained_path):
        """make_model.py:144-148."""
        param_dict = torch.load(trained_path)
        own = self.state_dict()
        for k in param_dict:
            own[k.replace("module.", "")].copy_(param_dict[k])
        print("Loading pretrained model from {}".format(trained_path))

    def engine(self):
        if self._engine is None:
            object.__setattr__(self, "_engine", _engine.EditorEngine(self))
        return self._engine

    def forward(self, x, cam_label=None, label=None, view_label=None, img_path=None, mode=1, writer=None, epoch=None):
        return self.engine().forward(x, cam_label, label, writer, epoch)


def make_model(cfg, num_class, camera_num):
    model = EDITOR(num_class, cfg, camera_num)
    print("===========Building EDITOR===========")
    return model


build_model = make_model
