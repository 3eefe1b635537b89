"""ViT / DeiT with the reference's module tree, factories, flag methods and forward contract
(reference: models/vit_fquant.py:70-1074, models/layers_quant.py:153-497).

    model = deit_small_patch16_224(cfg=Config())
    model.model_open_calibrate(); model.model_open_last_calibrate(); model(calib_images)
    model.model_close_calibrate(); model.model_quant()
    logits, FLOPs, global_distance = model(images, bit_config)          # bit_config: 1 + 4*depth + 1 entries of 4|8

Before `model_quant()` the forward is the calibration / FP forward, module by module (observers record,
values stay fp32).  After it, `forward` runs the integer engine (p2vit_b200/engine.py): weights packed to int8
once per bit_config, int8 activations end to end, fused tcgen05 GEMM epilogues.  `forward_eager` keeps the
module-by-module quantized evaluation (each Q-module = one kernel) for per-operator checks.
State-dict keys equal the reference's, so DeiT `.pth` / converted ViT `.npz` weights load unchanged.
"""
import math
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from .ptq import QAct, QConv2d, QIntLayerNorm, QIntSoftmax, QLinear
from .ptq.layers import fp_linear
from .ptq.observer.utils import allreduce_, pot_exponent

__all__ = ["deit_tiny_patch16_224", "deit_small_patch16_224", "deit_base_patch16_224", "vit_base_patch16_224",
           "vit_large_patch16_224", "VisionTransformer"]

ATTN_ALPHA_POOL = [0.35]   # vit_fquant.py:37
MLP_ALPHA_POOL = [0.5]     # layers_quant.py:14
BIT_POOL = [4, 8]          # vit_fquant.py:38


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    with torch.no_grad():
        return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def _qact(cfg, quant, calibrate, ln=False):
    return QAct(quant=quant, calibrate=calibrate, bit_type=cfg.BIT_TYPE_A,
                calibration_mode=cfg.CALIBRATION_MODE_A_LN if ln else cfg.CALIBRATION_MODE_A,
                observer_str=cfg.OBSERVER_A_LN if ln else cfg.OBSERVER_A,
                quantizer_str=cfg.QUANTIZER_A_LN if ln else cfg.QUANTIZER_A)


def _qlinear(cfg, quant, calibrate, fin, fout, bias=True):
    return QLinear(fin, fout, bias=bias, quant=quant, calibrate=calibrate, bit_type=cfg.BIT_TYPE_W,
                   calibration_mode=cfg.CALIBRATION_MODE_W, observer_str=cfg.OBSERVER_W, quantizer_str=cfg.QUANTIZER_W)


class _SmoothedLinear:
    """PoT channel smoothing in front of a QLinear (vit_fquant.py:232-353, layers_quant.py:255-360):
    cs[c] = 2^round_ln(max|x|_c^alpha / max|W|_c^(1-alpha)); x/cs feeds qact0, W*cs is what gets quantized."""

    def _smooth_calibrate(self, x, qact0, lin, alpha_pool, global_distance, bit_config, extra):
        gmax = torch.abs(x).max(axis=1).values.max(axis=0).values
        allreduce_(gmax, "max")
        wmax = torch.abs(lin.weight).max(axis=0).values
        pool, loss_pool = [], [[] for _ in BIT_POOL]
        act_scale, act_zp, w_scale, w_zp = [], [], [], []
        self.best_scale, self.best_act_scale, self.best_act_zp = [], [], []
        self.best_weight_scale, self.best_weight_zp = [], []
        gt = None
        for alpha in alpha_pool:
            cs = 2 ** pot_exponent(gmax ** alpha / (wmax ** (1 - alpha)), "round")
            pool.append(cs)
            xs = x / cs.reshape(1, 1, -1)
            ws = lin.weight * cs.reshape(1, -1)
            gt = fp_linear(xs, ws, lin.bias)
            mid = qact0(xs)
            if qact0.last_calibrate:
                act_scale.append(qact0.quantizer.scale)
                act_zp.append(qact0.quantizer.zero_point)
                lin(mid, global_distance, bit_config, ws, **extra)
                w_scale.append(lin.quantizer.dic_scale)
                w_zp.append(lin.quantizer.dic_zero_point)
                if len(alpha_pool) > 1:  # the per-alpha output error only matters when there is a choice
                    qact0.calibrate, qact0.quant, lin.calibrate, lin.quant = False, True, False, True
                    mid = qact0(xs)
                    for j, bit in enumerate(BIT_POOL):
                        loss_pool[j].append((gt - lin(mid, global_distance, bit, ws, **extra)).abs().pow(2.0).mean())
                    qact0.calibrate, qact0.quant, lin.calibrate, lin.quant = True, False, True, False
                else:
                    for lp in loss_pool:
                        lp.append(torch.zeros((), device=x.device))
        if qact0.last_calibrate:
            for loss in loss_pool:
                idx = loss.index(min(loss))
                self.channel_scale = pool[idx]
                self.best_scale.append(pool[idx])
                self.best_act_scale.append(act_scale[idx])
                self.best_act_zp.append(act_zp[idx])
                self.best_weight_scale.append(w_scale[idx])
                self.best_weight_zp.append(w_zp[idx])
        return gt

    def _smooth_quant(self, x, qact0, lin, global_distance, bit_config, extra):
        idx = BIT_POOL.index(bit_config)
        self.channel_scale = self.best_scale[idx]
        qact0.quantizer.scale = self.best_act_scale[idx]
        qact0.quantizer.zero_point = self.best_act_zp[idx]
        lin.quantizer.dic_scale = self.best_weight_scale[idx]
        lin.quantizer.dic_zero_point = self.best_weight_zp[idx]
        h = qact0(x / self.channel_scale.reshape(1, 1, -1))
        return lin(h, global_distance, bit_config, lin.weight * self.channel_scale.reshape(1, -1), **extra)


class Mlp(nn.Module, _SmoothedLinear):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0,
                 quant=False, calibrate=False, cfg=None):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.qact0 = _qact(cfg, quant, calibrate)
        self.fc1 = _qlinear(cfg, quant, calibrate, in_features, hidden_features)
        self.act = act_layer()
        self.qact1 = _qact(cfg, quant, calibrate)
        self.fc2 = _qlinear(cfg, quant, calibrate, hidden_features, out_features)
        self.qact2 = _qact(cfg, quant, calibrate, ln=True)
        self.drop = nn.Dropout(drop)
        self.channel_scale = None

    def forward(self, x, FLOPs, global_distance, ffn_bit_config, plot=False, quant=True, smoothquant=True,
                activation=[], hessian_statistic=False):
        B, N, C = x.shape
        bit_config = ffn_bit_config[0] if ffn_bit_config else None
        if smoothquant and not hessian_statistic:
            if self.channel_scale is None:
                x = self._smooth_calibrate(x, self.qact0, self.fc1, MLP_ALPHA_POOL, global_distance, bit_config, {})
            else:
                x = self._smooth_quant(x, self.qact0, self.fc1, global_distance, bit_config, {})
        else:
            x = self.fc1(self.qact0(x), global_distance, bit_config, None)
        FLOPs.append(N * C * x.shape[-1])
        x = self.qact1(self.act(x), asymmetric=False)
        x = self.drop(x)
        B, N, C = x.shape
        bit_config = ffn_bit_config[1] if ffn_bit_config else None
        x = self.fc2(x, global_distance, bit_config)
        FLOPs.append(N * C * x.shape[-1])
        return self.drop(self.qact2(x))


class PatchEmbed(nn.Module):
    """Image to patch embedding (layers_quant.py:396-497)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None, quant=False,
                 calibrate=False, cfg=None):
        super().__init__()
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = QConv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, quant=quant, calibrate=calibrate,
                            bit_type=cfg.BIT_TYPE_W, calibration_mode=cfg.CALIBRATION_MODE_W, observer_str=cfg.OBSERVER_W,
                            quantizer_str=cfg.QUANTIZER_W)
        if norm_layer:
            self.qact_before_norm = _qact(cfg, quant, calibrate)
            self.norm = norm_layer(embed_dim)
        else:
            self.qact_before_norm = nn.Identity()
            self.norm = nn.Identity()
        self.qact = _qact(cfg, quant, calibrate)

    def forward(self, x, FLOPs, bit_config):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        x = self.proj(x, bit_config)
        B, M, H, W = x.shape
        FLOPs.append(C * self.patch_size[0] * self.patch_size[0] * M * H * W)
        x = x.flatten(2).transpose(1, 2)
        x = self.qact_before_norm(x)
        if isinstance(self.norm, nn.Identity):
            x = self.norm(x)
        else:
            x = self.norm(x, self.qact_before_norm.quantizer, self.qact.quantizer)
        return self.qact(x)


class Attention(nn.Module, _SmoothedLinear):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0, quant=False,
                 calibrate=False, cfg=None):
        super().__init__()
        self.num_heads = num_heads
        self.calibrate = calibrate
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.qkv = _qlinear(cfg, quant, calibrate, dim, dim * 3, bias=qkv_bias)
        self.qact0 = _qact(cfg, quant, calibrate)
        self.qact1 = _qact(cfg, quant, calibrate)
        self.qact2 = _qact(cfg, quant, calibrate)
        self.proj = _qlinear(cfg, quant, calibrate, dim, dim)
        self.qact3 = _qact(cfg, quant, calibrate, ln=True)
        self.qact_attn1 = _qact(cfg, quant, calibrate)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj_drop = nn.Dropout(proj_drop)
        self.log_int_softmax = QIntSoftmax(log_i_softmax=cfg.INT_SOFTMAX, quant=quant, calibrate=calibrate, bit_type=cfg.BIT_TYPE_S,
                                           calibration_mode=cfg.CALIBRATION_MODE_S, observer_str=cfg.OBSERVER_S,
                                           quantizer_str=cfg.QUANTIZER_S)
        self.channel_scale = None

    def forward(self, x, FLOPs, global_distance, atten_bit_config, plot=False, quant=False, smoothquant=True,
                hessian_statistic=False):
        self.atten_bit_config = atten_bit_config
        B, N, C = x.shape
        bit_config = atten_bit_config[0] if atten_bit_config else None
        extra = dict(attn=False, attn_para=[self.num_heads, C, self.scale])
        if smoothquant and not hessian_statistic:
            if self.channel_scale is None:
                x = self._smooth_calibrate(x, self.qact0, self.qkv, ATTN_ALPHA_POOL, global_distance, bit_config, extra)
            else:
                x = self._smooth_quant(x, self.qact0, self.qkv, global_distance, bit_config, extra)
        else:
            x = self.qkv(self.qact0(x), global_distance, bit_config, None, **extra)
        B, N, M = x.shape
        FLOPs.append(N * C * M)
        x = self.qact1(x, **extra)
        qkv = x.reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = self.qact_attn1((q @ k.transpose(-2, -1)) * self.scale)
        attn = self.log_int_softmax(attn, self.qact_attn1.quantizer.scale)
        attn = self.attn_drop(attn)
        x = self.qact2((attn @ v).transpose(1, 2).reshape(B, N, C))
        bit_config = atten_bit_config[1] if atten_bit_config else None
        x = self.proj(x, global_distance, bit_config)
        FLOPs.append(N * C * x.shape[-1])
        return self.proj_drop(self.qact3(x))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, quant=False, calibrate=False, cfg=None):
        super().__init__()
        assert drop_path == 0.0, "inference only: stochastic depth is the identity in eval mode"
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop, cfg=cfg)
        self.drop_path = nn.Identity()
        self.qact2 = _qact(cfg, quant, calibrate, ln=True)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop, quant=quant,
                       calibrate=calibrate, cfg=cfg)
        self.qact4 = _qact(cfg, quant, calibrate, ln=True)

    def forward(self, x, last_quantizer=None, FLOPs=[], global_distance=[], local_bit_config=None, plot=False, quant=False,
                hessian_statistic=False):
        abits = local_bit_config[0:2] if local_bit_config else None
        fbits = local_bit_config[2:4] if local_bit_config else None
        h = self.norm1(x, last_quantizer, self.attn.qact0.quantizer, self.attn.channel_scale)
        x = self.qact2(x + self.drop_path(self.attn(h, FLOPs, global_distance, abits, plot=False, quant=quant,
                                                    hessian_statistic=hessian_statistic)))
        # norm2 is given the *attention's* channel scale as its output smoothing scale (vit_fquant.py:565-570)
        h = self.norm2(x, self.qact2.quantizer, self.mlp.qact0.quantizer, self.attn.channel_scale)
        x = self.qact4(x + self.drop_path(self.mlp(h, FLOPs, global_distance, fbits, plot, quant, activation=[],
                                                   hessian_statistic=hessian_statistic)))
        return x


_Q_TYPES = (QConv2d, QLinear, QAct, QIntSoftmax)


class QuantModelMixin:
    """flag protocol and calibrated-state exchange shared by the ViT and Swin model classes"""

    pixel_norm = None

    def set_pixel_normalization(self, mean, std):
        """Declare the ToTensor + Normalize constants of the data pipeline (test_quant.py:112-127; p2vit_b200.data.PREPROCESS) so
        the quantized forward also accepts the decoder's uint8 pixels [B,3,H,W]: `model(x_u8, bit_config)` then equals
        `model(((x_u8.float() / 255) - mean) / std, bit_config)` bit for bit, with a quarter of the host-to-device bytes."""
        self.pixel_norm = (tuple(float(m) for m in mean), tuple(float(v) for v in std))
        return self

    # ---- flag protocol (vit_fquant.py:797-828)
    def model_quant(self, flag="on"):
        if flag == "on":
            self.quant = True
        for m in self.modules():
            if type(m) in _Q_TYPES:
                m.quant = True
            if self.cfg.INT_NORM and type(m) is QIntLayerNorm and flag != "off":
                m.mode = "int"
        self._engine = None

    def model_dequant(self):
        self.quant = False
        for m in self.modules():
            if type(m) in _Q_TYPES:
                m.quant = False
            if type(m) is QIntLayerNorm:
                m.mode = "ln"

    def model_open_calibrate(self):
        for m in self.modules():
            if type(m) in _Q_TYPES:
                m.calibrate = True

    def model_open_last_calibrate(self):
        for m in self.modules():
            if type(m) in _Q_TYPES:
                m.last_calibrate = True

    def model_close_calibrate(self):
        for m in self.modules():
            if type(m) in _Q_TYPES:
                m.calibrate = False

    # ---- calibrated state exchange (names = module paths; shared with oracle/ and tests/golden)
    def export_quant_state(self):
        st = {}
        for name, m in self.named_modules():
            if isinstance(m, QAct) and m.quantizer.scale is not None:
                st[name + ".scale"] = m.quantizer.scale.detach().reshape(-1).float().cpu()
                st[name + ".zero_point"] = m.quantizer.zero_point.detach().reshape(-1).long().cpu()
            elif isinstance(m, (QLinear, QConv2d)):
                for bit, s in m.quantizer.dic_scale.items():
                    st["%s.scale.%s" % (name, bit)] = s.detach().reshape(-1).float().cpu()
                    st["%s.zero_point.%s" % (name, bit)] = m.quantizer.dic_zero_point[bit].detach().reshape(-1).long().cpu()
            if isinstance(m, _SmoothedLinear) and m.channel_scale is not None:
                st[name + ".channel_scale"] = m.channel_scale.detach().float().cpu()
        return st

    def load_quant_state(self, st):
        dev = next(self.parameters()).device
        t = lambda v, dt: torch.as_tensor(v).to(device=dev, dtype=dt)
        for name, m in self.named_modules():
            if isinstance(m, QAct) and (name + ".scale") in st:
                m.quantizer.scale = t(st[name + ".scale"], torch.float32)
                m.quantizer.zero_point = t(st[name + ".zero_point"], torch.int64)
            elif isinstance(m, (QLinear, QConv2d)):
                for bit in ("uint3", "uint4", "int4", "int8"):
                    k = "%s.scale.%s" % (name, bit)
                    if k in st:
                        m.quantizer.dic_scale[bit] = t(st[k], torch.float32)
                        m.quantizer.dic_zero_point[bit] = t(st["%s.zero_point.%s" % (name, bit)], torch.int64)
        for name, m in self.named_modules():   # second pass: the children's quantizers are filled now
            if isinstance(m, _SmoothedLinear) and (name + ".channel_scale") in st:
                cs = t(st[name + ".channel_scale"], torch.float32)
                lin, qa = (m.qkv, m.qact0) if hasattr(m, "qkv") else (m.fc1, m.qact0)
                m.channel_scale = cs
                m.best_scale = [cs, cs]
                m.best_act_scale = [qa.quantizer.scale] * 2
                m.best_act_zp = [qa.quantizer.zero_point] * 2
                m.best_weight_scale = [lin.quantizer.dic_scale] * 2
                m.best_weight_zp = [lin.quantizer.dic_zero_point] * 2
        self._engine = None


class VisionTransformer(nn.Module, QuantModelMixin):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4.0, qkv_bias=True, qk_scale=None, representation_size=None, drop_rate=0.0, attn_drop_rate=0.0,
                 drop_path_rate=0.0, hybrid_backbone=None, norm_layer=None, quant=False, calibrate=False, input_quant=False,
                 cfg=None):
        super().__init__()
        assert hybrid_backbone is None and not representation_size, "not part of the quantized path (unused by every factory)"
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_heads, self.patch_size, self.mlp_ratio = num_heads, patch_size, mlp_ratio
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.cfg = cfg
        self.quant = False
        self.input_quant = input_quant
        if input_quant:
            self.qact_input = _qact(cfg, quant, calibrate)
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim, quant=quant,
                                      calibrate=calibrate, cfg=cfg)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.qact_embed = _qact(cfg, quant, calibrate)
        self.qact_pos = _qact(cfg, quant, calibrate)
        self.qact1 = _qact(cfg, quant, calibrate, ln=True)
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                  attn_drop=attn_drop_rate, drop_path=0.0, norm_layer=norm_layer, quant=quant, calibrate=calibrate, cfg=cfg)
            for _ in range(depth)])
        self.depth = depth
        self.norm = norm_layer(embed_dim)
        self.qact2 = _qact(cfg, quant, calibrate)
        self.pre_logits = nn.Identity()
        self.head = _qlinear(cfg, quant, calibrate, self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        self.act_out = _qact(cfg, quant, calibrate)
        trunc_normal_(self.pos_embed, std=0.02)
        trunc_normal_(self.cls_token, std=0.02)
        self.apply(self._init_weights)
        self._engine = None

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed", "cls_token"}

    def get_classifier(self):
        return self.head

    # ---- module-by-module forward (calibration, FP, eager quantized)
    def forward_features(self, x, FLOPs, global_distance, bit_config, global_plot, hessian_statistic=False):
        B = x.shape[0]
        if self.input_quant:
            x = self.qact_input(x)
        x = self.patch_embed(x, FLOPs, bit_config[0] if bit_config else None)
        x = torch.cat((self.cls_token.expand(B, -1, -1), x), dim=1)
        x = self.qact_embed(x)
        x = x + self.qact_pos(self.pos_embed)
        x = self.pos_drop(self.qact1(x))
        for i, blk in enumerate(self.blocks):
            local = bit_config[i * 4 + 1: i * 4 + 5] if bit_config else None
            last_quantizer = self.qact1.quantizer if i == 0 else self.blocks[i - 1].qact4.quantizer
            x = blk(x, last_quantizer, FLOPs, global_distance, local, False, self.quant, hessian_statistic)
        x = self.norm(x, self.blocks[-1].qact4.quantizer, self.qact2.quantizer)[:, 0]
        return self.pre_logits(self.qact2(x))

    def forward_eager(self, x, bit_config=None, plot=False, hessian_statistic=False):
        FLOPs, global_distance = [], []
        x = self.forward_features(x, FLOPs, global_distance, bit_config, plot, hessian_statistic)
        C = x.shape[1]
        x = self.head(x, global_distance, bit_config[-1] if bit_config else None)
        FLOPs.append(C * x.shape[1])
        return self.act_out(x), FLOPs, global_distance

    def flops_list(self):
        """the MAC counts the reference's forward appends (patch-embed, 4 per block, head)."""
        D, N = self.embed_dim, self.patch_embed.num_patches + 1
        P, g = self.patch_size, self.patch_embed.grid_size
        Hd = int(D * self.mlp_ratio)
        out = [3 * P * P * D * g[0] * g[1]]
        for _ in range(self.depth):
            out += [N * D * 3 * D, N * D * D, N * D * Hd, N * Hd * D]
        return out + [D * self.num_classes]

    def forward(self, x, bit_config=None, plot=False, hessian_statistic=False):
        if not self.quant or hessian_statistic:
            return self.forward_eager(x, bit_config, plot, hessian_statistic)
        from .engine import VitEngine

        if bit_config is None:
            raise ValueError("the quantized forward needs bit_config (1 + 4*depth + 1 entries of 4 or 8), like the reference "
                             "(vit_fquant.py:335: bit_pool.index(None) fails)")
        if self._engine is None:
            self._engine = VitEngine(self)
        if self._engine is not False:
            from .engine import EngineNotApplicable
            try:
                return self._engine(x, bit_config), self.flops_list(), []
            except EngineNotApplicable as e:
                # configurations outside the integer program (Config(ptf=False), Config(lis=False), non-int8 activations): the
                # reference evaluates them through the same call, so does this model - module by module, same kernels
                import warnings
                warnings.warn("p2vit_b200: integer engine not applicable (%s); evaluating module by module (forward_eager)" % e)
                self._engine = False
        return self.forward_eager(x, bit_config, plot, hessian_statistic)



def _vit(embed_dim, depth, num_heads, input_quant, quant, calibrate, cfg, **kwargs):
    return VisionTransformer(patch_size=16, embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(QIntLayerNorm, eps=1e-6), quant=quant, calibrate=calibrate,
                             input_quant=input_quant, cfg=cfg, **kwargs)


def _no_pretrained(pretrained):
    if pretrained:
        raise RuntimeError("pretrained checkpoints need network access; load a local file with p2vit_b200.load_checkpoint(model, path) "
                           "(DeiT / Swin .pth, Flax ViT .npz; key names equal the reference's) or use p2vit_b200.synth for seeded synthetic weights")


def deit_tiny_patch16_224(pretrained=False, quant=False, calibrate=False, cfg=None, **kwargs):
    _no_pretrained(pretrained)
    return _vit(192, 12, 3, True, quant, calibrate, cfg, **kwargs)


def deit_small_patch16_224(pretrained=False, quant=False, calibrate=False, cfg=None, **kwargs):
    _no_pretrained(pretrained)
    return _vit(384, 12, 6, True, quant, calibrate, cfg, **kwargs)


def deit_base_patch16_224(pretrained=False, quant=False, calibrate=False, cfg=None, **kwargs):
    _no_pretrained(pretrained)
    return _vit(768, 12, 12, True, quant, calibrate, cfg, **kwargs)


def vit_base_patch16_224(pretrained=False, quant=False, calibrate=False, cfg=None, **kwargs):
    _no_pretrained(pretrained)
    return _vit(768, 12, 12, True, quant, calibrate, cfg, **kwargs)


def vit_large_patch16_224(pretrained=False, quant=False, calibrate=False, cfg=None, **kwargs):
    _no_pretrained(pretrained)
    return _vit(1024, 24, 16, False, quant, calibrate, cfg, **kwargs)
