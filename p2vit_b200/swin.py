"""Swin Transformer with the reference's module tree, factories and flag methods (reference: models/swin_quant.py:26-995),
with the four stale call sites of the reference repaired the way its own operators intend (PatchEmbed / Mlp
arguments, bias-free reduction, in_scale_expand slot, driver-compatible return).

    model = swin_tiny_patch4_window7_224(cfg=Config())
    model.model_open_calibrate(); model.model_open_last_calibrate(); model(calib_images)
    model.model_close_calibrate(); model.model_quant()
    logits, FLOPs, global_distance = model(images)        # weights are 8 bit at every Swin call site (SURVEY 8a'')

Before `model_quant()` the forward is the FP / calibration forward, module by module (observers record).  After it the
forward runs the integer engine (p2vit_b200/swin_engine.py): window partition / cyclic shift / window reverse / patch
merging are index remaps fused into the LayerNorm and GEMM kernels' row addressing, attention runs per 7x7 window with
the quantized relative-position bias and the shift mask inside the kernel.
"""
import torch
import torch.nn as nn

from .ptq import QAct, QIntLayerNorm, QIntSoftmax
from .vit import Mlp, PatchEmbed, QuantModelMixin, _qact, _qlinear, trunc_normal_

__all__ = ["SwinTransformer", "swin_tiny_patch4_window7_224", "swin_small_patch4_window7_224", "swin_base_patch4_window7_224"]


def window_partition(x, window_size):
    """[B, H, W, C] -> [B*nW, ws, ws, C]  (swin_quant.py:26-41)"""
    B, H, W, C = x.shape
    x = x.view(B, H // window_size, window_size, W // window_size, window_size, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, window_size, window_size, C)


def window_reverse(windows, window_size, H, W):
    """[B*nW, ws, ws, C] -> [B, H, W, C]  (swin_quant.py:44-59)"""
    B = int(windows.shape[0] / (H * W / window_size / window_size))
    x = windows.view(B, H // window_size, W // window_size, window_size, window_size, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def relative_position_index(ws):
    """[ws*ws, ws*ws] index into the (2ws-1)^2 bias table (swin_quant.py:100-115)"""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def shifted_window_mask(H, W, ws, shift):
    """[nW, ws*ws, ws*ws] of 0 / -100 for SW-MSA (swin_quant.py:365-395)"""
    img = torch.zeros((1, H, W, 1))
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, h, w, :] = cnt
            cnt += 1
    mw = window_partition(img, ws).view(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


class WindowAttention(nn.Module):
    def __init__(self, dim, window_size, num_heads, qkv_bias=True, attn_drop=0.0, proj_drop=0.0, quant=False, calibrate=False, cfg=None):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * window_size[0] - 1) * (2 * window_size[1] - 1), num_heads))
        assert window_size[0] == window_size[1]
        self.register_buffer("relative_position_index", relative_position_index(window_size[0]))
        self.qkv = _qlinear(cfg, quant, calibrate, dim, dim * 3, bias=qkv_bias)
        self.qact1 = _qact(cfg, quant, calibrate)
        self.qact_attn1 = _qact(cfg, quant, calibrate)
        self.qact_table = _qact(cfg, quant, calibrate)
        self.qact2 = _qact(cfg, quant, calibrate)
        self.attn_drop = nn.Dropout(attn_drop)
        self.log_int_softmax = QIntSoftmax(log_i_softmax=cfg.INT_SOFTMAX, quant=quant, calibrate=calibrate, bit_type=cfg.BIT_TYPE_S,
                                           calibration_mode=cfg.CALIBRATION_MODE_S, observer_str=cfg.OBSERVER_S, quantizer_str=cfg.QUANTIZER_S)
        self.qact3 = _qact(cfg, quant, calibrate)
        self.qact4 = _qact(cfg, quant, calibrate)
        self.proj = _qlinear(cfg, quant, calibrate, dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        trunc_normal_(self.relative_position_bias_table, std=0.02)

    def relative_position_bias(self, table):
        n = self.window_size[0] * self.window_size[1]
        return table[self.relative_position_index.view(-1)].view(n, n, -1).permute(2, 0, 1).contiguous()

    def forward(self, x, mask=None):
        B_, N, C = x.shape
        x = self.qact1(self.qkv(x))
        qkv = x.reshape(B_, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = self.qact_attn1((q * self.scale) @ k.transpose(-2, -1))
        attn = self.qact2(attn + self.relative_position_bias(self.qact_table(self.relative_position_bias_table)).unsqueeze(0))
        if mask is not None:
            nW = mask.shape[0]
            attn = (attn.view(B_ // nW, nW, self.num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, self.num_heads, N, N)
        attn = self.attn_drop(self.log_int_softmax(attn, self.qact2.quantizer.scale))
        x = self.qact3((attn @ v).transpose(1, 2).reshape(B_, N, C))
        return self.proj_drop(self.qact4(self.proj(x)))


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4.0, qkv_bias=True, drop=0.0,
                 attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, quant=False, calibrate=False, cfg=None):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, input_resolution, num_heads
        self.window_size, self.shift_size, self.mlp_ratio = window_size, shift_size, mlp_ratio
        if min(self.input_resolution) <= self.window_size:   # swin_quant.py:300-303
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size
        self.norm1 = norm_layer(dim)
        self.qact1 = _qact(cfg, quant, calibrate)
        self.attn = WindowAttention(dim, (self.window_size, self.window_size), num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop,
                                    proj_drop=drop, quant=quant, calibrate=calibrate, cfg=cfg)
        self.drop_path = nn.Identity()   # inference only
        self.qact2 = _qact(cfg, quant, calibrate, ln=True)
        self.norm2 = norm_layer(dim)
        self.qact3 = _qact(cfg, quant, calibrate)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop, quant=quant, calibrate=calibrate,
                       cfg=cfg)
        self.qact4 = _qact(cfg, quant, calibrate, ln=True)
        mask = shifted_window_mask(*self.input_resolution, self.window_size, self.shift_size) if self.shift_size > 0 else None
        self.register_buffer("attn_mask", mask)

    def forward(self, x, last_quantizer=None):
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        shortcut = x
        x = self.qact1(self.norm1(x, last_quantizer, self.qact1.quantizer)).view(B, H, W, C)
        if self.shift_size > 0:
            x = torch.roll(x, shifts=(-self.shift_size, -self.shift_size), dims=(1, 2))
        xw = window_partition(x, self.window_size).view(-1, self.window_size * self.window_size, C)
        aw = self.attn(xw, mask=self.attn_mask).view(-1, self.window_size, self.window_size, C)
        x = window_reverse(aw, self.window_size, H, W)
        if self.shift_size > 0:
            x = torch.roll(x, shifts=(self.shift_size, self.shift_size), dims=(1, 2))
        x = self.qact2(shortcut + self.drop_path(x.view(B, H * W, C)))
        h = self.qact3(self.norm2(x, self.qact2.quantizer, self.qact3.quantizer))
        return self.qact4(x + self.drop_path(self.mlp(h, [], [], [8, 8])))   # Mlp's W8 pair, SURVEY Q4a


class PatchMerging(nn.Module):
    def __init__(self, input_resolution, dim, norm_layer=nn.LayerNorm, quant=False, calibrate=False, cfg=None):
        super().__init__()
        self.input_resolution, self.dim = input_resolution, dim
        self.norm = norm_layer(4 * dim)
        self.qact1 = _qact(cfg, quant, calibrate)
        self.reduction = _qlinear(cfg, quant, calibrate, 4 * dim, 2 * dim, bias=False)
        self.qact2 = _qact(cfg, quant, calibrate, ln=True)

    def forward(self, x, last_quantizer=None):
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W and H % 2 == 0 and W % 2 == 0
        x = x.view(B, H, W, C)
        x = torch.cat([x[:, 0::2, 0::2, :], x[:, 1::2, 0::2, :], x[:, 0::2, 1::2, :], x[:, 1::2, 1::2, :]], -1).view(B, -1, 4 * C)
        x = self.qact1(self.norm(x, last_quantizer, self.qact1.quantizer, None, 4))   # in_scale_expand=4, SURVEY Q4c
        return self.qact2(self.reduction(x))


class BasicLayer(nn.Module):
    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4.0, qkv_bias=True, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False, quant=False, calibrate=False, cfg=None):
        super().__init__()
        self.dim, self.input_resolution, self.depth = dim, input_resolution, depth
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim=dim, input_resolution=input_resolution, num_heads=num_heads, window_size=window_size,
                                 shift_size=0 if (i % 2 == 0) else window_size // 2, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop,
                                 attn_drop=attn_drop, norm_layer=norm_layer, quant=quant, calibrate=calibrate, cfg=cfg)
            for i in range(depth)])
        self.downsample = downsample(input_resolution, dim=dim, norm_layer=norm_layer, quant=quant, calibrate=calibrate, cfg=cfg) \
            if downsample is not None else None

    def forward(self, x, last_quantizer=None):
        for i, blk in enumerate(self.blocks):
            x = blk(x, last_quantizer if i == 0 else self.blocks[i - 1].qact4.quantizer)
        if self.downsample is not None:
            x = self.downsample(x, self.blocks[-1].qact4.quantizer)
        return x


class SwinTransformer(nn.Module, QuantModelMixin):
    def __init__(self, img_size=224, patch_size=4, in_chans=3, num_classes=1000, embed_dim=96, depths=(2, 2, 6, 2),
                 num_heads=(3, 6, 12, 24), window_size=7, mlp_ratio=4.0, qkv_bias=True, drop_rate=0.0, attn_drop_rate=0.0,
                 drop_path_rate=0.1, norm_layer=nn.LayerNorm, ape=False, patch_norm=True, use_checkpoint=False, quant=False,
                 calibrate=False, input_quant=False, cfg=None, **kwargs):
        super().__init__()
        assert not ape, "absolute position embedding is not used by any factory of the reference"
        self.num_classes, self.num_layers, self.embed_dim = num_classes, len(depths), embed_dim
        self.depths, self.heads, self.window_size, self.patch_size = tuple(depths), tuple(num_heads), window_size, patch_size
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.mlp_ratio, self.input_quant, self.cfg = mlp_ratio, input_quant, cfg
        self.quant = False
        if input_quant:
            self.qact_input = _qact(cfg, quant, calibrate)
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      norm_layer=norm_layer if patch_norm else None, quant=quant, calibrate=calibrate, cfg=cfg)
        self.patch_grid = self.patch_embed.grid_size
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.layers = nn.Sequential(*[
            BasicLayer(dim=int(embed_dim * 2 ** i), input_resolution=(self.patch_grid[0] // 2 ** i, self.patch_grid[1] // 2 ** i),
                       depth=depths[i], num_heads=num_heads[i], window_size=window_size, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                       drop=drop_rate, attn_drop=attn_drop_rate, norm_layer=norm_layer,
                       downsample=PatchMerging if i < self.num_layers - 1 else None, quant=quant, calibrate=calibrate, cfg=cfg)
            for i in range(self.num_layers)])
        self.norm = norm_layer(self.num_features)
        self.qact2 = _qact(cfg, quant, calibrate)
        self.avgpool = nn.AdaptiveAvgPool1d(1)
        self.qact3 = _qact(cfg, quant, calibrate)
        self.head = _qlinear(cfg, quant, calibrate, self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        self.act_out = _qact(cfg, quant, calibrate)
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
        return {"absolute_pos_embed"}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        return {"relative_position_bias_table"}

    def get_classifier(self):
        return self.head

    def forward_features(self, x):
        if self.input_quant:
            x = self.qact_input(x)
        x = self.pos_drop(self.patch_embed(x, [], 8))
        for i, layer in enumerate(self.layers):
            x = layer(x, self.patch_embed.qact.quantizer if i == 0 else self.layers[i - 1].downsample.qact2.quantizer)
        x = self.qact2(self.norm(x, self.layers[-1].blocks[-1].qact4.quantizer, self.qact2.quantizer))
        x = self.qact3(self.avgpool(x.transpose(1, 2)))
        return torch.flatten(x, 1)

    def forward(self, x, bit_config=None, plot=False, hessian_statistic=False):
        """returns (logits, FLOPs, global_distance) like the ViT models so the driver's `validate` (test_quant.py:492) can unpack it
        (the reference's Swin returns the bare tensor and breaks there); bit_config is ignored: every Swin call site
        of the reference runs 8-bit weights."""
        if self.quant:
            from .swin_engine import SwinEngine

            if self._engine is None:
                self._engine = SwinEngine(self)
            return self._engine(x), [], []
        return self.act_out(self.head(self.forward_features(x))), [], []


def _swin(embed_dim, depths, num_heads, quant, calibrate, cfg, **kwargs):
    return SwinTransformer(patch_size=4, window_size=7, embed_dim=embed_dim, depths=depths, num_heads=num_heads, norm_layer=QIntLayerNorm,
                           quant=quant, calibrate=calibrate, input_quant=True, cfg=cfg, **kwargs)


def _no_pretrained(pretrained):
    if pretrained:
        raise RuntimeError("pretrained checkpoints need network access; load a state dict with model.load_state_dict() "
                           "(key names equal the reference's) or use p2vit_b200.synth for seeded synthetic weights")


def swin_tiny_patch4_window7_224(pretrained=False, quant=False, calibrate=False, cfg=None, **kwargs):
    _no_pretrained(pretrained)
    return _swin(96, (2, 2, 6, 2), (3, 6, 12, 24), quant, calibrate, cfg, **kwargs)


def swin_small_patch4_window7_224(pretrained=False, quant=False, calibrate=False, cfg=None, **kwargs):
    _no_pretrained(pretrained)
    return _swin(96, (2, 2, 18, 2), (3, 6, 12, 24), quant, calibrate, cfg, **kwargs)


def swin_base_patch4_window7_224(pretrained=False, quant=False, calibrate=False, cfg=None, **kwargs):
    _no_pretrained(pretrained)
    return _swin(128, (2, 2, 18, 2), (4, 8, 16, 32), quant, calibrate, cfg, **kwargs)
