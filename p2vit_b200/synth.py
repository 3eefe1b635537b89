"""Seeded synthetic weights and images (there is no network for checkpoints or datasets).

The recipe follows SURVEY.md section 8(d): the reference's `pretrained=False` init
(trunc-normal std .02 Linear weights, LN gamma=1 beta=0; vit_fquant.py:775-782) leaves the
quantized path degenerate (uniform attention, two distinct softmax codes, constant
top-1), so weights are drawn with larger, seeded gains chosen such that attention rows,
LayerNorm shifts, GELU and the classifier arg-max are all input dependent.

Everything is generated with numpy PCG64 streams keyed by (seed, crc32(tensor name)) so a
tensor's values do not depend on creation order, platform, or torch version; the same
dict is loaded into the reference model (golden generation), the CPU oracle and the
B200 model (state-dict key names follow vit_fquant.py:656-770).
"""
import zlib

import numpy as np
import torch

__all__ = ["VIT_CONFIGS", "SWIN_CONFIGS", "synth_vit_state_dict", "synth_swin_state_dict", "synth_images", "mixed_bit_config"]

# name -> dict(embed_dim, depth, num_heads, input_quant)      vit_fquant.py:942-1074
VIT_CONFIGS = {
    "vit_micro": dict(embed_dim=128, depth=2, num_heads=2, input_quant=True),  # test-only size
    "deit_tiny": dict(embed_dim=192, depth=12, num_heads=3, input_quant=True),
    "deit_small": dict(embed_dim=384, depth=12, num_heads=6, input_quant=True),
    "deit_base": dict(embed_dim=768, depth=12, num_heads=12, input_quant=True),
    "vit_base": dict(embed_dim=768, depth=12, num_heads=12, input_quant=True),
    "vit_large": dict(embed_dim=1024, depth=24, num_heads=16, input_quant=False),
}


def _rng(seed, name):
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def _normal(seed, name, shape, std, mean=0.0):
    x = _rng(seed, name).standard_normal(size=shape, dtype=np.float32) * np.float32(std)
    if mean:
        x = x + np.float32(mean)
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))


def synth_vit_state_dict(embed_dim, depth, num_heads=None, seed=0, patch=16, in_chans=3,
                         num_classes=1000, mlp_ratio=4, num_patches=196, **_):
    D = embed_dim
    Hd = int(D * mlp_ratio)
    sd = {}
    n = lambda name, shape, std, mean=0.0: sd.__setitem__(name, _normal(seed, name, shape, std, mean))
    n("cls_token", (1, 1, D), 0.1)
    n("pos_embed", (1, num_patches + 1, D), 0.15)
    n("patch_embed.proj.weight", (D, in_chans, patch, patch), 0.04)
    n("patch_embed.proj.bias", (D,), 0.1)
    for i in range(depth):
        p = "blocks.%d." % i
        n(p + "norm1.weight", (D,), 0.15, 1.0)
        n(p + "norm1.bias", (D,), 0.1)
        n(p + "attn.qkv.weight", (3 * D, D), 0.16)
        n(p + "attn.qkv.bias", (3 * D,), 0.1)
        n(p + "attn.proj.weight", (D, D), 0.04)
        n(p + "attn.proj.bias", (D,), 0.05)
        n(p + "norm2.weight", (D,), 0.15, 1.0)
        n(p + "norm2.bias", (D,), 0.1)
        n(p + "mlp.fc1.weight", (Hd, D), 0.05)
        n(p + "mlp.fc1.bias", (Hd,), 0.1)
        n(p + "mlp.fc2.weight", (D, Hd), 0.03)
        n(p + "mlp.fc2.bias", (D,), 0.05)
    n("norm.weight", (D,), 0.15, 1.0)
    n("norm.bias", (D,), 0.02)
    n("head.weight", (num_classes, D), 0.08)
    sd["head.bias"] = torch.zeros(num_classes, dtype=torch.float32)
    return sd


# name -> dict(embed_dim, depths, num_heads)      swin_quant.py:917-995 (window 7, patch 4, mlp_ratio 4, input_quant=True)
SWIN_CONFIGS = {
    "swin_micro": dict(embed_dim=32, depths=(2, 2), num_heads=(1, 2)),          # test-only size: stages 56x56 and 28x28
    "swin_tiny": dict(embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24)),
    "swin_small": dict(embed_dim=96, depths=(2, 2, 18, 2), num_heads=(3, 6, 12, 24)),
    "swin_base": dict(embed_dim=128, depths=(2, 2, 18, 2), num_heads=(4, 8, 16, 32)),
}


def synth_swin_state_dict(embed_dim, depths, num_heads, seed=0, patch=4, in_chans=3, num_classes=1000, mlp_ratio=4, window=7, **_):
    """parameter names of models/swin_quant.py (SwinTransformer :636-914); `downsample.reduction.bias` is the all-zero bias the
    reference needs to calibrate its bias-free QLinear."""
    sd = {}
    n = lambda name, shape, std, mean=0.0: sd.__setitem__(name, _normal(seed, name, shape, std, mean))
    C0 = embed_dim
    n("patch_embed.proj.weight", (C0, in_chans, patch, patch), 0.15)
    n("patch_embed.proj.bias", (C0,), 0.1)
    n("patch_embed.norm.weight", (C0,), 0.15, 1.0)
    n("patch_embed.norm.bias", (C0,), 0.1)
    for i, (depth, heads) in enumerate(zip(depths, num_heads)):
        C = C0 * 2 ** i
        Hd = int(C * mlp_ratio)
        for j in range(depth):
            p = "layers.%d.blocks.%d." % (i, j)
            n(p + "norm1.weight", (C,), 0.15, 1.0)
            n(p + "norm1.bias", (C,), 0.1)
            n(p + "attn.relative_position_bias_table", ((2 * window - 1) ** 2, heads), 0.5)
            n(p + "attn.qkv.weight", (3 * C, C), 0.9 / C ** 0.5)
            n(p + "attn.qkv.bias", (3 * C,), 0.1)
            n(p + "attn.proj.weight", (C, C), 0.5 / C ** 0.5)
            n(p + "attn.proj.bias", (C,), 0.05)
            n(p + "norm2.weight", (C,), 0.15, 1.0)
            n(p + "norm2.bias", (C,), 0.1)
            n(p + "mlp.fc1.weight", (Hd, C), 0.7 / C ** 0.5)
            n(p + "mlp.fc1.bias", (Hd,), 0.1)
            n(p + "mlp.fc2.weight", (C, Hd), 0.6 / Hd ** 0.5)
            n(p + "mlp.fc2.bias", (C,), 0.05)
        if i < len(depths) - 1:
            p = "layers.%d.downsample." % i
            n(p + "norm.weight", (4 * C,), 0.15, 1.0)
            n(p + "norm.bias", (4 * C,), 0.1)
            n(p + "reduction.weight", (2 * C, 4 * C), 0.7 / (4 * C) ** 0.5)
            sd[p + "reduction.bias"] = torch.zeros(2 * C, dtype=torch.float32)
    Cf = C0 * 2 ** (len(depths) - 1)
    n("norm.weight", (Cf,), 0.15, 1.0)
    n("norm.bias", (Cf,), 0.02)
    n("head.weight", (num_classes, Cf), 0.3)
    sd["head.bias"] = torch.zeros(num_classes, dtype=torch.float32)
    return sd


def synth_images(batch, seed=0, start=0, size=224):
    """`batch` images of the stream `seed`, beginning at image index `start` (so a batch
    sharded over ranks is the same data as the unsharded batch).  Unit-variance Gaussian data like the
    reference's `--mode 1` calibration input (test_quant.py:275-281), but with a 16x16-pixel
    block-constant component so patch tokens differ from each other and top-1 depends on
    the image (pure white noise averages out over the 196 tokens)."""
    out = np.empty((batch, 3, size, size), dtype=np.float32)
    cells = size // 16
    for i in range(batch):
        g = np.random.Generator(np.random.PCG64([seed, 0x696D67, start + i]))
        noise = g.standard_normal(size=(3, size, size), dtype=np.float32)
        coarse = g.standard_normal(size=(3, cells, cells), dtype=np.float32)
        out[i] = np.float32(0.6) * noise + np.float32(0.8) * np.kron(coarse, np.ones((16, 16), np.float32))
    return torch.from_numpy(out)


def mixed_bit_config(flops, depth, seed=0):
    """A 1 + 4*depth + 1 entry {4,8} bit_config drawn by the sampling rule of the reference's search (test_quant.py:323-341):
    first layer 8 bit, the attention pair and the MLP pair of a block share a width, sum(MACs_i * bits_i) <= 1.1 * sum(MACs_i * 4).
    `flops` = the per-layer MAC list the forward returns.  ViT-L (depth 24) gives the 98-entry config of BASELINE config C5."""
    import random
    rnd = random.Random(seed)
    budget = 1.1 * sum(f * 4 for f in flops)
    while True:
        cfg = [8]
        for _ in range(depth):
            a, m = rnd.choice([4, 8]), rnd.choice([4, 8])
            cfg += [a, a, m, m]
        cfg.append(rnd.choice([4, 8]))
        if sum(f * b for f, b in zip(flops, cfg)) <= budget:
            return cfg
