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

__all__ = ["VIT_CONFIGS", "synth_vit_state_dict", "synth_images"]

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
