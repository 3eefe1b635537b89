"""Checkpoint formats of the reference's model zoo -> this package's state-dict keys.

* DeiT / Swin `.pth`: `{"model": state_dict}` (or a bare state dict) with the reference's key names
  (vit_fquant.py:916-925, swin_quant.py:945-990) - `load_checkpoint` unwraps it and calls `load_state_dict(strict=False)`.
* ViT `.npz` from the Flax implementation (reference: models/utils.py:12-205 `load_weights_from_npz`):
  `flax_to_state_dict` is a table of (flax name -> key, layout rule) instead of the reference's imperative copy list, so
  the same table also runs backwards (`state_dict_to_flax`, used by the tests and to export).
  Layout rules: HWIO conv kernels -> OIHW, [in, out] dense kernels -> [out, in], the three [D, heads, hd] attention
  projections concatenated into qkv [3D, D], the [heads, hd, D] output projection flattened to [D, D]
  (models/utils.py:16-26,170-191).  Position embeddings of another grid are resized bicubically, class token kept
  (models/utils.py:77-100,147-154).  Hybrid (ResNet-stem) checkpoints are not on this path and raise.

No network: `url` arguments of the reference become local paths.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

__all__ = ["flax_to_state_dict", "state_dict_to_flax", "load_weights_from_npz", "load_checkpoint", "resize_pos_embed"]

# layout rules ------------------------------------------------------------------------------------------------------
_ID, _T2, _CONV = "id", "t2", "conv"      # as stored | [in, out] -> [out, in] | HWIO -> OIHW


def _to_torch(a, rule):
    a = np.asarray(a)
    if rule == _T2:
        a = a.transpose(1, 0)
    elif rule == _CONV:
        a = a.transpose(3, 2, 0, 1)
    return torch.from_numpy(np.ascontiguousarray(a))


def _to_flax(t, rule):
    a = t.detach().cpu().numpy()
    if rule == _T2:
        a = a.transpose(1, 0)
    elif rule == _CONV:
        a = a.transpose(2, 3, 1, 0)
    return np.ascontiguousarray(a)


_TOP = [  # (flax name, state-dict key, rule)
    ("embedding/kernel", "patch_embed.proj.weight", _CONV),
    ("embedding/bias", "patch_embed.proj.bias", _ID),
    ("cls", "cls_token", _ID),
    ("Transformer/posembed_input/pos_embedding", "pos_embed", _ID),
    ("Transformer/encoder_norm/scale", "norm.weight", _ID),
    ("Transformer/encoder_norm/bias", "norm.bias", _ID),
]
_HEAD = [("head/kernel", "head.weight", _T2), ("head/bias", "head.bias", _ID)]
_BLOCK = [  # relative to Transformer/encoderblock_{i}/ and blocks.{i}.
    ("LayerNorm_0/scale", "norm1.weight", _ID), ("LayerNorm_0/bias", "norm1.bias", _ID),
    ("LayerNorm_2/scale", "norm2.weight", _ID), ("LayerNorm_2/bias", "norm2.bias", _ID),
    ("MlpBlock_3/Dense_0/kernel", "mlp.fc1.weight", _T2), ("MlpBlock_3/Dense_0/bias", "mlp.fc1.bias", _ID),
    ("MlpBlock_3/Dense_1/kernel", "mlp.fc2.weight", _T2), ("MlpBlock_3/Dense_1/bias", "mlp.fc2.bias", _ID),
    ("MultiHeadDotProductAttention_1/out/bias", "attn.proj.bias", _ID),
]
_MHA = "MultiHeadDotProductAttention_1/"
_QKV = ("query", "key", "value")


def _depth(w, prefix):
    n = 0
    while f"{prefix}Transformer/encoderblock_{n}/LayerNorm_0/scale" in w:
        n += 1
    return n


def resize_pos_embed(posemb, ntok_new, num_tokens=1, gs_new=()):
    """[1, T + g*g, D] -> [1, T + gs_new[0]*gs_new[1], D]; bicubic on the grid part (models/utils.py:77-100)."""
    tok, grid = posemb[:, :num_tokens], posemb[0, num_tokens:]
    gs_old = int(math.sqrt(grid.shape[0]))
    if not len(gs_new):
        gs_new = [int(math.sqrt(ntok_new - num_tokens))] * 2
    grid = grid.reshape(1, gs_old, gs_old, -1).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=tuple(gs_new), mode="bicubic", align_corners=False)
    grid = grid.permute(0, 2, 3, 1).reshape(1, gs_new[0] * gs_new[1], -1)
    return torch.cat([tok, grid], dim=1)


def flax_to_state_dict(w, prefix=""):
    """Mapping of Flax parameter names -> arrays (an `np.load` handle or a dict) to a state dict with this package's keys."""
    if not prefix and "opt/target/embedding/kernel" in w:
        prefix = "opt/target/"                                                     # models/utils.py:106-107
    if f"{prefix}conv_root/kernel" in w:
        raise NotImplementedError("hybrid (ResNet stem) ViT checkpoints are outside the quantized path")
    sd = {}
    for name, key, rule in _TOP + [h for h in _HEAD if f"{prefix}{h[0]}" in w]:
        sd[key] = _to_torch(w[f"{prefix}{name}"], rule)
    for i in range(_depth(w, prefix)):
        bp, kp = f"{prefix}Transformer/encoderblock_{i}/", f"blocks.{i}."
        for name, key, rule in _BLOCK:
            sd[kp + key] = _to_torch(w[bp + name], rule)
        # [D, heads, hd] x 3 -> [3D, D];  [heads, hd] x 3 -> [3D]                    models/utils.py:170-187
        sd[kp + "attn.qkv.weight"] = torch.cat(
            [_to_torch(w[f"{bp}{_MHA}{n}/kernel"], _ID).flatten(1).T for n in _QKV]).contiguous()
        sd[kp + "attn.qkv.bias"] = torch.cat([_to_torch(w[f"{bp}{_MHA}{n}/bias"], _ID).reshape(-1) for n in _QKV])
        # [heads, hd, D] -> [D, heads*hd]                                           models/utils.py:188
        o = np.asarray(w[f"{bp}{_MHA}out/kernel"])
        sd[kp + "attn.proj.weight"] = torch.from_numpy(np.ascontiguousarray(o.transpose(2, 0, 1))).flatten(1)
    return sd


def state_dict_to_flax(sd, num_heads, prefix=""):
    """Inverse of `flax_to_state_dict` (float parameters only; quantizer buffers have no Flax counterpart)."""
    w = {}
    for name, key, rule in _TOP + [h for h in _HEAD if h[1] in sd]:
        w[prefix + name] = _to_flax(sd[key], rule)
    i = 0
    while f"blocks.{i}.norm1.weight" in sd:
        bp, kp = f"{prefix}Transformer/encoderblock_{i}/", f"blocks.{i}."
        for name, key, rule in _BLOCK:
            w[bp + name] = _to_flax(sd[kp + key], rule)
        qkv_w, qkv_b = sd[kp + "attn.qkv.weight"], sd[kp + "attn.qkv.bias"]
        D = qkv_w.shape[1]
        hd = D // num_heads
        for j, n in enumerate(_QKV):
            w[f"{bp}{_MHA}{n}/kernel"] = _to_flax(qkv_w[j * D:(j + 1) * D].T.reshape(D, num_heads, hd), _ID)
            w[f"{bp}{_MHA}{n}/bias"] = _to_flax(qkv_b[j * D:(j + 1) * D].reshape(num_heads, hd), _ID)
        w[f"{bp}{_MHA}out/kernel"] = _to_flax(sd[kp + "attn.proj.weight"].reshape(D, num_heads, hd).permute(1, 2, 0), _ID)
        i += 1
    return w


@torch.no_grad()
def load_weights_from_npz(model, path, prefix=""):
    """Reference signature minus the download arguments (models/utils.py:12): copy a Flax `.npz` into `model` in place.

    The head is copied only when its width matches (`models/utils.py:157-162`: a 21k-class checkpoint keeps the model's
    own head); the position embedding is resized to the model's grid when the shapes differ.
    """
    with np.load(path) as w:
        sd = flax_to_state_dict(w, prefix)
    own = model.state_dict()
    if "head.bias" in sd and ("head.bias" not in own or own["head.bias"].shape != sd["head.bias"].shape):
        del sd["head.weight"], sd["head.bias"]
    if sd["pos_embed"].shape != own["pos_embed"].shape:
        gs = getattr(model.patch_embed, "grid_size", ())
        sd["pos_embed"] = resize_pos_embed(sd["pos_embed"].float(), own["pos_embed"].shape[1],
                                           getattr(model, "num_tokens", 1), tuple(gs))
    n_blocks = sum(1 for k in own if k.endswith(".norm1.weight"))
    n_ckpt = sum(1 for k in sd if k.endswith(".norm1.weight"))
    if n_blocks != n_ckpt:
        raise ValueError(f"checkpoint has {n_ckpt} encoder blocks, the model {n_blocks}")
    for k, v in sd.items():
        if own[k].shape != v.shape:
            raise ValueError(f"{k}: checkpoint {tuple(v.shape)} vs model {tuple(own[k].shape)}")
        own[k].copy_(v)
    return sorted(sd)


def load_checkpoint(model, path, strict=False):
    """`.npz` -> `load_weights_from_npz`; anything else -> torch.load, unwrap `{"model": ...}`, load_state_dict
    (vit_fquant.py:916-925: `checkpoint["model"]`, `strict=False` because the Q-modules add observer buffers)."""
    if str(path).endswith(".npz"):
        return load_weights_from_npz(model, path)
    ckpt = torch.load(path, map_location="cpu", weights_only=True)
    sd = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt and isinstance(ckpt["model"], dict) else ckpt
    return model.load_state_dict(sd, strict=strict)
