"""Checkpoint formats: Flax `.npz` <-> state dict, `.pth` wrapper, position-embedding resize.
The cross-check against the reference's own loader runs only where /root/reference exists (the build container)."""
import os
import sys

import numpy as np
import pytest
import torch

from p2vit_b200 import Config, checkpoint, synth
from p2vit_b200.runner import build_model

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402


def _micro():
    return build_model("vit_micro", Config(), seed=3, device="cpu")


def _float_keys(model):
    return [k for k, _ in model.named_parameters()]


def test_flax_round_trip_is_exact(tmp_path):
    src = _micro()
    heads = synth.VIT_CONFIGS["vit_micro"]["num_heads"]
    w = checkpoint.state_dict_to_flax(src.state_dict(), heads)
    D = src.pos_embed.shape[-1]
    assert w["Transformer/encoderblock_0/MultiHeadDotProductAttention_1/query/kernel"].shape == (D, heads, D // heads)
    assert w["Transformer/encoderblock_0/MultiHeadDotProductAttention_1/out/kernel"].shape == (heads, D // heads, D)
    assert w["embedding/kernel"].shape == (16, 16, 3, D)
    path = tmp_path / "micro.npz"
    np.savez(path, **w)
    dst = build_model("vit_micro", Config(), seed=4, device="cpu")
    loaded = checkpoint.load_checkpoint(dst, str(path))
    assert set(loaded) == set(_float_keys(src))
    a, b = src.state_dict(), dst.state_dict()
    for k in _float_keys(src):
        assert torch.equal(a[k], b[k]), k


def test_prefix_head_mismatch_and_pos_embed_resize(tmp_path):
    src = _micro()
    heads = synth.VIT_CONFIGS["vit_micro"]["num_heads"]
    w = checkpoint.state_dict_to_flax(src.state_dict(), heads, prefix="opt/target/")
    D = src.pos_embed.shape[-1]
    # a 21k-style head and a 7x7 grid: the head is skipped, the grid resized, the class token kept
    w["opt/target/head/kernel"] = np.zeros((D, 21843), np.float32)
    w["opt/target/head/bias"] = np.zeros((21843,), np.float32)
    small = torch.randn(1, 1 + 49, D)
    w["opt/target/Transformer/posembed_input/pos_embedding"] = small.numpy()
    path = tmp_path / "prefixed.npz"
    np.savez(path, **w)
    dst = build_model("vit_micro", Config(), seed=5, device="cpu")
    head_before = dst.head.weight.clone()
    loaded = checkpoint.load_weights_from_npz(dst, str(path))
    assert "head.weight" not in loaded and torch.equal(dst.head.weight, head_before)
    assert dst.pos_embed.shape == src.pos_embed.shape
    assert torch.equal(dst.pos_embed[:, :1], small[:, :1])
    want = checkpoint.resize_pos_embed(small, dst.pos_embed.shape[1], 1, (14, 14))
    assert torch.equal(dst.pos_embed, want)
    assert torch.equal(dst.blocks[1].attn.qkv.weight, src.blocks[1].attn.qkv.weight)


def test_pth_wrapper_and_errors(tmp_path):
    src = _micro()
    path = tmp_path / "deit_style.pth"
    torch.save({"model": {k: v for k, v in src.state_dict().items() if k in set(_float_keys(src))}}, path)
    dst = build_model("vit_micro", Config(), seed=6, device="cpu")
    res = checkpoint.load_checkpoint(dst, str(path))
    assert not res.unexpected_keys
    assert torch.equal(dst.blocks[0].mlp.fc2.weight, src.blocks[0].mlp.fc2.weight)
    w = checkpoint.state_dict_to_flax(src.state_dict(), synth.VIT_CONFIGS["vit_micro"]["num_heads"])
    w["conv_root/kernel"] = np.zeros((7, 7, 3, 64), np.float32)
    with pytest.raises(NotImplementedError):
        checkpoint.flax_to_state_dict(w)
    del w["conv_root/kernel"]
    extra = {k.replace("encoderblock_0", "encoderblock_%d" % len(src.blocks)): v for k, v in w.items() if "encoderblock_0/" in k}
    np.savez(tmp_path / "deeper.npz", **w, **extra)
    with pytest.raises(ValueError):
        checkpoint.load_weights_from_npz(dst, str(tmp_path / "deeper.npz"))


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree only exists in the build container")
def test_same_parameters_as_the_reference_loader(monkeypatch):
    """models/utils.py:12-205 on the same arrays (its checkpoint path is hard-coded, so np.load is pointed at ours)."""
    ref_models, RefConfig = ref_loader.load_reference()
    import models.utils as ref_utils
    from models.vit_fquant import VisionTransformer as RefViT
    from functools import partial

    c = synth.VIT_CONFIGS["vit_micro"]
    src = _micro()
    w = checkpoint.state_dict_to_flax(src.state_dict(), c["num_heads"])
    rng = np.random.default_rng(0)
    w = {k: rng.standard_normal(v.shape).astype(np.float32) for k, v in w.items()}       # arbitrary, asymmetric values
    w["Transformer/posembed_input/pos_embedding"] = rng.standard_normal((1, 50, c["embed_dim"])).astype(np.float32)
    cfg = RefConfig(False, False, "minmax")
    ref = RefViT(patch_size=16, embed_dim=c["embed_dim"], depth=c["depth"], num_heads=c["num_heads"], mlp_ratio=4, qkv_bias=True,
                 norm_layer=partial(ref_models.QIntLayerNorm, eps=1e-6), input_quant=True, cfg=cfg)
    monkeypatch.setattr(ref_utils.np, "load", lambda *_a, **_k: w)
    ref_utils.load_weights_from_npz(ref, "unused")
    monkeypatch.undo()
    ours = build_model("vit_micro", Config(), seed=7, device="cpu")
    own = ours.state_dict()
    sd = checkpoint.flax_to_state_dict(w)
    sd["pos_embed"] = checkpoint.resize_pos_embed(sd["pos_embed"], own["pos_embed"].shape[1], 1, (14, 14))
    ours.load_state_dict(sd, strict=False)
    rsd = ref.state_dict()
    for k in _float_keys(ours):
        assert torch.equal(rsd[k], ours.state_dict()[k]), k
