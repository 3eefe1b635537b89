"""TEST INFRASTRUCTURE ONLY - generates tests/golden/*.npz by running the UNMODIFIED
reference (/root/reference via oracle/ref_loader.py) on seeded synthetic weights/images
(p2vit_b200/synth.py).  Runs only in the build container (the reference is not on the
GPU box); the produced fixtures are committed.

    python oracle/gen_golden.py vit_micro  --calib 4 --eval 2 --taps
    python oracle/gen_golden.py deit_tiny  --calib 8 --eval 8
    python oracle/gen_golden.py deit_small --calib 32 --eval 4 --method minmax

File content (all numpy arrays):
  meta.*                 model name, seeds, batch sizes, observer method, torch version
  state/<name>           calibrated quantizer state, names = reference module paths
                         (<qact>.scale/.zero_point, <qlinear>.scale.<bit>, <attn|mlp>.channel_scale)
  logits8 / logits4      act_out output for bit_config=[8]*n and [4]*n   (fp32 [B,1000])
  gd                     global_distance [n_linear,4] from the calibration forward
  tap8/<module>          (only --taps) output of every QAct / QIntLayerNorm / QIntSoftmax
                         module in the W8A8 forward, stored as integer codes
"""
import argparse
import os
import sys
import time
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference  # noqa: E402


def _load_synth():
    import importlib.util

    spec = importlib.util.spec_from_file_location("_p2v_synth", os.path.join(ROOT, "p2vit_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def build_reference_vit(name, method="minmax", seed=0):
    synth = _load_synth()
    ref, RefConfig = load_reference()
    from models.vit_fquant import VisionTransformer

    c = synth.VIT_CONFIGS[name]
    cfg = RefConfig(True, True, method)
    model = VisionTransformer(patch_size=16, embed_dim=c["embed_dim"], depth=c["depth"], num_heads=c["num_heads"],
                              mlp_ratio=4, qkv_bias=True, norm_layer=partial(ref.QIntLayerNorm, eps=1e-6),
                              input_quant=c["input_quant"], cfg=cfg)
    res = model.load_state_dict(synth.synth_vit_state_dict(**c, seed=seed), strict=False)
    assert not res.missing_keys and not res.unexpected_keys, res
    return model.eval(), c, synth, ref


def reference_calibrate(model, images):
    """test_quant.py:275-312 (mode 1)."""
    model.model_open_calibrate()
    with torch.no_grad():
        model.model_open_last_calibrate()
        _, _, gd = model(images)
    model.model_close_calibrate()
    model.model_quant()
    return np.array([[float(d) for d in row] for row in gd], dtype=np.float32)


def extract_state(model, ref):
    st = {}
    for name, m in model.named_modules():
        if isinstance(m, ref.QAct):
            st[name + ".scale"] = m.quantizer.scale.detach().reshape(-1).float().numpy()
            st[name + ".zero_point"] = m.quantizer.zero_point.detach().reshape(-1).long().numpy()
        elif isinstance(m, (ref.QLinear, ref.QConv2d)):
            for bit, s in m.quantizer.dic_scale.items():
                st["%s.scale.%s" % (name, bit)] = s.detach().reshape(-1).float().numpy()
                st["%s.zero_point.%s" % (name, bit)] = m.quantizer.dic_zero_point[bit].detach().reshape(-1).long().numpy()
        if getattr(m, "channel_scale", None) is not None and hasattr(m, "best_scale"):
            assert all(torch.equal(b, m.channel_scale) for b in m.best_scale)
            st[name + ".channel_scale"] = m.channel_scale.detach().float().numpy()
    return st


def encode_tap(v):
    """dequantized fp32 tensor on a (per-channel) grid -> smallest exact integer carrier."""
    v = v.detach().float().numpy()
    return v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model")
    ap.add_argument("--calib", type=int, default=4)
    ap.add_argument("--eval", type=int, default=2)
    ap.add_argument("--method", default="minmax")
    ap.add_argument("--taps", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=None)
    ap.add_argument("--mixed", action="store_true", help="also store logitsmix for the seeded {4,8} draw of synth.mixed_bit_config "
                                                         "(test_quant.py:323-341 sampling rule) and the draw itself as bits_mixed")
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    if args.threads:
        torch.set_num_threads(args.threads)

    torch.manual_seed(0)
    model, c, synth, ref = build_reference_vit(args.model, args.method, args.seed)
    t0 = time.time()
    gd = reference_calibrate(model, synth.synth_images(args.calib, seed=args.seed))
    t_cal = time.time() - t0
    out = {"gd": gd}
    for k, v in extract_state(model, ref).items():
        out["state/" + k] = v
    nb = 4 * c["depth"] + 2
    x = synth.synth_images(args.eval, seed=args.seed + 1)
    taps = {}
    hooks = []
    if args.taps:
        for name, m in model.named_modules():
            if isinstance(m, (ref.QAct, ref.QIntLayerNorm, ref.QIntSoftmax)):
                hooks.append(m.register_forward_hook(lambda mod, i, o, name=name: taps.__setitem__(name, o.detach().clone())))
    t0 = time.time()
    with torch.no_grad():
        logits8, flops, _ = model(x, [8] * nb, False)
    t_fwd = time.time() - t0
    for h in hooks:
        h.remove()
    for k, v in taps.items():
        out["tap8/" + k] = v.numpy().astype(np.float32)
    with torch.no_grad():
        logits4 = model(x, [4] * nb, False)[0]
    if args.mixed:
        mixed = synth.mixed_bit_config([int(f) for f in flops], c["depth"])
        with torch.no_grad():
            out["logitsmix"] = model(x, mixed, False)[0].numpy()
        out["bits_mixed"] = np.array(mixed, dtype=np.int64)
    out["logits8"] = logits8.numpy()
    out["logits4"] = logits4.numpy()
    out["meta.model"] = np.array(args.model)
    out["meta.method"] = np.array(args.method)
    out["meta.seed"] = np.array(args.seed)
    out["meta.calib"] = np.array(args.calib)
    out["meta.eval"] = np.array(args.eval)
    out["meta.torch"] = np.array(torch.__version__)
    out["meta.threads"] = np.array(torch.get_num_threads())   # the reference's fp32 reductions depend on it (tests pin it)
    out["meta.ref_calib_seconds"] = np.array(t_cal, dtype=np.float32)
    out["meta.ref_fwd_seconds"] = np.array(t_fwd, dtype=np.float32)
    path = args.out or os.path.join(ROOT, "tests", "golden", "%s_%s.npz" % (args.model, args.method))
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "calib %.1fs fwd %.2fs" % (t_cal, t_fwd),
          "top1", logits8.argmax(1).tolist())


if __name__ == "__main__":
    main()
