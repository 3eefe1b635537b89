"""Per-layer time of the attention kernel on the model's own qkv codes (golden-calibrated DeiT-S / ViT-B, synthetic images), both
probability paths, in a CUDA graph of 5 launches each.  usage: python tools/att_layers.py [deit_small] [minmax] [B]"""
import sys
import os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p2vit_b200 import Config, build_model, ops, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "deit_small"
method = sys.argv[2] if len(sys.argv) > 2 else "minmax"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "%s_%s.npz" % (name, method)))
m = build_model(name, Config(True, True, method), seed=int(g["meta.seed"]), device="cuda")
m.load_quant_state({k[6:]: g[k] for k in g.files if k.startswith("state/")})
m.model_quant()
x = synth.synth_images(B, seed=1).cuda()
bits = [8] * (4 * m.depth + 2)
m(x[:1].contiguous(), bits)          # creates the engine
eng = m._engine
prog = eng._program(tuple(bits), B)
prog["ws"]["img"].copy_(x)
steps = dict(prog["steps"])
for nm, fn in prog["steps"]:
    fn()
    if nm.endswith("attn.qact2"):
        res = []
        for mode in (0, 1):
            fn.args.prob_mode = mode
            gr = ops.capture_graph(lambda fn=fn: [fn() for _ in range(5)])
            gr.replay()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            gr.replay()
            b.record()
            b.synchronize()
            res.append(a.elapsed_time(b) * 1e3 / 5)
        fn.args.prob_mode = 0
        qkv = prog["ws"]["qkv"]
        print("%-26s mode 0 %6.1f us   mode 1 %6.1f us   |q| mean %.1f  |k| mean %.1f" % (
            nm, res[0], res[1], float(qkv.view(B, -1, 3, qkv.shape[-1] // 3)[:, :, 0].float().abs().mean()),
            float(qkv.view(B, -1, 3, qkv.shape[-1] // 3)[:, :, 1].float().abs().mean())))
