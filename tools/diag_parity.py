"""Per-op parity diagnosis on the GPU box: integer engine vs the CPU oracle (fp32-sum and exact-sum variants)
for a model with a golden calibrated state.  Prints, per engine step, the number of differing codes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle.port import VitOracle
from p2vit_b200 import Config, build_model, synth
from p2vit_b200.engine import VitEngine

name = sys.argv[1] if len(sys.argv) > 1 else "deit_tiny"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
wb = int(sys.argv[3]) if len(sys.argv) > 3 else 8
g = np.load("tests/golden/%s_minmax.npz" % name)
st = {k[6:]: g[k] for k in g.files if k.startswith("state/")}
c = synth.VIT_CONFIGS[name]
sd = synth.synth_vit_state_dict(**c, seed=0)
x = synth.synth_images(B, seed=1)
bits = [wb] * (4 * c["depth"] + 2)
res = {}
for exact in (False, True):
    o = VitOracle(sd, **c, exact_sums=exact); o.load_state(st)
    taps = {}
    t = time.time(); res[exact] = (o.forward_quant(x, bits, taps), taps); print("oracle exact=%s %.1fs" % (exact, time.time() - t))
print("oracle fp32-sum logits == golden:", np.array_equal(res[False][0].numpy(), g["logits%d" % wb][:B]))
m = build_model(name, Config(), seed=0, device="cuda"); m.load_quant_state(st); m.model_quant()
eng = VitEngine(m, use_graph=False)
et = {}
logits = eng(x.cuda(), bits, taps=et).cpu()
D = c["embed_dim"]
def codes(t, scale):
    return torch.round(t / torch.as_tensor(scale).reshape(1, 1, -1)).to(torch.int64)
pairs = [("cls", "qact1", st["qact1.scale"])]
for i in range(c["depth"]):
    p = "blocks.%d." % i
    pairs += [(p + "norm1", p + "attn.qact0", st[p + "attn.qact0.scale"]), (p + "attn.qact1", p + "attn.qact1", st[p + "attn.qact1.scale"]),
              (p + "attn.qact2", p + "attn.qact2", st[p + "attn.qact2.scale"]), (p + "qact2", p + "qact2", st[p + "qact2.scale"]),
              (p + "norm2", p + "mlp.qact0", st[p + "mlp.qact0.scale"]), (p + "mlp.qact1", p + "mlp.qact1", st[p + "mlp.qact1.scale"]),
              (p + "qact4", p + "qact4", st[p + "qact4.scale"])]
for exact in (False, True):
    print("---- engine vs oracle(exact_sums=%s)" % exact)
    taps = res[exact][1]
    for step, tap, sc in pairs:
        ref = codes(taps[tap], sc)
        got = et[step].cpu().reshape(ref.shape).to(torch.int64)
        d = (got != ref)
        if d.any():
            per_img = d.reshape(B, -1).sum(1).tolist()
            print("%-24s mismatches %7d / %d  maxabs %d  per-image %s" % (step, int(d.sum()), d.numel(), int((got - ref).abs().max()), per_img))
    print("logits mismatches:", int((logits != res[exact][0]).sum()), "top1 equal:", torch.equal(logits.argmax(1), res[exact][0].argmax(1)))
