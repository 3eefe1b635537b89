"""Which ops of the oracle depend on the torch backend?  Teacher-forced: the exact-sum oracle runs on the CPU, then every
LayerNorm / GELU+quant of the network is re-evaluated on CUDA from the CPU run's inputs and compared code by code."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

from oracle import port
from oracle.port import VitOracle
from p2vit_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "deit_tiny"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
g = np.load("tests/golden/%s_minmax.npz" % name)
st = {k[6:]: g[k] for k in g.files if k.startswith("state/")}
c = synth.VIT_CONFIGS[name]
sd = synth.synth_vit_state_dict(**c, seed=0)
x = synth.synth_images(B, seed=1)
bits = [8] * (4 * c["depth"] + 2)
o = VitOracle(sd, **c, exact_sums=True)
o.load_state(st)
taps = {}
o.forward_quant(x, bits, taps)
T = lambda k: torch.as_tensor(st[k]).float()
last = "qact1"
for i in range(c["depth"]):
    p = "blocks.%d." % i
    for norm, inq, outq, cs in ((p + "norm1", last, p + "attn.qact0", p + "attn"), (p + "norm2", p + "qact2", p + "mlp.qact0", p + "attn")):
        xin = taps[inq]
        out_scale = T(outq + ".scale") * T(cs + ".channel_scale")
        a = port.int_layernorm(xin, T(inq + ".scale"), out_scale, sd[norm + ".weight"], sd[norm + ".bias"], True)
        b = port.int_layernorm(xin.cuda(), T(inq + ".scale").cuda(), out_scale.cuda(), sd[norm + ".weight"].cuda(), sd[norm + ".bias"].cuda(), True).cpu()
        d = (a != b)
        if d.any():
            idx = d.nonzero()[:3]
            print(norm, "CPU vs CUDA LN differ:", int(d.sum()), "e.g.", [(tuple(j.tolist()), float(a[tuple(j)] / out_scale[j[-1]]), float(b[tuple(j)] / out_scale[j[-1]])) for j in idx])
    # GELU
    y = taps[p + "mlp.qact0"]
    last = p + "qact4"
# log2 / pow / sqrt on both backends over a dense set of floats near powers of two
v = torch.cat([(2.0 ** k) * (1 - torch.arange(1, 64) * 2.0 ** -24) for k in range(-8, 9)]).float()
print("log2 floor differs CPU/CUDA on", int((torch.floor(torch.log2(v)) != torch.floor(torch.log2(v.cuda())).cpu()).sum()), "of", v.numel(), "values just below powers of two")
print("log2(double).float() floor differs CPU/CUDA on", int((torch.floor(torch.log2(v.double()).float()) != torch.floor(torch.log2(v.cuda().double()).float()).cpu()).sum()))
print("CPU fp32 log2 floor vs CPU double->float floor differ on", int((torch.floor(torch.log2(v)) != torch.floor(torch.log2(v.double()).float())).sum()))
r = torch.rand(1 << 22) * 1000 + 1e-3
print("sqrt differs CPU/CUDA:", int((torch.sqrt(r) != torch.sqrt(r.cuda()).cpu()).sum()), " np.sqrt vs CUDA:", int((torch.from_numpy(np.sqrt(r.numpy())) != torch.sqrt(r.cuda()).cpu()).sum()))
n = torch.arange(0, 32).float()
print("pow(2,N) differs:", int((torch.pow(2, n) != torch.pow(2, n.cuda()).cpu()).sum()), int((torch.pow(2, -n) != torch.pow(2, -n.cuda()).cpu()).sum()))
yy = torch.randn(1 << 22) * 3
print("gelu differs CPU/CUDA:", int((F.gelu(yy) != F.gelu(yy.cuda()).cpu()).sum()), "of", yy.numel())
