"""Times p2v_layernorm_int alone (CUDA graph replay of 20 launches): rows x C int8 -> int8.  usage: python tools/ln_bench.py [C] [rows]"""
import sys
import torch
from p2vit_b200 import ops

C = int(sys.argv[1]) if len(sys.argv) > 1 else 384
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 256 * 197
dev = "cuda"
g = torch.Generator().manual_seed(0)
x = torch.randint(-128, 128, (rows, C), generator=g, dtype=torch.int32).to(torch.int8).to(dev)
fac = torch.tensor([1.0, 2.0, 4.0, 8.0])
in_mult = fac[torch.randint(0, 4, (C,), generator=g)].to(dev)
gamma = (1.0 + 0.1 * torch.randn(C, generator=g)).to(dev)
beta = (0.1 * torch.randn(C, generator=g)).to(dev)
cs = (2.0 ** torch.randint(-1, 2, (C,), generator=g).float()).to(dev)
out_scale = (2.0 ** -5) * cs
out = torch.empty(rows, C, dtype=torch.int8, device=dev)
args = ops.layernorm_args(x, rows, C, C, in_mult, 0.0171, gamma, beta, out_scale, cs, 2.0 ** -5, True, out_i8=out)
for _ in range(3):
    ops.layernorm(args)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    for _ in range(20):
        ops.layernorm(args)
graph.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
graph.replay()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 20
print("layernorm rows=%d C=%d: %.1f us  %.0f GB/s (2 B/element)  %.0f Gelem/s  checksum %d" % (rows, C, us, 2.0 * rows * C / us * 1e-3, rows * C / us * 1e-3,
                                                                                       int(out.int().sum())))
