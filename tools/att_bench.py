"""Times p2v_attention_i8 alone (CUDA graph replay of 10 launches) on B x H heads of 197 tokens.  usage: python tools/att_bench.py [H] [B]"""
import sys
import torch
from p2vit_b200 import intmath, ops

H = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
T, dh = 197, 64
dev = "cuda"
g = torch.Generator().manual_seed(0)
qkv = torch.randint(-40, 41, (B, T, 3, H, dh), generator=g, dtype=torch.int32).to(torch.int8).to(dev)
out = torch.empty(B, T, H * dh, dtype=torch.int8, device=dev)
s_q, s_attn, s_out = 2.0 ** -5, 2.0 ** -4, 2.0 ** -4
lut = intmath.lut_to_device(intmath.build_softmax_lut(s_attn), dev)
args = ops.attention_args(qkv, out, B, T, H, dh, s_q * s_q * dh ** -0.5 / s_attn, 2.0 ** -15 * s_q / s_out, lut)
for _ in range(3):
    ops.attention(args)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    for _ in range(10):
        ops.attention(args)
graph.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
graph.replay()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 10
print("attention B=%d H=%d T=%d: %.1f us  %.0f Gscore/s  checksum %d" % (B, H, T, us, B * H * T * T / us * 1e-3, int(out.int().sum())))
