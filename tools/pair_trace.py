"""Per-tile timeline of CTA pair 0 of the CTA-pair GEMM (library built with `make -C p2vit_b200/csrc trace`, loaded through P2V_LIB).
usage: P2V_LIB=p2vit_b200/csrc/libp2vit_b200_trace.so python tools/pair_trace.py [proj|fc2|qkv] [D]"""
import sys
import torch
from p2vit_b200 import ops

kind = sys.argv[1] if len(sys.argv) > 1 else "proj"
D = int(sys.argv[2]) if len(sys.argv) > 2 else 384
M = 256 * 197
K, N, epi = {"proj": (D, D, ops.EPI_RESIDUAL), "fc2": (4 * D, D, ops.EPI_RESIDUAL), "qkv": (D, 3 * D, ops.EPI_REQUANT), "fc1": (D, 4 * D, ops.EPI_GELU)}[kind]
dev = "cuda"
g = torch.Generator().manual_seed(0)
A = torch.randint(-128, 128, (M, K), generator=g, dtype=torch.int32).to(torch.int8).to(dev)
W = torch.randint(-20, 20, (N, K), generator=g, dtype=torch.int32).to(torch.int8).to(dev)
fac = torch.tensor([1.0, 2.0, 4.0, 8.0])
kw = dict(bias=(torch.randn(N, generator=g) * 0.5).to(dev), out_scale=(0.0171 * fac[torch.randint(0, 4, (N,), generator=g)]).to(dev),
          out_i8=torch.empty(M, N, dtype=torch.int8, device=dev), pot=True)
if epi == ops.EPI_RESIDUAL:
    kw.update(mid_scale=(0.00931 * fac[torch.randint(0, 4, (N,), generator=g)]).to(dev), res_scale=(0.0123 * fac[torch.randint(0, 4, (N,), generator=g)]).to(dev),
              res=torch.randint(-128, 128, (M, N), generator=g, dtype=torch.int32).to(torch.int8).to(dev))
else:
    kw["out_scale"] = torch.full((N,), 2.0 ** -5, device=dev)
    if epi == ops.EPI_GELU:
        kw["gelu_table"] = ops.gelu_table(2.0 ** -5, dev)
trace = torch.zeros(2 * 19 * 64 * 4, dtype=torch.int64, device=dev)
ops.set_gemm_variant(2)
args = ops.gemm_args(A, W, epi, torch.full((N,), 2.0 ** -13, device=dev), **kw)
for _ in range(3):
    ops.gemm(args)
torch.cuda.synchronize()
args.out_f32 = trace.data_ptr()
ops.gemm(args)
torch.cuda.synchronize()
t = trace.cpu().reshape(2, 19, 64, 4)
for rank in (0, 1):
    t0 = int(t[rank][t[rank] > 0].min())
    print("== CTA rank", rank, "(clock64 relative to first stamp)")
    for it in range(64):
        if t[rank, 3, it, 0] == 0:
            break
        r = lambda role, k: int(t[rank, role, it, k]) - t0 if t[rank, role, it, k] > 0 else -1
        print("tile %2d  prod: start %7d kblocks-issued %7d res-issued %7d | mma: wait-tempty %7d got %7d committed %7d | epi warp 0: start %7d rfull %7d tfull %7d stored %7d" % (
            it, r(2, 0), r(2, 1), r(2, 2), r(0, 0), r(0, 1), r(0, 2), r(3, 0), r(3, 1), r(3, 2), r(3, 3)))
        print("         epilogue warps tfull->stored: " + " ".join("%d:%d-%d" % (e, r(3 + e, 2), r(3 + e, 3)) for e in range(16)))
