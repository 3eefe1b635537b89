"""Times p2v_gemm_i8 alone on the block-GEMM shapes of a model (CUDA events, L2-warm loop of `iters` calls per shape).
usage: python tools/gemm_bench.py [deit_small|vit_base|deit_tiny|vit_large] [batch] [variant ...]"""
import os
import sys
import torch
from p2vit_b200 import ops

DIMS = {"deit_tiny": 192, "deit_small": 384, "vit_base": 768, "vit_large": 1024}


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "deit_small"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    variants = [int(v) for v in sys.argv[3:]] or [1, 2]
    D, M = DIMS[model], B * 197
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    fac = torch.tensor([1.0, 2.0, 4.0, 8.0])
    for name, K, N, epi in (("qkv", D, 3 * D, ops.EPI_REQUANT), ("proj", D, D, ops.EPI_RESIDUAL), ("fc1", D, 4 * D, ops.EPI_GELU),
                            ("fc2", 4 * D, D, ops.EPI_RESIDUAL)):
        A = torch.randint(-128, 128, (M, K), generator=g, dtype=torch.int32).to(torch.int8).to(dev)
        W = torch.randint(-20, 20, (N, K), generator=g, dtype=torch.int32).to(torch.int8).to(dev)
        bias = (torch.randn(N, generator=g) * 0.5).to(dev)
        acc_scale = torch.full((N,), 2.0 ** -13, device=dev)
        out_scale = torch.full((N,), 2.0 ** -5, device=dev) if epi != ops.EPI_RESIDUAL else (0.0171 * fac[torch.randint(0, 4, (N,), generator=g)]).to(dev)
        mid = (0.00931 * fac[torch.randint(0, 4, (N,), generator=g)]).to(dev)
        rs = (0.0123 * fac[torch.randint(0, 4, (N,), generator=g)]).to(dev)
        res = torch.randint(-128, 128, (M, N), generator=g, dtype=torch.int32).to(torch.int8).to(dev)
        outs = {}
        for v in variants:
            ops.set_gemm_variant(v)
            o8 = torch.empty(M, N, dtype=torch.int8, device=dev)
            kw = dict(bias=bias, out_scale=out_scale, out_i8=o8, pot=True)
            if epi == ops.EPI_RESIDUAL:
                kw.update(mid_scale=mid, res_scale=rs, res=res)
            if epi == ops.EPI_GELU:
                kw.update(gelu_table=ops.gelu_table(2.0 ** -5, dev))
            args = ops.gemm_args(A, W, epi, acc_scale, **kw)
            for _ in range(3 if not os.environ.get("GEMM_BENCH_ONCE") else 2):
                ops.gemm(args)
            torch.cuda.synchronize()
            if os.environ.get("GEMM_BENCH_ONCE"):     # profiling mode: two plain launches per shape
                outs[v] = o8
                continue
            iters = 20
            graph = torch.cuda.CUDAGraph()     # replayed graph: device time only, no host launch gaps
            with torch.cuda.graph(graph):
                for _ in range(iters):
                    ops.gemm(args)
            graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / iters
            outs[v] = o8
            print("%-10s %-5s M=%d N=%d K=%d variant %d: %8.1f us  %7.1f TOP/s  %6.0f Gelem/s" % (
                model, name, M, N, K, v, us, 2.0 * M * N * K / us * 1e-6, M * N / us * 1e-3), flush=True)
        if len(outs) == 2:
            a, b = list(outs.values())
            print("   identical:", bool(torch.equal(a, b)))
    ops.set_gemm_variant(0)


if __name__ == "__main__":
    main()
