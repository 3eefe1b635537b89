"""Experiment: one batch as K independent sub-batches on K streams inside one CUDA graph, so the persistent kernels of one
sub-batch fill the tail (wave quantisation) of the other's.  Prints img/s for K = 1 (the bench's path) and K = 2, 4 and checks
the logits against the single-program run.  Usage: python tools/two_stream.py [--model deit_small] [--batch 256]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p2vit_b200 import Config, build_model, calibrate_model, synth  # noqa: E402
from p2vit_b200.engine import VitEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="deit_small")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--stagger", type=int, default=0, help="launch-order offset (in steps) between consecutive sub-batches")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    model = build_model(args.model, Config(True, True, "minmax"), seed=0, device=dev)
    calibrate_model(model, synth.synth_images(32, seed=0).to(dev))
    bits = [8] * (4 * model.depth + 2)
    B = args.batch
    x = synth.synth_images(min(B, 64), seed=1).to(dev)
    x = x.repeat((B + x.shape[0] - 1) // x.shape[0], 1, 1, 1)[:B].contiguous()

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    eng = VitEngine(model, use_graph=True)
    eng.static_input(B, bits).copy_(x)
    ref = eng.run_static(B, bits).clone()
    ms = timed(lambda: eng.run_static(B, bits))
    print("K=1  %.3f ms  %.0f img/s" % (ms, B / ms * 1e3))

    for K in (2, 4):
        if B % K:
            continue
        engs = [VitEngine(model, use_graph=False) for _ in range(K)]
        progs = []
        for k, e in enumerate(engs):
            e.static_input(B // K, bits).copy_(x[k * (B // K):(k + 1) * (B // K)])
            e.run_static(B // K, bits)                           # eager warm-up: kernel attributes, tables
            progs.append(e._program(tuple(bits), B // K))
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream(device=dev) for _ in range(K)]
        nsteps = len(progs[0]["steps"])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            cur = torch.cuda.current_stream()
            for s in streams:
                s.wait_stream(cur)
            for i in range(nsteps + args.stagger * (K - 1)):
                for k in range(K):
                    j = i - args.stagger * k
                    if 0 <= j < nsteps:
                        with torch.cuda.stream(streams[k]):
                            progs[k]["steps"][j][1]()
            for s in streams:
                cur.wait_stream(s)
        g.replay()
        torch.cuda.synchronize()
        got = torch.cat([p["ws"]["logits"] for p in progs])
        same = torch.equal(got, ref)
        ms = timed(g.replay)
        print("K=%d  %.3f ms  %.0f img/s  logits equal: %s" % (K, ms, B / ms * 1e3, same))


if __name__ == "__main__":
    main()
