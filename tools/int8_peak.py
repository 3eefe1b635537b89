"""Measured dense INT8 tensor throughput on this box (BASELINE.md section 4 leaves the figure to the builder): cuBLASLt's
int8 x int8 -> int32 GEMM through torch._int_mm at 8192^3 and 16384x8192x8192, best of 10 single launches (burst) and back to
back for ~3 s (sustained), CUDA events.  Writes profiles/int8_peak.json; bench.py reads it for roofline.peak."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(M, N, K, sustain_s=3.0):
    a = torch.randint(-128, 127, (M, K), dtype=torch.int8, device="cuda")
    b = torch.randint(-128, 127, (K, N), dtype=torch.int8, device="cuda")     # row-major [K, N]
    bt = torch.randint(-128, 127, (N, K), dtype=torch.int8, device="cuda").t()  # column-major B (K-major, like nn.Linear weights)
    out = {}
    for name, bb in (("b_rowmajor", b), ("b_kmajor", bt)):
        for _ in range(3):
            torch._int_mm(a, bb)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, bb)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        ops = 2.0 * M * N * K
        burst = ops / (best * 1e-3) / 1e12
        n = max(10, int(sustain_s / (best * 1e-3)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            torch._int_mm(a, bb)
        e1.record()
        e1.synchronize()
        sus = ops * n / (e0.elapsed_time(e1) * 1e-3) / 1e12
        out[name] = {"burst_tops": round(burst, 1), "sustained_tops": round(sus, 1), "best_ms": round(best, 4), "launches_sustained": n}
    return out


def main():
    res = {"how": "torch._int_mm (cuBLASLt int8 x int8 -> int32), CUDA events; burst = best of 10 single launches, sustained = back to back ~3 s",
           "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    for shape in ((8192, 8192, 8192), (16384, 8192, 8192)):
        res["%dx%dx%d" % shape] = run(*shape)
    vals = [v[k] for s, d in res.items() if isinstance(d, dict) for v in d.values() for k in ("burst_tops",)]
    sus = [v["sustained_tops"] for s, d in res.items() if isinstance(d, dict) for v in d.values()]
    res["int8_tops_burst"] = max(vals)
    res["int8_tops_sustained"] = max(sus)
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "int8_peak.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
