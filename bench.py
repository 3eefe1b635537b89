#!/usr/bin/env python
"""Headline benchmark: images/s of the fully quantized (W8A8, power-of-two) DeiT-Small forward at 224^2 on
N B200s (BASELINE.json configs[1]); one "step" = one batch of synthetic images through the hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--model deit_small] [--batch 256]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1)
  python bench.py --impl reference ...      the reference's CPU algorithm (oracle port) on the host cores

Prints ONE JSON line on rank 0 (contract: task prompt, section 4): `value` = images/s with inputs resident in HBM
(CUDA events, max over ranks), `e2e` = the same through the public model call with pinned HOST images and a
device->host read of the logits inside the timed region, `roofline` for the dominant kernel family,
`cpu_baseline` = the oracle port timed on a bounded sample on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from p2vit_b200 import synth  # noqa: E402

# MACs per image = L*(12*N*D^2 + 2*N^2*D) + 196*768*D + 1000*D


SWIN_GMAC = {"swin_tiny": 4.4906e9, "swin_micro": None}


def macs_per_image(name):
    if name in synth.SWIN_CONFIGS:
        return SWIN_GMAC.get(name) or 0.0
    c = synth.VIT_CONFIGS[name]
    D, L, N = c["embed_dim"], c["depth"], 197
    return L * (12 * N * D * D + 2 * N * N * D) + 196 * 768 * D + 1000 * D


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def int8_peaks():
    """dense INT8 tensor throughput measured on this pool's B200s (tools/int8_peak.py: cuBLASLt int8 x int8 -> int32 at 8192^3 /
    16384 x 8192^2 through torch._int_mm; committed as profiles/int8_peak.json): (burst TOP/s for a kernel timed alone, sustained
    TOP/s for a seconds-long loop, source).  MEASURED_PEAKS.json has no int8 figure; without the file: 2 x the bf16 numbers."""
    p = os.path.join(ROOT, "profiles", "int8_peak.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["int8_tops_burst"]), float(d["int8_tops_sustained"]), "measured cuBLASLt int8 GEMM (profiles/int8_peak.json)"
    _, bf, bfs, src = peaks()
    return 2.0 * bf, 2.0 * bfs, "2 x %s bf16 cuBLAS (no int8 measurement found)" % src


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        try:      # NVML in-process: a sample costs ~0.1 ms, so even a 100 ms timed region gets tens of them
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = [0x8, 0x40, 0x20, 0x4]      # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
            while not self.stop_flag:
                r = get_reasons(h)
                self.rows.append([str(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), str(mx), str(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)] +
                                 ["Active" if r & b else "Not Active" for b in bits])
                time.sleep(0.004)
            return
        except Exception:
            pass
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def bind_to_gpu_numa_node(index):
    """Multi-GPU runs: keep this rank's threads (and so its pinned staging buffers, first touch) on the CPUs NVML names as local to
    its GPU - eight ranks pulling 154 MB of pixels per step across the socket interconnect is what held the 8-GPU e2e figure back.
    Returns the number of CPUs bound to (0 = left alone)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        n = os.cpu_count()
        mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index), (n + 63) // 64)
        cpus = [i for i in range(n) if (mask[i // 64] >> (i % 64)) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def cpu_reference_run(name, batch, steps, warmup, threads=None):
    """the reference's algorithm (oracle/port.py, fp32 fake-quant, torch CPU ops) on the host cores: BASELINE.md section 3 -
    all host threads, warm-up, then the BEST of `steps` forwards of `batch` images (W8A8)."""
    from oracle.port import VitOracle

    torch.set_num_threads(threads or os.cpu_count())
    c = synth.VIT_CONFIGS[name]
    o = VitOracle(synth.synth_vit_state_dict(**c, seed=0), **c)
    state = load_state(name)
    if state is None:
        t0 = time.time()
        o.calibrate(synth.synth_images(4, seed=0))
        calib_s = time.time() - t0
    else:
        o.load_state(state)
        calib_s = None
    x = synth.synth_images(batch, seed=1)
    bits = [8] * (4 * c["depth"] + 2)
    for _ in range(warmup):
        o.forward_quant(x, bits)
    best = None
    for _ in range(steps):
        t0 = time.time()
        out = o.forward_quant(x, bits)
        dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    return batch / best, best, torch.get_num_threads(), calib_s, out


def state_sha16(model):
    """fingerprint of the frozen quantizer state (every rank of a multi-GPU calibration must freeze the same one)"""
    import hashlib
    h = hashlib.sha256()
    for k, v in sorted(model.export_quant_state().items()):
        h.update(k.encode())
        h.update(v.contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def load_state(name):
    p = os.path.join(ROOT, "tests", "golden", "%s_minmax.npz" % name)
    if not os.path.isfile(p):
        return None
    g = np.load(p)
    return {k[6:]: g[k] for k in g.files if k.startswith("state/")}


def make_bit_config(kind, model):
    """1 + 4*depth + 1 entries.  'mixed': the sampling rule of the reference's search (test_quant.py:323-341): first layer 8 bit,
    the attention pair and the MLP pair of a block share a width, sum(MACs_i * bits_i) <= 1.1 * sum(MACs_i * 4), seed 0."""
    n = 4 * model.depth + 2
    if kind in ("8", "4"):
        return [int(kind)] * n
    return synth.mixed_bit_config(model.flops_list(), model.depth)


# The other BASELINE.json configs, measured in the same run as the headline (DeiT-S) line and reported under `configs`:
#   weak   = the per-GPU batch is fixed (like the headline);  strong = the job's batch is fixed and sharded over the GPUs
SECONDARY = [
    dict(key="C1_deit_tiny", model="deit_tiny", method="minmax", bits="8", batch_per_gpu=256),
    dict(key="C3_vit_base_percentile", model="vit_base", method="percentile", bits="8", batch_total=256),
    dict(key="C3_vit_base_omse", model="vit_base", method="omse", bits="8", batch_total=256),
    dict(key="C4_swin_tiny", model="swin_tiny", method="minmax", bits="8", batch_per_gpu=256),
    dict(key="C5_vit_large_mixed", model="vit_large", method="minmax", bits="mixed", batch_total=1024),
]


def run_secondary(spec, dev, rank, world, steps, warmup):
    """calibrate on the GPU(s), capture the forward, time `steps` graph replays (CUDA events, barrier on both sides, max over ranks)"""
    import torch.distributed as dist
    from p2vit_b200 import Config, build_model, calibrate_model
    from p2vit_b200.engine import VitEngine
    from p2vit_b200.runner import shard_range
    from p2vit_b200.swin_engine import SwinEngine

    name = spec["model"]
    is_swin = name in synth.SWIN_CONFIGS
    strong = "batch_total" in spec
    B = spec["batch_total"] // world if strong else spec["batch_per_gpu"]
    model = build_model(name, Config(True, True, spec["method"]), seed=0, device=dev)
    t0 = time.time()
    n_cal = 8 if world <= 8 else world
    a, b = shard_range(n_cal, rank, world)
    calibrate_model(model, synth.synth_images(b - a, seed=0, start=a).to(dev))
    calib_s = time.time() - t0
    sha = state_sha16(model)
    if world > 1:
        shas = [None] * world
        dist.all_gather_object(shas, sha)
        assert len(set(shas)) == 1, "%s: ranks froze different quantizer states: %s" % (spec["key"], shas)
    if is_swin:
        bits = [8]
        eng = SwinEngine(model, use_graph=True)
        img, run = eng.static_input(B), (lambda: eng.run_static(B))
        launches = eng.launches_per_forward()
    else:
        bits = make_bit_config(spec["bits"], model)
        eng = VitEngine(model, use_graph=True)
        img, run = eng.static_input(B, bits), (lambda: eng.run_static(B, bits))
        launches = eng.launches_per_forward(bits)
    src = synth.synth_images(min(B, 32), seed=1, start=rank * B).to(dev)
    img.copy_(src.repeat((B + src.shape[0] - 1) // src.shape[0], 1, 1, 1)[:B])
    for _ in range(warmup):
        run()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    value = B * world * steps / (ms * 1e-3)
    _, i8_sus, _ = int8_peaks()
    out = {"workload": "%s W%sA8 PoT weights, %s activations, 224x224" % (name, "4/8" if spec["bits"] == "mixed" else spec["bits"], spec["method"]),
           "value": value, "unit": "images/s", "ms_per_step": ms / steps, "steps": steps, "batch_per_gpu": B, "global_batch": B * world,
           "scaling": "strong" if strong else "weak", "calibration_seconds": round(calib_s, 2), "calibration_images": n_cal,
           "gpu_launches_per_step": launches, "state_sha16": sha,
           "tensor_fraction": 2.0 * macs_per_image(name) * value / world / (i8_sus * 1e12)}
    if spec["bits"] == "mixed":
        out["bit_config"] = "".join(str(x) for x in bits)
    del eng, model, img
    torch.cuda.empty_cache()
    return out


def main():
    # stdout carries the JSON line and nothing else: libraries that write to file descriptor 1 (NCCL prints its version banner there
    # when the box sets NCCL_DEBUG) go to stderr until the line is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--model", default="deit_small")
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--calib", type=int, default=32, help="calibration images (whole job)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--golden-state", action="store_true", help="load the reference-calibrated state instead of calibrating on the GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bits", default="8", help="'8', '4' or 'mixed' (seeded {4,8} draw under the 1.1 x 4-bit budget of test_quant.py:323-341)")
    ap.add_argument("--method", default="minmax", help="activation observer: minmax | ema | percentile | omse (config.py:19-27)")
    ap.add_argument("--configs", default="auto", help="'auto': a default (DeiT-S) run also measures the other BASELINE configs (DeiT-T, ViT-B "
                                                      "percentile / omse, Swin-T, ViT-L mixed at 1024 images per job) into `configs`; 'none' skips them")
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of back-to-back graph replay for the `sustained` figure (0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    workload = "%s W%sA8 PoT %s, 224x224, batch %d/GPU" % (args.model, args.bits if args.bits != "mixed" else "4/8", args.method, args.batch)

    if args.impl == "reference":
        if rank != 0:
            return
        cb = 32      # batch 32, warm-up 1, best of >= 3 forwards, all host threads
        ips, spstep, cores, calib_s, _ = cpu_reference_run(args.model, cb, max(3, min(args.steps, 5)), 1)
        line = {"impl": "reference", "metric": "images/sec (224^2, int8 PoT)", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
                "steps": max(3, min(args.steps, 5)), "warmup": 1, "ms_per_step": spstep * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 (fake-quant int8)", "data": "synthetic",
                "config": {"workload": workload, "sample": "batch %d per step on the host CPU, best step of the run" % cb},
                "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                 "sample": "oracle/port.py (the reference's algorithm op for op, bit-exact against the reference's golden vectors) quantized forward, batch %d, best of the steps" % cb},
                "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        _emit(line, real_stdout)
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else 0      # before any pinned allocation
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from p2vit_b200 import Config, build_model, calibrate_model, ops
    from p2vit_b200.engine import VitEngine
    from p2vit_b200.runner import shard_range
    from p2vit_b200.swin_engine import SwinEngine

    is_swin = args.model in synth.SWIN_CONFIGS

    model = build_model(args.model, Config(True, True, args.method), seed=0, device=dev)
    t0 = time.time()
    state = load_state(args.model) if args.golden_state else None
    if state is not None:
        model.load_quant_state(state)
        model.model_quant()
        calib_src = "reference-calibrated state (tests/golden)"
    else:
        s, e = shard_range(args.calib, rank, world)
        calibrate_model(model, synth.synth_images(e - s, seed=0, start=s).to(dev))
        calib_src = "calibrated on the GPU(s) from %d synthetic images" % args.calib
    calib_s = time.time() - t0
    # every rank of a data-parallel calibration must have frozen the same state (statistics all-reduced over NCCL)
    sha = state_sha16(model)
    if world > 1:
        shas = [None] * world
        dist.all_gather_object(shas, sha)
        assert len(set(shas)) == 1, "ranks froze different quantizer states: %s" % shas
    B = args.batch
    if is_swin:
        bits = [8]
        eng = SwinEngine(model, use_graph=True)
        eng_input = lambda: eng.static_input(B)
        eng_run = lambda: eng.run_static(B)
        eng_prog = lambda: eng._program(B)
    else:
        bits = make_bit_config(args.bits, model)
        eng = VitEngine(model, use_graph=True)
        eng_input = lambda: eng.static_input(B, bits)
        eng_run = lambda: eng.run_static(B, bits)
        eng_prog = lambda: eng._program(tuple(bits), B)
    model._engine = eng
    host = synth.synth_images(min(B, 64), seed=1, start=rank * B)
    host = host.repeat((B + host.shape[0] - 1) // host.shape[0], 1, 1, 1)[:B].contiguous().pin_memory()
    img = eng_input()
    img.copy_(host, non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM, CUDA-graph replay of the whole forward
    for _ in range(args.warmup):
        eng_run()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        eng_run()
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_value = float(ms)

    # ---------------- sustained: the same replay back to back for >= args.sustain seconds (clocks sampled on their own)
    sustained = None
    if args.sustain > 0:
        n_sus = max(args.steps, int(args.sustain * 1e3 / (ms_value / args.steps)) + 1)
        sus_sampler = ClockSampler(local_rank)
        if rank == 0:
            sus_sampler.start()
        barrier()
        ev0.record()
        for _ in range(n_sus):
            eng_run()
        ev1.record()
        barrier()
        sus_sampler.stop_flag = True
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sustained = {"value": B * world * n_sus / (float(ms) * 1e-3), "unit": "images/s", "steps": n_sus, "seconds": round(float(ms) * 1e-3, 3),
                     "ms_per_step": float(ms) / n_sus, "clocks": sus_sampler.summary() if rank == 0 else None}

    # ---------------- e2e: pinned host images -> H2D -> model(x, bits) -> logits D2H, every step
    out_host = torch.empty((B, 1000), dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    staging_f32 = [torch.empty_like(img) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def e2e_steps(n, host=host, staging=staging_f32):
        main_s = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            staging[0].copy_(host, non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(n):
            cur, nxt = i & 1, (i + 1) & 1
            if i + 1 < n:   # prefetch the next batch while this one computes
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(consumed[nxt])
                    staging[nxt].copy_(host, non_blocking=True)
                    ready[nxt].record(copy_stream)
            main_s.wait_event(ready[cur])
            logits = model(staging[cur], bits)[0]       # public API call: D2D into the program's input + graph replay
            consumed[cur].record(main_s)
            out_host.copy_(logits, non_blocking=True)
        main_s.synchronize()

    e2e_steps(args.warmup)
    barrier()
    ev0.record()
    e2e_steps(args.steps)
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms)

    # ---------------- e2e_u8: the same call with the decoder's 8-bit pixels (ToTensor + Normalize + qact_input through the
    # per-channel code table on the device, bit-identical logits: tests/test_gpu_model.py::test_uint8_pixels_equal_host_normalised_fp32)
    ms_e2e_u8 = None
    if getattr(model, "input_quant", True):
        from p2vit_b200.data import PREPROCESS
        fam = "swin" if is_swin else ("deit" if args.model.startswith("deit") else "vit")
        model.set_pixel_normalization(PREPROCESS[fam]["mean"], PREPROCESS[fam]["std"])
        mean = torch.tensor(PREPROCESS[fam]["mean"]).view(1, 3, 1, 1)
        std = torch.tensor(PREPROCESS[fam]["std"]).view(1, 3, 1, 1)
        host_u8 = host.mul(std).add(mean).mul(255).round().clamp(0, 255).to(torch.uint8).contiguous().pin_memory()
        staging_u8 = [torch.empty(host_u8.shape, dtype=torch.uint8, device=dev) for _ in range(2)]
        e2e_steps(args.warmup, host_u8, staging_u8)
        barrier()
        ev0.record()
        e2e_steps(args.steps, host_u8, staging_u8)
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_e2e_u8 = float(ms)
    sampler.stop_flag = True

    # ---------------- per-kernel-family device time (one CUDA graph per step, CUDA events on the launching stream)
    prog = eng_prog()
    fam_ms, fam_n = {}, {}

    def family(step):
        if is_swin:
            parts = step.split(".")
            tail2 = ".".join(parts[-2:])
            if step == "patchify" or parts[-1] == "gather" or step == "qact3":
                return "remap"
            if tail2 == "attn.qact3":
                return "attention"
            if step in ("patch_embed.qact", "qact2") or tail2 in ("mlp.qact0", "downsample.qact1") or \
                    (parts[-1] == "qact1" and len(parts) >= 2 and parts[-2].isdigit()):
                return "layernorm"
            if tail2 == "attn.qact1":
                return "gemm_qkv"
            if tail2 == "mlp.qact1":
                return "gemm_fc1"
            if parts[-1] == "qact2" and len(parts) >= 2 and parts[-2].isdigit():
                return "gemm_proj"
            if parts[-1] == "qact4":
                return "gemm_fc2"
            return "gemm_other"
        if step in ("patchify", "cls"):
            return step
        if "norm" in step or step == "qact2" and "blocks" not in step:
            return "layernorm"
        if step.endswith("attn.qact2"):
            return "attention"
        for suffix, kind in ((".attn.qact1", "gemm_qkv"), (".qact2", "gemm_proj"), (".mlp.qact1", "gemm_fc1"), (".qact4", "gemm_fc2")):
            if step.startswith("blocks.") and step.endswith(suffix):
                return kind
        return "gemm_other"

    # Each step `inner` times back to back (every step is idempotent) inside a CUDA graph of its own, the replay timed with CUDA
    # events: eager launches through ctypes cost the CPU longer than the 20-35 us kernels run (host-side tensor-map encoding, Python),
    # so an eager loop bills launch gaps to the short kernels (LayerNorm 29.6 us eager against 22 us in a graph, r2).
    from p2vit_b200 import ops as _ops
    reps, inner = 3, 4
    for step, fn in prog["steps"]:
        fn()
        n0 = _ops.launch_count() if hasattr(_ops, "launch_count") else 1
        fn()
        empty = hasattr(_ops, "launch_count") and _ops.launch_count() == n0      # (ViT-L has no patchify kernel: nothing to capture)
        g1 = None if empty else _ops.capture_graph(lambda fn=fn: [fn() for _i in range(inner)])
        if g1 is not None:
            g1.replay()
        f = family(step)
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if g1 is not None:
                g1.replay()
            b.record()
            b.synchronize()
            fam_ms[f] = fam_ms.get(f, 0.0) + a.elapsed_time(b) / inner
            fam_n[f] = fam_n.get(f, 0) + 1
        del g1
    fine_ms = {k: round(v / reps, 4) for k, v in fam_ms.items()}
    fam_ms = {k: v / reps for k, v in fam_ms.items() if not k.startswith("gemm")}
    fam_ms["gemm"] = sum(v for k, v in fine_ms.items() if k.startswith("gemm"))
    fam_n = {k: v // reps for k, v in fam_n.items()}
    fam_n["gemm"] = sum(v for k, v in fam_n.items() if k.startswith("gemm"))

    # ---------------- the other BASELINE configs (every rank takes part; rank 0 reports)
    launches_per_step = eng.launches_per_forward() if is_swin else eng.launches_per_forward(bits)
    main_state = model.export_quant_state() if (is_swin and not args.no_cpu_baseline and rank == 0) else None
    configs = None
    if args.configs == "auto" and args.model == "deit_small" and args.bits == "8" and args.method == "minmax":
        del prog, eng, staging_f32, img
        model._engine = None
        torch.cuda.empty_cache()
        configs = {}
        for spec in SECONDARY:
            try:
                configs[spec["key"]] = run_secondary(spec, dev, rank, world, 10, 3)
            except Exception as e:      # a secondary config must never cost the headline line
                configs[spec["key"]] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
                if world > 1:
                    raise

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_gbs, bf16_tf, bf16_tf_sus, peak_src = peaks()
    i8_burst, i8_sus, i8_src = int8_peaks()
    total_imgs = B * world * args.steps
    value = total_imgs / (ms_value * 1e-3)
    e2e = total_imgs / (ms_e2e * 1e-3)
    macs = macs_per_image(args.model)
    if is_swin:
        D, L, T1 = 0, 0, 0
    else:
        c = synth.VIT_CONFIGS[args.model]
        D, L, T1 = c["embed_dim"], c["depth"], 197
    top = max(fam_ms, key=fam_ms.get)
    share = {k: round(v / sum(fam_ms.values()), 4) for k, v in fam_ms.items()}
    if top == "gemm":
        lin_macs = 4.3504e9 if is_swin else L * 12 * T1 * D * D + 196 * 768 * D + 1000 * D     # Swin-T linear part: SURVEY 8(d)
        ach = 2.0 * lin_macs * B / (fam_ms["gemm"] * 1e-3) / 1e12
        peak = i8_burst
        roof = {"bound": "tensor", "kernel": "gemm_pair_kernel / gemm_tc_kernel (all %d GEMM launches of a step)" % fam_n["gemm"], "achieved": ach, "peak": peak,
                "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "note": "int8 TOP/s; each launch is timed on its own (CUDA events around a graph of 4 launches of the same step), so the peak is the BURST figure of the %s: %.0f "
                        "(sustained %.0f; nominal dense int8 4500; 2 x measured bf16 burst = %.0f)" % (i8_src, i8_burst, i8_sus, 2.0 * bf16_tf)}
    else:
        if is_swin:
            by = B * 56 * 56 * 96 * 4 * 6               # qkv codes in + attention codes out: 4*tokens*C bytes per block, ~constant per stage pair
        elif top == "attention":
            by = L * B * (T1 * 3 * D + T1 * D)          # qkv codes in, attention codes out
        elif top == "layernorm":
            by = (2 * L) * B * T1 * D * 2
        else:
            by = B * 3 * 224 * 224 * 5
        ach = by / (fam_ms[top] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "%s (all %d launches of a step)" % (top, fam_n[top]), "achieved": ach, "peak": hbm_gbs, "unit": "GB/s",
                "frac": ach / hbm_gbs, "traffic": None, "note": "algorithmic int8 bytes in+out; peak = %s copy bandwidth" % peak_src}
    # measured DRAM traffic per launch of the kernel family above (ncu capture of the same workload, profiles/ncu_traffic.json)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("%s_b%d" % (args.model, B)) if args.bits == "8" else None
    except OSError:
        tr = None
    if tr:
        kinds = [k for k in tr if k.startswith("gemm_")] if top == "gemm" else [top]
        if all(k in tr for k in kinds) and kinds:
            roof["traffic"] = sum(tr[k] for k in kinds) / len(kinds)
            roof["traffic_note"] = "bytes per launch, mean over %s; %s" % ("/".join(kinds), tr["source"])
    roof["per_launch"] = {"launches_per_step": fam_n[top], "avg_us": round(fam_ms[top] * 1e3 / fam_n[top], 2)}
    roof["device_ms_per_step_by_family"] = {k: round(v, 4) for k, v in fam_ms.items()}
    roof["share_of_step"] = share
    roof["gemm_ms_by_kind"] = {k: v for k, v in fine_ms.items() if k.startswith("gemm")}
    if not is_swin:      # achieved int8 TOP/s of each block GEMM kind (2 * MACs / device time)
        kind_macs = {"gemm_qkv": 3 * D * D, "gemm_proj": D * D, "gemm_fc1": 4 * D * D, "gemm_fc2": 4 * D * D}
        roof["gemm_tops_by_kind"] = {k: round(2.0 * L * B * T1 * m / (fine_ms[k] * 1e-3) / 1e12, 1) for k, m in kind_macs.items() if fine_ms.get(k)}
    if not is_swin:      # the row kernels against the HBM roofline: algorithmic int8 bytes in + out / device time / measured copy bandwidth
        hv = {}
        if fam_ms.get("attention"):
            by = L * B * T1 * 4 * D                       # qkv codes in (3D), attention codes out (D), per token and layer
            hv["attention"] = {"GB/s": round(by / (fam_ms["attention"] * 1e-3) / 1e9, 1), "frac": round(by / (fam_ms["attention"] * 1e-3) / 1e9 / hbm_gbs, 4)}
        if fam_ms.get("layernorm"):
            by = 2 * L * B * T1 * D * 2                   # two LayerNorms per block, codes in + codes out
            hv["layernorm"] = {"GB/s": round(by / (fam_ms["layernorm"] * 1e-3) / 1e9, 1), "frac": round(by / (fam_ms["layernorm"] * 1e-3) / 1e9 / hbm_gbs, 4)}
        if fam_ms.get("patchify"):
            by = B * 3 * 224 * 224 * 5                    # fp32 pixels in, int8 codes out
            hv["patchify"] = {"GB/s": round(by / (fam_ms["patchify"] * 1e-3) / 1e9, 1), "frac": round(by / (fam_ms["patchify"] * 1e-3) / 1e9 / hbm_gbs, 4)}
        roof["hbm_view"] = {"peak_GB/s": hbm_gbs, "note": "neither attention nor LayerNorm is HBM-bound: both are bound by the bit-exact integer arithmetic per element (DESIGN.md 3.2, 3.5)", **hv}
    roof["tensor_fraction_of_whole_forward"] = 2.0 * macs * value / world / (i8_sus * 1e12)      # whole step vs the SUSTAINED int8 peak
    roof["int8_peak"] = {"burst_tops": i8_burst, "sustained_tops": i8_sus, "source": i8_src}

    cpu = None
    if not args.no_cpu_baseline:
        if is_swin:
            from oracle.swin_port import SwinOracle
            torch.set_num_threads(os.cpu_count())
            cs = synth.SWIN_CONFIGS[args.model]
            o = SwinOracle(synth.synth_swin_state_dict(**cs, seed=0), **cs)
            o.load_state({k: v.numpy() for k, v in main_state.items()})
            xs = synth.synth_images(4, seed=1)
            o.forward_quant(xs)
            t0 = time.time()
            o.forward_quant(xs)
            cpu = {"value": 4 / (time.time() - t0), "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": "oracle/swin_port.py (reference algorithm, torch CPU fp32 fake-quant) forward, batch 4 x 1 step, GPU-calibrated state"}
        else:
            ips, spstep, cores, _, ref_logits = cpu_reference_run(args.model, 32, 3, 1)
            cpu = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                   "sample": "oracle/port.py (reference algorithm, torch CPU fp32 fake-quant) forward, batch 32, warm-up 1, best of 3 (BASELINE.md section 3)"}

    line = {"metric": "images/sec (224^2, int8 PoT)", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8 (int32 accumulate, fp32 requant epilogue)", "data": "synthetic",
            "config": {"workload": workload, "global_batch": B * world, "bit_config": ("[%s]*%d" % (args.bits, len(bits))) if args.bits in ("8", "4") else "".join(str(b) for b in bits), "calibration": calib_src,
                       "calibration_seconds": round(calib_s, 2), "calibration_state_sha16": sha, "ranks_froze_identical_state": True, "l2": "inputs larger than L2 (fp32 images %.0f MB + int8 workspace per step)" % (B * 3 * 224 * 224 * 4 / 1e6),
                       "parallelism": "dp%d (batch sharded, no collective in the forward)" % world, "cuda_graph": True,
                       "host_cpus_bound_to_gpu_numa_node": numa_cpus},
            # `value` / the fields above: the CONTRACT path - the call the reference's own driver makes, model(fp32 normalised pixels)
            # (test_quant.py:492).  u8_*: the same call on the decoder's 8-bit pixels (p2vit_b200 extension, bit-identical logits,
            # a quarter of the PCIe bytes) - what a deployment that owns its loader would use.
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": B * 3 * 224 * 224 * 4 * world, "d2h_bytes_per_step": B * 1000 * 4 * world,
                    "ms_per_step": ms_e2e / args.steps, "contract_path": "fp32",
                    "path": "pinned host fp32 images -> H2D (copy stream, double buffered) -> model(x, bit_config) -> logits D2H",
                    "u8_value": None if ms_e2e_u8 is None else total_imgs / (ms_e2e_u8 * 1e-3),
                    "u8_h2d_bytes_per_step": None if ms_e2e_u8 is None else B * 3 * 224 * 224 * world,
                    "u8_ms_per_step": None if ms_e2e_u8 is None else ms_e2e_u8 / args.steps,
                    "u8_path": "pinned host uint8 pixels -> H2D -> model(x_u8, bit_config) (code-table patchify) -> logits D2H; same logits as the fp32 path"},
            "sustained": sustained, "configs": configs,
            "gpu_launches": launches_per_step * args.steps,
            "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu}
    _emit(line, real_stdout)
    if world > 1:
        dist.destroy_process_group()


def _emit(line, real_stdout):
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
