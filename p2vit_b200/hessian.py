"""Hessian-trace sensitivity of the quantizable layers (reference: pyhessian/hessian.py:167-217 `trace`,
pyhessian/utils.py:61-100, test_quant.py:160-204).

The mixed-precision search ranks candidate bit configurations by  sum_i sensitivity_i * distance_i  (test_quant.py:350-368,
p2vit_b200/search.py); the reference ships the sensitivity vector as a hard-coded list per model and keeps the code that
produced it commented out.  This module regenerates it: Hutchinson's estimator  tr(H_i) ~ mean_v v^T H_i v  with Rademacher
probes, one weight tensor at a time, on the FP model (`hessian_statistic=True` forward: no smoothing, no quantizers), fp32
autograd double backward - the one part of the path that is not integer work and stays in PyTorch on the GPU.

    traces = [hessian_traces(model, criterion, x, y) for x, y in batches]          # 4*depth + 1 values each
    sensitivity = mean_normalised_sensitivity(traces)                               # what test_quant.py calls mean_hessian
"""
import numpy as np
import torch

__all__ = ["layer_parameters", "hessian_traces", "mean_normalised_sensitivity"]

_SKIP = ("norm", "bias", "cls_token", "pos_embed", "patch_embed")     # pyhessian/utils.py:72-79


def layer_parameters(model):
    """(names, parameters) of the tensors the reference takes Hessian traces of: every weight except norms, biases, class
    token, position and patch embedding - qkv / proj / fc1 / fc2 per block and the head, in module order (= bit_config[1:])."""
    names, params = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad or any(s in name for s in _SKIP):
            continue
        names.append(name)
        params.append(p)
    return names, params


def hessian_traces(model, criterion, inputs, targets, max_iter=150, tol=5e-3):
    """Per-layer Hutchinson trace estimates on one batch (pyhessian/hessian.py:167-217 with a single batch of data):
    probes are drawn with torch.randint_like, so torch.manual_seed fixes them; the running mean stops once it moves by less
    than `tol` relative.  Returns (names, traces)."""
    model.eval()
    names, params = layer_parameters(model)
    model.zero_grad()
    outputs = model(inputs, hessian_statistic=True)
    loss = criterion(outputs[0], targets)
    grads = torch.autograd.grad(loss, params, create_graph=True)
    traces = []
    for g, p in zip(grads, params):
        vhv, trace = [], 0.0
        for _ in range(max_iter):
            v = torch.randint_like(p, high=2)
            v[v == 0] = -1
            (hv,) = torch.autograd.grad(g, p, grad_outputs=v, only_inputs=True, retain_graph=True)
            vhv.append(float((hv * v).sum()))
            if abs(np.mean(vhv) - trace) / (abs(trace) + 1e-6) < tol:
                break
            trace = float(np.mean(vhv))
        traces.append(trace)
    return names, traces


def mean_normalised_sensitivity(trace_lists):
    """test_quant.py:184-201: per batch |trace| is min-max normalised over the layers, then averaged over the batches"""
    norm = []
    for tr in trace_lists:
        a = [abs(t) for t in tr]
        lo, hi = min(a), max(a)
        norm.append([(t - lo) / (hi - lo) for t in a])
    n = len(norm[0])
    return [sum(s[i] for s in norm) / len(norm) for i in range(n)]
