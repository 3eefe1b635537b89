"""CPU: the Hessian-trace sensitivity pass (p2vit_b200/hessian.py) against the reference's own pyhessian on the same weights,
batch and probe seed (build container only - needs /root/reference), plus the host logic without it."""
import sys

import pytest
import torch

from p2vit_b200 import Config, build_model, hessian, synth


def test_layer_selection_and_normalisation():
    m = build_model("vit_micro", Config(), seed=0, device="cpu")
    names, params = hessian.layer_parameters(m)
    assert len(names) == 4 * m.depth + 1 and names[-1] == "head.weight"
    assert names[:4] == ["blocks.0.attn.qkv.weight", "blocks.0.attn.proj.weight", "blocks.0.mlp.fc1.weight", "blocks.0.mlp.fc2.weight"]
    s = hessian.mean_normalised_sensitivity([[1.0, -3.0, 2.0], [10.0, 20.0, 30.0]])
    assert s == [(0.0 + 0.0) / 2, (1.0 + 0.5) / 2, (0.5 + 1.0) / 2]


def test_traces_are_deterministic_and_positive_on_average():
    m = build_model("vit_micro", Config(), seed=0, device="cpu")
    x, y = synth.synth_images(2, seed=3), torch.tensor([3, 7])
    crit = torch.nn.CrossEntropyLoss()
    torch.manual_seed(0)
    n1, t1 = hessian.hessian_traces(m, crit, x, y, max_iter=6)
    torch.manual_seed(0)
    n2, t2 = hessian.hessian_traces(m, crit, x, y, max_iter=6)
    assert n1 == n2 and t1 == t2 and len(t1) == 4 * m.depth + 1
    assert all(abs(t) > 0 for t in t1)


@pytest.mark.reference
def test_traces_equal_reference_pyhessian():
    """same weights, same batch, same torch seed -> the same Rademacher probes: traces equal the reference's (fp32 round-off)"""
    from oracle.gen_golden import build_reference_vit
    ref_model, c, _, _ = build_reference_vit("vit_micro")
    sys.path.insert(0, "/root/reference")
    from pyhessian import hessian as ref_hessian

    m = build_model("vit_micro", Config(), seed=0, device="cpu")
    x, y = synth.synth_images(2, seed=3), torch.tensor([3, 7])
    crit = torch.nn.CrossEntropyLoss()
    torch.manual_seed(5)
    names, got = hessian.hessian_traces(m, crit, x, y, max_iter=8)
    torch.manual_seed(5)
    ref_names, want = ref_hessian(ref_model, crit, data=(x, y), cuda=False).trace(maxIter=8)
    assert names == ref_names
    assert torch.allclose(torch.tensor(got, dtype=torch.float64), torch.tensor([float(w) for w in want], dtype=torch.float64), rtol=1e-3, atol=1e-6), (got, want)
