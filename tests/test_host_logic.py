"""CPU: host-side logic of the product - softmax table builder vs the oracle, data-parallel sharding, the drop-in
surface (names, signatures, defaults), seeded synthetic data, and the world_size-2 statistics all-reduce (gloo)."""
import inspect
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import port
from p2vit_b200 import Config, intmath, synth
from p2vit_b200.runner import accuracy, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lut_softmax(codes, lut):
    """numpy emulation of csrc softmax_kernel / attention: exact table sums + log2 rounding"""
    c = codes.astype(np.int64)
    d = c.max(-1, keepdims=True) - c
    e_int = (lut["hi"].astype(object) << 32) + lut["lo"].astype(object)
    tot = e_int[d].sum(-1, keepdims=True)
    tot_f = np.vectorize(lambda v: port._int_to_f32_rne(int(v)), otypes=[np.float32])(tot)
    x = np.rint((tot_f / lut["exp_f32"][d]).astype(np.float32))
    u = x.view(np.uint32).astype(np.int64)
    big = ((u + 0x00400000) >> 23) - 127
    p = np.where(big >= 16, 0.0, 2.0 ** (-np.minimum(big, 15).astype(np.float64))).astype(np.float32)
    return p


@pytest.mark.parametrize("log2s", [-2, -3, -4, -6, -9])
def test_softmax_table_reproduces_oracle(log2s):
    s = torch.tensor([2.0 ** log2s])
    g = torch.Generator().manual_seed(log2s + 50)
    codes = torch.randint(-128, 128, (3, 2, 30, 197), generator=g)
    ref = port.int_softmax_log2(codes.float() * s, s, 4, exact_sums=True).numpy()
    got = _lut_softmax(codes.numpy(), intmath.build_softmax_lut(s))
    assert np.array_equal(got, ref)


def test_softmax_table_rejects_unrepresentable_scale():
    with pytest.raises(NotImplementedError):
        intmath.build_softmax_lut(torch.tensor(2.0 ** -40))


def test_is_pot():
    assert intmath.is_pot(torch.tensor([0.5, 2.0 ** -9, 4.0])) and not intmath.is_pot(torch.tensor([0.5, 0.3]))


def test_shard_range_partitions_exactly():
    for n in (1, 7, 32, 256, 1000):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_synthetic_data_is_deterministic_and_shardable():
    a = synth.synth_images(4, seed=3)
    b = torch.cat([synth.synth_images(2, seed=3), synth.synth_images(2, seed=3, start=2)])
    assert torch.equal(a, b)
    sd1 = synth.synth_vit_state_dict(**synth.VIT_CONFIGS["vit_micro"], seed=0)
    sd2 = synth.synth_vit_state_dict(**synth.VIT_CONFIGS["vit_micro"], seed=0)
    assert all(torch.equal(sd1[k], sd2[k]) for k in sd1)
    assert float(sd1["blocks.0.attn.qkv.weight"].abs().sum()) != float(synth.synth_vit_state_dict(**synth.VIT_CONFIGS["vit_micro"], seed=1)["blocks.0.attn.qkv.weight"].abs().sum())


def test_config_defaults_match_reference():
    c = Config()
    assert (c.BIT_TYPE_W.name, c.BIT_TYPE_A.name, c.BIT_TYPE_S.name) == ("int4", "int8", "uint4")
    assert (c.OBSERVER_W, c.OBSERVER_A, c.OBSERVER_A_LN, c.OBSERVER_S) == ("minmax", "minmax", "ptf", "minmax")
    assert (c.CALIBRATION_MODE_W, c.CALIBRATION_MODE_A, c.CALIBRATION_MODE_A_LN) == ("channel_wise", "layer_wise", "channel_wise")
    assert c.INT_SOFTMAX and c.INT_NORM and c.QUANTIZER_S == "log2"
    c = Config(ptf=False, lis=False, quant_method="ema")
    assert not c.INT_NORM and not c.INT_SOFTMAX and c.OBSERVER_A_LN == "ema" and c.BIT_TYPE_S.name == "uint8"
    assert Config("False", "0").INT_NORM is False


def test_drop_in_surface():
    import p2vit_b200 as P
    from p2vit_b200.ptq import BIT_TYPE_DICT
    from p2vit_b200.ptq.observer import str2observer
    from p2vit_b200.ptq.quantizer import str2quantizer

    assert sorted(BIT_TYPE_DICT) == ["int4", "int8", "uint3", "uint4", "uint8"]
    assert (BIT_TYPE_DICT["int4"].lower_bound, BIT_TYPE_DICT["int4"].upper_bound, BIT_TYPE_DICT["uint4"].upper_bound) == (-8, 7, 15)
    assert sorted(str2observer) == ["ema", "minmax", "omse", "percentile", "ptf"] and sorted(str2quantizer) == ["log2", "uniform"]
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(P.QLinear.forward) == ["self", "x", "global_distance", "bit_config", "weight_smoothed", "attn", "attn_para"]
    assert sig(P.QAct.forward) == ["self", "x", "asymmetric", "attn", "attn_para"]
    assert sig(P.QIntLayerNorm.forward) == ["self", "x", "in_quantizer", "out_quantizer", "out_quantizer_scale", "in_scale_expand"]
    assert sig(P.QIntSoftmax.forward) == ["self", "x", "scale"] and sig(P.QConv2d.forward) == ["self", "x", "bit_config"]
    assert sig(P.QLinear.__init__)[1:] == ["in_features", "out_features", "bias", "quant", "calibrate", "last_calibrate", "bit_type",
                                           "calibration_mode", "observer_str", "quantizer_str"]
    m = P.deit_tiny_patch16_224(cfg=Config())
    ref_keys = set(synth.synth_vit_state_dict(**synth.VIT_CONFIGS["deit_tiny"]))
    assert set(m.state_dict()) == ref_keys
    for meth in ("model_quant", "model_dequant", "model_open_calibrate", "model_open_last_calibrate", "model_close_calibrate"):
        assert hasattr(m, meth)
    assert sig(m.forward) == ["x", "bit_config", "plot", "hessian_statistic"]
    assert len(m.flops_list()) == 50
    with pytest.raises(RuntimeError):
        P.deit_tiny_patch16_224(pretrained=True, cfg=Config())


def test_accuracy_counts():
    out = torch.tensor([[0.1, 0.9, 0.0], [0.8, 0.1, 0.1], [0.2, 0.3, 0.5]])
    c1, c2 = accuracy(out, torch.tensor([1, 2, 2]), (1, 2))
    assert float(c1) == 2 and float(c2) == 2


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from p2vit_b200.ptq.observer.utils import allreduce_
from p2vit_b200.runner import shard_range
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
full = torch.arange(40, dtype=torch.float32).reshape(10, 4) * (1 if True else 0)
a, b = shard_range(10, r, w)
mine = full[a:b]
mx = allreduce_(mine.max(0).values.clone(), "max"); mn = allreduce_(mine.min(0).values.clone(), "min")
sc = allreduce_((mine ** 2).sum(0).double(), "sum")
assert torch.equal(mx, full.max(0).values) and torch.equal(mn, full.min(0).values) and torch.equal(sc, (full ** 2).sum(0).double())
print("rank", r, "ok")
dist.destroy_process_group()
'''


def test_statistics_allreduce_world2_gloo(tmp_path):
    """calibration statistics are MAX/MIN/SUM all-reduced so every rank freezes identical scales (NCCL on GPUs, gloo here)"""
    script = tmp_path / "w.py"
    script.write_text(_WORKER % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29641", str(script)], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


# ------------------------------------------------------------------------------------------------ percentile observer (host logic)
def _np_hist_fn(x):
    """numpy stand-in for ops.radix_hist (the CUDA kernel; tests/test_gpu_ops.py checks the kernel against this very function)"""
    import numpy as np
    u = x.view(np.uint32)
    key = np.where(u & 0x80000000, ~u, u | 0x80000000).astype(np.uint32)

    def fn(mask, value, shift, nbits):
        sel = key[(key & np.uint32(mask)) == np.uint32(value)]
        return torch.from_numpy(np.bincount((sel >> np.uint32(shift)) & np.uint32((1 << nbits) - 1), minlength=1 << nbits).astype(np.int64))
    return fn


def test_percentile_select_and_interpolation_equal_the_reference_libraries():
    """observer/percentile.py:26-55 sorts (torch.quantile; np.percentile above 2^24 elements).  The radix select + the restated
    interpolation arithmetic return the same fp32 value, bit for bit, in both size classes - ties, negative values, both tails."""
    import numpy as np
    from p2vit_b200.ptq.observer.percentile import quantile_from_order_statistics, select_kth
    rng = np.random.default_rng(0)
    local = lambda t, op: t
    for n in (1000, 123457, 2_400_000):
        x = (rng.standard_normal(n) * 3).astype(np.float32)
        x[::7] = x[3]
        fn, srt = _np_hist_fn(x), np.sort(x)
        for k in (0, n // 3, n - 2, n - 1):
            assert select_kth(fn, k, allreduce=local) == srt[k]
        kth = lambda k: select_kth(fn, k, allreduce=local)
        for q in (0.99999, 1 - 0.99999, 0.5):
            assert quantile_from_order_statistics(kth, n, q) == float(torch.quantile(torch.from_numpy(x), q)), (n, q)
    n = 16_777_300                                  # torch.quantile refuses: the reference's numpy branch (percentile.py:36-43)
    x = rng.standard_normal(n).astype(np.float32)
    with pytest.raises(RuntimeError):
        torch.quantile(torch.from_numpy(x), 0.5)
    srt = np.sort(x)
    for q in (0.99999, 1 - 0.99999):
        want = float(torch.tensor(np.percentile(x, q * 100), dtype=torch.float32))
        assert quantile_from_order_statistics(lambda k: srt[k], n, q) == want, q


_PCT_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
sys.path.insert(0, os.path.join(%r, "tests"))
from test_host_logic import _np_hist_fn
from p2vit_b200.ptq.observer.percentile import quantile_from_order_statistics, select_kth
from p2vit_b200.ptq.observer.utils import allreduce_
from p2vit_b200.runner import shard_range
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
full = (np.random.default_rng(5).standard_normal(300007) * 2).astype(np.float32)
a, b = shard_range(full.size, r, w)
fn = _np_hist_fn(full[a:b].copy())                       # this rank's shard only
n = int(allreduce_(torch.tensor([b - a]), "sum"))
kth = lambda k: select_kth(fn, k)                        # histograms all-reduced (SUM) pass by pass
for q in (0.99999, 1 - 0.99999):
    got = quantile_from_order_statistics(kth, n, q)
    want = float(torch.quantile(torch.from_numpy(full), q))      # one process on the concatenated batch
    assert got == want, (r, q, got, want)
print("rank", r, "ok")
dist.destroy_process_group()
'''


def test_percentile_world2_gloo_equals_single_process(tmp_path):
    """every rank of a data-parallel percentile calibration freezes the quantile of the WHOLE batch"""
    script = tmp_path / "p.py"
    script.write_text(_PCT_WORKER % (ROOT, ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29643", str(script)], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


# ------------------------------------------------------------------------------------------------ mixed-precision search
def _vit_flops(depth=12, D=192, T=197):
    f = [196 * 768 * D]
    for _ in range(depth):
        f += [T * D * 3 * D, T * D * D, T * D * 4 * D, T * 4 * D * D]
    return f + [D * 1000]


def test_search_candidates_follow_reference_sampling_rule():
    """test_quant.py:323-341: first layer 8 bit, consecutive layers share a width pairwise, budget 1.1 x the 4-bit cost, no repeats"""
    import random
    from p2vit_b200 import search
    flops = _vit_flops()
    cands = search.sample_candidates(flops, random.Random(0))
    assert len(cands) == 51 and len({tuple(c) for c in cands}) == 51
    lim = 1.1 * sum(f * 4 for f in flops)
    for c in cands:
        assert len(c) == 50 and c[0] == 8 and set(c) <= {4, 8}
        assert all(c[1 + 2 * k] == c[2 + 2 * k] for k in range(24))
        assert search.model_cost(flops, c) <= lim


def test_search_ranking_and_distance_columns():
    from p2vit_b200 import search
    gd = [[9.0, 8.0, 4.0, 1.0]] * 3                       # uint3, uint4, int4, int8 per layer
    assert search.distance_columns(gd) == [[4.0, 1.0]] * 3                                   # fixed mapping: int4 / int8 entries
    assert search.distance_columns(gd, replicate_reference=True) == [[9.0, 8.0]] * 3         # test_quant.py:351-354 as written
    ranked = search.rank_by_sensitivity([[8, 4, 4, 8], [8, 8, 8, 4], [8, 4, 4, 4]], search.distance_columns(gd), [1.0, 2.0, 0.5])
    assert [r[0] for r in ranked] == [[8, 8, 8, 4], [8, 4, 4, 8], [8, 4, 4, 4]]
    assert ranked[0][1] == 1.0 * 1 + 2.0 * 1 + 0.5 * 4 and ranked[2][1] == 4 + 8 + 2


def test_evolutionary_search_improves_and_respects_budget():
    """toy objective: accuracy = weighted count of 8-bit layers; the population's best must not get worse, every survivor fits
    the budget, and no configuration is evaluated twice"""
    import random
    from p2vit_b200 import search
    flops = _vit_flops()
    w = [((7 * i) % 11 + 1) / 10.0 for i in range(50)]
    calls = []

    def evaluate(cfg):
        calls.append(tuple(cfg))
        return sum(wi for wi, b in zip(w, cfg) if b == 8)

    rng = random.Random(1)
    seeds = search.sample_candidates(flops, rng)
    first_best = max(evaluate(c) for c in seeds[:25])
    calls.clear()
    popu, n_eval = search.evolutionary_search(evaluate, seeds, flops, rng, evo_iter=4)
    assert len(popu) == 25 and popu[0][1] >= first_best
    assert [s for _, s in popu] == sorted((s for _, s in popu), reverse=True)
    lim = search.budget(flops)
    assert all(search.model_cost(flops, c) <= lim for c, _ in popu)
    assert len(calls) == len(set(calls)) == n_eval


# ------------------------------------------------------------------------------------------------ real-data plumbing
def test_build_transform_and_sharded_imagefolder(tmp_path):
    """test_quant.py:112-157, 565-597 on a tiny synthetic ImageFolder tree: transform geometry / normalisation per model family,
    ordered validation batches, disjoint per-rank shards that cover the set"""
    import numpy as np
    from PIL import Image
    from p2vit_b200 import data
    rng = np.random.RandomState(0)
    for split, per_class in (("train", 3), ("val", 5)):
        for c in ("n01", "n02"):
            d = tmp_path / split / c
            d.mkdir(parents=True)
            for i in range(per_class):
                Image.fromarray(rng.randint(0, 256, (300 + 10 * i, 280, 3), dtype=np.uint8)).save(d / ("%d.png" % i))
    assert data.preprocess_for("deit_small")["crop_pct"] == 0.875 and data.preprocess_for("vit_base")["mean"] == (0.5, 0.5, 0.5)
    assert data.preprocess_for("swin_tiny")["crop_pct"] == 0.9
    with pytest.raises(NotImplementedError):
        data.preprocess_for("resnet50")
    tf = data.build_transform(**data.preprocess_for("vit_base"))
    assert [type(t).__name__ for t in tf.transforms] == ["Resize", "CenterCrop", "ToTensor", "Normalize"]
    assert tf.transforms[0].size == 248            # floor(224 / 0.9)
    x = tf(Image.fromarray(np.full((300, 280, 3), 255, dtype=np.uint8)))
    assert x.shape == (3, 224, 224) and torch.allclose(x, torch.ones_like(x))      # (1 - 0.5) / 0.5
    train, val = data.build_loaders(str(tmp_path), "deit_tiny", calib_batchsize=4, val_batchsize=4, num_workers=0)
    xb, yb = next(iter(train))
    assert xb.shape == (4, 3, 224, 224) and len(train) == 1          # 6 images, drop_last
    labels = torch.cat([y for _, y in val])
    assert labels.tolist() == [0] * 5 + [1] * 5
    seen = []
    for r in range(3):
        _, v = data.build_loaders(str(tmp_path), "deit_tiny", 4, 4, 0, rank=r, world_size=3)
        seen += torch.cat([y for _, y in v]).tolist()
    assert seen == labels.tolist()
