"""GPU: the quantized model (integer engine and module-by-module path) against the golden vectors of the
unmodified reference (tests/golden, calibrated state loaded) and against the CPU oracle on fresh inputs."""
import numpy as np
import pytest
import torch

from oracle.port import VitOracle
from p2vit_b200 import Config, build_model, synth
from p2vit_b200.engine import VitEngine

pytestmark = pytest.mark.gpu


def _state(g):
    return {k[6:]: g[k] for k in g.files if k.startswith("state/")}


def _model(name, g):
    m = build_model(name, Config(), seed=int(g["meta.seed"]), device="cuda")
    m.load_quant_state(_state(g))
    m.model_quant()
    return m


def _tap_codes(golden_tap, scale):
    return np.round(golden_tap / scale.reshape(1, 1, -1)).astype(np.int64)


def test_micro_engine_per_op_codes_match_reference(golden):
    g = golden("vit_micro_minmax")
    m = _model("vit_micro", g)
    st = _state(g)
    x = synth.synth_images(int(g["meta.eval"]), seed=1).cuda()
    bits = [8] * (4 * m.depth + 2)
    eng = VitEngine(m, use_graph=False)
    taps = {}
    logits = eng(x, bits, taps=taps).cpu().numpy()
    B, T1, D = x.shape[0], 197, m.embed_dim
    report = []
    # residual stream after the stem
    got = taps["cls"].cpu().numpy().reshape(B, T1, D).astype(np.int64)
    report.append(("qact1", (got != _tap_codes(g["tap8/qact1"], st["qact1.scale"])).sum()))
    for i in range(m.depth):
        p = "blocks.%d." % i
        for step, tap, sc in ((p + "norm1", p + "attn.qact0", st[p + "attn.qact0.scale"]),
                              (p + "attn.qact1", p + "attn.qact1", st[p + "attn.qact1.scale"]),
                              (p + "attn.qact2", p + "attn.qact2", st[p + "attn.qact2.scale"]),
                              (p + "qact2", p + "qact2", st[p + "qact2.scale"]),
                              (p + "norm2", p + "mlp.qact0", st[p + "mlp.qact0.scale"]),
                              (p + "mlp.qact1", p + "mlp.qact1", st[p + "mlp.qact1.scale"]),
                              (p + "qact4", p + "qact4", st[p + "qact4.scale"])):
            ref = _tap_codes(g["tap8/" + tap], sc)
            got = taps[step].cpu().numpy().reshape(ref.shape).astype(np.int64)
            report.append((step, int((got != ref).sum())))
    bad = [(k, v) for k, v in report if v]
    assert not bad, "code mismatches vs reference taps: %s" % bad
    assert np.array_equal(logits, g["logits8"])


@pytest.mark.parametrize("wbits", [8, 4])
def test_engine_logits_match_reference_golden(golden, wbits):
    g = golden("vit_micro_minmax")
    m = _model("vit_micro", g)
    x = synth.synth_images(int(g["meta.eval"]), seed=1).cuda()
    bits = [wbits] * (4 * m.depth + 2)
    logits, flops, gd = m(x, bits)
    ref = g["logits%d" % wbits]
    got = logits.cpu().numpy()
    assert np.array_equal(got.argmax(1), ref.argmax(1)), "top-1 differs"
    assert np.array_equal(got, ref), "logit codes differ in %d of %d entries" % ((got != ref).sum(), ref.size)
    assert len(flops) == 4 * m.depth + 2


def _teacher_forced(name, g, wbits, B):
    """golden-state wrapper of _teacher_forced_run"""
    st = _state(g)
    c = synth.VIT_CONFIGS[name]
    o = VitOracle(synth.synth_vit_state_dict(**c, seed=0), **c, exact_sums=True)
    o.load_state(st)
    return _teacher_forced_run(_model(name, g), o, st, c, synth.synth_images(B, seed=1), [wbits] * (4 * c["depth"] + 2))


def _teacher_forced_run(m, o, st, c, x, bits):
    """Runs every engine step on the ORACLE's input codes for that step and compares with the oracle's output codes, so a
    rounding-tie flip in one op cannot cascade into the next comparison.  Returns {step: (mismatches, max |delta|, numel)}."""
    B = x.shape[0]
    taps = {}
    ref_logits = o.forward_quant(x, bits, taps)
    eng = VitEngine(m, use_graph=False)
    prog = eng._program(tuple(bits), B)
    ws = prog["ws"]
    ws["img"].copy_(x.cuda())
    T1, D = 197, c["embed_dim"]

    def codes(tap, scale_key):
        shape = (1, 1, -1) if taps[tap].dim() == 3 else (1, -1)
        s = torch.as_tensor(st[scale_key]).reshape(shape)
        zp = torch.as_tensor(st[scale_key[:-len("scale")] + "zero_point"]).reshape(shape)      # non-zero for asymmetric observers only
        return (torch.round(taps[tap] / s) + zp).to(torch.int8)

    def put(buf, t):
        ws[buf].copy_(t.reshape(ws[buf].shape).cuda())

    report = {}

    def check(step, buf, ref):
        got = ws[buf].cpu().reshape(ref.shape).to(torch.int64)
        d = (got - ref.to(torch.int64)).abs()
        report[step] = (int((d != 0).sum()), int(d.max()), d.numel())

    steps = dict(prog["steps"])
    steps["patchify"]()
    steps["embed"]()
    steps["cls"]()
    check("stem", "ra", codes("qact1", "qact1.scale"))
    r_tap, r_key = "qact1", "qact1.scale"
    for i in range(c["depth"]):
        p = "blocks.%d." % i
        put("ra", codes(r_tap, r_key))
        steps[p + "norm1"]()
        check(p + "norm1", "ln", codes(p + "attn.qact0", p + "attn.qact0.scale"))
        put("ln", codes(p + "attn.qact0", p + "attn.qact0.scale"))
        steps[p + "attn.qact1"]()
        check(p + "qkv", "qkv", codes(p + "attn.qact1", p + "attn.qact1.scale"))
        put("qkv", codes(p + "attn.qact1", p + "attn.qact1.scale"))
        steps[p + "attn.qact2"]()
        check(p + "attention", "ao", codes(p + "attn.qact2", p + "attn.qact2.scale"))
        put("ao", codes(p + "attn.qact2", p + "attn.qact2.scale"))
        steps[p + "qact2"]()
        check(p + "proj+res", "rb", codes(p + "qact2", p + "qact2.scale"))
        put("rb", codes(p + "qact2", p + "qact2.scale"))
        steps[p + "norm2"]()
        check(p + "norm2", "ln", codes(p + "mlp.qact0", p + "mlp.qact0.scale"))
        put("ln", codes(p + "mlp.qact0", p + "mlp.qact0.scale"))
        steps[p + "mlp.qact1"]()
        check(p + "fc1+gelu", "hid", codes(p + "mlp.qact1", p + "mlp.qact1.scale"))
        put("hid", codes(p + "mlp.qact1", p + "mlp.qact1.scale"))
        put("rb", codes(p + "qact2", p + "qact2.scale"))
        steps[p + "qact4"]()
        check(p + "fc2+res", "ra", codes(p + "qact4", p + "qact4.scale"))
        r_tap, r_key = p + "qact4", p + "qact4.scale"
    put("ra", codes(r_tap, r_key))
    steps["qact2"]()
    check("final_norm", "cls", codes("qact2", "qact2.scale"))
    put("cls", codes("qact2", "qact2.scale"))
    steps["act_out"]()
    lsb = float(torch.as_tensor(st["act_out.scale"]).reshape(-1)[0])
    d = ((ws["logits"].cpu() - ref_logits) / lsb).round().abs()
    report["head"] = (int((d != 0).sum()), int(d.max()), d.numel())
    return report


@pytest.mark.parametrize("name,wbits,B", [("deit_tiny", 8, 8), ("deit_tiny", 4, 4), ("deit_small", 8, 4)])
def test_per_op_parity_at_model_scale(golden, name, wbits, B):
    """DeiT-Tiny / DeiT-Small with the reference-calibrated state: every fused op, fed the oracle's codes, reproduces the
    oracle's codes bit for bit; only the erf-GELU epilogue may differ, by one LSB at rounding ties."""
    rep = _teacher_forced(name, golden(name + "_minmax"), wbits, B)
    bad = {k: v for k, v in rep.items() if v[0] and "gelu" not in k}
    assert not bad, "non-GELU steps differ from the oracle: %s" % bad
    gelu = [v for k, v in rep.items() if "gelu" in k]
    tot, n = sum(v[0] for v in gelu), sum(v[2] for v in gelu)
    assert all(v[1] <= 1 for v in gelu) and tot / n < 1e-5, "GELU epilogue: %d of %d codes differ" % (tot, n)


def _engine_vs_cuda_oracle(m, o, st, x, bits, force_stem=False):
    """free-running comparison (no teacher forcing): every engine step's int8 buffer against the oracle's tap of the same
    run, then the logits.  Returns ({step: mismatches}, engine logits, oracle logits) - all on the CPU.  force_stem: compare
    the stem separately ("stem": (bad, max, n)) and continue from the oracle's stem codes."""
    taps, ref_taps = {}, {}
    ref = o.forward_quant(x, bits, ref_taps).cpu()

    def ref_codes(tap):
        r = ref_taps[tap]
        shape = [1] * (r.dim() - 1) + [-1]
        sc = torch.as_tensor(st[tap + ".scale"]).to(r.device).reshape(shape)
        zp = torch.as_tensor(st[tap + ".zero_point"]).to(r.device).reshape(shape)
        return (torch.round(r / sc) + zp).to(torch.int64)

    eng = VitEngine(m, use_graph=False)
    rep = {}
    if force_stem:
        prog = eng._program(tuple(bits), x.shape[0])
        ws = prog["ws"]
        ws["img"].copy_(x.cuda())
        steps = prog["steps"]
        for name, fn in steps[:3]:
            fn()
        rc = ref_codes("qact1")
        d = (ws["ra"].reshape(rc.shape).to(torch.int64) - rc).abs()
        rep["stem"] = (int((d != 0).sum()), int(d.max()), d.numel())
        ws["ra"].copy_(rc.reshape(ws["ra"].shape).to(torch.int8))
        for name, fn in steps[3:]:
            fn()
            taps[name] = ws[prog["outs"][name]].clone()
        taps["cls"] = rc.to(torch.int8)
        got = ws["logits"].cpu()
    else:
        got = eng(x.cuda(), bits, taps=taps).cpu()
    pairs = [("cls", "qact1")]
    for i in range(m.depth):
        p = "blocks.%d." % i
        pairs += [(p + "norm1", p + "attn.qact0"), (p + "attn.qact1", p + "attn.qact1"), (p + "attn.qact2", p + "attn.qact2"),
                  (p + "qact2", p + "qact2"), (p + "norm2", p + "mlp.qact0"), (p + "mlp.qact1", p + "mlp.qact1"), (p + "qact4", p + "qact4")]
    pairs.append(("qact2", "qact2"))
    for step, tap in pairs:
        rc = ref_codes(tap)
        rep[step] = int((_as_codes(taps[step], st, tap).reshape(rc.shape) != rc).sum())
    return rep, got, ref


def _as_codes(buf, st, tap):
    """engine buffer -> integer codes on the quantizer's grid (int8 [-128,127] for every activation bit type on this path; an
    asymmetric observer only adds a zero point inside that range, observer/omse.py:46-47)"""
    return buf.to(torch.int64)


def _golden_bits(g, kind, depth):
    return [int(b) for b in g["bits_mixed"]] if kind == "mixed" else [int(kind)] * (4 * depth + 2)


STRICT_CASES = [("deit_tiny_minmax", 8, 8), ("deit_small_minmax", 8, 8), ("deit_small_minmax", 8, 4), ("deit_tiny_minmax", 32, 8),
                ("vit_base_percentile", 4, 8), ("vit_base_ema", 4, 8), ("vit_base_ema", 2, 4), ("vit_base_omse", 4, 8),
                ("vit_large_minmax", 2, "mixed"), ("vit_large_minmax", 2, 8)]


@pytest.mark.parametrize("gname,B,wbits", STRICT_CASES)
def test_model_scale_strict_parity_vs_cuda_oracle(golden, gname, B, wbits):
    """The strict model-scale bar (north_star: codes bit-exact, top-1 identical) on every BASELINE config that is a ViT: DeiT-Tiny
    (C1), DeiT-Small (C2), ViT-Base with the percentile / ema / omse activation observers (C3: raw fp32 scales, omse with zero
    points) and ViT-Large with the 98-entry mixed {4,8} bit_config (C5), each with the state the UNMODIFIED reference calibrated.
    Free running over the whole network against the oracle evaluated by torch's CUDA backend (same erff and IEEE sqrt
    as the kernels, so no backend-dependent rounding tie is left): EVERY step's codes, EVERY logit of EVERY image and top-1 must
    be identical.  ViT-Large only: its stem multiplies fp32 pixels (no input quantizer, vit_fquant.py:1063) - an fp32 dot product
    of 768 terms whose rounding depends on the summation order in the reference too - so the stem is compared on its own
    (<= 1 LSB, rare) and the oracle's stem codes are then fed to the engine for the strict comparison of everything after it."""
    g = golden(gname)
    name = str(g["meta.model"])
    st = _state(g)
    c = synth.VIT_CONFIGS[name]
    o = VitOracle(synth.synth_vit_state_dict(**c, seed=0), **c, method=str(g["meta.method"]), exact_sums=True, device="cuda")
    o.load_state(st)
    m = build_model(name, Config(True, True, str(g["meta.method"])), seed=int(g["meta.seed"]), device="cuda")
    m.load_quant_state(st)
    m.model_quant()
    x = synth.synth_images(B, seed=1)
    bits = _golden_bits(g, wbits, c["depth"])
    rep, got, ref = _engine_vs_cuda_oracle(m, o, st, x, bits, force_stem=not c["input_quant"])
    if not c["input_quant"]:
        bad, mx, n = rep.pop("stem")
        assert mx <= 1 and bad / n < 2e-3, "fp32 stem: %d of %d codes differ, max %d LSB" % (bad, n, mx)
    bad = {k: v for k, v in rep.items() if v}
    assert not bad, "engine steps differ from the CUDA-evaluated oracle: %s" % bad
    assert torch.equal(got, ref), "%d logits differ" % int((got != ref).sum())
    assert torch.equal(got.argmax(1), ref.argmax(1))
    if c["input_quant"]:
        assert torch.equal(m(x.cuda(), bits)[0].cpu(), ref), "graph path differs"
    assert len(set(ref.argmax(1).tolist())) > 1, "top-1 must depend on the image for the check to mean anything"


@pytest.mark.parametrize("name", ["deit_tiny", "deit_small"])
def test_end_to_end_vs_reference_golden_logits(golden, name):
    """Second, documented bound: the reference as run on a CPU (tests/golden/<name>_minmax.npz = its logits in the build
    container).  The same algorithm evaluated by another backend leaves that run's trajectory in exactly one kind of place:
    erf (Sleef on the CPU, CUDA erff in the kernels and in torch-CUDA; they disagree in the last bit on a third of all
    arguments, tools/diag_backend.py), which matters when GELU's output lands on a rounding tie of mlp.qact1 - about once per
    million GELU outputs, i.e. in most DeiT-S images (3.6 M GELU outputs each).  These random-weight networks amplify one flipped
    code to the whole image, so an image either reproduces the CPU logits bit for bit or diverges completely.  Asserted here, per
    image, against the CPU oracle's free-running taps (canonical variant; tests/test_oracle_golden.py ties it to the golden
    logits): the FIRST step at which the kernels leave the CPU trajectory is a GELU step, there by one LSB in a handful of
    codes - nothing else ever differs first - and an image without such a flip reproduces the reference's golden logits
    whenever the CPU oracle does.  Kernels == reference algorithm on the CUDA backend, every image, every code:
    test_model_scale_strict_parity_vs_cuda_oracle."""
    g = golden(name + "_minmax")
    st = _state(g)
    m = _model(name, g)
    c = synth.VIT_CONFIGS[name]
    B = int(g["meta.eval"])
    x = synth.synth_images(B, seed=1)
    bits = [8] * (4 * c["depth"] + 2)
    o = VitOracle(synth.synth_vit_state_dict(**c, seed=0), **c, exact_sums=True)
    o.load_state(st)
    ref_taps, taps = {}, {}
    ref = o.forward_quant(x, bits, ref_taps)
    got = VitEngine(m, use_graph=False)(x.cuda(), bits, taps=taps).cpu()
    order = [("cls", "qact1")]
    for i in range(m.depth):
        p = "blocks.%d." % i
        order += [(p + "norm1", p + "attn.qact0"), (p + "attn.qact1", p + "attn.qact1"), (p + "attn.qact2", p + "attn.qact2"),
                  (p + "qact2", p + "qact2"), (p + "norm2", p + "mlp.qact0"), (p + "mlp.qact1", p + "mlp.qact1"), (p + "qact4", p + "qact4")]
    first = {}
    for step, tap in order:
        r = ref_taps[tap]
        rc = torch.round(r / torch.as_tensor(st[tap + ".scale"]).reshape(1, 1, -1)).to(torch.int64)
        d = (taps[step].cpu().reshape(rc.shape).to(torch.int64) - rc).abs().reshape(B, -1)
        for b in range(B):
            if b not in first and int(d[b].max()) > 0:
                first[b] = (step, int(d[b].max()), int((d[b] > 0).sum()))
    same = (got == ref).all(dim=1)
    assert {b for b in range(B) if not bool(same[b])} <= set(first), (first, same)      # (a late flip may die out before the logits)
    for b, (step, mx, n) in first.items():
        assert step.endswith("mlp.qact1") and mx == 1 and n <= 8, "image %d leaves the CPU trajectory at %s (%d codes, max %d LSB)" % (b, step, n, mx)
    gold = torch.from_numpy(g["logits8"])
    cpu_is_gold = (ref == gold).all(dim=1)
    for b in range(B):
        if bool(same[b]) and bool(cpu_is_gold[b]):
            assert torch.equal(got[b], gold[b])
    print("%s: %d of %d images bit-identical to the reference's CPU logits; GELU-tie divergences: %s" %
          (name, int((got == gold).all(dim=1).sum()), B, first))


def test_graph_replay_and_simt_cross_check(golden):
    g = golden("vit_micro_minmax")
    m = _model("vit_micro", g)
    bits = [8] * (4 * m.depth + 2)
    x = synth.synth_images(3, seed=5).cuda()
    a = VitEngine(m, use_graph=True)(x, bits)
    a2 = m(x, bits)[0]           # graph replay through the model's own engine (second call replays)
    a3 = m(x, bits)[0]
    b = VitEngine(m, use_graph=False, simt_gemm=True)(x, bits)
    assert torch.equal(a, a2) and torch.equal(a, a3) and torch.equal(a, b)


def test_batch_split_invariance(golden):
    """sharding the batch (the data-parallel partition) cannot change any logit."""
    g = golden("vit_micro_minmax")
    m = _model("vit_micro", g)
    bits = [8] * (4 * m.depth + 2)
    x = synth.synth_images(7, seed=11).cuda()
    full = m(x, bits)[0]
    parts = torch.cat([m(x[:3].contiguous(), bits)[0], m(x[3:].contiguous(), bits)[0]])
    assert torch.equal(full, parts)


def test_engine_vs_oracle_fresh_inputs(golden):
    """new images, W8A8 and mixed bit_config, against the exact-sum oracle."""
    g = golden("vit_micro_minmax")
    m = _model("vit_micro", g)
    c = synth.VIT_CONFIGS["vit_micro"]
    o = VitOracle(synth.synth_vit_state_dict(**c, seed=0), **c, exact_sums=True)
    o.load_state(_state(g))
    x = synth.synth_images(4, seed=123)
    for bits in ([8] * 10, [8, 4, 4, 8, 8, 8, 8, 4, 4, 8]):
        ref = o.forward_quant(x, bits)
        got = m(x.cuda(), bits)[0].cpu()
        assert torch.equal(got, ref), "bits=%s: %d mismatches" % (bits, int((got != ref).sum()))


def test_eager_modules_match_engine(golden):
    g = golden("vit_micro_minmax")
    m = _model("vit_micro", g)
    bits = [8] * (4 * m.depth + 2)
    x = synth.synth_images(2, seed=1).cuda()
    eager = m.forward_eager(x, bits)[0]
    assert np.array_equal(eager.cpu().numpy(), g["logits8"])
    assert torch.equal(eager, m(x, bits)[0])


def test_quantized_forward_requires_bit_config_and_cuda(golden):
    g = golden("vit_micro_minmax")
    m = _model("vit_micro", g)
    with pytest.raises(ValueError):
        m(synth.synth_images(1).cuda())
    with pytest.raises(RuntimeError):
        m(synth.synth_images(1), [8] * 10)
    with pytest.raises(ValueError):
        m(synth.synth_images(1).cuda(), [8] * 9)


def test_calibration_matches_reference_state(golden):
    """the B200 calibration (block-reduce observers, batched weight search) freezes the same PoT exponents as the
    reference; raw fp32 scales (PTF base) agree to fp32 rounding of the GPU's own FP forward."""
    from p2vit_b200 import calibrate_model

    g = golden("vit_micro_minmax")
    m = build_model("vit_micro", Config(), seed=0, device="cuda")
    calibrate_model(m, synth.synth_images(int(g["meta.calib"]), seed=0).cuda())
    st, ref = m.export_quant_state(), _state(g)
    assert set(st) == set(ref)
    exp_bad, ptf_bad = [], []
    for k, v in ref.items():
        a = st[k].numpy().astype(np.float64)
        v = v.astype(np.float64)
        if "zero_point" in k:
            assert np.array_equal(a, v), k
        elif k.endswith(".scale") and v.size > 1:   # PTF: fp32 base scale x {1,2,4,8}
            if not (np.array_equal(np.round(np.log2(a / a.min())), np.round(np.log2(v / v.min()))) and abs(a.min() / v.min() - 1) < 1e-5):
                ptf_bad.append(k)
        elif not np.array_equal(a, v):
            exp_bad.append(k)
    assert not exp_bad, "PoT scales differ: %s" % exp_bad[:10]
    assert not ptf_bad, "PTF scales differ: %s" % ptf_bad[:10]


# ------------------------------------------------------------------------------------------------ other observers / configs
def _calibrated_pair(method, input_quant=True, calib=8, seed=0):
    """vit_micro calibrated on the GPU with `method`; the CPU oracle loaded with the very same frozen state"""
    from functools import partial

    from p2vit_b200 import calibrate_model
    from p2vit_b200.ptq import QIntLayerNorm
    from p2vit_b200.vit import VisionTransformer

    c = dict(synth.VIT_CONFIGS["vit_micro"], input_quant=input_quant)
    cfg = Config(True, True, method)
    m = VisionTransformer(patch_size=16, embed_dim=c["embed_dim"], depth=c["depth"], num_heads=c["num_heads"], mlp_ratio=4, qkv_bias=True,
                          norm_layer=partial(QIntLayerNorm, eps=1e-6), input_quant=input_quant, cfg=cfg)
    m.load_state_dict(synth.synth_vit_state_dict(**c, seed=seed), strict=False)
    m = m.cuda().eval()
    calibrate_model(m, synth.synth_images(calib, seed=0).cuda())
    o = VitOracle(synth.synth_vit_state_dict(**c, seed=seed), **c, method=method, exact_sums=True)
    o.load_state(m.export_quant_state())
    return m, o


def _assert_codes_close(got, ref, lsb, what, max_rate):
    d = ((got - ref) / lsb).round().abs()
    assert float(d.max()) <= 1.0, "%s: max code difference %g" % (what, float(d.max()))
    rate = float((d > 0).float().mean())
    assert rate <= max_rate, "%s: %.4f of the codes differ" % (what, rate)


def _assert_per_op(rep, what, max_rate):
    """non power-of-two activation scales / fp32 stems: the reference accumulates dequantized fp32 products (order dependent),
    the kernels accumulate integers exactly, so isolated codes may sit on the other side of a rounding tie: per op, fed the
    oracle's own inputs, |diff| <= 1 LSB and rare."""
    for step, (bad, mx, n) in rep.items():
        assert mx <= 1, "%s %s: max code difference %d" % (what, step, mx)
        assert bad / n <= max_rate, "%s %s: %d of %d codes differ" % (what, step, bad, n)


@pytest.mark.parametrize("method", ["ema", "percentile"])
def test_engine_non_pot_observers_vs_oracle(method):
    """OBSERVER_A = ema / percentile (observer/ema.py, percentile.py): raw fp32 activation scales -> the GEMM / LayerNorm /
    attention epilogues run their general (non power-of-two) variants; per-op parity with teacher forcing"""
    m, o = _calibrated_pair(method)
    bits = [8] * (4 * m.depth + 2)
    x = synth.synth_images(4, seed=5)
    c = synth.VIT_CONFIGS["vit_micro"]
    rep = _teacher_forced_run(m, o, {k: v.numpy() for k, v in m.export_quant_state().items()}, c, x, bits)
    bad = {k: v for k, v in rep.items() if v[0] and "gelu" not in k}
    assert not bad, "steps differ from the canonical (exact-accumulation) oracle: %s" % bad
    _assert_per_op({k: v for k, v in rep.items() if "gelu" in k}, method, 1e-4)
    got = m(x.cuda(), bits)[0]
    assert torch.equal(m.forward_eager(x.cuda(), bits)[0], got), "module-by-module path differs from the engine"


def test_omse_engine_vs_oracle():
    """OBSERVER_A = omse (observer/omse.py:30-57; crashes in the reference as shipped): asymmetric activations with
    integer zero points.  The integer engine carries them: zp_corr columns in the GEMMs, zero points in the QAct epilogues, the
    LayerNorm output and the attention core (constant-tile MMAs).  Per-op parity with teacher forcing against the canonical
    oracle, then engine == module-by-module path."""
    m, o = _calibrated_pair("omse")
    bits = [8] * (4 * m.depth + 2)
    x = synth.synth_images(3, seed=6)
    st = {k: v.numpy() for k, v in m.export_quant_state().items()}
    assert sum(int(np.abs(v).max()) != 0 for k, v in st.items() if k.endswith(".zero_point")) >= 10, "omse should be asymmetric"
    c = synth.VIT_CONFIGS["vit_micro"]
    rep = _teacher_forced_run(m, o, st, c, x, bits)
    bad = {k: v for k, v in rep.items() if v[0] and "gelu" not in k}
    assert not bad, "steps differ from the canonical (exact-accumulation) oracle: %s" % bad
    _assert_per_op({k: v for k, v in rep.items() if "gelu" in k}, "omse", 1e-4)
    got = m(x.cuda(), bits)[0]
    assert torch.equal(m.forward_eager(x.cuda(), bits)[0], got), "module-by-module path differs from the engine"
    assert torch.equal(VitEngine(m, use_graph=False, simt_gemm=True)(x.cuda(), bits), got), "dp4a cross-check kernels differ"


@pytest.mark.parametrize("method", ["omse", "minmax"])
def test_engine_reproduces_reference_logits_vit_micro(golden, method):
    """vit_micro with the state the unmodified reference calibrated (minmax: powers of two; omse: raw fp32 scales + zero points):
    the engine's logits equal the reference's own, bit for bit (minmax: W8A8 and W4A8; omse: W8A8 - with raw fp32 activation
    scales AND per-row int4 weight scales the reference's fp32 GEMM rounds order-dependently, so only the canonical oracle is a
    bit-exact target there: test_omse_engine_vs_oracle, test_model_scale_strict_parity_vs_cuda_oracle[vit_base_omse])"""
    g = golden("vit_micro_" + method)
    m = build_model("vit_micro", Config(True, True, method), seed=0, device="cuda")
    m.load_quant_state(_state(g))
    m.model_quant()
    x = synth.synth_images(int(g["meta.eval"]), seed=1).cuda()
    for wb in ((8, 4) if method == "minmax" else (8,)):
        got = m(x, [wb] * 10)[0].cpu().numpy()
        assert np.array_equal(got, g["logits%d" % wb]), "W%dA8: %d logits differ" % (wb, int((got != g["logits%d" % wb]).sum()))


def test_engine_falls_back_to_eager_when_not_applicable(golden):
    """Config(ptf=False): no integer LayerNorm -> the integer program does not apply; model(x, bits) then evaluates module by
    module like the reference does through the same call (one warning), instead of raising"""
    import warnings
    from p2vit_b200 import calibrate_model
    m = build_model("vit_micro", Config(False, True, "minmax"), seed=0, device="cuda")
    calibrate_model(m, synth.synth_images(4, seed=0).cuda())
    x = synth.synth_images(2, seed=3).cuda()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        a = m(x, [8] * 10)[0]
        b = m(x, [8] * 10)[0]
    assert len([i for i in w if "not applicable" in str(i.message)]) == 1
    assert torch.equal(a, b) and torch.equal(a, m.forward_eager(x, [8] * 10)[0])


def test_engine_caches_are_bounded(golden):
    """a search visits hundreds of bit_configs - plans / programs / graphs are LRU-bounded, the int8 workspace is
    shared per batch size, and evicted configurations are rebuilt with identical results"""
    g = golden("vit_micro_minmax")
    m = _model("vit_micro", g)
    x = synth.synth_images(2, seed=4).cuda()
    import itertools
    cfgs = [[8] + list(c) + [8] for c in itertools.product([4, 8], repeat=8)][:40]
    first = [m(x, c)[0].clone() for c in cfgs]
    eng = m._engine
    assert len(eng.plans) <= eng.MAX_PLANS and len(eng.programs) <= eng.MAX_PROGRAMS and len(eng.graphs) <= 2 * eng.MAX_PROGRAMS
    assert len(eng.workspaces) == 1
    for c, want in zip(cfgs[:5], first[:5]):
        assert torch.equal(m(x, c)[0], want)
    eng.clear()
    assert not eng.plans and not eng.graphs
    assert torch.equal(m(x, cfgs[0])[0], first[0])


def test_input_quant_false_engine_vs_oracle():
    """ViT-L style stem (vit_fquant.py:1063): fp32 pixels into the quantized-weight patch embedding (an fp32 GEMM in the
    reference too), int8 from patch_embed.qact on"""
    m, o = _calibrated_pair("minmax", input_quant=False)
    bits = [8] * (4 * m.depth + 2)
    x = synth.synth_images(3, seed=7)
    c = dict(synth.VIT_CONFIGS["vit_micro"], input_quant=False)
    rep = _teacher_forced_run(m, o, {k: v.numpy() for k, v in m.export_quant_state().items()}, c, x, bits)
    stem = rep.pop("stem")
    assert stem[1] <= 1 and stem[0] / stem[2] < 2e-3, "fp32 stem: %s" % (stem,)
    bad = {k: v for k, v in rep.items() if v[0] and "gelu" not in k}
    assert not bad, "integer steps differ from the oracle: %s" % bad
    got = m(x.cuda(), bits)[0]
    assert torch.equal(m.forward_eager(x.cuda(), bits)[0], got)


def test_mixed_precision_search_flow():
    """test_quant.py:316-463 end to end on the GPU: calibrate vit_micro, take FLOPs / global_distance from the calibration
    forward, rank sampled {4,8} configurations and run a short evolutionary search whose fitness is top-1 agreement with the
    all-8-bit model (synthetic labels).  Checks the plumbing: per-layer bit_config switching through the engine, budget,
    ordering, memoised evaluations."""
    from p2vit_b200 import runner, search
    m = build_model("vit_micro", Config(), seed=0, device="cuda")
    calib = synth.synth_images(16, seed=3).cuda()
    (_, flops, gdist), _ = runner.calibrate_model(m, calib)
    n = 4 * m.depth + 2
    assert len(flops) == n and len(gdist) == n - 1
    imgs = [synth.synth_images(16, seed=10 + i).cuda() for i in range(2)]
    batches = [(x, m(x, [8] * n)[0].argmax(1)) for x in imgs]          # labels = the W8A8 model's own top-1
    assert runner.validate(m, batches, [8] * n)[1] == 100.0
    res = search.mixed_precision_search(m, flops, gdist, batches, seed=0, top_validate=2, ratio=1.6, pop_size=6, evo_iter=2,
                                        mutate_size=3, crossover_size=3)
    lim = search.budget(res["flops"], 1.6)
    assert len(res["ranked"]) >= 6 and [o for _, o in res["ranked"]] == sorted(o for _, o in res["ranked"])
    assert len(res["validated"]) == 2 and all(0.0 <= a <= 100.0 for _, _, a in res["validated"])
    popu = res["population"]
    assert len(popu) == 6 and [s for _, s in popu] == sorted((s for _, s in popu), reverse=True)
    for cfg, acc in popu:
        assert len(cfg) == n and set(cfg) <= {4, 8} and search.model_cost(res["flops"], cfg) <= lim
        assert runner.validate(m, batches, cfg)[1] == acc            # deterministic: the memoised score is reproducible


def test_calibration_forward_is_batch_split_invariant():
    """The FP forward that feeds the observers (test_quant.py:275-281) gives every image the same activations whatever the
    batch around it - own fp32 GEMM with a fixed k order instead of a library GEMM whose kernel choice depends on M - so the
    statistics of a calibration sharded over N GPUs are those of one GPU (raw fp32 scales of ema / percentile / omse included)"""
    m = build_model("vit_micro", Config(), seed=0, device="cuda")
    x = synth.synth_images(7, seed=9).cuda()
    with torch.no_grad():
        full = m(x)[0]
        parts = torch.cat([m(x[:3].contiguous())[0], m(x[3:].contiguous())[0]])
    assert torch.equal(full, parts)


def test_hessian_sensitivity_feeds_the_search():
    """Hutchinson traces of the FP model on the GPU (fp32 autograd), normalised as test_quant.py:184-201, are
    the `sensitivity` vector of the mixed-precision ranking (test_quant.py:350-368)"""
    from p2vit_b200 import hessian, runner, search
    m = build_model("vit_micro", Config(), seed=0, device="cuda")
    crit = torch.nn.CrossEntropyLoss()
    torch.manual_seed(0)
    traces = []
    for i in range(2):
        x = synth.synth_images(4, seed=40 + i).cuda()
        names, tr = hessian.hessian_traces(m, crit, x, torch.tensor([1, 2, 3, 4], device="cuda"), max_iter=6)
        traces.append(tr)
    sens = hessian.mean_normalised_sensitivity(traces)
    n = 4 * m.depth + 2
    assert len(sens) == n - 1 and min(sens) >= 0.0 and max(sens) <= 1.0
    (_, flops, gdist), _ = runner.calibrate_model(m, synth.synth_images(8, seed=3).cuda())
    cands = [[8] * n, [8] + [4] * (n - 1), [8, 4, 4] + [8] * (n - 3)]
    ranked = search.rank_by_sensitivity(cands, search.distance_columns(gdist), sens)
    assert [o for _, o in ranked] == sorted(o for _, o in ranked) and ranked[0][0] == [8] * n


def test_uint8_pixels_equal_host_normalised_fp32(golden):
    """`model(x_u8)` after set_pixel_normalization == `model(Normalize(ToTensor(x_u8)))`: logits bit for bit, through the graph
    and the eager engine; without the normalisation constants the byte input is refused"""
    from p2vit_b200.data import PREPROCESS
    g = golden("deit_tiny_minmax")
    m = _model("deit_tiny", g)
    bits = [8] * (4 * m.depth + 2)
    gen = torch.Generator().manual_seed(11)
    x8 = torch.randint(0, 256, (6, 3, 224, 224), generator=gen, dtype=torch.uint8)
    x8[:, :, ::16, :] //= 3                                            # some structure across patches
    with pytest.raises(RuntimeError):
        m(x8.cuda(), bits)
    pp = PREPROCESS["deit"]
    m.set_pixel_normalization(pp["mean"], pp["std"])
    mean, std = torch.tensor(pp["mean"]).view(1, 3, 1, 1), torch.tensor(pp["std"]).view(1, 3, 1, 1)
    xf = x8.float().div(255).sub(mean).div(std)
    want = m(xf.cuda(), bits)[0].cpu()
    got = m(x8.cuda(), bits)[0].cpu()
    assert torch.equal(got, want)
    assert torch.equal(m(x8.cuda(), bits)[0].cpu(), want)                # graph replay
    eng = VitEngine(m, use_graph=False)
    assert torch.equal(eng(x8.cuda(), bits).cpu(), want)
    assert len(set(want.argmax(1).tolist())) > 1


@pytest.mark.parametrize("mode,kinds", [("1", "15"), ("0", "15"), ("2", "15")])
def test_programmatic_dependent_launch_modes_keep_the_logits(golden, mode, kinds):
    """P2VIT_PDL / P2VIT_PDL_KINDS are read once per process: a child runs DeiT-T through the graph in the always-on, the off and the
    all-families graph mode; its logits must be this process's (default mode: block GEMMs + LayerNorm inside graphs) bit for bit"""
    import hashlib
    import os
    import subprocess
    import sys
    g = golden("deit_tiny_minmax")
    m = _model("deit_tiny", g)
    x = synth.synth_images(8, seed=21).cuda()
    bits = [8] * (4 * m.depth + 2)
    m(x, bits)
    want = hashlib.sha256(m(x, bits)[0].cpu().numpy().tobytes()).hexdigest()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import hashlib, numpy as np, torch\n"
        "from p2vit_b200 import Config, build_model, synth\n"
        "g = np.load('tests/golden/deit_tiny_minmax.npz')\n"
        "m = build_model('deit_tiny', Config(), seed=int(g['meta.seed']), device='cuda')\n"
        "m.load_quant_state({k[6:]: g[k] for k in g.files if k.startswith('state/')}); m.model_quant()\n"
        "x = synth.synth_images(8, seed=21).cuda()\n"
        "bits = [8] * (4 * m.depth + 2)\n"
        "hs = {hashlib.sha256(m(x, bits)[0].cpu().numpy().tobytes()).hexdigest() for _ in range(4)}\n"
        "assert len(hs) == 1, hs\n"
        "print('sha', hs.pop())\n")
    env = dict(os.environ, P2VIT_PDL=mode, P2VIT_PDL_KINDS=kinds, PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
    assert ("sha " + want) in r.stdout, (want, r.stdout[-200:])
