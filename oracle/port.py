"""TEST INFRASTRUCTURE ONLY - CPU oracle for the fully-quantized ViT/DeiT path of P2-ViT.

A restatement (not a copy) of the reference's algorithm in plain torch-CPU fp32 ops, in
the same op order as the reference so results are bit-identical to it on the same host.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import this file; the product (`p2vit_b200/`) never does.

Pinning: `tests/golden/*.npz` were produced by running the unmodified reference
(/root/reference, through oracle/ref_loader.py) with oracle/gen_golden.py;
tests/test_oracle_golden.py checks this port against them (calibrated scales, per-op
codes and logits).  The reference itself ships no tests or golden vectors.

Reference map (file:line relative to /root/reference):
  fake_quant            models/ptq/quantizer/uniform.py:48-126, base.py:14-31
  round_ln              models/ptq/observer/minmax.py:50-64
  MinmaxObs             models/ptq/observer/minmax.py:15-237
  PtfObs                models/ptq/observer/ptf.py:13-152
  EmaObs / PercentileObs / OmseObs   observer/ema.py, percentile.py, omse.py
  int_layernorm         models/ptq/layers.py:270-337
  int_softmax_log2      models/ptq/layers.py:376-428
  VitOracle.calibrate / .forward_quant
                        models/vit_fquant.py:177-407,489-596,830-939,
                        models/layers_quant.py:225-393,462-497, test_quant.py:262-312

`exact_sums=True` is the canonical, backend-independent variant the CUDA kernels implement: the two
order-dependent fp32 row reductions of the reference (sum of squares in LayerNorm, layers.py:316-318;
sum of exp_int in softmax, :416) become exactly-rounded sums (fp64/int accumulation, one final rounding
to fp32), LayerNorm's sqrt is the IEEE correctly-rounded one (torch's CPU sqrt is MKL VML, off by one
ulp on ~0.1% of inputs; torch-CUDA's is IEEE), and the matmuls accumulate the integer codes exactly before
the fp32 scale / bias operations (identical to the reference's fp32 matmul for power-of-two scales, where
every partial sum is exact; order independent for the raw fp32 scales of ema / percentile).  Batch-split
invariant; see DESIGN.md "tie adjudication".
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BITS = {  # models/ptq/bit_type.py:17-57 : name -> (lower_bound, upper_bound, signed)
    "uint3": (0, 7, False),
    "uint4": (0, 15, False),
    "int4": (-8, 7, True),
    "int8": (-128, 127, True),
    "uint8": (0, 255, False),
}
W_BIT_ORDER = ["uint3", "uint4", "int4", "int8"]  # BIT_TYPE_LIST minus uint8 (layers.py:178-180)
EPS = torch.finfo(torch.float32).eps
_LN2 = torch.log(torch.tensor([2.0]))


def round_ln(x, mode=None):
    """floor/ceil/nearest-in-linear-distance power-of-two exponent (minmax.py:50-64)."""
    y = torch.div(torch.log(x), _LN2.to(x.device))
    if mode == "ceil":
        return torch.ceil(y)
    y = torch.floor(y)
    if mode == "floor":
        return y
    return torch.gt(x - 2 ** y, 2 ** (y + 1) - x) + y


def tdiv(x, c):
    """x / c as an IEEE division on every backend: torch-CUDA turns `tensor / python_scalar` into a multiplication by the
    reciprocal (ATen BinaryDivTrueKernel.cu, is_cpu_scalar branch), torch-CPU divides.  The golden vectors come from the
    reference's CPU run, so the division is the arithmetic to restate; a 0-dim tensor divisor on x's device keeps it one."""
    return x / torch.tensor(float(c), dtype=x.dtype, device=x.device)


def act_shape(x):
    return {2: (1, -1), 3: (1, 1, -1), 4: (1, -1, 1, 1)}[x.dim()]


def quantize(x, scale, zp, lo, hi, shape):
    return (x / scale.reshape(shape) + zp.reshape(shape)).round().clamp(lo, hi)


def fake_quant(x, scale, zp, lo, hi, shape):
    q = quantize(x, scale, zp, lo, hi, shape)
    return (q - zp.reshape(shape)) * scale.reshape(shape)


def lp2(a, b):
    return (a - b).abs().pow(2.0).mean()


def _as_channels_first(v, kind):
    # observer/base.py:16-29
    v = v.detach()
    if kind in ("conv_weight", "linear_weight"):
        return v.reshape(v.shape[0], -1)
    if v.dim() == 4:
        v = v.permute(0, 2, 3, 1)
    return v.reshape(-1, v.shape[-1]).transpose(0, 1)


class _RangeObs:
    def __init__(self, kind, bit, mode):
        self.kind, self.bit, self.mode = kind, bit, mode
        self.max_val = self.min_val = None
        self.symmetric = BITS[bit][2]

    def update(self, v):
        self.v = v
        c = _as_channels_first(v, self.kind)
        cur_max, cur_min = c.max(axis=1).values, c.min(axis=1).values
        self.max_val = cur_max if self.max_val is None else torch.max(cur_max, self.max_val)
        self.min_val = cur_min if self.min_val is None else torch.min(cur_min, self.min_val)
        if self.mode == "layer_wise":
            self.max_val, self.min_val = self.max_val.max(), self.min_val.min()


class MinmaxObs(_RangeObs):
    """Running min/max + power-of-two scale search over 4 exponents (minmax.py:145-207)."""

    def params(self, x, others=None):
        lo, hi, _ = BITS[self.bit]
        max_val, min_val = self.max_val, self.min_val
        if self.symmetric:
            zp = torch.zeros_like(max_val, dtype=torch.int64)
            scale = torch.max(-min_val, max_val) / (float(hi - lo) / 2)
            zp_f = torch.tensor([0.0])
        else:
            scale = (max_val - min_val) / float(hi - lo)
            zp = lo - torch.round(min_val / scale)
            zp.clamp_(lo, hi)
            zp_f = zp
        alpha = round_ln(scale, "round")
        floor = round_ln(scale, "floor")
        dim = 1 if self.mode == "layer_wise" else scale.shape[0]
        for j in range(dim):
            if dim == 1:
                w = x if self.kind == "activation" else self.v
            else:
                w = self.v[j, ...].unsqueeze(0)
            ref_out = self._out(w, j, x, others, False)
            scores = []
            for d in (-1, 0, 1, 2):
                s = 2 ** (floor[j] + d)
                wq = ((w / s + zp_f).round().clamp(lo, hi) - zp_f) * s
                scores.append(lp2(ref_out, self._out(wq, j, x, others, True)))
            alpha[j] = floor[j] - 1 + scores.index(min(scores))
        scale = 2 ** alpha
        scale.clamp_(EPS)
        return scale, zp

    def _out(self, w, j, x, others, quant):
        if self.kind == "activation":
            return w if quant else x
        if not quant:
            w = self.v[j, ...].unsqueeze(0) if self.mode == "channel_wise" else self.v
        bias = None
        if others and others[0] is not None:
            bias = others[0][j].unsqueeze(0) if self.mode == "channel_wise" else others[0]
        if self.kind == "conv_weight":
            return F.conv2d(x, w, bias, *others[1:5])
        return F.linear(x, w, bias)


class PtfObs(_RangeObs):
    """Power-of-two-factor per-channel scale {1,2,4,8} * s1 (ptf.py:32-152)."""

    def params(self, x, others=None):
        lo, hi, _ = BITS[self.bit]
        max_t = torch.max(-self.min_val.min(), self.max_val.max())
        s8 = 2 * max_t / float(hi - lo)
        s8.clamp_(EPS)
        s4 = s8 / 2
        s2 = s4 / 2
        s1 = s2 / 2
        zp = torch.zeros_like(self.max_val.max(), dtype=torch.int64)
        mask = torch.ones_like(self.max_val)
        for j in range(x.shape[2]):
            d = x[..., j].unsqueeze(-1)
            scores = [lp2(d, ((d / s + zp).round().clamp(lo, hi) - zp) * s) for s in (s1, s2, s4, s8)]
            mask[j] *= 2 ** scores.index(min(scores))
        return s1 * mask, zp


class EmaObs(_RangeObs):
    def update(self, v, sigma=0.01):
        c = _as_channels_first(v, self.kind)
        cur_max, cur_min = c.max(axis=1).values, c.min(axis=1).values
        self.max_val = cur_max if self.max_val is None else self.max_val + sigma * (cur_max - self.max_val)
        self.min_val = cur_min if self.min_val is None else self.min_val + sigma * (cur_min - self.min_val)
        if self.mode == "layer_wise":
            self.max_val, self.min_val = self.max_val.max(), self.min_val.min()

    def params(self, x=None, others=None):
        return _plain_range_params(self)


class PercentileObs(_RangeObs):
    def update(self, v, sigma=0.01, alpha=0.99999):
        assert self.mode == "layer_wise"
        flat = _as_channels_first(v, self.kind).reshape(-1)
        try:
            cur_max, cur_min = torch.quantile(flat, alpha), torch.quantile(flat, 1.0 - alpha)
        except RuntimeError:  # > 16.7M elements (percentile.py:33-43)
            cur_max = torch.tensor(np.percentile(flat.numpy(), alpha * 100), dtype=torch.float32)
            cur_min = torch.tensor(np.percentile(flat.numpy(), (1 - alpha) * 100), dtype=torch.float32)
        self.max_val = cur_max if self.max_val is None else self.max_val + sigma * (cur_max - self.max_val)
        self.min_val = cur_min if self.min_val is None else self.min_val + sigma * (cur_min - self.min_val)

    def params(self, x=None, others=None):
        return _plain_range_params(self)


def _plain_range_params(o):
    lo, hi, _ = BITS[o.bit]
    if o.symmetric:
        scale = torch.max(-o.min_val, o.max_val) / (float(hi - lo) / 2)
        scale.clamp_(EPS)
        return scale, torch.zeros_like(o.max_val, dtype=torch.int64)
    scale = (o.max_val - o.min_val) / float(hi - lo)
    scale.clamp_(EPS)
    zp = lo - torch.round(o.min_val / scale)
    zp.clamp_(lo, hi)
    return scale, zp


class OmseObs(_RangeObs):
    """90-step range shrink, asymmetric, non-PoT (omse.py:30-57; kwargs-tolerant)."""

    def params(self, x, others=None):
        lo, hi, _ = BITS[self.bit]
        best = 1e10
        for i in range(90):
            new_max = self.max_val * (1.0 - (i * 0.01))
            new_min = self.min_val * (1.0 - (i * 0.01))
            s = (new_max - new_min) / float(hi - lo)
            s.clamp_(EPS)
            z = lo - torch.round(new_min / s)
            z.clamp_(lo, hi)
            score = lp2(x, ((x / s + z).round().clamp(lo, hi) - z) * s)
            if score < best:
                best, scale, zp = score, s, z
        return scale, zp


OBSERVERS = {"minmax": MinmaxObs, "ema": EmaObs, "percentile": PercentileObs, "omse": OmseObs, "ptf": PtfObs}


# --------------------------------------------------------------------------- integer ops
def int_layernorm(x, in_scale, out_scale, weight, bias, exact_sums=False, in_scale_expand=1):
    """QIntLayerNorm 'int' mode (layers.py:294-337).  x: dequantized fp32 [.., C]."""
    if in_scale_expand != 1:
        in_scale = in_scale.unsqueeze(-1).expand(-1, in_scale_expand).T.reshape(-1)
    C = x.shape[-1]
    in_scale = in_scale.reshape(1, 1, -1)
    out_scale = out_scale.reshape(1, 1, -1)
    x_q = (x / in_scale).round()
    s1 = in_scale.min()
    x_q = x_q * (in_scale / s1).round()
    if exact_sums:
        xd = x_q.double()
        sum_x = xd.sum(dim=-1).float()
        sum_sq = (xd * xd).sum(dim=-1).float()
        mean = tdiv(sum_x, C) * s1
    else:
        sum_x = x_q.sum(dim=-1)
        sum_sq = (x_q ** 2).sum(dim=-1)
        mean = x_q.mean(dim=-1) * s1
    var = C * sum_sq - sum_x ** 2
    # torch's CPU sqrt goes through MKL VML (<= 1 ulp, not correctly rounded: ~0.1% of inputs differ from IEEE sqrt, which is
    # what torch-CUDA and the kernels compute); the canonical (exact) variant uses the correctly rounded one
    # (on a CUDA device torch.sqrt is the IEEE one already)
    root = torch.from_numpy(np.sqrt(var.numpy())) if (exact_sums and not var.is_cuda) else torch.sqrt(var)
    std = tdiv(s1, C) * root
    g = weight.reshape(1, 1, -1)
    A = (s1 / std).unsqueeze(-1) * g / out_scale
    N = torch.clamp(7 - torch.floor(torch.log2(A.abs())), 0, 31)
    M = torch.clamp(torch.floor(A.abs() * torch.pow(2, N)), 0, 255)
    B = ((bias.reshape(1, 1, -1) - (mean / std).unsqueeze(-1) * g) / out_scale * torch.pow(2, N)).round()
    y_q = ((A.sign() * M * x_q + B) / torch.pow(2, N)).round()
    return y_q * out_scale


def _log_round(x):
    big = x.log2().floor()
    extra = (x - 2 ** big) >= 2 ** (big - 1)
    big[extra] = big[extra] + 1
    return big


def int_softmax_log2(x, s, bits=4, exact_sums=False, codes=None):
    """QIntSoftmax.forward with log_i_softmax (layers.py:384-428).  x: dequantized scores.  `codes` (canonical variant
    for raw fp32 scales): the integer codes themselves replace fl(fl(code*s)/s), which is off by one ulp from the
    integer for ~15 % of the codes when s is not a power of two."""
    x_int = x / s if codes is None else codes
    x_int = x_int - x_int.max(dim=-1, keepdim=True).values
    n = 32
    x0 = torch.floor(-0.6931 / s)
    x_int = torch.max(x_int, n * x0)
    q = torch.floor(x_int / x0)
    r = x_int - x0 * q
    c0, c1, c2 = 0.35815147, 0.96963238 / 0.35815147, 1.0 / 0.35815147
    z = r * (r + torch.floor(c1 / s)) + torch.floor(c2 / s ** 2)
    e = torch.clamp(torch.floor(z * 2 ** (n - q)), min=0)
    if exact_sums:
        tot = _exact_row_sum_f32(e)
    else:
        tot = e.sum(dim=-1, keepdim=True)
    rounds = _log_round(torch.round(tot / e))
    mask = rounds >= 2 ** bits
    out = 2 ** (-torch.clamp(rounds, 0, 2 ** bits - 1))
    out[mask] = 0
    return out


def _exact_row_sum_f32(e):
    """Exactly-rounded fp32 of the row sum of integer-valued fp32 numbers (< 2^64)."""
    if e.is_cuda:
        # device form: hi / lo 32-bit halves summed in int64, recombined in int64 (needs the total < 2^63), int64 -> fp32 is
        # round-to-nearest-even (cvt.rn.f32.s64); larger values take the host path below
        if float(e.max()) * e.shape[-1] < 2.0 ** 62:
            a = e.double()
            hi = torch.floor(a / 4294967296.0)
            lo = a - hi * 4294967296.0
            tot = (hi.long().sum(dim=-1, keepdim=True) << 32) + lo.long().sum(dim=-1, keepdim=True)
            return tot.float()
        return _exact_row_sum_f32(e.cpu()).to(e.device)
    a = e.detach().numpy().astype(np.float64)  # exact: fp32 -> fp64
    hi = np.floor(a / 4294967296.0)
    lo = a - hi * 4294967296.0
    hs = hi.astype(np.uint64).sum(axis=-1, keepdims=True)
    ls = lo.astype(np.uint64).sum(axis=-1, keepdims=True)
    out = np.empty(hs.shape, dtype=np.float32)
    flat_h, flat_l, flat_o = hs.reshape(-1), ls.reshape(-1), out.reshape(-1)
    for i in range(flat_h.size):  # python ints are exact; int -> float32 via numpy rounds to nearest even
        v = (int(flat_h[i]) << 32) + int(flat_l[i])
        flat_o[i] = _int_to_f32_rne(v)
    return torch.from_numpy(out)


def _int_to_f32_rne(v):
    if v < (1 << 24):
        return np.float32(v)
    sh = v.bit_length() - 24
    m, rem = v >> sh, v & ((1 << sh) - 1)
    half = 1 << (sh - 1)
    if rem > half or (rem == half and (m & 1)):
        m += 1
    return np.float32(math.ldexp(m, sh))


# --------------------------------------------------------------------------- the model
class _ActQ:
    def __init__(self, bit="int8", mode="layer_wise", observer="minmax"):
        self.bit, self.obs = bit, OBSERVERS[observer]("activation", bit, mode)
        self.scale = self.zp = None

    def calibrate(self, x):
        self.obs.update(x)
        self.scale, self.zp = self.obs.params(x)

    def __call__(self, x):
        lo, hi, _ = BITS[self.bit]
        return fake_quant(x, self.scale, self.zp, lo, hi, act_shape(x))


class _WeightQ:
    def __init__(self, kind):
        self.kind = kind
        self.obs = MinmaxObs(kind, "int4", "channel_wise")  # Config: W=int4, channel_wise, minmax
        self.scale, self.zp = {}, {}

    def calibrate(self, w, x, others):
        """layers.py:62-85 / 175-201: all four bit types; returns the weight-MSE list."""
        dist = []
        shape = (-1, 1, 1, 1) if self.kind == "conv_weight" else (-1, 1)
        for bit in W_BIT_ORDER:
            self.obs.bit = bit
            self.obs.mode = "layer_wise" if bit == "int8" else "channel_wise"
            self.obs.update(w)
            self.scale[bit], self.zp[bit] = self.obs.params(x, others)
            dist.append(lp2(w, self.fq(w, bit)))
        return dist

    def fq(self, w, bit):
        lo, hi, _ = BITS[bit]
        shape = (-1, 1, 1, 1) if self.kind == "conv_weight" else (-1, 1)
        return fake_quant(w, self.scale[bit], self.zp[bit], lo, hi, shape)


ATTN_ALPHA, MLP_ALPHA = 0.35, 0.5  # vit_fquant.py:37, layers_quant.py:14 (single-entry pools)


class VitOracle:
    """Functional ViT/DeiT: calibrate() then forward_quant(); state in self.q (name -> quantizer)."""

    def __init__(self, sd, embed_dim, depth, num_heads, input_quant=True, method="minmax",
                 ptf=True, exact_sums=False, patch=16, device="cpu", **_):
        # device="cuda": the same restatement evaluated by torch's CUDA backend (IEEE sqrt and CUDA erff instead of
        # MKL VML sqrt / Sleef erf) - what the -m gpu tests use for strict all-image comparisons with the kernels
        self.dev = torch.device(device)
        self.sd = {k: v.float().to(self.dev) for k, v in sd.items()}
        self.D, self.L, self.H, self.P = embed_dim, depth, num_heads, patch
        self.input_quant, self.exact = input_quant, exact_sums
        ln_obs, ln_mode = ("ptf", "channel_wise") if ptf else (method, "layer_wise")
        A = lambda: _ActQ("int8", "layer_wise", method)
        LN = lambda: _ActQ("int8", ln_mode, ln_obs)
        q = self.q = {}
        if input_quant:
            q["qact_input"] = A()
        q["patch_embed.proj"] = _WeightQ("conv_weight")
        for nm in ("patch_embed.qact", "qact_embed", "qact_pos", "qact2", "act_out"):
            q[nm] = A()
        q["qact1"] = LN()
        q["head"] = _WeightQ("linear_weight")
        for i in range(depth):
            p = "blocks.%d." % i
            for nm in ("attn.qact0", "attn.qact1", "attn.qact2", "attn.qact_attn1", "mlp.qact0", "mlp.qact1"):
                q[p + nm] = A()
            for nm in ("attn.qact3", "qact2", "mlp.qact2", "qact4"):
                q[p + nm] = LN()
            for nm in ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2"):
                q[p + nm] = _WeightQ("linear_weight")
        self.cs = {}  # smoothing channel scales: "blocks.i.attn" / "blocks.i.mlp" -> [D]

    # ---- state (de)serialisation, names shared with the golden files and the product
    def export_state(self):
        out = {}
        for nm, qq in self.q.items():
            if isinstance(qq, _ActQ):
                out[nm + ".scale"] = qq.scale.reshape(-1).clone()
                out[nm + ".zero_point"] = qq.zp.reshape(-1).clone()
            else:
                for bit in qq.scale:
                    out["%s.scale.%s" % (nm, bit)] = qq.scale[bit].reshape(-1).clone()
                    out["%s.zero_point.%s" % (nm, bit)] = qq.zp[bit].reshape(-1).clone()
        for nm, v in self.cs.items():
            out[nm + ".channel_scale"] = v.clone()
        return out

    def load_state(self, st):
        for nm, qq in self.q.items():
            if isinstance(qq, _ActQ):
                qq.scale = torch.as_tensor(st[nm + ".scale"]).float().to(self.dev)
                qq.zp = torch.as_tensor(st[nm + ".zero_point"]).long().to(self.dev)
            else:
                for bit in W_BIT_ORDER:
                    k = "%s.scale.%s" % (nm, bit)
                    if k in st:
                        qq.scale[bit] = torch.as_tensor(st[k]).float().to(self.dev)
                        qq.zp[bit] = torch.as_tensor(st["%s.zero_point.%s" % (nm, bit)]).long().to(self.dev)
        for k, v in st.items():
            if k.endswith(".channel_scale"):
                self.cs[k[: -len(".channel_scale")]] = torch.as_tensor(v).float().to(self.dev)

    # ---- shared pieces
    def _ln(self, x, name, eps=1e-6):
        return F.layer_norm(x, (self.D,), self.sd[name + ".weight"], self.sd[name + ".bias"], eps)

    def _int_ln(self, x, name, in_q, out_q, cs=None):
        out_scale = out_q.scale if cs is None else out_q.scale * cs
        return int_layernorm(x, in_q.scale, out_scale, self.sd[name + ".weight"], self.sd[name + ".bias"], self.exact)

    def _smooth_scale(self, x, w, alpha):
        gmax = torch.abs(x).max(axis=1).values.max(axis=0).values
        wmax = torch.abs(w).max(axis=0).values
        return 2 ** round_ln(gmax ** alpha / (wmax ** (1 - alpha)), "round")

    def _heads(self, x):
        B, N, _ = x.shape
        qkv = x.reshape(B, N, 3, self.H, self.D // self.H).permute(2, 0, 3, 1, 4)
        return qkv[0], qkv[1], qkv[2]

    # ---- canonical (exact_sums) arithmetic for the matmuls: exact integer accumulation of the codes, then the same two fp32
    #      operations the reference performs on the accumulated value.  Identical to the reference's fp32 matmul whenever the
    #      scales are powers of two (every product and partial sum is then exact); for raw fp32 scales (ema / percentile /
    #      omse observers) the reference's result depends on the BLAS summation order, this one does not.
    def _codes(self, h, aq):
        return torch.round(h / aq.scale.reshape(-1))

    def _lin(self, h, aq, wq, w, bit, bias):
        w_hat = wq.fq(w, bit)
        wm = w_hat.reshape(w.shape[0], -1)
        if not self.exact or aq is None or aq.scale.numel() != 1:
            return F.linear(h, wm, bias)
        # (with a zero point, round(h / s) is the centred code q - zp: the accumulation below is sum_k (q - zp) w, exact)
        ws = wq.scale[bit].reshape(-1, 1).float()
        wc = torch.round(wm / ws)
        acc = (self._codes(h, aq).double() @ wc.double().T).float()
        acc_scale = (aq.scale.reshape(-1).float() * ws.reshape(-1)).reshape(*([1] * (h.dim() - 1)), -1)
        y = acc * acc_scale
        return y if bias is None else y + bias

    def _attention(self, h, q1, qs, q2, B):
        """h: dequantized qkv (attn.qact1 output) -> dequantized attn.qact2 output, plus the score / probability taps"""
        dh = self.D // self.H
        qh, kh, vh = self._heads(h)
        if not self.exact:
            a = qs((qh @ kh.transpose(-2, -1)) * dh ** -0.5)
            p = int_softmax_log2(a, qs.scale, 4, self.exact)
            return a, p, q2((p @ vh).transpose(1, 2).reshape(B, -1, self.D))
        s1, sa, s2 = q1.scale.reshape(()).double(), qs.scale.reshape(()).double(), q2.scale.reshape(()).double()
        cq, ck, cv = (torch.round(t / q1.scale.reshape(())).double() for t in (qh, kh, vh))
        S = (cq @ ck.transpose(-2, -1)).float()
        mult = (s1 * s1 * dh ** -0.5 / sa).float()
        # zero points (asymmetric observers; all 0 otherwise): cq / ck / cv are the centred codes q - zp, the score code is
        # clamp(RNE(fl(fl(S * mult) + zp_s))), the softmax works on the codes (a zero point cancels in x - rowmax)
        zs, z2 = qs.zp.reshape(()).float(), q2.zp.reshape(()).float()
        ca = torch.clamp(torch.round(S * mult + zs), -128, 127)
        a = (ca - zs) * qs.scale.reshape(())
        p = int_softmax_log2(a, qs.scale, 4, True, codes=ca)
        O = ((p.double() * 32768.0) @ cv).float()
        omult = (s1 / s2 / 32768.0).float()
        c2 = torch.clamp(torch.round(O * omult + z2), -128, 127)
        return a, p, ((c2 - z2) * q2.scale.reshape(())).transpose(1, 2).reshape(B, -1, self.D)

    # ---- calibration forward (FP values + observers), test_quant.py:275-312
    @torch.no_grad()
    def calibrate(self, x):
        q, sd = self.q, self.sd
        gd = []
        B = x.shape[0]
        if self.input_quant:
            q["qact_input"].calibrate(x)
        w, b = sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"]
        q["patch_embed.proj"].calibrate(w, x, [b, (self.P, self.P), (0, 0), (1, 1), 1])
        x = F.conv2d(x, w, b, (self.P, self.P)).flatten(2).transpose(1, 2)
        q["patch_embed.qact"].calibrate(x)
        x = torch.cat((sd["cls_token"].expand(B, -1, -1), x), dim=1)
        q["qact_embed"].calibrate(x)
        q["qact_pos"].calibrate(sd["pos_embed"])
        x = x + sd["pos_embed"]
        q["qact1"].calibrate(x)
        for i in range(self.L):
            p = "blocks.%d." % i
            # attention (vit_fquant.py:232-333 calibrate branch)
            h = self._ln(x, p + "norm1")
            w, b = sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]
            cs = self.cs[p + "attn"] = self._smooth_scale(h, w, ATTN_ALPHA)
            hs, ws = h / cs.reshape(1, 1, -1), w * cs.reshape(1, -1)
            gt = F.linear(hs, ws, b)
            q[p + "attn.qact0"].calibrate(hs)
            gd.append(q[p + "attn.qkv"].calibrate(ws, hs, [b]))
            h = gt
            q[p + "attn.qact1"].calibrate(h)
            qh, kh, vh = self._heads(h)
            a = (qh @ kh.transpose(-2, -1)) * (self.D // self.H) ** -0.5
            q[p + "attn.qact_attn1"].calibrate(a)
            a = int_softmax_log2(a, q[p + "attn.qact_attn1"].scale, 4, self.exact)
            h = (a @ vh).transpose(1, 2).reshape(B, -1, self.D)
            q[p + "attn.qact2"].calibrate(h)
            w, b = sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"]
            gd.append(q[p + "attn.proj"].calibrate(w, h, [b]))
            h = F.linear(h, w, b)
            q[p + "attn.qact3"].calibrate(h)
            x = x + h
            q[p + "qact2"].calibrate(x)
            # mlp (layers_quant.py:255-347)
            h = self._ln(x, p + "norm2")
            w, b = sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]
            cs = self.cs[p + "mlp"] = self._smooth_scale(h, w, MLP_ALPHA)
            hs, ws = h / cs.reshape(1, 1, -1), w * cs.reshape(1, -1)
            gt = F.linear(hs, ws, b)
            q[p + "mlp.qact0"].calibrate(hs)
            gd.append(q[p + "mlp.fc1"].calibrate(ws, hs, [b]))
            h = F.gelu(gt)
            q[p + "mlp.qact1"].calibrate(h)
            w, b = sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"]
            gd.append(q[p + "mlp.fc2"].calibrate(w, h, [b]))
            h = F.linear(h, w, b)
            q[p + "mlp.qact2"].calibrate(h)
            x = x + h
            q[p + "qact4"].calibrate(x)
        x = self._ln(x, "norm")[:, 0]
        q["qact2"].calibrate(x)
        w, b = sd["head.weight"], sd["head.bias"]
        gd.append(q["head"].calibrate(w, x, [b]))
        x = F.linear(x, w, b)
        q["act_out"].calibrate(x)
        return x, gd

    # ---- quantized forward (fake-quant fp32, the reference's inference path)
    @torch.no_grad()
    def forward_quant(self, x, bit_config=None, taps=None):
        """bit_config: list of 1+4L+1 ints in {4,8} (vit_fquant.py:876-880,928-932)."""
        q, sd = self.q, self.sd
        L = self.L
        bits = list(bit_config) if bit_config else [8] * (4 * L + 2)
        tap = (lambda n, v: taps.__setitem__(n, v.clone())) if taps is not None else (lambda n, v: None)
        wname = lambda b: "int%d" % b
        B = x.shape[0]
        x = x.to(self.dev)
        if self.input_quant:
            x = q["qact_input"](x)
            tap("qact_input", x)
        if self.input_quant and self.exact:
            P, gs = self.P, x.shape[-1] // self.P
            rows = x.reshape(B, 3, gs, P, gs, P).permute(0, 2, 4, 1, 3, 5).reshape(B, gs * gs, 3 * P * P)
            x = self._lin(rows, q["qact_input"], q["patch_embed.proj"], sd["patch_embed.proj.weight"], wname(bits[0]), sd["patch_embed.proj.bias"])
        else:
            w = q["patch_embed.proj"].fq(sd["patch_embed.proj.weight"], wname(bits[0]))
            x = F.conv2d(x, w, sd["patch_embed.proj.bias"], (self.P, self.P)).flatten(2).transpose(1, 2)
        x = q["patch_embed.qact"](x)
        tap("patch_embed.qact", x)
        x = torch.cat((sd["cls_token"].expand(B, -1, -1), x), dim=1)
        x = q["qact_embed"](x)
        x = x + q["qact_pos"](sd["pos_embed"])
        x = q["qact1"](x)
        tap("qact1", x)
        last = q["qact1"]
        for i in range(L):
            p = "blocks.%d." % i
            b4 = bits[4 * i + 1: 4 * i + 5]
            cs_a, cs_m = self.cs[p + "attn"], self.cs[p + "mlp"]
            h = self._int_ln(x, p + "norm1", last, q[p + "attn.qact0"], cs_a)
            tap(p + "norm1", h)
            h = q[p + "attn.qact0"](h / cs_a.reshape(1, 1, -1))
            tap(p + "attn.qact0", h)
            h = q[p + "attn.qact1"](self._lin(h, q[p + "attn.qact0"], q[p + "attn.qkv"], sd[p + "attn.qkv.weight"] * cs_a.reshape(1, -1),
                                              wname(b4[0]), sd[p + "attn.qkv.bias"]))
            tap(p + "attn.qact1", h)
            a, pr, h = self._attention(h, q[p + "attn.qact1"], q[p + "attn.qact_attn1"], q[p + "attn.qact2"], B)
            tap(p + "attn.qact_attn1", a)
            tap(p + "attn.log_int_softmax", pr)
            tap(p + "attn.qact2", h)
            h = q[p + "attn.qact3"](self._lin(h, q[p + "attn.qact2"], q[p + "attn.proj"], sd[p + "attn.proj.weight"], wname(b4[1]),
                                              sd[p + "attn.proj.bias"]))
            tap(p + "attn.qact3", h)
            x = q[p + "qact2"](x + h)
            tap(p + "qact2", x)
            h = self._int_ln(x, p + "norm2", q[p + "qact2"], q[p + "mlp.qact0"], cs_a)  # attn's scale: SURVEY Q7
            tap(p + "norm2", h)
            h = q[p + "mlp.qact0"](h / cs_m.reshape(1, 1, -1))
            tap(p + "mlp.qact0", h)
            h = q[p + "mlp.qact1"](F.gelu(self._lin(h, q[p + "mlp.qact0"], q[p + "mlp.fc1"], sd[p + "mlp.fc1.weight"] * cs_m.reshape(1, -1),
                                                    wname(b4[2]), sd[p + "mlp.fc1.bias"])))
            tap(p + "mlp.qact1", h)
            h = q[p + "mlp.qact2"](self._lin(h, q[p + "mlp.qact1"], q[p + "mlp.fc2"], sd[p + "mlp.fc2.weight"], wname(b4[3]),
                                             sd[p + "mlp.fc2.bias"]))
            tap(p + "mlp.qact2", h)
            x = q[p + "qact4"](x + h)
            tap(p + "qact4", x)
            last = q[p + "qact4"]
        x = self._int_ln(x, "norm", last, q["qact2"])[:, 0]
        x = q["qact2"](x)
        tap("qact2", x)
        x = q["act_out"](self._lin(x, q["qact2"], q["head"], sd["head.weight"], wname(bits[-1]), sd["head.bias"]))
        return x
