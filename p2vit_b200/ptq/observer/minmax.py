"""Min/max observer with the power-of-two scale search (reference: models/ptq/observer/minmax.py:9-237).

The reference evaluates, per output channel and in a Python loop, four candidate exponents
floor(log2 s) + {-1,0,1,2} by fake-quantising and re-running the layer (5 tiny F.linear per channel).
Here the same scores are produced for all channels at once: activations through the block-reduce kernel
`p2v_quant_mse_scores`, weights through one launch of the fp32 GEMM + column-square-sum kernel (csrc/sgemm.cu) on
the stacked candidate differences W - fq_k(W).
Scores are all-reduced over ranks before the arg-min so every rank picks identical exponents.
"""
import torch

from ... import ops
from .base import BaseObserver
from .utils import allreduce_, pot_exponent


class MinmaxObserver(BaseObserver):
    def __init__(self, module_type, bit_type, calibration_mode):
        super().__init__(module_type, bit_type, calibration_mode)
        self.symmetric = self.bit_type.signed

    def update(self, v):
        self.v = v
        self._running_range(v, torch.max, torch.min)
        self.allreduce_range()

    def get_quantization_params(self, x, others=None, attn=False, attn_para=None, *args, **kwargs):
        qmax, qmin = self.bit_type.upper_bound, self.bit_type.lower_bound
        max_val, min_val = self.max_val, self.min_val
        if self.symmetric:
            zero_point = torch.zeros_like(max_val, dtype=torch.int64)
            scale = torch.max(-min_val, max_val) / (float(qmax - qmin) / 2)
            zp_f = None
        else:
            scale = (max_val - min_val) / float(qmax - qmin)
            zero_point = qmin - torch.round(min_val / scale)
            zero_point.clamp_(qmin, qmax)
            zp_f = zero_point.float()
        floor = pot_exponent(scale, "floor").reshape(-1)                       # [1] or [C]
        cand = torch.stack([floor + d for d in (-1, 0, 1, 2)])                 # [4, 1|C] exponents
        if self.module_type == "activation":
            zps = None if zp_f is None else zp_f.reshape(1, -1).expand(4, -1).contiguous()
            scores = ops.quant_mse_scores(x, 2 ** cand, qmin, qmax, zps, per_channel_out=False).reshape(4, 1)
        else:
            scores = self._weight_scores(x, others, 2 ** cand, zp_f, qmin, qmax)
        allreduce_(scores, "sum")
        alpha = floor - 1 + torch.argmin(scores, dim=0).to(floor.dtype)        # first minimum, like list.index(min)
        scale = 2 ** alpha
        scale.clamp_(self.eps)
        return scale, zero_point

    def _weight_scores(self, x, others, cand_scales, zp_f, qmin, qmax):
        """score[k, j] = sum over calibration rows of (layer(x; W)[., j] - layer(x; fq_k(W))[., j])^2, reduced over j when
        layer_wise (minmax.py:82-141,165-201).  The two layer outputs differ by x . (W - fq_k(W))[j, :] (the bias cancels), so
        the four candidates' difference rows are stacked and ONE launch of the fp32 GEMM + column-square-sum kernel
        (csrc/sgemm.cu) returns all scores; the [rows, 4 Cout] product is never written."""
        w = self.v.detach().float()
        wm = w.reshape(w.shape[0], -1)
        patch = 0
        if self.module_type == "conv_weight":
            stride, k = others[1], w.shape[-1]
            assert tuple(stride) == (k, k) and tuple(others[2]) == (0, 0), "QConv2d is the patch-embed conv (kernel == stride)"
            patch, xm = k, x
        else:
            xm = x.reshape(-1, x.shape[-1])
        zp = 0.0 if zp_f is None else zp_f.reshape(-1, 1)
        K4 = cand_scales.shape[0]
        diffs = []
        for kidx in range(K4):
            s = cand_scales[kidx].reshape(-1, 1)
            diffs.append(wm - ((wm / s + zp).round().clamp(qmin, qmax) - zp) * s)
        sc = ops.linear_sqerr_scores(xm, torch.cat(diffs, dim=0), patch=patch).reshape(K4, wm.shape[0])
        return sc if self.calibration_mode == "channel_wise" else sc.sum(dim=1, keepdim=True)
