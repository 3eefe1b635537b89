"""Power-of-two-factor observer for LayerNorm inputs (reference: models/ptq/observer/ptf.py:8-152):
one fp32 base scale s1 = 2*max|x|/255/8 and a per-channel factor in {1,2,4,8} chosen by MSE.  The reference
loops over channels in Python; the four per-channel score vectors come from one kernel launch here."""
import torch

from ... import ops
from .base import BaseObserver
from .utils import allreduce_


class PtfObserver(BaseObserver):
    def update(self, v):
        self.v = v
        self._running_range(v, torch.max, torch.min)
        self.allreduce_range()

    def get_quantization_params(self, inputs, *args, **kwargs):
        qmax, qmin = self.bit_type.upper_bound, self.bit_type.lower_bound
        max_val_t = torch.max(-self.min_val.min(), self.max_val.max())
        scale8 = 2 * max_val_t / float(qmax - qmin)
        scale8.clamp_(self.eps)
        scale4 = scale8 / 2
        scale2 = scale4 / 2
        scale1 = scale2 / 2
        zero_point = torch.zeros_like(self.max_val.max(), dtype=torch.int64)
        cand = torch.stack([scale1, scale2, scale4, scale8]).reshape(4, 1)
        scores = ops.quant_mse_scores(inputs, cand, qmin, qmax, None, per_channel_out=True)    # [4, C]
        allreduce_(scores, "sum")
        self.scale_mask = (2 ** torch.argmin(scores, dim=0)).to(torch.float32)
        return scale1 * self.scale_mask, zero_point
