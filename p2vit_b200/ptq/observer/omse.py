"""OMSE observer (reference: models/ptq/observer/omse.py:7-57): 90-step range shrink, asymmetric, raw fp32
scale.  The 90 candidate scores come from one `p2v_quant_mse_scores` launch instead of 90 elementwise
passes.  Accepts and ignores the kwargs QAct passes (the reference's signature does not)."""
import torch

from ... import ops
from .base import BaseObserver
from .utils import allreduce_


class OmseObserver(BaseObserver):
    def update(self, v):
        self._running_range(v, torch.max, torch.min)
        self.allreduce_range()

    def get_quantization_params(self, inputs, *args, **kwargs):
        assert self.calibration_mode == "layer_wise" and self.module_type == "activation"
        qmax, qmin = self.bit_type.upper_bound, self.bit_type.lower_bound
        steps = torch.arange(90, device=inputs.device, dtype=torch.float32)
        shrink = 1.0 - (steps * 0.01)
        new_max = self.max_val * shrink
        new_min = self.min_val * shrink
        scales = (new_max - new_min) / float(qmax - qmin)
        scales.clamp_(self.eps)
        zps = qmin - torch.round(new_min / scales)
        zps.clamp_(qmin, qmax)
        scores = ops.quant_mse_scores(inputs, scales.reshape(90, 1), qmin, qmax, zps.reshape(90, 1), per_channel_out=False).reshape(-1)
        allreduce_(scores, "sum")
        i = int(torch.argmin(scores))  # first strict minimum == the reference's `score < best_score` scan
        self.max_val, self.min_val = new_max[i], new_min[i]
        return scales[i], zps[i]
