"""EMA min/max observer (reference: models/ptq/observer/ema.py:7-51); raw fp32 scale, not power of two."""
import torch

from .base import BaseObserver


def plain_range_params(obs):
    qmax, qmin = obs.bit_type.upper_bound, obs.bit_type.lower_bound
    if obs.symmetric:
        scale = torch.max(-obs.min_val, obs.max_val) / (float(qmax - qmin) / 2)
        scale.clamp_(obs.eps)
        return scale, torch.zeros_like(obs.max_val, dtype=torch.int64)
    scale = (obs.max_val - obs.min_val) / float(qmax - qmin)
    scale.clamp_(obs.eps)
    zero_point = qmin - torch.round(obs.min_val / scale)
    zero_point.clamp_(qmin, qmax)
    return scale, zero_point


class EmaObserver(BaseObserver):
    def __init__(self, module_type, bit_type, calibration_mode, ema_sigma=0.01):
        super().__init__(module_type, bit_type, calibration_mode)
        self.ema_sigma = ema_sigma
        self.symmetric = self.bit_type.signed

    def update(self, v):
        sig = self.ema_sigma
        self._running_range(v, lambda cur, old: old + sig * (cur - old), lambda cur, old: old + sig * (cur - old))
        self.allreduce_range()

    def get_quantization_params(self, *args, **kwargs):
        return plain_range_params(self)
