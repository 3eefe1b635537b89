"""Percentile observer (reference: models/ptq/observer/percentile.py:9-77): alpha = 0.99999 quantiles, EMA'd,
layer-wise only.  torch.quantile refuses > 16.7M elements (the reference then falls back to numpy on the CPU,
percentile.py:33-43); here large tensors use an exact device-side order statistic (kthvalue) with the same
linear interpolation, so nothing leaves the GPU.  Multi-GPU: the quantile of a sharded batch is not
decomposable; ranks exchange their top/bottom tails instead (see calibrate.py)."""
import torch

from .base import BaseObserver
from .ema import plain_range_params


def _quantile(flat, q):
    n = flat.numel()
    if n <= 16_000_000:
        return torch.quantile(flat, q)
    pos = q * (n - 1)
    lo = int(pos)
    hi = min(lo + 1, n - 1)
    a = torch.kthvalue(flat, lo + 1).values
    b = torch.kthvalue(flat, hi + 1).values
    return a + (b - a) * (pos - lo)


class PercentileObserver(BaseObserver):
    def __init__(self, module_type, bit_type, calibration_mode, percentile_sigma=0.01, percentile_alpha=0.99999):
        super().__init__(module_type, bit_type, calibration_mode)
        self.percentile_sigma = 0.01
        self.percentile_alpha = 0.99999
        self.symmetric = self.bit_type.signed

    def update(self, v):
        assert self.calibration_mode == "layer_wise"  # channel-wise needs too much time (percentile.py:27-28)
        flat = self.reshape_tensor(v).reshape(-1).float()
        cur_max = _quantile(flat, self.percentile_alpha)
        cur_min = _quantile(flat, 1.0 - self.percentile_alpha)
        sig = self.percentile_sigma
        self.max_val = cur_max if self.max_val is None else self.max_val + sig * (cur_max - self.max_val)
        self.min_val = cur_min if self.min_val is None else self.min_val + sig * (cur_min - self.min_val)

    def get_quantization_params(self, *args, **kwargs):
        return plain_range_params(self)
