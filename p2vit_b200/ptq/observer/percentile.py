"""Percentile observer (reference: models/ptq/observer/percentile.py:9-77): alpha = 0.99999 quantiles, EMA'd, layer-wise only.

The reference sorts: torch.quantile, and np.percentile on the CPU for more than 16 777 216 elements (percentile.py:33-43).
Here the two order statistics a quantile interpolates between come from a most-significant-digit radix select
(csrc/rowops.cu: radix_hist_kernel, three histogram passes of 12 + 12 + 8 key bits, no sort, nothing leaves the GPU).
The histograms are integer counts, so in a data-parallel calibration they are all-reduced (SUM) pass by pass and every rank
finds the order statistics of the WHOLE batch: N ranks freeze exactly the scale one process would freeze on the concatenated
batch.  The interpolation restates the reference's arithmetic for each size class (fp32 rank and torch.lerp below 2^24
elements, numpy's float32 virtual index and _lerp above), so the result equals the reference's bit for bit."""
import numpy as np
import torch

from .base import BaseObserver
from .ema import plain_range_params
from .utils import allreduce_

TORCH_QUANTILE_LIMIT = 16_777_216      # torch.quantile's input size limit; the reference's except-branch takes over above it


def _key_to_float(key):
    """inverse of the order-preserving fp32 -> uint32 map of the kernel (float_order_key)"""
    bits = (key ^ 0x80000000) if key & 0x80000000 else (~key & 0xFFFFFFFF)
    return np.array([bits], dtype=np.uint32).view(np.float32)[0]


def select_kth(hist_fn, k, allreduce=allreduce_):
    """k-th smallest (0-based) element of the union of all ranks' data.  hist_fn(prefix_mask, prefix_value, shift, nbits) returns
    this rank's int64 digit counts (ops.radix_hist on the GPU; the CPU tests pass a numpy equivalent)."""
    mask, value = 0, 0
    for shift, nbits in ((20, 12), (8, 12), (0, 8)):
        hist = allreduce(hist_fn(mask, value, shift, nbits), "sum")
        cum = torch.cumsum(hist, 0)
        digit = int(torch.searchsorted(cum, torch.tensor([k], dtype=cum.dtype, device=cum.device), right=True)[0])
        k -= int(cum[digit - 1]) if digit else 0
        mask |= ((1 << nbits) - 1) << shift
        value |= digit << shift
    return _key_to_float(value)


def quantile_from_order_statistics(kth, n, q):
    """the reference's quantile of n elements given kth(i) = i-th smallest: torch.quantile's arithmetic (rank and weight in fp32,
    torch.lerp) up to TORCH_QUANTILE_LIMIT elements, np.percentile's (float64 virtual index, linear) above it"""
    if n <= TORCH_QUANTILE_LIMIT:
        rank = np.float32(q) * np.float32(n - 1)
        lo = int(np.floor(rank))
        w = np.float32(rank - np.float32(lo))
        a = kth(lo)
        b = kth(min(lo + 1, n - 1)) if w > 0 else a
        return float(torch.lerp(torch.tensor(a), torch.tensor(b), torch.tensor(w)))
    # np.percentile on a float32 array (numpy >= 2: q is divided by float32(100) and the virtual index (n - 1) * q stays in the
    # array's dtype - with ulp 2 above 2^24 - and so do the weight and the interpolation; percentile.py:36-43 passes alpha * 100)
    q32 = np.float32(q * 100.0) / np.float32(100)
    virt = np.float32(n - 1) * q32
    lo = int(np.floor(virt))
    g = np.float32(virt - np.float32(lo))
    a = np.float32(kth(min(lo, n - 1)))
    b = np.float32(kth(min(lo + 1, n - 1))) if g > 0 else a
    d = np.float32(b - a)
    return float(a + d * g if g < 0.5 else b - d * (np.float32(1) - g))     # numpy's _lerp


class PercentileObserver(BaseObserver):
    def __init__(self, module_type, bit_type, calibration_mode, percentile_sigma=0.01, percentile_alpha=0.99999):
        super().__init__(module_type, bit_type, calibration_mode)
        self.percentile_sigma = 0.01
        self.percentile_alpha = 0.99999
        self.symmetric = self.bit_type.signed

    def update(self, v):
        from ... import ops

        assert self.calibration_mode == "layer_wise"  # channel-wise needs too much time (percentile.py:27-28)
        flat = self.reshape_tensor(v).reshape(-1).float().contiguous()
        n = allreduce_(torch.tensor([flat.numel()], dtype=torch.int64, device=flat.device), "sum")
        n = int(n)
        hist_fn = lambda mask, value, shift, nbits: ops.radix_hist(flat, mask, value, shift, nbits)
        kth = lambda k: select_kth(hist_fn, k)
        dev = flat.device
        cur_max = torch.tensor(quantile_from_order_statistics(kth, n, self.percentile_alpha), device=dev)
        cur_min = torch.tensor(quantile_from_order_statistics(kth, n, 1.0 - self.percentile_alpha), device=dev)
        sig = self.percentile_sigma
        self.max_val = cur_max if self.max_val is None else self.max_val + sig * (cur_max - self.max_val)
        self.min_val = cur_min if self.min_val is None else self.min_val + sig * (cur_min - self.min_val)

    def get_quantization_params(self, *args, **kwargs):
        return plain_range_params(self)
