"""Host-side integer-arithmetic tables for the kernels.

`build_softmax_lut(scale)`: QIntSoftmax's exp_int (reference: models/ptq/layers.py:386-410) only depends on
d = rowmax_code - code in [0,255] once its input is a QAct code (x/scale is then an exact integer), so the
polynomial, the floor divisions and the 2^(32-q) shift are tabulated once per attention layer with the
reference's own fp32 operation sequence; the kernels (csrc/rowops.cu softmax_kernel, csrc/attention.cu)
sum the table entries exactly (hi*2^32 + lo) and apply the log2 rounding.
"""
import numpy as np
import torch

LUT_DTYPE = np.dtype([("hi", np.uint32, (256,)), ("lo", np.uint32, (256,)), ("exp_f32", np.float32, (256,))])


def softmax_exp_table(scale):
    """fp32 tensor [256]: exp_int for x_int = -d, d = 0..255 (same op order as int_exp/int_polynomial)."""
    s = torch.as_tensor(scale, dtype=torch.float32).reshape(()).cpu()
    n = 32
    x_int = -torch.arange(256, dtype=torch.float32)
    x0_int = torch.floor(-0.6931 / s)
    x_int = torch.max(x_int, n * x0_int)
    q = torch.floor(x_int / x0_int)
    r = x_int - x0_int * q
    coef = [0.35815147, 0.96963238, 1.0]
    coef[1] /= coef[0]
    coef[2] /= coef[0]
    b_int = torch.floor(coef[1] / s)
    c_int = torch.floor(coef[2] / s ** 2)
    z = r + b_int
    z = r * z
    z = z + c_int
    return torch.clamp(torch.floor(z * 2 ** (n - q)), min=0)


def build_softmax_lut(scale, max_row_len=1024):
    e = softmax_exp_table(scale)
    lut = np.zeros((), dtype=LUT_DTYPE)
    vals = [int(v) for v in e.double().tolist()]  # fp32 -> python int is exact (integer valued)
    if max(vals) * max_row_len >= 1 << 95 or max(vals) >= 1 << 55:
        # the tcgen05 attention kernel sums a row (<= 256 keys) exactly in 64 bits; 2^55 is reached only below scale ~2^-10,
        # far finer than an int8 score grid ever needs (include/p2vit_b200.h: p2v_softmax_lut)
        raise NotImplementedError("attention score scale %g is too small for the exact integer row sum" % float(scale))
    lut["hi"] = np.array([v >> 32 for v in vals], dtype=np.uint64).astype(np.uint32)
    lut["lo"] = np.array([v & 0xFFFFFFFF for v in vals], dtype=np.uint64).astype(np.uint32)
    lut["exp_f32"] = e.numpy()
    return lut


def lut_to_device(lut, device):
    return torch.from_numpy(np.frombuffer(lut.tobytes(), dtype=np.uint8).copy()).to(device)


def is_pot(t):
    """True when every element is an exact power of two (mantissa bits zero)."""
    t = torch.as_tensor(t, dtype=torch.float32).reshape(-1).cpu()
    m, _ = torch.frexp(t)
    return bool(((m == 0.5) & (t > 0)).all())
