"""Uniform quantizer (reference: models/ptq/quantizer/uniform.py:8-126).

q = clamp(round(x / scale + zero_point), lo, hi);  x_hat = (q - zero_point) * scale, with the same
broadcasting rules.  Activations keep their parameters in `.scale/.zero_point`, weights keep one entry per
bit type in `.dic_scale/.dic_zero_point` (other modules read these fields, vit_fquant.py:385,413-415,519-524).
The arithmetic runs in csrc/rowops.cu (`p2v_fake_quant_f32`, `p2v_quantize_f32`); weights (tiny, quantized
once per bit_config by the engine) use the same kernels through their [Cout, rest] view.
"""
import torch

from ... import ops
from .base import BaseQuantizer


class UniformQuantizer(BaseQuantizer):
    def __init__(self, bit_type, observer, module_type):
        super().__init__(bit_type, observer, module_type)
        self.scale = None
        self.zero_point = None
        self.dic_scale = {}
        self.dic_zero_point = {}

    def update_quantization_params(self, *args, **kwargs):
        scale, zero_point = self.observer.get_quantization_params(*args, **kwargs)
        if self.module_type == "activation":
            self.scale, self.zero_point = scale, zero_point
        else:
            self.dic_scale[self.bit_type.name] = scale
            self.dic_zero_point[self.bit_type.name] = zero_point

    def _params(self, scale, zero_point):
        if scale is None:
            scale = self.scale if self.module_type == "activation" else self.dic_scale[self.bit_type.name]
        if zero_point is None:
            zero_point = self.zero_point if self.module_type == "activation" else self.dic_zero_point[self.bit_type.name]
        return scale, zero_point

    def _geometry(self, inputs):
        """view of `inputs` whose channel axis follows ops._channel_geometry"""
        if self.module_type in ("conv_weight", "linear_weight"):
            return inputs.reshape(1, inputs.shape[0], -1, 1)
        return inputs

    @staticmethod
    def _zp_scalar(zero_point):
        zp = zero_point.reshape(-1).float()
        if zp.numel() > 1 and not bool((zp == zp[0]).all()):
            raise NotImplementedError("per-channel zero points are not produced by any observer of this path")
        return float(zp[0])

    def quant(self, inputs, scale=None, zero_point=None):
        scale, zero_point = self._params(scale, zero_point)
        lo, hi = self.bit_type.lower_bound, self.bit_type.upper_bound
        x = self._geometry(inputs.float())
        if lo >= -128 and hi <= 127:
            q = ops.quantize(x, scale, self._zp_scalar(zero_point), lo, hi).float()
        else:  # uint8 codes do not fit the int8 carrier: derive them from the fake-quant kernel output
            zp = self._zp_scalar(zero_point)
            y = ops.fake_quant(x, scale, zp, lo, hi)
            s = scale.reshape(-1).float()
            s = s.reshape(1, -1, 1, 1) if x.dim() == 4 else s
            q = torch.round(y / s + zp)
        return q.reshape(inputs.shape)

    def dequantize(self, inputs, scale=None, zero_point=None):
        scale, zero_point = self._params(scale, zero_point)
        shape = self.get_reshape_range(inputs)
        return (inputs - zero_point.reshape(shape).to(inputs.device)) * scale.reshape(shape).to(inputs.device)

    def forward(self, inputs):
        scale, zero_point = self._params(None, None)
        lo, hi = self.bit_type.lower_bound, self.bit_type.upper_bound
        y = ops.fake_quant(self._geometry(inputs.float()), scale, self._zp_scalar(zero_point), lo, hi)
        return y.reshape(inputs.shape)
