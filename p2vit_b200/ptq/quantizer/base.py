"""Quantizer base (reference: models/ptq/quantizer/base.py:6-45)."""
import torch.nn as nn


class BaseQuantizer(nn.Module):
    def __init__(self, bit_type, observer, module_type):
        super().__init__()
        self.bit_type = bit_type
        self.observer = observer
        self.module_type = module_type

    def get_reshape_range(self, inputs):
        if self.module_type == "conv_weight":
            return (-1, 1, 1, 1)
        if self.module_type == "linear_weight":
            return (-1, 1)
        if self.module_type == "activation":
            try:
                return {2: (1, -1), 3: (1, 1, -1), 4: (1, -1, 1, 1)}[inputs.dim()]
            except KeyError:
                raise NotImplementedError
        raise NotImplementedError

    def update_quantization_params(self, *args, **kwargs):
        pass

    def quant(self, inputs, scale=None, zero_point=None):
        raise NotImplementedError

    def dequantize(self, inputs, scale=None, zero_point=None):
        raise NotImplementedError

    def forward(self, inputs):
        return self.dequantize(self.quant(inputs))
