"""Log2 quantizer (reference: models/ptq/quantizer/log2.py:7-26).  Constructed for QIntSoftmax by Config
(QUANTIZER_S = "log2") but its quant() is never called on the active path (layers.py:446 is commented out):
the log2 codes come from QIntSoftmax.forward itself.  Kept for API parity; pure tensor ops."""
import torch

from .base import BaseQuantizer


class Log2Quantizer(BaseQuantizer):
    def __init__(self, bit_type, observer, module_type):
        super().__init__(bit_type, observer, module_type)
        self.softmax_mask = None

    def quant(self, inputs):
        rounds = torch.round(-1 * inputs.log2())
        self.softmax_mask = rounds >= 2 ** self.bit_type.bits
        return torch.clamp(rounds, 0, 2 ** self.bit_type.bits - 1)

    def dequantize(self, inputs):
        outputs = 2 ** (-1 * inputs)
        outputs[self.softmax_mask] = 0
        return outputs
