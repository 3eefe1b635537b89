"""Bit types of the quantized operators (reference: models/ptq/bit_type.py:7-57)."""
from dataclasses import dataclass


@dataclass(frozen=True)
class BitType:
    bits: int
    signed: bool
    name: str = ""

    def __post_init__(self):
        if not self.name:
            object.__setattr__(self, "name", ("int" if self.signed else "uint") + str(self.bits))

    @property
    def upper_bound(self):
        return 2 ** (self.bits - 1) - 1 if self.signed else 2 ** self.bits - 1

    @property
    def lower_bound(self):
        return -(2 ** (self.bits - 1)) if self.signed else 0

    @property
    def range(self):
        return 2 ** self.bits


# same registry and order as the reference (the calibration loop iterates it, layers.py:178-188)
BIT_TYPE_LIST = [BitType(3, False), BitType(4, False), BitType(4, True), BitType(8, True), BitType(8, False)]
BIT_TYPE_DICT = {b.name: b for b in BIT_TYPE_LIST}
