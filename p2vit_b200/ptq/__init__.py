"""Drop-in surface of the reference's `models.ptq` package (models/ptq/__init__.py:2-3)."""
from .bit_type import BIT_TYPE_DICT, BIT_TYPE_LIST, BitType  # noqa: F401
from .layers import QAct, QConv2d, QIntLayerNorm, QIntSoftmax, QLinear  # noqa: F401
