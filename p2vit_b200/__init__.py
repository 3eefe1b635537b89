"""p2vit_b200 - B200-native (sm_100a) implementation of P2-ViT's fully quantized ViT/DeiT inference path.

Public surface = the reference's (`from models import *`, `from config import Config`):
    Config, QAct, QConv2d, QLinear, QIntLayerNorm, QIntSoftmax, BIT_TYPE_DICT,
    deit_{tiny,small,base}_patch16_224, vit_{base,large}_patch16_224, swin_{tiny,small,base}_patch4_window7_224
plus `calibrate_model` / `validate` (the calibrate -> quant -> validate flow of test_quant.py) and
`build_model` / `synth` helpers for seeded synthetic weights and images.
"""
from .config import Config  # noqa: F401
from .ptq import BIT_TYPE_DICT, BIT_TYPE_LIST, QAct, QConv2d, QIntLayerNorm, QIntSoftmax, QLinear  # noqa: F401
from .vit import (VisionTransformer, deit_base_patch16_224, deit_small_patch16_224, deit_tiny_patch16_224,  # noqa: F401
                  vit_base_patch16_224, vit_large_patch16_224)
from .swin import (SwinTransformer, swin_base_patch4_window7_224, swin_small_patch4_window7_224,  # noqa: F401
                   swin_tiny_patch4_window7_224)
from .runner import build_model, calibrate_model, str2model, validate  # noqa: F401
from . import checkpoint, data, hessian, search, synth  # noqa: F401
from .checkpoint import load_checkpoint, load_weights_from_npz  # noqa: F401

__version__ = "0.1.0"
