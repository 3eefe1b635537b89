"""The calibrate -> quant -> validate flow of the reference driver (test_quant.py:262-312, 464-527) as library
functions, data-parallel over the image batch (one process per GPU, torch.distributed/NCCL).

Inference shards by batch with no collective in the forward (every op is per image once the scales are frozen);
`validate` all-reduces [top1, top5, count, loss_sum] once at the end.  Calibration all-reduces the observers'
statistics (ptq/observer/*: MAX/MIN of ranges, SUM of candidate scores), so every rank freezes identical scales.
"""
import time
from functools import partial

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import synth
from .config import Config
from .ptq import QIntLayerNorm
from .swin import SwinTransformer, swin_base_patch4_window7_224, swin_small_patch4_window7_224, swin_tiny_patch4_window7_224
from .vit import (VisionTransformer, deit_base_patch16_224, deit_small_patch16_224, deit_tiny_patch16_224,
                  vit_base_patch16_224, vit_large_patch16_224)


def str2model(name):
    """test_quant.py:69-81"""
    return {
        "deit_tiny": deit_tiny_patch16_224,
        "deit_small": deit_small_patch16_224,
        "deit_base": deit_base_patch16_224,
        "vit_base": vit_base_patch16_224,
        "vit_large": vit_large_patch16_224,
        "swin_tiny": swin_tiny_patch4_window7_224,
        "swin_small": swin_small_patch4_window7_224,
        "swin_base": swin_base_patch4_window7_224,
    }[name]


def build_model(name, cfg=None, seed=0, device="cuda"):
    """Model `name` (a factory name or the test-only 'vit_micro') with seeded synthetic weights on `device`."""
    cfg = cfg or Config()
    if name in synth.SWIN_CONFIGS:
        c = synth.SWIN_CONFIGS[name]
        if name == "swin_micro":
            model = SwinTransformer(patch_size=4, window_size=7, embed_dim=c["embed_dim"], depths=c["depths"], num_heads=c["num_heads"],
                                    norm_layer=QIntLayerNorm, input_quant=True, cfg=cfg)
        else:
            model = str2model(name)(pretrained=False, cfg=cfg)
        sd = synth.synth_swin_state_dict(**c, seed=seed)
        res = model.load_state_dict({k: v for k, v in sd.items() if not k.endswith("reduction.bias")}, strict=False)
        assert not res.unexpected_keys and all("relative_position_index" in k or "attn_mask" in k for k in res.missing_keys), res
        return model.to(device).eval()
    c = synth.VIT_CONFIGS[name]
    if name == "vit_micro":
        model = VisionTransformer(patch_size=16, embed_dim=c["embed_dim"], depth=c["depth"], num_heads=c["num_heads"], mlp_ratio=4,
                                  qkv_bias=True, norm_layer=partial(QIntLayerNorm, eps=1e-6), input_quant=c["input_quant"], cfg=cfg)
    else:
        model = str2model(name)(pretrained=False, cfg=cfg)
    res = model.load_state_dict(synth.synth_vit_state_dict(**c, seed=seed), strict=False)
    assert not res.missing_keys and not res.unexpected_keys, res
    return model.to(device).eval()


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, world_size):
    """contiguous [start, stop) slice of n items for `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


@torch.no_grad()
def calibrate_model(model, images):
    """One calibration forward with last_calibrate already on (test_quant.py:275-281,306-307), then freeze:
    model_close_calibrate() + model_quant().  `images`: this rank's shard of the calibration batch (CUDA)."""
    t0 = time.time()
    model.model_open_calibrate()
    model.model_open_last_calibrate()
    out = model(images)
    model.model_close_calibrate()
    model.model_quant()
    if images.is_cuda:
        torch.cuda.synchronize()
    return out, time.time() - t0


def accuracy(output, target, topk=(1,)):
    """test_quant.py:549-562, returning correct counts (not percentages) so they can be summed over ranks"""
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.reshape(1, -1).expand_as(pred.t()))
    return [correct[:k].reshape(-1).float().sum() for k in topk]


@torch.no_grad()
def validate(model, batches, bit_config, device="cuda"):
    """batches: iterable of (images, targets) for THIS rank.  Returns (loss_avg, top1 %, top5 %, images, seconds)."""
    model.eval()
    stats = torch.zeros(4, dtype=torch.float64, device=device)  # top1, top5, count, loss_sum
    t0 = time.time()
    for data, target in batches:
        data, target = data.to(device, non_blocking=True), target.to(device, non_blocking=True)
        output = model(data, bit_config, False)[0]
        c1, c5 = accuracy(output, target, (1, 5))
        stats += torch.stack([c1.double(), c5.double(), torch.tensor(float(data.shape[0]), device=device, dtype=torch.float64),
                              F.cross_entropy(output, target, reduction="sum").double()])
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(stats)
    torch.cuda.synchronize() if torch.cuda.is_available() else None
    n = max(float(stats[2]), 1.0)
    return float(stats[3]) / n, 100.0 * float(stats[0]) / n, 100.0 * float(stats[1]) / n, int(stats[2]), time.time() - t0
