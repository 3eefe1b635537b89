"""Real-data plumbing of the reference's driver: per-family preprocessing constants, `build_transform`
and the ImageFolder loaders, sharded by rank for the data-parallel validation (`runner.validate`).

Reference: test_quant.py:112-157 (mean / std / crop_pct per model family, ImageFolder + DataLoader), :565-597 (build_transform).
The synthetic-data path (`synth.synth_images`) is what the tests and bench.py use; this module is for a real ImageNet tree.
"""
import math
import os

import torch

PREPROCESS = {      # test_quant.py:113-127
    "deit": dict(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), crop_pct=0.875),
    "vit": dict(mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), crop_pct=0.9),
    "swin": dict(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), crop_pct=0.9),
}


def preprocess_for(model_name):
    """constants for a factory name such as 'deit_small' / 'vit_base' / 'swin_tiny' (test_quant.py:112: the prefix decides)"""
    family = model_name.split("_")[0]
    if family not in PREPROCESS:
        raise NotImplementedError(model_name)
    return PREPROCESS[family]


def build_transform(input_size=224, interpolation="bicubic", mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), crop_pct=0.875):
    """test_quant.py:565-597: Resize(floor(input_size / crop_pct)) -> CenterCrop(input_size) -> ToTensor -> Normalize"""
    from PIL import Image
    from torchvision import transforms
    interp = {"bicubic": Image.BICUBIC, "lanczos": Image.LANCZOS, "hamming": Image.HAMMING}.get(interpolation, Image.BILINEAR)
    t = []
    if input_size > 32:
        t.append(transforms.Resize(int(math.floor(input_size / crop_pct)), interpolation=interp))
        t.append(transforms.CenterCrop(input_size))
    t.append(transforms.ToTensor())
    t.append(transforms.Normalize(mean, std))
    return transforms.Compose(t)


class _Shard(torch.utils.data.Dataset):
    """contiguous slice of a dataset (runner.shard_range): every rank validates its own part, the counts are all-reduced"""

    def __init__(self, base, start, stop):
        self.base, self.start, self.stop = base, start, stop

    def __len__(self):
        return self.stop - self.start

    def __getitem__(self, i):
        return self.base[self.start + i]


def build_loaders(data_root, model_name, calib_batchsize=32, val_batchsize=256, num_workers=8, rank=0, world_size=1, seed=0):
    """(train_loader, val_loader) as test_quant.py:133-157 builds them: ImageFolder(<root>/train | val) with the family's
    transform, shuffled drop-last calibration batches, ordered validation batches, pinned memory.  With world_size > 1 the
    validation set is cut into contiguous per-rank shards; the calibration loader is seeded identically on every rank and
    each rank takes its slice of the batch (runner.calibrate_model + the observers' all-reduce)."""
    from torchvision import datasets
    from . import runner
    tf = build_transform(**preprocess_for(model_name))
    val = datasets.ImageFolder(os.path.join(data_root, "val"), tf)
    if world_size > 1:
        val = _Shard(val, *runner.shard_range(len(val), rank, world_size))
    val_loader = torch.utils.data.DataLoader(val, batch_size=val_batchsize, shuffle=False, num_workers=num_workers, pin_memory=True)
    train = datasets.ImageFolder(os.path.join(data_root, "train"), tf)
    g = torch.Generator()
    g.manual_seed(seed)
    train_loader = torch.utils.data.DataLoader(train, batch_size=calib_batchsize, shuffle=True, num_workers=num_workers, pin_memory=True,
                                               drop_last=True, generator=g)
    return train_loader, val_loader
