"""Helpers shared by the observers."""
import torch


def lp_loss(pred, tgt, p=2.0, reduction="none"):
    """L_p distance (reference: models/ptq/observer/utils.py:2-9)."""
    d = (pred - tgt).abs().pow(p)
    return d.sum(1).mean() if reduction == "none" else d.mean()


def pot_exponent(x, mode=None):
    """Power-of-two exponent of x with the reference's fp32 formula floor(log(x)/log(2)) and its
    linear-distance nearest rule (observer/minmax.py:50-64) - NOT log2f, so exact powers of
    two land on the same side as in the reference."""
    ln2 = torch.log(torch.tensor([2.0], device=x.device))
    y = torch.div(torch.log(x), ln2)
    if mode == "ceil":
        return torch.ceil(y)
    y = torch.floor(y)
    if mode == "floor":
        return y
    return torch.gt(x - 2 ** y, 2 ** (y + 1) - x) + y


def allreduce_(t, op="sum"):
    """In-place all-reduce of a calibration statistic over the data-parallel ranks (NCCL on GPUs, gloo in
    the CPU tests); no-op in a single process.  Messages are tiny (<= tens of KB): latency bound."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[op])
    return t
