"""Observer base (reference: models/ptq/observer/base.py:5-36)."""
import torch


class BaseObserver:
    def __init__(self, module_type, bit_type, calibration_mode):
        self.module_type = module_type
        self.bit_type = bit_type
        self.calibration_mode = calibration_mode
        self.max_val = None
        self.min_val = None
        self.eps = torch.finfo(torch.float32).eps

    def reshape_tensor(self, v):
        """[C, everything-else] view with the reference's channel rule."""
        if not isinstance(v, torch.Tensor):
            v = torch.tensor(v)
        v = v.detach()
        if self.module_type in ("conv_weight", "linear_weight"):
            return v.reshape(v.shape[0], -1)
        if self.module_type == "activation":
            if v.dim() == 4:
                v = v.permute(0, 2, 3, 1)
            return v.reshape(-1, v.shape[-1]).transpose(0, 1)
        raise NotImplementedError

    def channel_minmax(self, v):
        """per-channel (min, max) through the block-reduce kernel (csrc/rowops.cu: minmax_partial_kernel)."""
        from ... import ops

        v = v.detach()
        if self.module_type in ("conv_weight", "linear_weight"):
            mm = ops.minmax_per_channel(v.reshape(1, v.shape[0], -1, 1).float())  # [outer=1, C=Cout, inner=Cin*k*k]
        elif self.module_type == "activation":
            mm = ops.minmax_per_channel(v.float())
        else:
            raise NotImplementedError
        return mm[0], mm[1]

    def _running_range(self, v, combine_max, combine_min):
        cur_min, cur_max = self.channel_minmax(v)
        self.max_val = cur_max if self.max_val is None else combine_max(cur_max, self.max_val)
        self.min_val = cur_min if self.min_val is None else combine_min(cur_min, self.min_val)
        if self.calibration_mode == "layer_wise":
            self.max_val = self.max_val.max()
            self.min_val = self.min_val.min()

    def update(self, v):
        raise NotImplementedError

    def get_quantization_params(self, *args, **kwargs):
        raise NotImplementedError

    # ---- multi-GPU calibration: statistics are combined over ranks with NCCL
    def allreduce_range(self):
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and self.max_val is not None:
            dist.all_reduce(self.max_val, op=dist.ReduceOp.MAX)
            dist.all_reduce(self.min_val, op=dist.ReduceOp.MIN)
