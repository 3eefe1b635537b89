"""The quantized operator surface of P2-ViT - QConv2d, QLinear, QAct, QIntLayerNorm, QIntSoftmax - with the
reference's constructor signatures, flag protocol (.quant/.calibrate/.last_calibrate, .mode) and forward
arguments (reference: models/ptq/layers.py:14-447), on B200 kernels instead of ATen fake-quant chains.

Two execution levels:
  * these modules, called one by one (calibration, FP evaluation, per-operator use): fp32 in / fp32 out like
    the reference; in quant mode each forward is one or two launches of csrc kernels and the QAct output
    carries its int8 codes (attribute `_p2v_codes`) so the following QLinear/QConv2d runs the tcgen05 int8 GEMM;
  * the whole-model integer engine (p2vit_b200/engine.py) that the model classes switch to after
    `model_quant()`: same numbers, fused epilogues, int8 tensors end to end.
There is no CPU path: quant-mode forwards raise on non-CUDA tensors.
"""
import torch
import torch.nn as nn
from torch.nn import functional as F

from .. import intmath, ops
from .bit_type import BIT_TYPE_DICT, BIT_TYPE_LIST
from .observer import build_observer
from .observer.utils import lp_loss
from .quantizer import build_quantizer


def fp_linear(x, weight, bias):
    """The FP layer.  Calibration / FP evaluation on the GPU (no autograd): the library's own fp32 GEMM - deterministic and
    batch-split invariant, so a calibration sharded over N GPUs sees exactly the activations one GPU would.  With autograd
    (Hessian sensitivity pass) or on the CPU (host-logic tests): F.linear."""
    if x.is_cuda and not (torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad)) and weight.shape[-1] % 4 == 0:
        return ops.linear_f32(x, weight, bias)
    return F.linear(x, weight, bias)


def _attach_codes(y, codes, scale, zero_point):
    y._p2v_codes = (codes, scale, zero_point)
    return y


def _require_codes(x, who):
    c = getattr(x, "_p2v_codes", None)
    if c is None:
        raise RuntimeError(
            "%s in quant mode runs an int8 tensor-core GEMM and needs the integer codes of its input: feed it the "
            "direct output of a QAct (as every call site of the reference does, vit_fquant.py:342-346,389-397); "
            "there is no fp32 fallback path" % who)
    codes, scale, zp = c
    if scale.numel() != 1:
        raise NotImplementedError("%s: channel-wise input scales cannot be factored out of the GEMM" % who)
    return codes, scale, zp


class _QWeightMixin:
    """weight-side calibration loop and int8 packing shared by QLinear and QConv2d"""

    def _init_q(self, quant, calibrate, last_calibrate, bit_type, calibration_mode, observer_str, quantizer_str, module_type):
        self.quant = quant
        self.calibrate = calibrate
        self.last_calibrate = last_calibrate
        self.bit_type = bit_type
        self.calibration_mode = calibration_mode
        self.observer_str = observer_str
        self.quantizer_str = quantizer_str
        self.module_type = module_type
        self.observer = build_observer(observer_str, module_type, bit_type, calibration_mode)
        self.quantizer = build_quantizer(quantizer_str, bit_type, self.observer, module_type)

    def _calibrate_all_bit_types(self, weight, x, others, need_params=True, kwargs=None):
        """layers.py:62-85,175-201: every registered weight bit type gets its own PoT scale(s);
        int8 is layer-wise, the narrower types channel-wise."""
        distance = []
        for bit_type in BIT_TYPE_LIST:
            if bit_type == BIT_TYPE_DICT["uint8"]:
                continue
            self.quantizer.bit_type = bit_type
            self.observer.bit_type = bit_type
            self.observer.calibration_mode = "layer_wise" if bit_type == BIT_TYPE_DICT["int8"] else "channel_wise"
            self.quantizer.observer.update(weight)
            if need_params:
                self.quantizer.update_quantization_params(x, others=others, **(kwargs or {}))
                distance.append(lp_loss(weight, self.quantizer(weight), p=2.0, reduction="all"))
        return distance

    def _set_bits(self, bit_config):
        if bit_config:
            bt = BIT_TYPE_DICT["int" + str(bit_config)]
            self.quantizer.bit_type = bt
            self.observer.bit_type = bt

    def weight_codes(self, weight=None):
        """int8 codes [Cout, K] and fp32 scales [Cout] of `weight` (default: self.weight) at the current bit type."""
        w = self.weight if weight is None else weight
        name = self.quantizer.bit_type.name
        scale = self.quantizer.dic_scale[name].reshape(-1).float()
        zp = self.quantizer.dic_zero_point[name]
        if bool((zp != 0).any()):
            raise NotImplementedError("asymmetric weights are not produced by this path (OBSERVER_W is minmax/symmetric)")
        w2 = w.detach().float().reshape(1, w.shape[0], -1, 1)
        codes = ops.quantize(w2, scale, 0.0, self.quantizer.bit_type.lower_bound, self.quantizer.bit_type.upper_bound)
        return codes.reshape(w.shape[0], -1), scale.expand(w.shape[0]).contiguous() if scale.numel() == 1 else scale

    def _int8_forward(self, a_codes, a_scale, a_zp, weight):
        wq, ws = self.weight_codes(weight)
        M, K = a_codes.shape
        N = wq.shape[0]
        acc_scale = (a_scale.reshape(-1).float() * ws).contiguous()
        bias = None if self.bias is None else self.bias.detach().float().contiguous()
        zp_corr = None
        zp = float(a_zp.reshape(-1)[0]) if a_zp is not None else 0.0
        if zp != 0.0:
            zp_corr = (wq.to(torch.int32).sum(dim=1) * int(zp)).to(torch.int32).contiguous()
        out = torch.empty((M, N), dtype=torch.float32, device=a_codes.device)
        args = ops.gemm_args(a_codes, wq, ops.EPI_F32, acc_scale, bias=bias, out_f32=out, zp_corr=zp_corr)
        ops.gemm(args)
        return out


class QConv2d(nn.Conv2d, _QWeightMixin):
    """Patch-embedding convolution (kernel == stride, no padding) with quantized weights (layers.py:14-103)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 quant=False, calibrate=False, last_calibrate=False, bit_type=BIT_TYPE_DICT["int8"],
                 calibration_mode="layer_wise", observer_str="minmax", quantizer_str="uniform"):
        super().__init__(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size, stride=stride,
                         padding=padding, dilation=dilation, groups=groups, bias=bias)
        self._init_q(quant, calibrate, last_calibrate, bit_type, calibration_mode, observer_str, quantizer_str, "conv_weight")

    def _patch_rows(self, x):
        k = self.kernel_size[0]
        B, Cin, H, W = x.shape
        return x.reshape(B, Cin, H // k, k, W // k, k).permute(0, 2, 4, 1, 3, 5).reshape(-1, Cin * k * k)

    def _fp_conv(self, x, weight):
        # k == stride: the convolution is a GEMM over gathered patches; plain fp32 (no TF32 cuDNN path) so the
        # calibration statistics do not depend on cuDNN's algorithm choice
        k = self.kernel_size[0]
        assert self.kernel_size == self.stride and self.padding == (0, 0), "only the patch-embed form is supported"
        B, _, H, W = x.shape
        if x.is_cuda and k % 4 == 0 and not (torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad)):
            y = ops.linear_f32(x, weight, self.bias, patch=k)                 # patch gather inside the kernel
        else:
            y = F.linear(self._patch_rows(x), weight.reshape(weight.shape[0], -1), self.bias)
        return y.reshape(B, H // k, W // k, -1).permute(0, 3, 1, 2)

    def forward(self, x, bit_config):
        if self.calibrate:
            self._calibrate_all_bit_types(
                self.weight, x, [self.bias, self.stride, self.padding, self.dilation, self.groups], need_params=self.last_calibrate)
        if not self.quant:
            return self._fp_conv(x, self.weight)
        self._set_bits(bit_config)
        if getattr(x, "_p2v_codes", None) is None:
            # input_quant=False models (ViT-L, vit_fquant.py:1063): raw fp32 pixels meet fake-quantized weights, so
            # this convolution is an fp32 GEMM in the reference as well - not a fallback of the int8 path
            wq, ws = self.weight_codes(self.weight)
            return self._fp_conv(x, (wq.float() * ws.reshape(-1, 1)).reshape(self.weight.shape))
        codes, a_scale, a_zp = _require_codes(x, "QConv2d")
        k = self.kernel_size[0]
        B, Cin, H, W = x.shape
        rows = self._patch_rows(codes.reshape(B, Cin, H, W)).contiguous()
        y = self._int8_forward(rows, a_scale, a_zp, self.weight)
        return y.reshape(B, H // k, W // k, -1).permute(0, 3, 1, 2)


class QLinear(nn.Linear, _QWeightMixin):
    """Linear layer with quantized weights (layers.py:119-209)."""

    def __init__(self, in_features, out_features, bias=True, quant=False, calibrate=False, last_calibrate=False,
                 bit_type=BIT_TYPE_DICT["int8"], calibration_mode="layer_wise", observer_str="minmax", quantizer_str="uniform"):
        super().__init__(in_features, out_features, bias)
        self._init_q(quant, calibrate, last_calibrate, bit_type, calibration_mode, observer_str, quantizer_str, "linear_weight")

    def forward(self, x, global_distance=[], bit_config=None, weight_smoothed=None, attn=False, attn_para=None):
        if weight_smoothed is None:
            weight_smoothed = self.weight
        if not self.quant:
            y = fp_linear(x, weight_smoothed, self.bias)
        if self.calibrate:
            distance = self._calibrate_all_bit_types(weight_smoothed, x, [self.bias], kwargs=dict(attn=attn, attn_para=attn_para))
            global_distance.append(distance)
        if not self.quant:
            return y
        self._set_bits(bit_config)
        codes, a_scale, a_zp = _require_codes(x, "QLinear")
        y = self._int8_forward(codes.reshape(-1, codes.shape[-1]), a_scale, a_zp, weight_smoothed)
        return y.reshape(*x.shape[:-1], -1)


class QAct(nn.Module):
    """Activation quantizer (layers.py:212-257)."""

    def __init__(self, quant=False, calibrate=False, last_calibrate=False, bit_type=BIT_TYPE_DICT["int8"],
                 calibration_mode="layer_wise", observer_str="minmax", quantizer_str="uniform"):
        super().__init__()
        self.quant = quant
        self.calibrate = calibrate
        self.last_calibrate = last_calibrate
        self.bit_type = bit_type
        self.calibration_mode = calibration_mode
        self.observer_str = observer_str
        self.quantizer_str = quantizer_str
        self.module_type = "activation"
        self.observer = build_observer(observer_str, self.module_type, bit_type, calibration_mode)
        self.quantizer = build_quantizer(quantizer_str, bit_type, self.observer, self.module_type)

    def forward(self, x, asymmetric=False, attn=False, attn_para=None):
        if self.calibrate:
            if asymmetric:
                self.quantizer.bit_type = BIT_TYPE_DICT["uint8"]
                self.observer.bit_type = BIT_TYPE_DICT["uint8"]
                self.observer.symmetric = False
            self.quantizer.observer.update(x)
            if self.last_calibrate:
                self.quantizer.update_quantization_params(x, attn=attn, attn_para=attn_para)
        if not self.quant:
            return x
        q = self.quantizer
        bt = q.bit_type
        if bt.lower_bound >= -128 and bt.upper_bound <= 127:
            zp = q._zp_scalar(q.zero_point)
            y, codes = ops.fake_quant(x.float(), q.scale, zp, bt.lower_bound, bt.upper_bound, return_codes=True)
            return _attach_codes(y, codes, q.scale, q.zero_point)
        return q(x)


class QIntLayerNorm(nn.LayerNorm):
    """LayerNorm with the integer ('int') evaluation mode (layers.py:263-339)."""

    def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True):
        super().__init__(normalized_shape, eps, elementwise_affine)
        assert isinstance(normalized_shape, int)
        self.mode = "ln"

    def get_MN(self, x):
        bit = 7
        N = torch.clamp(bit - torch.floor(torch.log2(x)), 0, 31)
        M = torch.clamp(torch.floor(x * torch.pow(2, N)), 0, 2 ** (bit + 1) - 1)
        return M, N

    def forward(self, x, in_quantizer=None, out_quantizer=None, out_quantizer_scale=None, in_scale_expand=1):
        if self.mode == "ln":
            return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        if self.mode != "int":
            raise NotImplementedError
        in_scale = in_quantizer.scale
        if in_scale_expand != 1:
            in_scale = in_scale.unsqueeze(-1).expand(-1, in_scale_expand).T.reshape(-1)
        out_scale_global = out_quantizer.scale
        assert in_scale is not None and out_scale_global is not None
        C = x.shape[-1]
        dev = x.device
        in_scale = in_scale.reshape(-1).float().to(dev)
        if in_scale.numel() == 1:
            in_scale = in_scale.expand(C)
        out_scale = out_scale_global if out_quantizer_scale is None else out_scale_global * out_quantizer_scale
        out_scale = out_scale.reshape(-1).float().to(dev)
        if out_scale.numel() == 1:
            out_scale = out_scale.expand(C)
        out_scale = out_scale.contiguous()
        s1 = in_scale.min()
        in_mult = (in_scale / s1).round().contiguous()
        codes = ops.quantize(x.float().contiguous(), in_scale.contiguous(), 0.0, -128, 127)  # x_q = round(x / in_scale)
        rows = codes.numel() // C
        y = torch.empty(x.shape, dtype=torch.float32, device=dev)
        ones = torch.ones(C, dtype=torch.float32, device=dev)
        args = ops.layernorm_args(codes, rows, C, C, in_mult, float(s1), self.weight.detach().float().contiguous(),
                                  self.bias.detach().float().contiguous(), out_scale, ones, 1.0,
                                  intmath.is_pot(out_scale), out_f32=y)
        ops.layernorm(args)
        return y


class QIntSoftmax(nn.Module):
    """Integer log2 softmax (layers.py:343-447)."""

    def __init__(self, log_i_softmax=False, quant=False, calibrate=False, last_calibrate=False,
                 bit_type=BIT_TYPE_DICT["int8"], calibration_mode="layer_wise", observer_str="minmax", quantizer_str="uniform"):
        super().__init__()
        self.log_i_softmax = log_i_softmax
        self.quant = quant
        self.calibrate = calibrate
        self.last_calibrate = last_calibrate
        self.bit_type = bit_type
        self.calibration_mode = calibration_mode
        self.observer_str = observer_str
        self.quantizer_str = quantizer_str
        self.module_type = "activation"
        self.observer = build_observer(observer_str, self.module_type, bit_type, calibration_mode)
        self.quantizer = build_quantizer(quantizer_str, bit_type, self.observer, self.module_type)

    @staticmethod
    def log_round(x):
        big = x.log2().floor()
        extra_mask = (x - 2 ** big) >= 2 ** (big - 1)
        big[extra_mask] = big[extra_mask] + 1
        return big

    @staticmethod
    def int_softmax(x, scaling_factor):
        """tensor-op form, used on un-quantized scores during the calibration forward (x/scale is not an integer
        there, so the code table of the kernels does not apply)."""
        n = 32
        coef = [0.35815147, 0.96963238, 1.0]
        coef[1] /= coef[0]
        coef[2] /= coef[0]
        x_int = x / scaling_factor
        x_int = x_int - x_int.max(dim=-1, keepdim=True).values
        x0_int = torch.floor(-0.6931 / scaling_factor)
        x_int = torch.max(x_int, n * x0_int)
        q = torch.floor(x_int / x0_int)
        r = x_int - x0_int * q
        z = r * (r + torch.floor(coef[1] / scaling_factor)) + torch.floor(coef[2] / scaling_factor ** 2)
        exp_int = torch.clamp(torch.floor(z * 2 ** (n - q)), min=0)
        return exp_int, exp_int.sum(dim=-1, keepdim=True)

    def forward(self, x, scale):
        if self.log_i_softmax and scale is not None:
            bits = self.bit_type.bits
            codes = getattr(x, "_p2v_codes", None)
            if self.quant and codes is not None and bits == 4:
                lut = intmath.lut_to_device(intmath.build_softmax_lut(scale), x.device)
                c = ops.int_softmax_log2(codes[0].reshape(x.shape), lut)
                out = torch.pow(2.0, -c.float())
                out[c == 255] = 0
                return out
            exp_int, exp_int_sum = self.int_softmax(x, scale)
            rounds = self.log_round(torch.round(exp_int_sum / exp_int))
            mask = rounds >= 2 ** bits
            out = 2 ** (-torch.clamp(rounds, 0, 2 ** bits - 1))
            out[mask] = 0
            return out
        return x.softmax(dim=-1)
