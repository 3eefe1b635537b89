"""Tensor-level wrappers over the C ABI (include/p2vit_b200.h).  Every function enqueues on
torch's current CUDA stream and allocates outputs with torch; inputs must be CUDA tensors."""
import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_DEQUANT, EPI_EMBED, EPI_F32, EPI_GELU, EPI_REQUANT, EPI_RESIDUAL,  # noqa: F401
                   AttentionArgs, GemmArgs, LayerNormArgs, WindowAttentionArgs, check, ptr, stream)


def _scale_vec(scale, device):
    s = scale if isinstance(scale, torch.Tensor) else torch.tensor([float(scale)])
    return s.detach().reshape(-1).to(device=device, dtype=torch.float32).contiguous()


def _channel_geometry(x):
    """(C, inner) of the reference's activation channel rule (quantizer/base.py:14-31)."""
    if x.dim() == 4:
        return x.shape[1], x.shape[2] * x.shape[3]
    return x.shape[-1], 1


def quantize(x, scale, zero_point=0.0, lo=-128, hi=127):
    """QAct codes: clamp(RNE(x/scale + zp), lo, hi) as int8."""
    x = x.contiguous()
    s = _scale_vec(scale, x.device)
    Cn, inner = _channel_geometry(x)
    q = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    check(_lib.load().p2v_quantize_f32(ptr(x), ptr(q), x.numel(), Cn, inner, ptr(s), s.numel(), float(zero_point), lo, hi, stream()),
          "quantize_f32")
    return q


def fake_quant(x, scale, zero_point=0.0, lo=-128, hi=127, return_codes=False):
    """BaseQuantizer.forward: dequantize(quant(x)) in one pass (quantizer/base.py:42-45)."""
    x = x.contiguous()
    s = _scale_vec(scale, x.device)
    Cn, inner = _channel_geometry(x)
    y = torch.empty_like(x)
    q = torch.empty(x.shape, dtype=torch.int8, device=x.device) if return_codes else None
    check(_lib.load().p2v_fake_quant_f32(ptr(x), ptr(y), ptr(q), x.numel(), Cn, inner, ptr(s), s.numel(), float(zero_point), lo, hi,
                                         stream()), "fake_quant_f32")
    return (y, q) if return_codes else y


def dequantize(q, scale, zero_point=0.0):
    q = q.contiguous()
    s = _scale_vec(scale, q.device)
    Cn, inner = _channel_geometry(q)
    y = torch.empty(q.shape, dtype=torch.float32, device=q.device)
    check(_lib.load().p2v_dequantize_i8(ptr(q), ptr(y), q.numel(), Cn, inner, ptr(s), s.numel(), float(zero_point), stream()),
          "dequantize_i8")
    return y


def quantize_patchify(img, patch, scale, zero_point=0.0, lo=-128, hi=127, out=None):
    img = img.contiguous()
    B, Cin, H, W = img.shape
    rows, K = B * (H // patch) * (W // patch), Cin * patch * patch
    if out is None:
        out = torch.empty((rows, K), dtype=torch.int8, device=img.device)
    check(_lib.load().p2v_quantize_patchify(ptr(img), ptr(out), B, Cin, H, W, patch, float(scale), float(zero_point), lo, hi, stream()),
          "quantize_patchify")
    return out


def patchify_u8(img_u8, lut, patch, out=None):
    """uint8 pixels [B,Cin,H,W] + per-channel code table [Cin,256] -> int8 patch rows (same layout as quantize_patchify)."""
    if img_u8.dtype != torch.uint8 or lut.dtype != torch.int8 or not img_u8.is_contiguous():
        raise ValueError("patchify_u8: contiguous uint8 image and int8 table expected")
    B, Cin, H, W = img_u8.shape
    if tuple(lut.shape) != (Cin, 256):
        raise ValueError("patchify_u8: table must be [Cin, 256]")
    rows, K = B * (H // patch) * (W // patch), Cin * patch * patch
    if out is None:
        out = torch.empty((rows, K), dtype=torch.int8, device=img_u8.device)
    check(_lib.load().p2v_patchify_u8_lut(ptr(img_u8), ptr(lut), ptr(out), B, Cin, H, W, patch, stream()), "patchify_u8_lut")
    return out


def pixel_code_table(mean, std, patch_scale, device, zero_point=0.0):
    """[Cin,256] int8: qact_input code of ToTensor + Normalize of every byte value (x/255, then (x - mean)/std in fp32, the order
    torchvision applies; reference test_quant.py:565-597), produced by the fp32 quantizer kernel itself."""
    mean = torch.as_tensor(mean, dtype=torch.float32, device=device).reshape(-1, 1)
    std = torch.as_tensor(std, dtype=torch.float32, device=device).reshape(-1, 1)
    v = torch.arange(256, dtype=torch.float32, device=device).div(255).reshape(1, 256)
    x = v.sub(mean).div(std)                                     # [Cin, 256]
    Cin = x.shape[0]
    return quantize_patchify(x.reshape(1, Cin, 16, 16).contiguous(), 16, patch_scale, zero_point).reshape(Cin, 256).contiguous()


def gemm_args(A, W, epilogue, acc_scale, bias=None, out_scale=None, mid_scale=None, res_scale=None, res=None, pos=None,
              aux_scale=0.0, tokens_per_image=0, out_i8=None, out_f32=None, pot=False, zp_corr=None, row_map=None, gelu_table=None,
              out_zp=0.0, mid_zp=0.0, aux_zp=0.0):
    M, K = A.shape
    N = W.shape[0]
    assert W.shape[1] == K
    a = GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.A, a.W = ptr(A), ptr(W)
    a.epilogue = epilogue
    a.acc_scale, a.bias, a.zp_corr = ptr(acc_scale), ptr(bias), ptr(zp_corr)
    a.out_scale, a.mid_scale, a.res_scale = ptr(out_scale), ptr(mid_scale), ptr(res_scale)
    a.res, a.pos = ptr(res), ptr(pos)
    a.aux_scale, a.tokens_per_image = float(aux_scale), int(tokens_per_image)
    a.out_i8, a.out_f32 = ptr(out_i8), ptr(out_f32)
    a.row_map = ptr(row_map)
    a.gelu_table = ptr(gelu_table)
    a.pot_scales = 1 if pot else 0
    a.out_zp, a.mid_zp, a.aux_zp = float(out_zp), float(mid_zp), float(aux_zp)
    if (a.out_zp or a.mid_zp or a.aux_zp) and pot:
        raise ValueError("gemm: zero points go with the general (non power-of-two) epilogues")
    return a


_gelu_tables = {}


def gelu_table(out_scale, device, zp=0.0):
    """device-resident step table of y -> qact(gelu(y)) for an output quantizer (scale, zero point); None if it is not tabulable.
    A scale that is not a power of two or a zero point gets the CTA-pair kernel's form only; cached per (scale, zp, device)"""
    key = (float(out_scale), float(zp), str(device))
    if key not in _gelu_tables:
        t = torch.empty(_lib.GELU_TABLE_BYTES, dtype=torch.uint8, device=device)
        if zp:
            rc = _lib.load().p2v_build_gelu_table_zp(float(out_scale), float(zp), ptr(t), stream())
        else:
            rc = _lib.load().p2v_build_gelu_table(float(out_scale), ptr(t), stream())
        if rc not in (0, 3):
            check(rc, "build_gelu_table")
        _gelu_tables[key] = t if rc == 0 else None
    return _gelu_tables[key]


def set_gemm_variant(variant):
    """0 = automatic, 1 = csrc/gemm_tc.cu always, 2 = csrc/gemm_pair.cu whenever it applies (tests cross-check the two)"""
    _lib.load().p2v_set_gemm_variant(int(variant))


def gemm(args, simt=False):
    lib = _lib.load()
    fn = lib.p2v_gemm_i8_simt if simt else lib.p2v_gemm_i8
    check(fn(C.byref(args), stream()), "gemm_i8_simt" if simt else "gemm_i8")


def fill_cls_rows(out, cls_row, B, T, N):
    check(_lib.load().p2v_fill_cls_rows(ptr(out), ptr(cls_row), B, T, N, stream()), "fill_cls_rows")


def layernorm_args(x, rows, Cn, row_stride, in_mult, in_scale_min, gamma, beta, out_scale, post_div, next_scale, pot,
                   out_i8=None, out_f32=None, out_row_map=None, clamp_mid=False, next_zp=0.0, in_gather=None, gather_segs=0):
    a = LayerNormArgs()
    a.rows, a.C = rows, Cn
    a.x, a.x_row_stride = ptr(x), row_stride
    a.in_mult, a.in_scale_min = ptr(in_mult), float(in_scale_min)
    a.gamma, a.beta = ptr(gamma), ptr(beta)
    a.out_scale, a.post_div = ptr(out_scale), ptr(post_div)
    a.next_scale, a.pot_scales = float(next_scale), 1 if pot else 0
    a.out_i8, a.out_f32 = ptr(out_i8), ptr(out_f32)
    a.out_row_map, a.clamp_mid = ptr(out_row_map), 1 if clamp_mid else 0
    a.next_zp = float(next_zp)
    a.in_gather, a.gather_segs = ptr(in_gather), int(gather_segs)
    if a.next_zp and pot:
        raise ValueError("layernorm: a zero point goes with the general (non power-of-two) kernel")
    return a


def layernorm(args):
    check(_lib.load().p2v_layernorm_int(C.byref(args), stream()), "layernorm_int")


def int_softmax_log2(scores_i8, lut_dev):
    scores_i8 = scores_i8.contiguous()
    n = scores_i8.shape[-1]
    rows = scores_i8.numel() // n
    out = torch.empty(scores_i8.shape, dtype=torch.uint8, device=scores_i8.device)
    check(_lib.load().p2v_int_softmax_log2(ptr(scores_i8), ptr(out), rows, n, ptr(lut_dev), stream()), "int_softmax_log2")
    return out


def attention_args(qkv, out, B, T, H, dh, score_mult, out_mult, lut_dev, probs=None, scores=None, zp_qkv=0, zp_score=0.0, zp_out=0.0,
                   prob_mode=0):
    a = AttentionArgs()
    a.B, a.T, a.H, a.dh = B, T, H, dh
    a.qkv, a.out = ptr(qkv), ptr(out)
    a.score_mult, a.out_mult = float(score_mult), float(out_mult)
    a.lut_dev = ptr(lut_dev)
    a.probs_or_null, a.scores_or_null = ptr(probs), ptr(scores)
    a.zp_qkv, a.zp_score, a.zp_out = int(zp_qkv), float(zp_score), float(zp_out)
    a.prob_mode = int(prob_mode)
    return a


def attention(args, simt=False):
    lib = _lib.load()
    fn = lib.p2v_attention_i8_simt if simt else lib.p2v_attention_i8
    check(fn(C.byref(args), stream()), "attention_i8_simt" if simt else "attention_i8")


BIAS_PITCH = 80      # row pitch of WindowAttentionArgs.bias_codes (include/p2vit_b200.h)


def window_bias_codes(codes):
    """int8 [H, T, T] qact_table codes gathered through relative_position_index -> the padded [H, T, 80] layout of the ABI"""
    H, T, _ = codes.shape
    out = torch.zeros((H, T, BIAS_PITCH), dtype=torch.int8, device=codes.device)
    out[:, :, :T] = codes
    return out.contiguous()


def window_mask_bits(labels):
    """SW-MSA region labels int8 [windows_per_image, T] -> int64 [windows_per_image, T]: bit j of word (w, i) = [label_i != label_j]"""
    lab = labels.long()
    ne = (lab.unsqueeze(2) != lab.unsqueeze(1)).long()                      # [wpi, T(i), T(j)]
    sh = torch.arange(lab.shape[1], device=lab.device, dtype=torch.int64)
    return (ne << sh.reshape(1, 1, -1)).sum(dim=2).contiguous()              # T <= 63 bits set: no sign issue below bit 63


def window_attention_args(qkv, out, n_windows, T, H, dh, windows_per_image, score_mult, s_attn1, s_attn2, bias, labels, mask_code,
                          mask_exp_int, out_mult, lut_dev, out_row_map=None, bias_codes=None, bias_scale=0.0, mask_bits=None):
    a = WindowAttentionArgs()
    a.n_windows, a.T, a.H, a.dh, a.windows_per_image = n_windows, T, H, dh, windows_per_image
    a.qkv, a.out = ptr(qkv), ptr(out)
    a.score_mult, a.s_attn1, a.s_attn2 = float(score_mult), float(s_attn1), float(s_attn2)
    a.bias, a.labels, a.mask_code, a.mask_exp_int = ptr(bias), ptr(labels), int(mask_code), int(mask_exp_int)
    a.out_mult, a.lut_dev, a.out_row_map = float(out_mult), ptr(lut_dev), ptr(out_row_map)
    a.bias_codes, a.bias_scale, a.mask_bits = ptr(bias_codes), float(bias_scale), ptr(mask_bits)
    return a


def window_attention(args, simt=False):
    lib = _lib.load()
    fn = lib.p2v_window_attention_i8_simt if simt else lib.p2v_window_attention_i8
    check(fn(C.byref(args), stream()), "window_attention_i8")


def gather_rows(x, out, src_rows, rows_out, segs, Cn):
    check(_lib.load().p2v_gather_rows_i8(ptr(x), ptr(out), ptr(src_rows), rows_out, segs, Cn, stream()), "gather_rows_i8")


def avgpool_quant(x, out, B, T, Cn, s_in, s_out):
    check(_lib.load().p2v_avgpool_quant_i8(ptr(x), ptr(out), B, T, Cn, float(s_in), float(s_out), stream()), "avgpool_quant_i8")


def minmax_per_channel(x):
    """[2,C] tensor: row 0 per-channel min, row 1 per-channel max (observer/base.py:16-29 layout rule)."""
    x = x.detach().contiguous().float()
    Cn, inner = _channel_geometry(x)
    lib = _lib.load()
    out = torch.empty((2, Cn), dtype=torch.float32, device=x.device)
    scratch = torch.empty(max(1, lib.p2v_minmax_scratch_bytes(x.numel(), Cn, inner) // 4), dtype=torch.float32, device=x.device)
    check(lib.p2v_minmax_per_channel(ptr(x), ptr(out), x.numel(), Cn, inner, ptr(scratch), stream()), "minmax_per_channel")
    return out


def quant_mse_scores(x, scales, lo, hi, zero_points=None, per_channel_out=False):
    """sum((x - fq_k(x))^2) for K candidate scales [K, 1 or C] -> float64 [K, 1 or C] (bit-reproducible: fixed reduction order)."""
    x = x.detach().contiguous().float()
    Cn, inner = _channel_geometry(x)
    scales = scales.detach().to(device=x.device, dtype=torch.float32).contiguous()
    K, n_scale = scales.shape
    zps = None if zero_points is None else zero_points.detach().to(device=x.device, dtype=torch.float32).contiguous()
    lib = _lib.load()
    pco = 1 if per_channel_out else 0
    out = torch.empty((K, Cn if per_channel_out else 1), dtype=torch.float64, device=x.device)
    scratch = torch.empty(max(1, lib.p2v_quant_mse_scratch_bytes(x.numel(), Cn, inner, K, pco) // 8), dtype=torch.float64, device=x.device)
    check(lib.p2v_quant_mse_scores(ptr(x), x.numel(), Cn, inner, ptr(scales), ptr(zps), K, n_scale, pco, lo, hi, ptr(out), ptr(scratch),
                                   stream()), "quant_mse_scores")
    return out


def linear_sqerr_scores(x, D, patch=0):
    """double [n]: sum over the rows of x of (x . D[j, :])^2 for every row j of D (fp32 [n, K]).  x: fp32 [M, K], or with patch > 0
    an NCHW image whose k = stride = patch patches are the rows.  (csrc/sgemm.cu; observer/minmax.py:145-207)"""
    x = x.detach().contiguous().float()
    D = D.detach().contiguous().float()
    n, K = D.shape
    if patch:
        B, Cin, H, W = x.shape
        M = B * (H // patch) * (W // patch)
    else:
        Cin = H = W = 0
        M = x.numel() // K
    lib = _lib.load()
    out = torch.empty(n, dtype=torch.float64, device=x.device)
    scratch = torch.empty(max(1, lib.p2v_linear_sqerr_scratch_bytes(M, n) // 8), dtype=torch.float64, device=x.device)
    check(lib.p2v_linear_sqerr_scores(ptr(x), M, K, int(patch), Cin, H, W, ptr(D), n, ptr(out), ptr(scratch), stream()), "linear_sqerr_scores")
    return out


def linear_f32(x, weight, bias=None, patch=0):
    """x W^T + bias in fp32 on the library's own GEMM (csrc/sgemm.cu): the FP forward of QLinear / QConv2d during calibration.
    Batch-split invariant (one ascending-k FMA chain per output).  x: [..., K], or with patch > 0 an NCHW image -> [B*T, N]."""
    x = x.detach().contiguous().float()
    w = weight.detach().float().reshape(weight.shape[0], -1).contiguous()
    N, K = w.shape
    if patch:
        B, Cin, H, W = x.shape
        M = B * (H // patch) * (W // patch)
        shape = (M, N)
    else:
        Cin = H = W = 0
        M = x.numel() // K
        shape = tuple(x.shape[:-1]) + (N,)
    b = None if bias is None else bias.detach().float().contiguous()
    out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    check(_lib.load().p2v_linear_f32(ptr(x), M, K, int(patch), Cin, H, W, ptr(w), ptr(b), N, ptr(out), stream()), "linear_f32")
    return out.reshape(shape)


def embed_f32(img, patch, w_hat, bias, mid_scale, mid_zp, aux_scale, aux_zp, pos, out_scale, out):
    """ViT-L stem: fp32 image -> int8 residual rows (class rows excluded), see include/p2vit_b200.h: p2v_embed_f32"""
    B, Cin, H, W = img.shape
    check(_lib.load().p2v_embed_f32(ptr(img), B, Cin, H, W, int(patch), ptr(w_hat), ptr(bias), w_hat.shape[0], float(mid_scale), float(mid_zp),
                                    float(aux_scale), float(aux_zp), ptr(pos), ptr(out_scale), ptr(out), stream()), "embed_f32")


def radix_hist(flat, prefix_mask, prefix_value, shift, nbits):
    """one radix-select pass over a flat fp32 CUDA tensor: int64 [2^nbits] counts of the digit (key >> shift) among the elements
    whose order key matches the prefix (include/p2vit_b200.h: p2v_radix_hist_f32)"""
    hist = torch.zeros(1 << nbits, dtype=torch.int64, device=flat.device)
    check(_lib.load().p2v_radix_hist_f32(ptr(flat), flat.numel(), int(prefix_mask), int(prefix_value), int(shift), int(nbits), ptr(hist),
                                         stream()), "radix_hist_f32")
    return hist


def launch_count(reset=False):
    lib = _lib.load()
    n = lib.p2v_launch_count()
    if reset:
        lib.p2v_reset_launch_count()
    return n


def capture_graph(launch):
    """CUDA graph of the launches `launch()` makes.  The cyclic garbage collector stays off while the stream is capturing: a model
    of an earlier (bit_config, batch) - engine, programs and their CUDAGraph objects form reference cycles - may be collected at any
    allocation, and destroying a CUDAGraph is not permitted while another capture is open (it invalidates the capture)."""
    import gc
    g = torch.cuda.CUDAGraph()
    was_enabled = gc.isenabled()
    gc.collect()
    gc.disable()
    try:
        with torch.cuda.graph(g):
            launch()
    finally:
        if was_enabled:
            gc.enable()
    return g
