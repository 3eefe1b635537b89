"""ctypes binding of csrc/libp2vit_b200.so (C ABI declared in include/p2vit_b200.h).

There is no CPU fallback: a missing library or a non-CUDA tensor raises.  The library is
built in-tree by `__graft_entry__.build()` / `make -C p2vit_b200/csrc`.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("P2V_LIB") or os.path.join(_HERE, "csrc", "libp2vit_b200.so")   # P2V_LIB: experiment builds (tools/)

EPI_REQUANT, EPI_GELU, EPI_RESIDUAL, EPI_EMBED, EPI_DEQUANT, EPI_F32 = range(6)
GELU_TABLE_BYTES = 16 + 8 * 4096 + 64 + 8 * 64 + 4 * 512


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("A", C.c_void_p), ("W", C.c_void_p),
        ("epilogue", C.c_int),
        ("acc_scale", C.c_void_p), ("bias", C.c_void_p), ("zp_corr", C.c_void_p),
        ("out_scale", C.c_void_p), ("mid_scale", C.c_void_p), ("res_scale", C.c_void_p),
        ("res", C.c_void_p), ("pos", C.c_void_p),
        ("aux_scale", C.c_float), ("tokens_per_image", C.c_int),
        ("out_i8", C.c_void_p), ("out_f32", C.c_void_p),
        ("gelu_table", C.c_void_p), ("row_map", C.c_void_p),
        ("pot_scales", C.c_int),
        ("out_zp", C.c_float), ("mid_zp", C.c_float), ("aux_zp", C.c_float),
    ]


class LayerNormArgs(C.Structure):
    _fields_ = [
        ("rows", C.c_int), ("C", C.c_int),
        ("x", C.c_void_p), ("x_row_stride", C.c_int64),
        ("in_mult", C.c_void_p), ("in_scale_min", C.c_float),
        ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("out_scale", C.c_void_p), ("post_div", C.c_void_p),
        ("next_scale", C.c_float), ("pot_scales", C.c_int),
        ("out_i8", C.c_void_p), ("out_f32", C.c_void_p),
        ("out_row_map", C.c_void_p), ("clamp_mid", C.c_int), ("next_zp", C.c_float),
        ("in_gather", C.c_void_p), ("gather_segs", C.c_int),
    ]


class WindowAttentionArgs(C.Structure):
    _fields_ = [
        ("n_windows", C.c_int), ("T", C.c_int), ("H", C.c_int), ("dh", C.c_int), ("windows_per_image", C.c_int),
        ("qkv", C.c_void_p), ("out", C.c_void_p),
        ("score_mult", C.c_float), ("s_attn1", C.c_float), ("s_attn2", C.c_float),
        ("bias", C.c_void_p), ("labels", C.c_void_p), ("mask_code", C.c_int), ("mask_exp_int", C.c_uint32),
        ("out_mult", C.c_float), ("lut_dev", C.c_void_p), ("out_row_map", C.c_void_p),
        ("bias_codes", C.c_void_p), ("bias_scale", C.c_float), ("mask_bits", C.c_void_p),
    ]


class AttentionArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("T", C.c_int), ("H", C.c_int), ("dh", C.c_int),
        ("qkv", C.c_void_p), ("out", C.c_void_p),
        ("score_mult", C.c_float), ("out_mult", C.c_float),
        ("lut_dev", C.c_void_p),
        ("probs_or_null", C.c_void_p), ("scores_or_null", C.c_void_p),
        ("zp_qkv", C.c_int), ("zp_score", C.c_float), ("zp_out", C.c_float),
        ("prob_mode", C.c_int),
    ]


# every symbol include/p2vit_b200.h declares: name -> (restype, argtypes)
_I, _I64, _F, _P = C.c_int, C.c_int64, C.c_float, C.c_void_p
SYMBOLS = {
    "p2v_abi_version": (_I, []),
    "p2v_last_error": (C.c_char_p, []),
    "p2v_launch_count": (_I64, []),
    "p2v_reset_launch_count": (None, []),
    "p2v_quantize_f32": (_I, [_P, _P, _I64, _I, _I64, _P, _I, _F, _I, _I, _P]),
    "p2v_fake_quant_f32": (_I, [_P, _P, _P, _I64, _I, _I64, _P, _I, _F, _I, _I, _P]),
    "p2v_dequantize_i8": (_I, [_P, _P, _I64, _I, _I64, _P, _I, _F, _P]),
    "p2v_quantize_patchify": (_I, [_P, _P, _I, _I, _I, _I, _I, _F, _F, _I, _I, _P]),
    "p2v_patchify_u8_lut": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "p2v_build_gelu_table": (_I, [_F, _P, _P]),
    "p2v_build_gelu_table_zp": (_I, [_F, _F, _P, _P]),
    "p2v_gemm_i8": (_I, [C.POINTER(GemmArgs), _P]),
    "p2v_gemm_i8_simt": (_I, [C.POINTER(GemmArgs), _P]),
    "p2v_set_gemm_variant": (None, [_I]),
    "p2v_fill_cls_rows": (_I, [_P, _P, _I, _I, _I, _P]),
    "p2v_layernorm_int": (_I, [C.POINTER(LayerNormArgs), _P]),
    "p2v_int_softmax_log2": (_I, [_P, _P, _I64, _I, _P, _P]),
    "p2v_attention_i8": (_I, [C.POINTER(AttentionArgs), _P]),
    "p2v_attention_i8_simt": (_I, [C.POINTER(AttentionArgs), _P]),
    "p2v_window_attention_i8": (_I, [C.POINTER(WindowAttentionArgs), _P]),
    "p2v_window_attention_i8_simt": (_I, [C.POINTER(WindowAttentionArgs), _P]),
    "p2v_gather_rows_i8": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "p2v_avgpool_quant_i8": (_I, [_P, _P, _I, _I, _I, _F, _F, _P]),
    "p2v_minmax_scratch_bytes": (_I64, [_I64, _I, _I64]),
    "p2v_minmax_per_channel": (_I, [_P, _P, _I64, _I, _I64, _P, _P]),
    "p2v_quant_mse_scratch_bytes": (_I64, [_I64, _I, _I64, _I, _I]),
    "p2v_quant_mse_scores": (_I, [_P, _I64, _I, _I64, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "p2v_linear_sqerr_scratch_bytes": (_I64, [_I, _I]),
    "p2v_linear_sqerr_scores": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P]),
    "p2v_linear_f32": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P]),
    "p2v_embed_f32": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _F, _F, _F, _F, _P, _P, _P, _P]),
    "p2v_radix_hist_f32": (_I, [_P, _I64, C.c_uint32, C.c_uint32, _I, _I, _P, _P]),
}

_lib = None


def load():
    """Loads the shared library (no GPU needed for loading) and types every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "p2vit_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C p2vit_b200/csrc`; there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library diverge
        fn.restype, fn.argtypes = res, args
    if lib.p2v_abi_version() != 3:
        raise RuntimeError("p2vit_b200: ABI version mismatch")
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        raise RuntimeError("p2vit_b200.%s failed: %s" % (what, load().p2v_last_error().decode()))


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("p2vit_b200 kernels need CUDA tensors (got %s); there is no CPU fallback" % t.device)
    if not t.is_contiguous():
        raise RuntimeError("p2vit_b200 kernels need contiguous tensors")
    return t.data_ptr()
