/*
 * p2vit_b200 - C ABI of the B200 (sm_100a) kernels behind the P2-ViT quantized operator surface.
 *
 * The reference has no FFI: its hot path sits behind Python classes (models/ptq/__init__.py:2-3)
 * whose forward() methods dispatch ATen ops.  Each entry point below replaces the ATen sequence
 * of one of those call sites (cited per function, paths relative to the reference root); the
 * Python classes in p2vit_b200/ptq bind them with ctypes (p2vit_b200/_lib.py).  See INTEGRATION.md
 * for the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all memory;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value: 0 = ok, non-zero = error (p2v_last_error() returns a static description);
 *   - activations are int8 codes, token-major [rows, channels], channels contiguous;
 *     weights are int8 codes [out_features, in_features] (nn.Linear layout), in_features contiguous;
 *   - RNE = round-half-to-even (torch.round); "sat" = clamp to [-128,127] (bit_type.py:17-27);
 *   - fp32 epilogue arithmetic is op-for-op IEEE (no FMA contraction, correctly rounded division)
 *     so integer codes equal the reference's fake-quant results (quantizer/uniform.py:83-86,125).
 */
#ifndef P2VIT_B200_H_
#define P2VIT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P2V_ABI_VERSION 3

int p2v_abi_version(void);
const char* p2v_last_error(void);
/* number of kernels this library has launched since load / since the last reset (bench: gpu_launches) */
int64_t p2v_launch_count(void);
void p2v_reset_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * QAct  (models/ptq/layers.py:242-257, quantizer/uniform.py:48-126)
 * q = sat(RNE(x / scale[c] + zp)), lo/hi given;  x_hat = (q - zp) * scale[c].
 * `n_scale` is 1 (layer_wise) or C (channel_wise); `inner` = number of contiguous elements that
 * share one channel (1 for [..,C] activations, H*W for NCHW inputs - base.py:14-31).
 * ------------------------------------------------------------------------------------------- */
int p2v_quantize_f32(const float* x, int8_t* q, int64_t n, int C, int64_t inner,
                     const float* scale, int n_scale, float zp, int lo, int hi, void* stream);
int p2v_fake_quant_f32(const float* x, float* y, int8_t* q_or_null, int64_t n, int C, int64_t inner,
                       const float* scale, int n_scale, float zp, int lo, int hi, void* stream);
int p2v_dequantize_i8(const int8_t* q, float* y, int64_t n, int C, int64_t inner,
                      const float* scale, int n_scale, float zp, void* stream);

/* qact_input fused with the patch gather of the k=stride=P convolution
 * (vit_fquant.py:842-851, layers_quant.py:486-489, layers.py:96-103): fp32 image [B,Cin,H,W]
 * -> int8 codes [B*(H/P)*(W/P), Cin*P*P] with K ordered (c,py,px) = QConv2d weight.reshape(D,-1). */
int p2v_quantize_patchify(const float* img, int8_t* out, int B, int Cin, int H, int W, int P,
                          float scale, float zp, int lo, int hi, void* stream);

/* The same step for 8-bit pixels [B,Cin,H,W] (what an image decoder produces, a quarter of the PCIe bytes): ToTensor (x/255),
 * Normalize ((x-mean)/std; test_quant.py:112-127,565-597) and qact_input depend only on (channel, byte), so `lut` [Cin,256] holds the
 * int8 code of each byte per channel - tabulated by the caller with p2v_quantize_patchify on the 256 normalised values, which makes
 * the result identical to the fp32 entry point on the normalised image. */
int p2v_patchify_u8_lut(const uint8_t* img, const int8_t* lut, int8_t* out, int B, int Cin, int H, int W, int P, void* stream);

/* ---------------------------------------------------------------------------------------------
 * QLinear / QConv2d(patch-embed) + the QAct(s) that follow it      (layers.py:202-209, 96-103)
 *   acc[m,n] = sum_k A[m,k] * W[n,k]            int8 x int8 -> int32, tcgen05.mma kind::i8
 *   y        = fl(acc * acc_scale[n] + bias[n])  (acc*acc_scale is exact for power-of-two scales)
 * followed by one of the fused epilogues below.  Column vectors have N entries.
 * ------------------------------------------------------------------------------------------- */
typedef enum {
  P2V_EPI_REQUANT = 0,   /* out_i8 = sat(RNE(y / out_scale[n]))            qkv->qact1 (vit_fquant.py:346-371), generic */
  P2V_EPI_GELU = 1,      /* out_i8 = sat(RNE(gelu_erf(y) / out_scale[n]))  fc1->GELU->qact1 (layers_quant.py:360-375) */
  P2V_EPI_RESIDUAL = 2,  /* c = sat(RNE(y/mid_scale[n])); z = fl(res[m,n]*res_scale[n]) + fl(c*mid_scale[n]);
                            out_i8 = sat(RNE(z / out_scale[n]))            proj->qact3->+x->qact2, fc2->qact2->+x->qact4
                            (vit_fquant.py:397-401,514-534,561-580; layers_quant.py:384-388) */
  P2V_EPI_EMBED = 3,     /* c = sat(RNE(y/mid_scale[0])); e = sat(RNE(fl(c*mid_scale[0]) / aux_scale));
                            v = fl(e*aux_scale) + pos[tok+1,n]; out row (b*(T+1)+tok+1) = sat(RNE(v / out_scale[n]))
                            patch_embed.qact -> qact_embed -> +qact_pos(pos) -> qact1 (vit_fquant.py:851-869) */
  P2V_EPI_DEQUANT = 4,   /* out_f32 = sat(RNE(y / out_scale[n])) * out_scale[n]; optional out_i8 codes
                            head->act_out (vit_fquant.py:932-936) */
  P2V_EPI_F32 = 5        /* out_f32 = y  (eager QLinear.forward result before the next QAct) */
} p2v_epilogue_t;

typedef struct {
  int M, N, K;
  const int8_t* A;          /* [M,K] */
  const int8_t* W;          /* [N,K] */
  int epilogue;             /* p2v_epilogue_t */
  const float* acc_scale;   /* [N]  s_in * s_w[n] */
  const float* bias;        /* [N] or NULL */
  const int32_t* zp_corr;   /* [N] or NULL: zp_in * sum_k W[n,k], subtracted from acc (asymmetric inputs) */
  const float* out_scale;   /* [N] */
  const float* mid_scale;   /* [N] RESIDUAL: scale of the QAct directly after the GEMM; EMBED: [1] */
  const float* res_scale;   /* [N] RESIDUAL: scale of the residual stream codes */
  const int8_t* res;        /* [M,N] RESIDUAL: residual codes */
  const float* pos;         /* EMBED: [(T+1),N] dequantized qact_pos(pos_embed) */
  float aux_scale;          /* EMBED: qact_embed scale */
  int tokens_per_image;     /* EMBED: T (196) */
  int8_t* out_i8;           /* [M,N] (EMBED: [B*(T+1),N]) */
  float* out_f32;           /* [M,N] DEQUANT / F32 */
  const void* gelu_table;   /* GELU with pot_scales: optional p2v_gelu_table (device) built by p2v_build_gelu_table for this
                               out_scale; NULL = evaluate erf per element */
  const int32_t* row_map;   /* [M] or NULL (REQUANT/GELU/RESIDUAL): output row (and residual row) of GEMM row m - Swin window
                               reverse + inverse cyclic shift fused into the store (swin_quant.py:426-436) */
  int pot_scales;           /* REQUANT/GELU/DEQUANT: 1 = acc_scale and out_scale are exact powers of two (division == exact
                               multiply); RESIDUAL: 1 = acc_scale is (acc*acc_scale exact; mid/out scales stay general) */
  /* Zero points of asymmetric activation quantizers (observer/omse.py:30-57, quantizer/uniform.py:83-86,125; integer valued,
   * within [-128,127]; all 0 for the symmetric observers).  The INPUT zero point enters through zp_corr.  With a zero point the
   * output code is  sat(RNE(fl(fl(y / scale) + zp)))  and a dequantized value is  fl((code - zp) * scale):
   *   out_zp  REQUANT / GELU / DEQUANT: zero point of the output QAct
   *   mid_zp  EMBED: zero point of patch_embed.qact        aux_zp  EMBED: zero point of qact_embed
   * (RESIDUAL's mid / out quantizers and EMBED's out quantizer are the channel-wise PTF ones: symmetric by construction,
   * ptf.py:120).  Non-zero values need pot_scales = 0 and run on csrc/gemm_tc.cu. */
  float out_zp, mid_zp, aux_zp;
} p2v_gemm_args;

/* Step table of  y -> sat(RNE(gelu_erf(y) / out_scale))  for a power-of-two out_scale (layers_quant.py:373-375).  The
 * code is piecewise constant in y; the table cuts [-8.5, 128*out_scale + 0.5) into segments of width out_scale/2 (at most
 * one code change each) and stores the code below / above the change and the exact fp32 threshold, found by bisection on
 * the kernels' own gelu_erf.  The epilogue looks the code up and re-evaluates erf only for y within 8 ulps of a threshold
 * (and for the rare segments flagged non-monotone), so its result is identical to the direct evaluation.
 * Layout: header {float y0, inv_w; int32 n, reserved} then n entries {float threshold; uint32 below | above<<8 | slow<<31}. */
#define P2V_GELU_TABLE_MAX_ENTRIES 4096
/* A second form of the same function follows the first in the buffer (used by the CTA-pair GEMM, csrc/common.cuh:
 * GeluStepsHeader): a linear map per y-segment that pins the code down to two candidates, and one exact threshold per code
 * boundary on each side of GELU's minimum, laid out so that the epilogue's lookups are free of bank conflicts.  The builder
 * checks it against the direct evaluation on ~2.3 M arguments (every threshold +-256 ulps, a dense grid, 200 binades). */
#define P2V_GELU_TABLE_BYTES (16 + 8 * P2V_GELU_TABLE_MAX_ENTRIES + 64 + 8 * 64 + 4 * 512)
/* returns 0 and fills `table_dev` (P2V_GELU_TABLE_BYTES bytes, 16-byte aligned), or 3 if out_scale is outside the tabulated
 * range 2^-7 .. 2^-2 or a form failed its self-check (the caller then passes gelu_table = NULL and the kernels evaluate erf per
 * element).  An out_scale that is not a power of two (ema / percentile observers, zero point 0) gets the second form only
 * (first form: n = 0), with thresholds found on the reference's division gelu(y) / out_scale; the kernels that read the first
 * form ignore the table unless pot_scales is set.  Synchronises `stream` once (the verdict of the self-check is read back). */
int p2v_build_gelu_table(float out_scale, void* table_dev, void* stream);
/* the same for an asymmetric output quantizer (omse): code = sat(RNE(fl(fl(gelu(y) / out_scale) + out_zp))), out_zp an integer in
 * [-128, 127]; second form only.  The caller passes the table with the matching out_scale / out_zp of p2v_gemm_args. */
int p2v_build_gelu_table_zp(float out_scale, float out_zp, void* table_dev, void* stream);

int p2v_gemm_i8(const p2v_gemm_args* args_host, void* stream);
/* Two tcgen05 kernels implement p2v_gemm_i8 with identical results: csrc/gemm_pair.cu (CTA pairs, cta_group::2, TMA-staged
 * output; REQUANT / GELU / RESIDUAL with int8 output, N % 16 == 0, no row_map / zero points) and csrc/gemm_tc.cu (everything
 * else and small M).  variant: 0 = automatic (default), 1 = always gemm_tc.cu, 2 = gemm_pair.cu whenever it applies
 * (tests cross-check the two). Process-wide, not thread safe. */
void p2v_set_gemm_variant(int variant);
/* same contract on CUDA cores (dp4a); used by tests to cross-check the tcgen05 kernel */
int p2v_gemm_i8_simt(const p2v_gemm_args* args_host, void* stream);
/* EMBED helper: writes the B class-token rows: out[b*(T+1), n] = cls_row[n] */
int p2v_fill_cls_rows(int8_t* out, const int8_t* cls_row, int B, int T, int N, void* stream);

/* ---------------------------------------------------------------------------------------------
 * QIntLayerNorm 'int' mode + the QAct that consumes it     (layers.py:270-337; vit_fquant.py:519-524,
 * 565-570, 905; layers_quant.py:349-356)
 *   x      = codes[m,c] * in_mult[c]          in_mult = RNE(in_scale/min(in_scale)) in {1,2,4,8}
 *   mean, std, A, M, N, B as the reference (fp32, op for op); row sums are exact integers
 *   y_q    = RNE((sign(A)*M*x + B) / 2^N)                       (not clamped by the reference)
 *   out_i8 = sat(RNE(fl(fl(y_q*out_scale[c]) / post_div[c]) / next_scale))
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int rows, C;
  const int8_t* x;          /* [rows,C] */
  int64_t x_row_stride;     /* bytes between rows (C, or (T+1)*C to pick the class token only) */
  const float* in_mult;     /* [C] */
  float in_scale_min;       /* s1 */
  const float* gamma;       /* [C] */
  const float* beta;        /* [C] */
  const float* out_scale;   /* [C] LN output grid: next_qact.scale * channel_scale (or next_qact.scale) */
  const float* post_div;    /* [C] smoothing divisor applied before the next QAct (1 if none) */
  float next_scale;         /* scale of the QAct after the LN */
  int pot_scales;           /* 1: out_scale, post_div, next_scale are powers of two */
  int8_t* out_i8;           /* [rows,C] */
  float* out_f32;           /* optional [rows,C]: y_q * out_scale (eager QIntLayerNorm.forward result) */
  const int32_t* out_row_map; /* [rows] or NULL: destination row of out_i8 for input row r - Swin cyclic shift + window partition
                                 fused into the store (swin_quant.py:408-419) */
  int clamp_mid;            /* 1: y_q is clamped to [-128,127] first - a QAct at the LayerNorm's own output scale sits between the
                               LayerNorm and the smoothing divide (Swin: norm2 -> qact3 -> Mlp, swin_quant.py:439-446) */
  float next_zp;            /* zero point of the QAct after the LN (0 unless its observer is asymmetric; needs pot_scales = 0):
                               out_i8 = sat(RNE(fl(fl(fl(y_q*out_scale[c]) / post_div[c]) / next_scale) + next_zp)) */
  const int32_t* in_gather; /* [rows * gather_segs] or NULL.  Swin patch merging (swin_quant.py:512-519) as the LayerNorm's input
                               row map: row r is the concatenation of gather_segs source rows of C / gather_segs channels,
                               segment k = x[in_gather[r * gather_segs + k]] (x_row_stride = bytes of one SOURCE row), so the 2x2
                               neighbourhood gather costs no kernel and no [rows, 4C] buffer of its own */
  int gather_segs;          /* 4 for patch merging; C / gather_segs must be a multiple of 4 */
} p2v_layernorm_args;

int p2v_layernorm_int(const p2v_layernorm_args* args_host, void* stream);

/* ---------------------------------------------------------------------------------------------
 * QIntSoftmax (log2, 4 bit)                                         (layers.py:376-428)
 * A 256-entry table of exp_int for x_int = -d (d = rowmax - code), built on the host with the
 * reference's fp32 formulas (p2vit_b200/engine.py: build_softmax_lut), drives both the stand-alone
 * kernel and the fused attention.  exp_int = hi*2^32 + lo exactly; exp_f32 = the same value in fp32.
 * code = clamp(log_round(RNE(fl(sum)/exp_f32)), 0, 15); probability 2^-code, 0 when log_round >= 16
 * (code 255 on the wire).  Every entry must be below 2^55 (the fused attention kernel sums up to 256 of them in 64 bits);
 * p2vit_b200/intmath.py: build_softmax_lut refuses scales that would exceed it (below ~2^-10).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t hi[256];
  uint32_t lo[256];
  float exp_f32[256];
} p2v_softmax_lut;

/* scores: int8 codes [rows, n]; out: u8 codes [rows, n] (0..15, 255 = zero probability) */
int p2v_int_softmax_log2(const int8_t* scores, uint8_t* out, int64_t rows, int n,
                         const p2v_softmax_lut* lut_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Attention core between qact1 and qact2                     (vit_fquant.py:373-389)
 *   S = q k^T (int32);  c = sat(RNE(S * score_mult))  with score_mult = s_q^2 * head_scale / s_attn
 *   p = int_softmax_log2(c);  O = sum_j 2^(15-code_j) * v_j (int32, exact)
 *   out_i8 = sat(RNE(O * out_mult)),  out_mult = 2^-15 * s_v / s_out
 * qkv: int8 [B, T, 3, H, dh] (the qkv QLinear output, qact1 codes); out: int8 [B, T, H*dh].
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int B, T, H, dh;
  const int8_t* qkv;
  int8_t* out;
  float score_mult;
  float out_mult;
  const p2v_softmax_lut* lut_dev;
  uint8_t* probs_or_null;   /* optional dump of softmax codes [B,H,T,T] (tests) */
  int8_t* scores_or_null;   /* optional dump of qact_attn1 codes [B,H,T,T] (tests) */
  /* asymmetric qact1 / qact_attn1 / qact2 (omse): S = sum_d (q - zp_qkv)(k - zp_qkv);  c = sat(RNE(fl(S*score_mult) + zp_score));
   * the softmax works on c (a zero point cancels in x - rowmax);  O = sum_j 2^(15-code_j) (v_j - zp_qkv);
   * out_i8 = sat(RNE(fl(O*out_mult) + zp_out)).  All integer valued; 0 for symmetric observers. */
  int zp_qkv;
  float zp_score, zp_out;
  /* how the tcgen05 kernel gets log_round(RNE(fl(sum / exp))); the codes are identical, only the speed differs with the data:
   *   0  product with the table's reciprocal + a guard band; a 16-score unit with a score next to a decision boundary is redone
   *      with the IEEE division (fastest when that is rare)
   *   1  the exactly rounded quotient for every score (reciprocal + two residual corrections), no guard, no redo - for coarse score
   *      scales whose quotients land on the ties x.5 all the time (ViT-B: up to a fifth of the units took the redo)
   * p2vit_b200/engine.py times both on the first batch of a program and keeps the faster one per layer. */
  int prob_mode;
} p2v_attention_args;

/* Head dim 64, T <= 224 and no debug dumps: tcgen05 kernel (csrc/attention_tc.cu: TMA-fed S = q k^T and O = P v on
 * the tensor cores, accumulators in TMEM, softmax per TMEM row); otherwise the dp4a kernel (csrc/attention.cu). */
int p2v_attention_i8(const p2v_attention_args* args_host, void* stream);
/* same contract, always on CUDA cores (dp4a); used by tests to cross-check the tcgen05 kernel */
int p2v_attention_i8_simt(const p2v_attention_args* args_host, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Swin window attention between attn.qact1 and attn.qact3            (swin_quant.py:211-249)
 *   S = q k^T (int32, head scale factored out);  c1 = sat(RNE(S * score_mult))           qact_attn1
 *   c2 = sat(RNE(fl(fl(c1*s_attn1) + bias[h,i,j]) / s_attn2))                             + quantized rel-pos bias -> qact2
 *   x_int = c2 + mask_code * [label_i != label_j]      (SW-MSA mask -100 added after qact2, as integer -100/s_attn2)
 *   p = int_softmax_log2(x_int);  O = sum_j 2^(15-code_j) v_j;  out = sat(RNE(O * out_mult))              qact3
 * qkv: int8 [nWin_total, T, 3, H, dh] in window order; out: int8 [nWin_total, T, H*dh]; T = ws*ws <= 64, dh = 32.
 * bias: fp32 [H, T, T] dequantized qact_table(relative_position_bias_table)[relative_position_index];
 * labels: int8 [windows_per_image, T] SW-MSA region labels or NULL (no shift); window w uses labels[w % windows_per_image].
 * lut_dev: table for s_attn2; masked entries use the table's clamped tail (the host checks -100/s_attn2 reaches it).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int n_windows, T, H, dh, windows_per_image;
  const int8_t* qkv;
  int8_t* out;
  float score_mult;         /* s_q^2 * dh^-0.5 / s_attn1 */
  float s_attn1, s_attn2;
  const float* bias;
  const int8_t* labels;
  int mask_code;            /* RNE(-100 / s_attn2) (negative) */
  uint32_t mask_exp_int;    /* exp_int of a masked entry: the clamped tail floor((1/0.35815147)/s_attn2^2) of int_exp
                               (layers.py:396-410 with x_int = 32*x0) */
  float out_mult;           /* 2^-15 * s_v / s_attn3 */
  const p2v_softmax_lut* lut_dev;
  const int32_t* out_row_map;   /* optional [n_windows * T]: row (window order) -> destination row of `out`; NULL = same row.
                                   window_reverse + roll (swin_quant.py:426-436) applied by the store, so the proj GEMM that
                                   follows reads and writes token order and needs no scatter */
  /* The same two tables in the form the tensor-core kernel stages (both optional; without them the dp4a kernel runs):
   *   bias_codes  int8 [H, T, 80]: qact_table codes gathered through relative_position_index, rows padded to 80 bytes,
   *               16-byte aligned;  bias[h,i,j] == fl(bias_codes[h,i,j] * bias_scale) (bias_scale = qact_table.scale)
   *   mask_bits   uint64 [windows_per_image, T]: bit j of word (w, i) = [label_i != label_j]; required with `labels` */
  const int8_t* bias_codes;
  float bias_scale;
  const uint64_t* mask_bits;
} p2v_window_attention_args;

/* Head dim 32, T <= 64, bias_codes (and mask_bits when labels are given): tcgen05 kernel (csrc/swin_attention_tc.cu: two
 * windows per 128-row tile, S = q k^T and O = P v on the tensor cores, softmax per TMEM row); otherwise the dp4a kernel
 * (csrc/swin_ops.cu). */
int p2v_window_attention_i8(const p2v_window_attention_args* args_host, void* stream);
/* same contract, always on CUDA cores (dp4a); used by tests to cross-check the tcgen05 kernel */
int p2v_window_attention_i8_simt(const p2v_window_attention_args* args_host, void* stream);

/* Patch merging gather (swin_quant.py:512-519): out[r, k*C:(k+1)*C] = in[src_rows[r*segs + k], :]  (int8 rows of C bytes) */
int p2v_gather_rows_i8(const int8_t* in, int8_t* out, const int32_t* src_rows, int rows_out, int segs, int C, void* stream);

/* Token average pooling + QAct (swin_quant.py:904-905): codes [B, T, C] at scale s_in ->
 * out[b, c] = sat(RNE(fl(fl(float(sum_t codes) * s_in) / T) / s_out))   (the sum of codes is exact) */
int p2v_avgpool_quant_i8(const int8_t* in, int8_t* out, int B, int T, int C, float s_in, float s_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Calibration observers                           (observer/minmax.py:15-32, ptf.py:13-30, base.py:16-29)
 * per-channel min and max of x viewed as [C, n/C] with the reference's channel rule
 * (`inner` as in p2v_quantize_f32).  minmax: float [2,C] (row 0 = min, row 1 = max),
 * initialised by the kernel (not accumulated).
 * ------------------------------------------------------------------------------------------- */
/* `scratch`: caller-owned device buffer of at least p2v_minmax_scratch_bytes(...) bytes (block partials; the library holds
 * no device memory of its own, so concurrent streams cannot collide on it) */
int64_t p2v_minmax_scratch_bytes(int64_t n, int C, int64_t inner);
int p2v_minmax_per_channel(const float* x, float* minmax, int64_t n, int C, int64_t inner, float* scratch, void* stream);
/* sum over all elements of (x - fq_k(x))^2 for K candidate scales (minmax.py:165-201 activation case,
 * omse.py:30-57, ptf.py:123-149): scales [K, n_scale]; out double [K, n_scale_out] where
 * n_scale_out = C if per_channel_out else 1.  fq_k(x) = (sat_lo_hi(RNE(x/s + zp_k)) - zp_k) * s.
 * Block partials go to `scratch` (>= p2v_quant_mse_scratch_bytes(...) bytes) and are folded in a fixed order: the scores are
 * bit-reproducible, so the argmin of near-tied candidates is too. */
int64_t p2v_quant_mse_scratch_bytes(int64_t n, int C, int64_t inner, int K, int per_channel_out);
int p2v_quant_mse_scores(const float* x, int64_t n, int C, int64_t inner, const float* scales,
                         const float* zps_or_null, int K, int n_scale, int per_channel_out,
                         int lo, int hi, double* out, double* scratch, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Percentile observer                                              (observer/percentile.py:26-55)
 * One pass of a most-significant-digit radix select over the order-preserving integer image of fp32
 * (key = bits ^ 0x80000000 for non-negative values, ~bits for negative ones): among the elements whose key
 * satisfies (key & prefix_mask) == prefix_value, hist[(key >> shift) & (2^nbits - 1)] += 1.  `hist` holds 2^nbits
 * unsigned 64-bit counters and is ACCUMULATED into (the caller zeroes it).  Integer counts add exactly across the
 * ranks of a data-parallel calibration (all-reduce SUM), so three passes (12 + 12 + 8 bits) give the global k-th
 * smallest element - what torch.quantile / np.percentile sort for in the reference - bit for bit, with no sort.
 * ------------------------------------------------------------------------------------------- */
int p2v_radix_hist_f32(const float* x, int64_t n, uint32_t prefix_mask, uint32_t prefix_value, int shift, int nbits,
                       unsigned long long* hist, void* stream);

/* ---------------------------------------------------------------------------------------------
 * fp32 GEMMs on the CUDA cores (csrc/sgemm.cu) - the two places whose operands are genuinely fp32.
 *
 * Weight power-of-two search of MinmaxObserver (observer/minmax.py:145-207): the reference's score of candidate k for output
 * channel j is  sum_rows (layer(x; W)[., j] - layer(x; fq_k(W))[., j])^2  =  sum_rows (x . D[j, :])^2  with D = W - fq_k(W).
 * D: fp32 [n, K] (the difference rows of every candidate stacked); out: double [n] = per-row-of-D sums over the M calibration
 * rows.  x: fp32 [M, K], or (patch > 0) an NCHW image [B, Cin, H, W] whose k = stride = patch patches are the rows (QConv2d).
 * scratch: >= p2v_linear_sqerr_scratch_bytes(M, n) bytes; row blocks are folded in a fixed order (bit-reproducible).
 * ------------------------------------------------------------------------------------------- */
int64_t p2v_linear_sqerr_scratch_bytes(int M, int n);
int p2v_linear_sqerr_scores(const float* x, int M, int K, int patch, int Cin, int H, int W, const float* D, int n, double* out,
                            double* scratch, void* stream);
/* The FP layer itself, out[M, N] = x W^T + bias (bias may be NULL): QLinear / QConv2d before model_quant() - the calibration
 * forward (layers.py:87,173; test_quant.py:275-281).  Every output element is one ascending-k FMA chain whatever M is, so the
 * activations - and with them every observer statistic - do not depend on how the calibration batch is split over GPUs. */
int p2v_linear_f32(const float* x, int M, int K, int patch, int Cin, int H, int W, const float* Wt, const float* bias, int N, float* out,
                   void* stream);
/* ViT-Large stem (vit_fquant.py:1063 input_quant=False; layers_quant.py:486-489): fp32 pixels x dequantized weights w_hat [N, K]
 * + bias, then the P2V_EPI_EMBED chain (patch_embed.qact -> qact_embed -> + pos -> qact1) with IEEE divisions;
 * out: int8 [B*(T+1), N], rows b*(T+1) + tok + 1 (the class rows are p2v_fill_cls_rows'). */
int p2v_embed_f32(const float* img, int B, int Cin, int H, int W, int P, const float* w_hat, const float* bias, int N, float mid_scale,
                  float mid_zp, float aux_scale, float aux_zp, const float* pos, const float* out_scale, int8_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* P2VIT_B200_H_ */
