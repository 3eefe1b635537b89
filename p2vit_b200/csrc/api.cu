// extern "C" surface of libp2vit_b200.so (include/p2vit_b200.h): argument validation + launch.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <atomic>
#include "common.cuh"

namespace p2v {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
// 0 = never, 1 = every launch, 2 (default) = launches that are being captured into a CUDA graph.  The engines capture the
// forward once per (bit_config, batch) and every constant a kernel reads ahead of its dependency wait was written at plan time;
// eager launches (per-operator modules, taps) may follow a torch kernel that has just produced such a constant, so they keep the
// plain stream order.
static int pdl_mode() {
  static const int mode = [] {
    const char* e = std::getenv("P2VIT_PDL");
    return e && e[0] == '0' ? 0 : (e && e[0] == '1' ? 1 : 2);
  }();
  return mode;
}
static thread_local int t_pdl_kind = PDL_OTHER;
void pdl_next_kind(int kind) { t_pdl_kind = kind; }
int pdl_take_kind() {
  const int k = t_pdl_kind;
  t_pdl_kind = PDL_OTHER;
  return k;
}
bool pdl_enabled(cudaStream_t stream, int kind) {
  static const int kinds = [] {
    const char* e = std::getenv("P2VIT_PDL_KINDS");
    return e ? std::atoi(e) : (PDL_GEMM | PDL_LAYERNORM);     // measured (tools/ab_pdl_kinds.sh): attention loses 1 % with it, the small kernels gain nothing
  }();
  if (!(kinds & kind)) return false;
  const int m = pdl_mode();
  if (m != 2) return m == 1;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  return cudaStreamIsCapturing(stream, &st) == cudaSuccess && st == cudaStreamCaptureStatusActive;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return 2;
  }
  return 0;
}

int launch_gemm_simt(const p2v_gemm_args& a, cudaStream_t stream);
int launch_build_gelu_table(float out_scale, float out_zp, void* table_dev, cudaStream_t stream);
int launch_gemm_tc(const p2v_gemm_args& a, cudaStream_t stream);
void set_gemm_variant(int v);
int launch_quantize(const float* x, int8_t* q, float* y, int64_t n, int C, int64_t inner, const float* scale, int n_scale,
                    float zp, int lo, int hi, cudaStream_t stream);
int launch_dequantize(const int8_t* q, float* y, int64_t n, int C, int64_t inner, const float* scale, int n_scale, float zp,
                      cudaStream_t stream);
int launch_patchify_u8(const uint8_t* img, const int8_t* lut, int8_t* out, int B, int Cin, int H, int W, int P, cudaStream_t stream);
int launch_patchify(const float* img, int8_t* out, int B, int Cin, int H, int W, int P, float scale, float zp, int lo, int hi,
                    cudaStream_t stream);
int launch_fill_cls(int8_t* out, const int8_t* cls_row, int B, int T, int N, cudaStream_t stream);
int launch_layernorm(const p2v_layernorm_args& a, cudaStream_t stream);
int launch_softmax(const int8_t* scores, uint8_t* out, int64_t rows, int n, const p2v_softmax_lut* lut, cudaStream_t stream);
int launch_attention(const p2v_attention_args& a, cudaStream_t stream);
int launch_attention_tc(const p2v_attention_args& a, cudaStream_t stream);
bool attention_tc_supported(const p2v_attention_args& a);
int launch_window_attention(const p2v_window_attention_args& a, uint32_t e_mask, cudaStream_t stream);
int launch_window_attention_tc(const p2v_window_attention_args& a, uint32_t e_mask, cudaStream_t stream);
bool window_attention_tc_supported(const p2v_window_attention_args& a);
int launch_gather_rows(const int8_t* in, int8_t* out, const int32_t* src, int rows_out, int segs, int C, cudaStream_t stream);
int launch_avgpool_quant(const int8_t* in, int8_t* out, int B, int T, int C, float s_in, float s_out, cudaStream_t stream);
int launch_minmax(const float* x, float* minmax, int64_t n, int C, int64_t inner, float* part, cudaStream_t stream);
int64_t minmax_scratch_bytes(int64_t n, int C, int64_t inner);
int launch_mse_scores(const float* x, int64_t n, int C, int64_t inner, const float* scales, const float* zps, int K, int n_scale,
                      int per_channel_out, int lo, int hi, double* out, double* part, cudaStream_t stream);
int64_t mse_scratch_bytes(int64_t n, int C, int64_t inner, int K, int per_channel_out);
int launch_radix_hist(const float* x, int64_t n, uint32_t prefix_mask, uint32_t prefix_value, int shift, int nbits,
                      unsigned long long* hist, cudaStream_t stream);

struct SgemmEmbed { const float* bias; float* out_f32; const float* pos; const float* out_scale; float mid_scale, mid_zp, aux_scale, aux_zp; int tokens_per_image; int8_t* out; };
int64_t linear_sqerr_scratch_bytes(int M, int n);
int launch_linear_sqerr(const float* x, int M, int K, int patch, int Cin, int H, int W, const float* D, int n, double* out, double* scratch,
                        cudaStream_t stream);
int launch_embed_f32(const float* img, int B, int Cin, int H, int W, int P, const float* w_hat, int N, const SgemmEmbed& ep, cudaStream_t stream);
int launch_linear_f32(const float* x, int M, int K, int patch, int Cin, int H, int W, const float* Wt, const float* bias, int N, float* out,
                      cudaStream_t stream);

static int validate_gemm(const p2v_gemm_args* a) {
  P2V_REQUIRE(a != nullptr, "gemm: null args");
  P2V_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "gemm: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  P2V_REQUIRE(a->K % 16 == 0, "gemm: K=%d must be a multiple of 16 (TMA row pitch)", a->K);
  P2V_REQUIRE(a->A && a->W && a->acc_scale, "gemm: A, W and acc_scale are required");
  P2V_REQUIRE((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->W) & 15) == 0,
              "gemm: A and W must be 16-byte aligned");
  const int e = a->epilogue;
  P2V_REQUIRE(e >= P2V_EPI_REQUANT && e <= P2V_EPI_F32, "gemm: unknown epilogue %d", e);
  if (e != P2V_EPI_F32) P2V_REQUIRE(a->out_scale != nullptr, "gemm: out_scale required");
  if (e == P2V_EPI_F32 || e == P2V_EPI_DEQUANT) P2V_REQUIRE(a->out_f32 != nullptr, "gemm: out_f32 required");
  else P2V_REQUIRE(a->out_i8 != nullptr, "gemm: out_i8 required");
  if (e == P2V_EPI_RESIDUAL) P2V_REQUIRE(a->mid_scale && a->res_scale && a->res, "gemm: residual epilogue needs mid_scale, res_scale, res");
  if (a->out_zp != 0.f || a->mid_zp != 0.f || a->aux_zp != 0.f) {
    P2V_REQUIRE(!a->pot_scales, "gemm: zero points need the general epilogues (pot_scales = 0)");
    P2V_REQUIRE(e != P2V_EPI_RESIDUAL && e != P2V_EPI_F32, "gemm: this epilogue's quantizers are symmetric (no zero point)");
    P2V_REQUIRE(a->out_zp == rintf(a->out_zp) && a->mid_zp == rintf(a->mid_zp) && a->aux_zp == rintf(a->aux_zp) && fabsf(a->out_zp) <= 128.f &&
                fabsf(a->mid_zp) <= 128.f && fabsf(a->aux_zp) <= 128.f, "gemm: zero points must be integers within [-128,127]");
  }
  if (e == P2V_EPI_EMBED) P2V_REQUIRE(a->mid_scale && a->pos && a->tokens_per_image > 0 && a->M % a->tokens_per_image == 0,
                                      "gemm: embed epilogue needs mid_scale, pos and M %% tokens_per_image == 0");
  return 0;
}

}  // namespace p2v

using namespace p2v;

extern "C" {

int p2v_abi_version(void) { return P2V_ABI_VERSION; }
const char* p2v_last_error(void) { return g_err; }
int64_t p2v_launch_count(void) { return g_launches.load(); }
void p2v_reset_launch_count(void) { g_launches.store(0); }

int p2v_quantize_f32(const float* x, int8_t* q, int64_t n, int C, int64_t inner, const float* scale, int n_scale, float zp,
                     int lo, int hi, void* stream) {
  P2V_REQUIRE(x && q && scale && n >= 0 && C > 0 && inner > 0 && (n_scale == 1 || n_scale == C), "quantize: bad arguments");
  P2V_REQUIRE(lo >= -128 && hi <= 127, "quantize: int8 carrier holds [-128,127] only (lo=%d hi=%d)", lo, hi);
  return launch_quantize(x, q, nullptr, n, C, inner, scale, n_scale, zp, lo, hi, (cudaStream_t)stream);
}
int p2v_fake_quant_f32(const float* x, float* y, int8_t* q, int64_t n, int C, int64_t inner, const float* scale, int n_scale,
                       float zp, int lo, int hi, void* stream) {
  P2V_REQUIRE(x && y && scale && n >= 0 && C > 0 && inner > 0 && (n_scale == 1 || n_scale == C), "fake_quant: bad arguments");
  return launch_quantize(x, q, y, n, C, inner, scale, n_scale, zp, lo, hi, (cudaStream_t)stream);
}
int p2v_dequantize_i8(const int8_t* q, float* y, int64_t n, int C, int64_t inner, const float* scale, int n_scale, float zp,
                      void* stream) {
  P2V_REQUIRE(q && y && scale && n >= 0 && C > 0 && inner > 0 && (n_scale == 1 || n_scale == C), "dequantize: bad arguments");
  return launch_dequantize(q, y, n, C, inner, scale, n_scale, zp, (cudaStream_t)stream);
}
int p2v_quantize_patchify(const float* img, int8_t* out, int B, int Cin, int H, int W, int P, float scale, float zp, int lo,
                          int hi, void* stream) {
  P2V_REQUIRE(img && out && B > 0 && Cin > 0 && P > 0 && H % P == 0 && W % P == 0, "patchify: bad shape");
  P2V_REQUIRE(P % 4 == 0 && W % 4 == 0, "patchify: P and W must be multiples of 4");
  P2V_REQUIRE(lo >= -128 && hi <= 127, "patchify: int8 carrier only");
  return launch_patchify(img, out, B, Cin, H, W, P, scale, zp, lo, hi, (cudaStream_t)stream);
}
int p2v_patchify_u8_lut(const uint8_t* img, const int8_t* lut, int8_t* out, int B, int Cin, int H, int W, int P, void* stream) {
  P2V_REQUIRE(img && lut && out && B > 0 && Cin > 0 && Cin <= 64 && P > 0 && H % P == 0 && W % P == 0, "patchify_u8: bad shape");
  P2V_REQUIRE(P % 4 == 0 && W % 4 == 0 && (reinterpret_cast<uintptr_t>(img) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
              "patchify_u8: P and W must be multiples of 4, buffers 4-byte aligned");
  return launch_patchify_u8(img, lut, out, B, Cin, H, W, P, (cudaStream_t)stream);
}
int p2v_build_gelu_table(float out_scale, void* table_dev, void* stream) {
  P2V_REQUIRE(table_dev != nullptr && (reinterpret_cast<uintptr_t>(table_dev) & 15) == 0, "build_gelu_table: table must be 16-byte aligned");
  return launch_build_gelu_table(out_scale, 0.f, table_dev, (cudaStream_t)stream);
}
int p2v_build_gelu_table_zp(float out_scale, float out_zp, void* table_dev, void* stream) {
  P2V_REQUIRE(table_dev != nullptr && (reinterpret_cast<uintptr_t>(table_dev) & 15) == 0, "build_gelu_table: table must be 16-byte aligned");
  return launch_build_gelu_table(out_scale, out_zp, table_dev, (cudaStream_t)stream);
}
int p2v_gemm_i8(const p2v_gemm_args* a, void* stream) {
  if (int r = validate_gemm(a)) return r;
  return launch_gemm_tc(*a, (cudaStream_t)stream);
}
void p2v_set_gemm_variant(int variant) { set_gemm_variant(variant); }
int p2v_gemm_i8_simt(const p2v_gemm_args* a, void* stream) {
  if (int r = validate_gemm(a)) return r;
  return launch_gemm_simt(*a, (cudaStream_t)stream);
}
int p2v_fill_cls_rows(int8_t* out, const int8_t* cls_row, int B, int T, int N, void* stream) {
  P2V_REQUIRE(out && cls_row && B > 0 && T > 0 && N > 0, "fill_cls_rows: bad arguments");
  return launch_fill_cls(out, cls_row, B, T, N, (cudaStream_t)stream);
}
int p2v_layernorm_int(const p2v_layernorm_args* a, void* stream) {
  P2V_REQUIRE(a && a->x && a->in_mult && a->gamma && a->beta && a->out_scale && a->post_div, "layernorm: missing pointers");
  P2V_REQUIRE(a->rows > 0 && a->C > 0 && a->C % 4 == 0 && a->C <= 4096, "layernorm: C=%d must be a multiple of 4, <= 4096", a->C);
  P2V_REQUIRE(a->out_i8 || a->out_f32, "layernorm: no output");
  P2V_REQUIRE(a->x_row_stride % 4 == 0, "layernorm: row stride must be a multiple of 4 bytes");
  P2V_REQUIRE(a->next_zp == 0.f || !a->pot_scales, "layernorm: a zero point needs the general kernel (pot_scales = 0)");
  if (a->in_gather) P2V_REQUIRE(a->gather_segs > 0 && a->C % (4 * a->gather_segs) == 0 && !a->out_f32,
                                "layernorm: gathered input needs C %% (4 * gather_segs) == 0 and int8 output");
  return launch_layernorm(*a, (cudaStream_t)stream);
}
int p2v_int_softmax_log2(const int8_t* scores, uint8_t* out, int64_t rows, int n, const p2v_softmax_lut* lut, void* stream) {
  P2V_REQUIRE(scores && out && lut && rows > 0 && n > 0 && n <= 1024, "int_softmax: bad arguments (n <= 1024)");
  return launch_softmax(scores, out, rows, n, lut, (cudaStream_t)stream);
}
static int validate_attention(const p2v_attention_args* a) {
  P2V_REQUIRE(a && a->qkv && a->out && a->lut_dev, "attention: missing pointers");
  P2V_REQUIRE(a->B > 0 && a->H > 0 && a->T > 0 && a->T <= 256, "attention: T=%d unsupported (1..256)", a->T);
  P2V_REQUIRE(a->dh == 64 || a->dh == 32, "attention: head dim %d unsupported (32 or 64)", a->dh);
  return 0;
}
int p2v_attention_i8(const p2v_attention_args* a, void* stream) {
  if (int r = validate_attention(a)) return r;
  if (attention_tc_supported(*a)) return launch_attention_tc(*a, (cudaStream_t)stream);
  return launch_attention(*a, (cudaStream_t)stream);
}
int p2v_attention_i8_simt(const p2v_attention_args* a, void* stream) {
  if (int r = validate_attention(a)) return r;
  return launch_attention(*a, (cudaStream_t)stream);
}
static int validate_window_attention(const p2v_window_attention_args* a) {
  P2V_REQUIRE(a && a->qkv && a->out && a->bias && a->lut_dev, "window_attention: missing pointers");
  P2V_REQUIRE(a->n_windows > 0 && a->H > 0 && a->T > 0 && a->T <= 64, "window_attention: T=%d unsupported (1..64)", a->T);
  P2V_REQUIRE(a->dh == 32 || a->dh == 64, "window_attention: head dim %d unsupported (32 or 64)", a->dh);
  P2V_REQUIRE(a->windows_per_image > 0 && a->mask_code <= 0, "window_attention: bad mask arguments");
  P2V_REQUIRE((reinterpret_cast<uintptr_t>(a->qkv) & 15) == 0, "window_attention: qkv must be 16-byte aligned");
  return 0;
}
int p2v_window_attention_i8(const p2v_window_attention_args* a, void* stream) {
  if (int r = validate_window_attention(a)) return r;
  if (window_attention_tc_supported(*a)) return launch_window_attention_tc(*a, a->mask_exp_int, (cudaStream_t)stream);
  return launch_window_attention(*a, a->mask_exp_int, (cudaStream_t)stream);
}
int p2v_window_attention_i8_simt(const p2v_window_attention_args* a, void* stream) {
  if (int r = validate_window_attention(a)) return r;
  return launch_window_attention(*a, a->mask_exp_int, (cudaStream_t)stream);
}
int p2v_gather_rows_i8(const int8_t* in, int8_t* out, const int32_t* src_rows, int rows_out, int segs, int C, void* stream) {
  P2V_REQUIRE(in && out && src_rows && rows_out > 0 && segs > 0 && C > 0 && C % 16 == 0, "gather_rows: bad arguments (C %% 16 == 0)");
  P2V_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "gather_rows: 16-byte alignment");
  return launch_gather_rows(in, out, src_rows, rows_out, segs, C, (cudaStream_t)stream);
}
int p2v_avgpool_quant_i8(const int8_t* in, int8_t* out, int B, int T, int C, float s_in, float s_out, void* stream) {
  P2V_REQUIRE(in && out && B > 0 && T > 0 && C > 0 && s_in > 0.f && s_out > 0.f, "avgpool_quant: bad arguments");
  return launch_avgpool_quant(in, out, B, T, C, s_in, s_out, (cudaStream_t)stream);
}
int64_t p2v_minmax_scratch_bytes(int64_t n, int C, int64_t inner) {
  if (n <= 0 || C <= 0 || inner <= 0) return 0;
  return minmax_scratch_bytes(n, C, inner);
}
int p2v_minmax_per_channel(const float* x, float* minmax, int64_t n, int C, int64_t inner, float* scratch, void* stream) {
  P2V_REQUIRE(x && minmax && scratch && n > 0 && C > 0 && inner > 0 && n % (int64_t(C) * inner) == 0, "minmax: bad arguments");
  return launch_minmax(x, minmax, n, C, inner, scratch, (cudaStream_t)stream);
}
int64_t p2v_quant_mse_scratch_bytes(int64_t n, int C, int64_t inner, int K, int per_channel_out) {
  if (n <= 0 || C <= 0 || inner <= 0 || K <= 0) return 0;
  return mse_scratch_bytes(n, C, inner, K, per_channel_out);
}
int p2v_quant_mse_scores(const float* x, int64_t n, int C, int64_t inner, const float* scales, const float* zps, int K,
                         int n_scale, int per_channel_out, int lo, int hi, double* out, double* scratch, void* stream) {
  P2V_REQUIRE(x && scales && out && scratch && n > 0 && C > 0 && inner > 0 && K > 0 && K <= 96, "mse_scores: bad arguments (K <= 96)");
  P2V_REQUIRE(n_scale == 1 || n_scale == C, "mse_scores: n_scale must be 1 or C");
  return launch_mse_scores(x, n, C, inner, scales, zps, K, n_scale, per_channel_out, lo, hi, out, scratch, (cudaStream_t)stream);
}
int64_t p2v_linear_sqerr_scratch_bytes(int M, int n) { return (M > 0 && n > 0) ? linear_sqerr_scratch_bytes(M, n) : 0; }
int p2v_linear_sqerr_scores(const float* x, int M, int K, int patch, int Cin, int H, int W, const float* D, int n, double* out,
                            double* scratch, void* stream) {
  P2V_REQUIRE(x && D && out && scratch && M > 0 && K > 0 && n > 0 && K % 4 == 0, "linear_sqerr: bad arguments (K %% 4 == 0)");
  P2V_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(D)) & 15) == 0, "linear_sqerr: 16-byte alignment");
  if (patch > 0) P2V_REQUIRE(patch % 4 == 0 && Cin > 0 && H % patch == 0 && W % patch == 0 && K == Cin * patch * patch &&
                             M % ((H / patch) * (W / patch)) == 0, "linear_sqerr: bad patch geometry");
  return launch_linear_sqerr(x, M, K, patch, Cin, H, W, D, n, out, scratch, (cudaStream_t)stream);
}
int p2v_linear_f32(const float* x, int M, int K, int patch, int Cin, int H, int W, const float* Wt, const float* bias, int N, float* out,
                   void* stream) {
  P2V_REQUIRE(x && Wt && out && M > 0 && K > 0 && N > 0 && K % 4 == 0, "linear_f32: bad arguments (K %% 4 == 0)");
  P2V_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(Wt)) & 15) == 0, "linear_f32: 16-byte alignment");
  if (patch > 0) P2V_REQUIRE(patch % 4 == 0 && Cin > 0 && H % patch == 0 && W % patch == 0 && K == Cin * patch * patch &&
                             M % ((H / patch) * (W / patch)) == 0, "linear_f32: bad patch geometry");
  return launch_linear_f32(x, M, K, patch, Cin, H, W, Wt, bias, N, out, (cudaStream_t)stream);
}
int p2v_embed_f32(const float* img, int B, int Cin, int H, int W, int P, const float* w_hat, const float* bias, int N, float mid_scale,
                  float mid_zp, float aux_scale, float aux_zp, const float* pos, const float* out_scale, int8_t* out, void* stream) {
  P2V_REQUIRE(img && w_hat && bias && pos && out_scale && out && B > 0 && Cin > 0 && P > 0 && P % 4 == 0 && H % P == 0 && W % P == 0 && N % 4 == 0,
              "embed_f32: bad arguments (P %% 4 == 0, N %% 4 == 0)");
  P2V_REQUIRE(((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(w_hat)) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
              "embed_f32: alignment");
  SgemmEmbed ep{bias, nullptr, pos, out_scale, mid_scale, mid_zp, aux_scale, aux_zp, (H / P) * (W / P), out};
  return launch_embed_f32(img, B, Cin, H, W, P, w_hat, N, ep, (cudaStream_t)stream);
}
int p2v_radix_hist_f32(const float* x, int64_t n, uint32_t prefix_mask, uint32_t prefix_value, int shift, int nbits,
                       unsigned long long* hist, void* stream) {
  P2V_REQUIRE(x && hist && n > 0 && nbits >= 1 && nbits <= 12 && shift >= 0 && shift + nbits <= 32, "radix_hist: bad arguments");
  P2V_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "radix_hist: x must be 16-byte aligned");
  return launch_radix_hist(x, n, prefix_mask, prefix_value, shift, nbits, hist, (cudaStream_t)stream);
}

}  // extern "C"
