// Fused GEMM epilogues: one thread owns NC consecutive output columns of one row (int32 accumulators in
// registers, read from TMEM by the tcgen05 kernel or produced by dp4a in the SIMT cross-check kernel).
// Reference call sites per mode are listed in include/p2vit_b200.h (p2v_epilogue_t).
#pragma once
#include "common.cuh"

namespace p2v {

struct EpiParams {  // device-side copy of p2v_gemm_args (pointers only)
  int M, N, K;
  const float* acc_scale;
  const float* bias;
  const int32_t* zp_corr;
  const float* out_scale;
  const float* mid_scale;
  const float* res_scale;
  const int8_t* res;
  const float* pos;
  float aux_scale;
  int tokens_per_image;
  int8_t* out_i8;
  float* out_f32;
};

inline EpiParams make_epi_params(const p2v_gemm_args& a) {
  EpiParams p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.acc_scale = a.acc_scale; p.bias = a.bias; p.zp_corr = a.zp_corr; p.out_scale = a.out_scale;
  p.mid_scale = a.mid_scale; p.res_scale = a.res_scale; p.res = a.res; p.pos = a.pos;
  p.aux_scale = a.aux_scale; p.tokens_per_image = a.tokens_per_image;
  p.out_i8 = a.out_i8; p.out_f32 = a.out_f32;
  return p;
}

// y = fl(acc * acc_scale + bias): acc*acc_scale is one rounding (exact for PoT), + bias a second one
__device__ __forceinline__ float acc_to_y(int acc, float s, float b) { return fadd(fmul(float(acc), s), b); }

// Processes columns [col0, col0+NC) of row `row`; NC is a multiple of 4; col0 % 4 == 0; N % 4 == 0 is NOT
// required (tail columns are masked element-wise on store).
template <int EPI, bool POT, int NC>
__device__ __forceinline__ void epilogue_row(const EpiParams& p, int row, int col0, const int (&acc)[NC]) {
  if (row >= p.M) return;
  const int N = p.N;
  int q[NC];
  float f[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int n = col0 + j;
    if (n >= N) { q[j] = 0; f[j] = 0.f; continue; }
    int a = acc[j];
    if (p.zp_corr) a -= __ldg(p.zp_corr + n);
    const float y = acc_to_y(a, __ldg(p.acc_scale + n), p.bias ? __ldg(p.bias + n) : 0.f);
    if (EPI == P2V_EPI_F32) { f[j] = y; continue; }
    const float so = __ldg(p.out_scale + n);
    const float rso = POT ? fdiv(1.f, so) : 0.f;
    if (EPI == P2V_EPI_REQUANT) {
      q[j] = quant_s8<POT>(y, so, rso);
    } else if (EPI == P2V_EPI_GELU) {
      q[j] = quant_s8<POT>(gelu_erf(y), so, rso);
    } else if (EPI == P2V_EPI_DEQUANT) {
      q[j] = quant_s8<POT>(y, so, rso);
      f[j] = fmul(float(q[j]), so);
    } else if (EPI == P2V_EPI_RESIDUAL) {
      const float sm = __ldg(p.mid_scale + n);
      const int c = sat_s8(fdiv(y, sm));
      const float t = fmul(float(c), sm);
      const int r = int(__ldg(p.res + size_t(row) * N + n));
      const float z = fadd(fmul(float(r), __ldg(p.res_scale + n)), t);
      q[j] = sat_s8(fdiv(z, so));
    } else if (EPI == P2V_EPI_EMBED) {
      const float sm = __ldg(p.mid_scale);
      const int c = sat_s8(fdiv(y, sm));
      const int e = sat_s8(fdiv(fmul(float(c), sm), p.aux_scale));
      const int tok = row % p.tokens_per_image;
      const float v = fadd(fmul(float(e), p.aux_scale), __ldg(p.pos + size_t(tok + 1) * N + n));
      q[j] = sat_s8(fdiv(v, so));
    }
  }
  size_t orow = size_t(row);
  if (EPI == P2V_EPI_EMBED) {
    const int T = p.tokens_per_image;
    orow = size_t(row / T) * (T + 1) + (row % T) + 1;
  }
  if (EPI == P2V_EPI_F32 || EPI == P2V_EPI_DEQUANT) {
    float* o = p.out_f32 + orow * N + col0;
    if ((N & 3) == 0 && col0 + NC <= N) {
#pragma unroll
      for (int j = 0; j < NC; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j) if (col0 + j < N) o[j] = f[j];
    }
    if (EPI == P2V_EPI_F32 || p.out_i8 == nullptr) return;
  }
  int8_t* o8 = p.out_i8 + orow * N + col0;
  if ((N & 15) == 0 && (NC & 15) == 0 && col0 + NC <= N) {
#pragma unroll
    for (int j = 0; j < NC; j += 16) {
      uint4 v;
      v.x = pack4_s8(q[j], q[j + 1], q[j + 2], q[j + 3]);
      v.y = pack4_s8(q[j + 4], q[j + 5], q[j + 6], q[j + 7]);
      v.z = pack4_s8(q[j + 8], q[j + 9], q[j + 10], q[j + 11]);
      v.w = pack4_s8(q[j + 12], q[j + 13], q[j + 14], q[j + 15]);
      *reinterpret_cast<uint4*>(o8 + j) = v;
    }
  } else if ((N & 3) == 0 && col0 + NC <= N) {
#pragma unroll
    for (int j = 0; j < NC; j += 4) *reinterpret_cast<uint32_t*>(o8 + j) = pack4_s8(q[j], q[j + 1], q[j + 2], q[j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < NC; ++j) if (col0 + j < N) o8[j] = int8_t(q[j]);
  }
}

// run-time (EPI, POT) -> compile-time dispatch used by both GEMM kernels' launchers
#define P2V_DISPATCH_EPI(EPI_RT, POT_RT, ...)                                                         \
  do {                                                                                                 \
    switch (EPI_RT) {                                                                                  \
      case P2V_EPI_REQUANT:  if (POT_RT) { constexpr int EPI = P2V_EPI_REQUANT;  constexpr bool POT = true;  __VA_ARGS__ } \
                             else        { constexpr int EPI = P2V_EPI_REQUANT;  constexpr bool POT = false; __VA_ARGS__ } break; \
      case P2V_EPI_GELU:     if (POT_RT) { constexpr int EPI = P2V_EPI_GELU;     constexpr bool POT = true;  __VA_ARGS__ } \
                             else        { constexpr int EPI = P2V_EPI_GELU;     constexpr bool POT = false; __VA_ARGS__ } break; \
      case P2V_EPI_RESIDUAL: { constexpr int EPI = P2V_EPI_RESIDUAL; constexpr bool POT = false; __VA_ARGS__ } break; \
      case P2V_EPI_EMBED:    { constexpr int EPI = P2V_EPI_EMBED;    constexpr bool POT = false; __VA_ARGS__ } break; \
      case P2V_EPI_DEQUANT:  if (POT_RT) { constexpr int EPI = P2V_EPI_DEQUANT;  constexpr bool POT = true;  __VA_ARGS__ } \
                             else        { constexpr int EPI = P2V_EPI_DEQUANT;  constexpr bool POT = false; __VA_ARGS__ } break; \
      case P2V_EPI_F32:      { constexpr int EPI = P2V_EPI_F32;      constexpr bool POT = false; __VA_ARGS__ } break; \
      default: p2v::set_error("unknown epilogue %d", int(EPI_RT)); return 1;                           \
    }                                                                                                  \
  } while (0)

}  // namespace p2v
