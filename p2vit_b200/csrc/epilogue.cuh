// Fused GEMM epilogues: one thread owns NC consecutive output columns of one row (int32 accumulators in
// registers, read from TMEM by the tcgen05 kernel or produced by dp4a in the SIMT cross-check kernel).
// Reference call sites per mode are listed in include/p2vit_b200.h (p2v_epilogue_t).
//
// Per-column parameters are staged once per output tile into shared memory as a small struct-of-arrays
// (ColParams) together with derived values (reciprocals, folded scales), so the per-element work is pure
// register math + broadcast LDS.128.  Divisions by non-power-of-two scales use quant_div(): an exact
// shortcut (multiply by the correctly rounded reciprocal, round) that falls back to the IEEE division only
// when the quotient is within 1e-3 of a rounding tie - the result is always identical to
// rint(__fdiv_rn(y, s)) (see DESIGN.md "exact fast requantisation").
#pragma once
#include <cstdlib>
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
  const int32_t* row_map;
  const float* pos;
  float aux_scale;
  int tokens_per_image;
  int8_t* out_i8;
  float* out_f32;
  const void* gelu_table;   // device p2v_gelu_table or NULL
  float out_zp, mid_zp, aux_zp;   // zero points of asymmetric QActs (0 otherwise; general epilogues only)
  int dbg;   // P2V_DBG experiments (perf triage only): bit 0 = skip the global stores, bit 1 = skip residual loads
};

inline EpiParams make_epi_params(const p2v_gemm_args& a) {
  EpiParams p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.acc_scale = a.acc_scale; p.bias = a.bias; p.zp_corr = a.zp_corr; p.out_scale = a.out_scale;
  p.mid_scale = a.mid_scale; p.res_scale = a.res_scale; p.res = a.res; p.pos = a.pos;
  p.row_map = a.row_map;
  p.gelu_table = a.gelu_table;
  p.aux_scale = a.aux_scale; p.tokens_per_image = a.tokens_per_image;
  p.out_i8 = a.out_i8; p.out_f32 = a.out_f32;
  p.out_zp = a.out_zp; p.mid_zp = a.mid_zp; p.aux_zp = a.aux_zp;
  static const int dbg = getenv("P2V_DBG") ? atoi(getenv("P2V_DBG")) : 0;
  p.dbg = dbg;
  return p;
}

// sat(RNE(fl(y / s))) with rs = RN(1/s).  |y*rs - fl(y/s)| <= |q| * 3 * 2^-24 (< 2.4e-5 for |q| < 129), so away from
// a tie the rounded integers agree; within the guard band (`slow` is raised; ~6e-5 of the quotients, plus harmless
// false alarms for saturated |q|) the caller redoes the span with EXACT = true, i.e. with the IEEE quotient.  No branch per element: the fast pass is straight-line code
// (RNE through the 1.5*2^23 magic constant - exact for |q| < 2^22, saturating beyond), so the compiler can interleave
// the 16 / 32 independent columns a thread owns.
// `zp`: zero point of an asymmetric quantizer (uniform.py:83-86: (x / scale + zp).round()), an integer in [-128,127]: |q| < 256
// before saturation, so |y*rs - fl(y/s)| < 4.6e-5 and the sum fl(. + zp) adds at most one ulp of 256 (1.5e-5) on either side - the
// guard band is 1e-4 wide.  zp = 0 adds nothing.
template <bool EXACT>
__device__ __forceinline__ float quant_div(float y, float s, float rs, bool& slow, float zp = 0.f) {
  float k;
  if (EXACT) {
    k = rintf(fadd(fdiv(y, s), zp));
  } else {
    const float qa = fadd(fmul(y, rs), zp);
    k = rintf(qa);
    slow |= fabsf(fsub(qa, k)) > 0.4999f;
  }
  return fminf(fmaxf(k, -128.f), 127.f);
}
// same, result as int8 code (one saturating conversion instead of round + clamp + convert)
template <bool EXACT>
__device__ __forceinline__ int quant_div_s8(float y, float s, float rs, bool& slow, float zp = 0.f) {
  if (EXACT) return sat_s8(fadd(fdiv(y, s), zp));
  const float qa = fadd(fmul(y, rs), zp);
  slow |= fabsf(fsub(qa, rintf(qa))) > 0.4999f;
  return sat_s8(qa);
}

// column-parameter tile in shared memory: CP_ROWS arrays of BN floats
constexpr int CP_ROWS = 8;
enum { CP_S = 0, CP_B = 1, CP_O = 2, CP_RO = 3, CP_M = 4, CP_RM = 5, CP_RS = 6, CP_Z = 7 };

// thread `t` of a group stages column n = n0 + t (t < BN)
template <int EPI, bool POT, int BN>
__device__ __forceinline__ void stage_col_params(const EpiParams& p, float* cp, int n0, int t) {
  if (t >= BN) return;
  const int n = n0 + t;
  const bool ok = n < p.N;
  const float s = ok ? __ldg(p.acc_scale + n) : 0.f;
  const float b = (ok && p.bias) ? __ldg(p.bias + n) : 0.f;
  const float o = (ok && EPI != P2V_EPI_F32) ? __ldg(p.out_scale + n) : 1.f;
  const float ro = fdiv(1.f, o);
  if (POT && (EPI == P2V_EPI_REQUANT || EPI == P2V_EPI_DEQUANT)) {
    // y/o == fma(acc, s*ro, b*ro) exactly: ro is a power of two, scaling commutes with rounding
    cp[CP_S * BN + t] = fmul(s, ro);
    cp[CP_B * BN + t] = fmul(b, ro);
  } else {
    cp[CP_S * BN + t] = s;
    cp[CP_B * BN + t] = b;
  }
  cp[CP_O * BN + t] = o;
  cp[CP_RO * BN + t] = ro;
  if (EPI == P2V_EPI_RESIDUAL) {
    const float m = ok ? __ldg(p.mid_scale + n) : 1.f;
    cp[CP_M * BN + t] = m;
    cp[CP_RM * BN + t] = fdiv(1.f, m);
    cp[CP_RS * BN + t] = ok ? __ldg(p.res_scale + n) : 0.f;
  }
  cp[CP_Z * BN + t] = (ok && p.zp_corr) ? __int_as_float(__ldg(p.zp_corr + n)) : __int_as_float(0);
}

// Epilogue arithmetic for columns [col0, col0+NC) of one row; EXACT = false is the branch-free fast pass.
template <int EPI, bool POT, int BN, int NC, bool EXACT>
__device__ __forceinline__ void epilogue_math(const EpiParams& p, const float* cp, int row, int col0, int c0, const int (&acc)[NC],
                                              const uint32_t (&resw)[NC / 4], int (&q)[NC], float (&f)[NC], bool& slow, const GeluTab& gt) {
  const int N = p.N;
  const bool has_zp = p.zp_corr != nullptr;
  float e_sm = 0.f, e_rsm = 0.f, e_raux = 0.f;
  int tok = 0;
  if (EPI == P2V_EPI_EMBED) {
    e_sm = __ldg(p.mid_scale);
    e_rsm = fdiv(1.f, e_sm);
    e_raux = fdiv(1.f, p.aux_scale);
    tok = row % p.tokens_per_image;
  }
#pragma unroll
  for (int j4 = 0; j4 < NC; j4 += 4) {
    const float4 S4 = *reinterpret_cast<const float4*>(cp + CP_S * BN + c0 + j4);
    const float4 B4 = *reinterpret_cast<const float4*>(cp + CP_B * BN + c0 + j4);
    const float4 O4 = *reinterpret_cast<const float4*>(cp + CP_O * BN + c0 + j4);
    const float4 R4 = *reinterpret_cast<const float4*>(cp + CP_RO * BN + c0 + j4);
    const float Sv[4] = {S4.x, S4.y, S4.z, S4.w}, Bv[4] = {B4.x, B4.y, B4.z, B4.w};
    const float Ov[4] = {O4.x, O4.y, O4.z, O4.w}, Rv[4] = {R4.x, R4.y, R4.z, R4.w};
    float Mv[4], RMv[4], RSv[4];
    if (EPI == P2V_EPI_RESIDUAL) {
      const float4 M4 = *reinterpret_cast<const float4*>(cp + CP_M * BN + c0 + j4);
      const float4 RM4 = *reinterpret_cast<const float4*>(cp + CP_RM * BN + c0 + j4);
      const float4 RS4 = *reinterpret_cast<const float4*>(cp + CP_RS * BN + c0 + j4);
      Mv[0] = M4.x; Mv[1] = M4.y; Mv[2] = M4.z; Mv[3] = M4.w;
      RMv[0] = RM4.x; RMv[1] = RM4.y; RMv[2] = RM4.z; RMv[3] = RM4.w;
      RSv[0] = RS4.x; RSv[1] = RS4.y; RSv[2] = RS4.z; RSv[3] = RS4.w;
    }
    int4 Z4 = make_int4(0, 0, 0, 0);
    if (has_zp) Z4 = *reinterpret_cast<const int4*>(cp + CP_Z * BN + c0 + j4);
    const int Zv[4] = {Z4.x, Z4.y, Z4.z, Z4.w};
    float Pv[4] = {0.f, 0.f, 0.f, 0.f};
    if (EPI == P2V_EPI_EMBED) {     // position embedding of this token, four columns per 16-byte load (a lane = a row: one line per lane)
      const float* pp = p.pos + size_t(tok + 1) * N + col0 + j4;
      if ((N & 3) == 0 && col0 + j4 + 4 <= N) {
        const float4 P4 = __ldg(reinterpret_cast<const float4*>(pp));
        Pv[0] = P4.x; Pv[1] = P4.y; Pv[2] = P4.z; Pv[3] = P4.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col0 + j4 + e < N) Pv[e] = __ldg(pp + e);
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = j4 + e;
      const float af = float(acc[j] - Zv[e]);
      if (POT && (EPI == P2V_EPI_REQUANT || EPI == P2V_EPI_DEQUANT)) {
        const float yq = __fmaf_rn(af, Sv[e], Bv[e]);        // == fl(fl(acc*s) + b) / o  (s, o powers of two)
        q[j] = sat_s8(yq);
        if (EPI == P2V_EPI_DEQUANT) f[j] = fmul(float(q[j]), Ov[e]);
        continue;
      }
      // RESIDUAL with POT: acc_scale is a power of two, so acc*s is exact and fl(acc*s + b) is one fused rounding
      const float y = (EPI == P2V_EPI_RESIDUAL && POT) ? __fmaf_rn(af, Sv[e], Bv[e]) : fadd(fmul(af, Sv[e]), Bv[e]);
      if (EPI == P2V_EPI_F32) {
        f[j] = y;
      } else if (EPI == P2V_EPI_REQUANT) {
        q[j] = quant_div_s8<EXACT>(y, Ov[e], Rv[e], slow, p.out_zp);
      } else if (EPI == P2V_EPI_DEQUANT) {
        const float k = quant_div<EXACT>(y, Ov[e], Rv[e], slow, p.out_zp);
        q[j] = int(k);
        f[j] = fmul(fsub(k, p.out_zp), Ov[e]);
      } else if (EPI == P2V_EPI_GELU) {
        if (POT && !EXACT && gt.entries != nullptr) {
          q[j] = gelu_code_table(y, gt.entries[gelu_segment(y, gt.inv_w, gt.off, gt.n)], slow);
        } else {
          const float g = gelu_erf(y);
          q[j] = POT ? sat_s8(fmul(g, Rv[e])) : quant_div_s8<EXACT>(g, Ov[e], Rv[e], slow, p.out_zp);
        }
      } else if (EPI == P2V_EPI_RESIDUAL) {
        const float c = quant_div<EXACT>(y, Mv[e], RMv[e], slow);
        const float t = fmul(c, Mv[e]);
        const float r = float(int(int8_t((resw[j >> 2] >> (8 * (j & 3))) & 0xffu)));
        const float z = fadd(fmul(r, RSv[e]), t);
        q[j] = quant_div_s8<EXACT>(z, Ov[e], Rv[e], slow);
      } else if (EPI == P2V_EPI_EMBED) {
        const float c = quant_div<EXACT>(y, e_sm, e_rsm, slow, p.mid_zp);
        const float ecode = quant_div<EXACT>(fmul(fsub(c, p.mid_zp), e_sm), p.aux_scale, e_raux, slow, p.aux_zp);
        const float v = fadd(fmul(fsub(ecode, p.aux_zp), p.aux_scale), Pv[e]);
        q[j] = quant_div_s8<EXACT>(v, Ov[e], Rv[e], slow);
      }
    }
  }
}

// residual codes of columns [col0, col0+NC) of `row` as packed words (RESIDUAL epilogue); issued by the caller ahead of
// use so that the global-load latency overlaps the previous chunk's arithmetic
template <int NC>
__device__ __forceinline__ void load_residual(const EpiParams& p, int row, int col0, uint32_t (&resw)[NC / 4]) {
  const int N = p.N;
  if (row >= p.M || col0 >= N || (p.dbg & 2)) {
#pragma unroll
    for (int j = 0; j < NC / 4; ++j) resw[j] = 0u;
    return;
  }
  const int8_t* rp = p.res + size_t(p.row_map ? __ldg(p.row_map + row) : row) * N + col0;
  if ((N & 15) == 0 && (NC & 15) == 0 && col0 + NC <= N) {
#pragma unroll
    for (int j = 0; j < NC / 16; ++j) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(rp) + j);
      resw[4 * j] = v.x; resw[4 * j + 1] = v.y; resw[4 * j + 2] = v.z; resw[4 * j + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < NC / 4; ++j) {
      uint32_t w = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (col0 + 4 * j + e < N) w |= (uint32_t(uint8_t(rp[4 * j + e])) << (8 * e));
      resw[j] = w;
    }
  }
}

// Processes columns [col0, col0+NC) of row `row`, col0 = n0 + c0 with c0 the offset inside the staged tile.
// NC is a multiple of 4.  Tail columns (>= N) are masked on store.  TMEM_WAIT: `acc` is the destination of an
// in-flight tcgen05.ld; the wait sits right before the first use.  `resw`: load_residual() of the same span.
template <int EPI, bool POT, int BN, int NC, bool TMEM_WAIT = false>
__device__ __forceinline__ void epilogue_row(const EpiParams& p, const float* cp, int row, int n0, int c0, const int (&acc)[NC],
                                             const uint32_t (&resw)[NC / 4], const GeluTab& gt) {
  const int N = p.N;
  const int col0 = n0 + c0;
  int q[NC];
  float f[NC];
  if (TMEM_WAIT) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (row >= p.M) return;
  bool slow = false;
  epilogue_math<EPI, POT, BN, NC, false>(p, cp, row, col0, c0, acc, resw, q, f, slow, gt);
  if (slow) epilogue_math<EPI, POT, BN, NC, true>(p, cp, row, col0, c0, acc, resw, q, f, slow, gt);
  size_t orow = size_t(p.row_map ? __ldg(p.row_map + row) : row);
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
  if (p.dbg & 1) { if (q[0] == 12345) p.out_i8[0] = 1; return; }
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
      case P2V_EPI_RESIDUAL: if (POT_RT) { constexpr int EPI = P2V_EPI_RESIDUAL; constexpr bool POT = true;  __VA_ARGS__ } \
                             else        { constexpr int EPI = P2V_EPI_RESIDUAL; constexpr bool POT = false; __VA_ARGS__ } break; \
      case P2V_EPI_EMBED:    { constexpr int EPI = P2V_EPI_EMBED;    constexpr bool POT = false; __VA_ARGS__ } break; \
      case P2V_EPI_DEQUANT:  if (POT_RT) { constexpr int EPI = P2V_EPI_DEQUANT;  constexpr bool POT = true;  __VA_ARGS__ } \
                             else        { constexpr int EPI = P2V_EPI_DEQUANT;  constexpr bool POT = false; __VA_ARGS__ } break; \
      case P2V_EPI_F32:      { constexpr int EPI = P2V_EPI_F32;      constexpr bool POT = false; __VA_ARGS__ } break; \
      default: p2v::set_error("unknown epilogue %d", int(EPI_RT)); return 1;                           \
    }                                                                                                  \
  } while (0)

}  // namespace p2v
