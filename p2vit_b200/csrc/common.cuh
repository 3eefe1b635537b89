// Shared device helpers for the p2vit_b200 kernels (sm_100a).
//
// Bit-exactness rules (DESIGN.md "numerics"): every fp32 operation that the reference performs as a
// separate ATen op is performed here as one correctly rounded IEEE operation through the __f*_rn
// intrinsics (never contracted into an FMA by the compiler); torch.round == rintf (half to even).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/p2vit_b200.h"

namespace p2v {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);

#define P2V_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      p2v::set_error(__VA_ARGS__);  \
      return 1;                     \
    }                               \
  } while (0)

// Programmatic dependent launch.  Every kernel of the forward is launched with the programmatic-stream-serialization attribute
// (launch_pdl) and runs  <prologue on constants: barrier init, TMEM allocation, tables> ; pdl_wait() ; pdl_trigger() ; <work>.
// pdl_wait returns when the preceding kernel has completed and its stores are visible, so everything that reads activations or
// writes a buffer comes after it; pdl_trigger (after the wait: at most one kernel runs ahead) lets the next kernel's CTAs take an SM
// as soon as this kernel's CTA on it exits, and run their prologue under this kernel's tail.  Every PDL-launched kernel must reach
// pdl_wait - a kernel that finished without it would let its successor overtake its predecessor.  The attribute is set for launches
// captured into a CUDA graph (api.cu: pdl_mode; P2VIT_PDL=0 never, =1 always); without it the device instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
enum { PDL_GEMM = 1, PDL_ATTENTION = 2, PDL_LAYERNORM = 4, PDL_OTHER = 8 };     // P2VIT_PDL_KINDS masks kernel families (experiments)
bool pdl_enabled(cudaStream_t stream, int kind);
void pdl_next_kind(int kind);      // family of the next launch_pdl on this thread (reset to PDL_OTHER by the launch)
int pdl_take_kind();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled(stream, pdl_take_kind()) ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// clamp(RNE(v), -128, 127) as int; NaN -> 0 like cvt.sat
__device__ __forceinline__ int sat_s8(float v) {
  int r;
  asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}

// q = sat(RNE(x / s)) : POT => s is a power of two and rs = 1/s, so x*rs == x/s exactly
template <bool POT>
__device__ __forceinline__ int quant_s8(float x, float s, float rs) {
  return sat_s8(POT ? fmul(x, rs) : fdiv(x, s));
}

// Exactly rounded a / b from rb = RN(1 / b) without the division sequence: q0 = a * rb, then two residual corrections
// q <- q + (a - b q) * rb with the residual exact in one FFMA.  The first leaves q within half an ulp (+ 2^-40) of a / b, the
// second then rounds correctly (Markstein's theorem for a correctly rounded reciprocal and a faithful quotient).  Zero numerators
// give zero; the operands must be far from the ends of the exponent range (scales and activations are).
__device__ __forceinline__ float div_rb(float a, float b, float rb) {
  const float q0 = fmul(a, rb);
  const float q1 = __fmaf_rn(__fmaf_rn(-b, q0, a), rb, q0);
  return __fmaf_rn(__fmaf_rn(-b, q1, a), rb, q1);
}

__device__ __forceinline__ uint32_t pack4_s8(int a, int b, int c, int d) {
  return (uint32_t(a) & 0xffu) | ((uint32_t(b) & 0xffu) << 8) | ((uint32_t(c) & 0xffu) << 16) | (uint32_t(d) << 24);
}

// ---- rounding / saturation without the XU pipe (F2I and FRND run on 16 lanes per SM)
constexpr float RMAGIC = 12582912.f;        // 1.5 * 2^23: x + RMAGIC has RNE(x) in its low mantissa bits for |x| < 2^22
// four magic-biased floats (RMAGIC + RNE(x), not clamped) -> four saturated int8 codes in one word.  The bit pattern of
// x + RMAGIC is monotone in x over all finite x (the sum is positive above -1.5*2^23, sign bit set below), so
// sat_s8(bits - bits(RMAGIC)) == clamp(RNE(x), -128, 127) whatever the magnitude of x; cvt.pack.sat saturates two
// values per instruction (a -> byte 0, b -> byte 1, c's low half -> the upper half of the result).
__device__ __forceinline__ uint32_t pack4_sat(float a, float b, float c, float d) {
  const int ia = int(__float_as_uint(a)) - 0x4B400000, ib = int(__float_as_uint(b)) - 0x4B400000;
  const int ic = int(__float_as_uint(c)) - 0x4B400000, id = int(__float_as_uint(d)) - 0x4B400000;
  uint32_t hi, out;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(id), "r"(ic), "r"(0));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(ib), "r"(ia), "r"(hi));
  return out;
}


// torch's nn.GELU() (erf form): x * 0.5 * (1 + erf(x * sqrt(1/2)))   (ATen cpu/Activation: vectorized
// x * kAlpha -> erf -> +1 -> * x * 0.5).  erff differs from Sleef's by <= 1 ulp on rare inputs; those
// only matter at rounding ties (DESIGN.md "tie adjudication").
__device__ __forceinline__ float gelu_erf(float x) {
  const float kAlpha = 0.70710678118654752440f;
  float e = erff(fmul(x, kAlpha));
  return fmul(fmul(x, 0.5f), fadd(1.0f, e));
}

// ---- GELU step table (include/p2vit_b200.h: p2v_build_gelu_table)
struct GeluTabHeader { float y0, inv_w; int n, reserved; };
struct GeluTab { const uint2* entries; int n; float inv_w, off; };   // entries may live in shared or global memory
// the reference arithmetic: qact1(gelu(y)) for a power-of-two output scale (ro = 1/out_scale)
__device__ __forceinline__ int gelu_code_direct(float y, float ro) { return sat_s8(fmul(gelu_erf(y), ro)); }
// the same for any output scale: the reference's division (equal to the product above when the scale is a power of two)
__device__ __forceinline__ int gelu_code_div(float y, float so, float zp = 0.f) { return sat_s8(fadd(fdiv(gelu_erf(y), so), zp)); }
// segment of y; the SAME expression builds the table and looks it up, so its own rounding is immaterial
__device__ __forceinline__ int gelu_segment(float y, float inv_w, float off, int n) {
  return min(max(__float2int_rd(__fmaf_rn(y, inv_w, off)), 0), n - 1);
}
// code of y from its table entry; `slow` is raised when y is within 8 ulps of the entry's threshold or the entry is flagged
__device__ __forceinline__ int gelu_code_table(float y, uint2 e, bool& slow) {
  const int d = __float_as_int(y) - int(e.x);
  slow |= (unsigned(d + 8) <= 16u) | (int(e.y) < 0);
  const uint32_t sel = y >= __uint_as_float(e.x) ? 0x9991u : 0x8880u;     // prmt selectors: sign-extended byte 1 / byte 0
  return int(__byte_perm(e.y, 0u, sel));
}

// ---- GELU step tables, second form (CTA-pair GEMM epilogue; include/p2vit_b200.h: p2v_build_gelu_table)
// code(y) = sat(RNE(gelu_erf(y) / so)) is a step function: non-increasing for y < y* (the minimum of GELU) and non-decreasing
// above.  A per-segment linear map P(y) ~ gelu(y)/so (error < 0.4) gives f = RNE(P - 1/2), so the code is f or f + 1, and ONE
// exact threshold decides: code = f + [code(y) >= f + 1], where [code(y) >= c] is  y >= thrR[c]  right of y* and  y < thrL[c]
// left of it.  Thresholds are exact fp32 values found by bisection on the kernels' own gelu_erf; y within 8 ulps of the
// threshold consulted is sent to the direct evaluation (erff is not monotone at the ulp level).  Both tables sit in shared
// memory replicated per lane (entry i of lane l at i * stride + 4 l), so the two data-dependent lookups are free of bank
// conflicts - the reason the first form (one 8-byte entry per y-segment, ~5-way conflicts) gained almost nothing.
struct GeluStepsHeader {   // 64 bytes, then float2 seg[P2V_GELU_STEPS_MAX_SEG], float thr[P2V_GELU_STEPS_MAX_THR]
  float ymin, ymax, inv_w, soff, ystar;
  int nseg, nr, nl, k1, rep_log2, ok;
  float seg_scale;         // nseg - 1 + 0.49: segment = RNE(sat((y * inv_w + soff) / seg_scale) * seg_scale)
  float f_scale;           // 126 - f0 + 0.49 (f0 = -k1): f = f0 + RNE(sat(A' y + B') * f_scale), seg = (A', B') = (A, B - f0) / f_scale
  int clean;               // 1: the step code equals the direct evaluation also within 8 ulps of every threshold (no distance test needed)
  float zp;                // zero point of the output quantizer the table was built for (0 for symmetric observers)
  int pad[1];
};
bool gelu_table_is_clean(const void* table_dev);     // host (gelu_table.cu): verdict of the table's self-check, by device address
constexpr int P2V_GELU_STEPS_MAX_SEG = 64, P2V_GELU_STEPS_MAX_THR = 512;
constexpr int P2V_GELU_STEPS_OFFSET = 16 + 8 * P2V_GELU_TABLE_MAX_ENTRIES;                 // byte offset inside a p2v gelu table buffer
constexpr int P2V_GELU_STEPS_SMEM_MAX = 256 * P2V_GELU_STEPS_MAX_SEG + 26 * 1024;          // replicated tables: segments + thresholds
struct GeluSteps {         // per-lane view of the replicated tables (32-bit shared addresses, pre-biased by the magic constant)
  uint32_t seg_addr, thr_r, thr_l, thr_mul;
  float ymin, sa, sb, seg_scale, f_scale, f_magic, ystar;
};
__host__ __device__ inline uint32_t gelu_steps_smem_bytes(const GeluStepsHeader& h) {
  return uint32_t(h.nseg) * 256u + (uint32_t(h.nr + h.nl) << (2 + h.rep_log2));
}
// every thread of the block copies its share of the global table into the replicated shared-memory layout at `smem`
__device__ __forceinline__ void gelu_steps_fill_smem(const void* table, uint32_t smem, int tid, int nthreads) {
  const char* base = reinterpret_cast<const char*>(table) + P2V_GELU_STEPS_OFFSET;
  const GeluStepsHeader h = *reinterpret_cast<const GeluStepsHeader*>(base);
  const float2* seg = reinterpret_cast<const float2*>(base + sizeof(GeluStepsHeader));
  const float* thr = reinterpret_cast<const float*>(base + sizeof(GeluStepsHeader) + 8 * P2V_GELU_STEPS_MAX_SEG);
  for (int i = tid; i < h.nseg * 32; i += nthreads) {
    const float2 v = __ldg(seg + (i >> 5));
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(smem + uint32_t(i) * 8u), "f"(v.x), "f"(v.y) : "memory");
  }
  const uint32_t thr0 = smem + uint32_t(h.nseg) * 256u;
  const int n = (h.nr + h.nl) << h.rep_log2;
  for (int i = tid; i < n; i += nthreads) {
    const float v = __ldg(thr + (i >> h.rep_log2));
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(thr0 + uint32_t(i) * 4u), "f"(v) : "memory");
  }
}
__device__ __forceinline__ float fma_sat(float a, float b, float c) {     // clamp(fl(a*b + c), 0, 1) in one FMA-pipe instruction
  float r;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ GeluSteps gelu_steps_view(const void* table, uint32_t smem, int lane) {
  const GeluStepsHeader h = *reinterpret_cast<const GeluStepsHeader*>(reinterpret_cast<const char*>(table) + P2V_GELU_STEPS_OFFSET);
  GeluSteps t;
  const uint32_t shift = uint32_t(2 + h.rep_log2);
  t.thr_mul = 1u << shift;
  t.seg_addr = smem + uint32_t(lane) * 8u - (0x4B400000u << 8);
  t.thr_r = smem + uint32_t(h.nseg) * 256u + (uint32_t(lane) & ((1u << h.rep_log2) - 1u)) * 4u + ((uint32_t(h.k1) - 0x4B400000u) << shift);
  t.thr_l = t.thr_r + (uint32_t(h.nr) << shift);
  // the biased addresses stay opaque to the compiler: it would otherwise split the bias off and add it back per column
  asm volatile("" : "+r"(t.seg_addr), "+r"(t.thr_r), "+r"(t.thr_l), "+r"(t.thr_mul));
  t.ymin = h.ymin; t.ystar = h.ystar;
  t.seg_scale = h.seg_scale; t.sa = h.inv_w / h.seg_scale; t.sb = h.soff / h.seg_scale;
  t.f_scale = h.f_scale; t.f_magic = RMAGIC - float(h.k1);             // RMAGIC + f0, f0 = cr0 - 1 = -k1
  return t;
}
// RMAGIC-biased code of y: the low byte of the result is the int8 code (0x4B400000 has a zero low byte and the code lies in
// [-128, 127] by construction: f is capped at 126 by the FFMA.SAT, the threshold of code 127 adds the last step).
// Written for the ALU / FMA pipe split: both clamps of the table indices are FFMA.SATs (the segment entry holds A / f_scale and
// (B - f0) / f_scale), the indices leave an FFMA already biased and become addresses with one IMAD each.
// near_min collects min(bits(y) - bits(threshold) + 8) as unsigned: a value <= 16 means some y was within 8 ulps of the
// threshold consulted and the caller must evaluate erf directly.
// GUARD = false: the table's self-check found every threshold a clean step of the direct evaluation (GeluStepsHeader.clean), so
// the distance test - three ALU-pipe instructions per output - is not needed.
template <bool GUARD = true>
__device__ __forceinline__ uint32_t gelu_steps_code(float y, const GeluSteps& t, uint32_t& near_min) {
  const float yc = fmaxf(y, t.ymin);                     // left of ymin the code is constant; right of ymax the SATs hold the indices
  const uint32_t sb = __float_as_uint(__fmaf_rn(fma_sat(yc, t.sa, t.sb), t.seg_scale, RMAGIC));      // 0x4B400000 + segment
  float A, B;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(A), "=f"(B) : "r"(sb * 256u + t.seg_addr));
  const uint32_t fb = __float_as_uint(__fmaf_rn(fma_sat(A, yc, B), t.f_scale, t.f_magic));            // 0x4B400000 + f
  const bool left = yc < t.ystar;
  float thr;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(thr) : "r"(fb * t.thr_mul + (left ? t.thr_l : t.thr_r)));
  if (GUARD) near_min = min(near_min, __float_as_uint(yc) - __float_as_uint(thr) + 8u);
  return fb + (((yc >= thr) != left) ? 1u : 0u);
}
// low bytes of four words -> one word
__device__ __forceinline__ uint32_t pack4_low_bytes(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}
// four int32 codes -> four saturated int8 codes in one word (see pack4_sat)
__device__ __forceinline__ uint32_t pack4_sat_int(int a, int b, int c, int d) {
  uint32_t hi, out;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(b), "r"(a), "r"(hi));
  return out;
}

// floor(log2(|a|)) for finite non-zero a (normal or subnormal)
__device__ __forceinline__ int ilog2f(float a) {
  uint32_t u = __float_as_uint(a) & 0x7fffffffu;
  int e = int(u >> 23);
  if (e == 0) return -118 - __clz(u);  // subnormal: value = u * 2^-149 -> floor(log2) = 31 - clz(u) - 149
  return e - 127;
}

// floor(fl32(log2(a))) - what `torch.floor(torch.log2(a))` evaluates to (layers.py:272): the fp32 log2 of a value a few
// ulps below 2^k rounds UP to exactly k, so the floor is k, not k-1 (measured: torch's CPU log2 is correctly rounded at
// these points).  Only mantissas within 16 ulps of the next power of two can be affected; they take the fp64 path.
__device__ __forceinline__ int floor_log2_as_fp32(float a) {
  int e = ilog2f(a);
  if ((__float_as_uint(a) & 0x007fffffu) >= 0x007ffff0u) e = int(floorf(__double2float_rn(log2(double(a)))));
  return e;
}

__device__ __forceinline__ float pow2i(int e) {  // 2^e for -126 <= e <= 127
  return __uint_as_float(uint32_t(e + 127) << 23);
}

// QIntSoftmax tail (layers.py:376-381, 422-427): x = RNE(sum/exp); big = floor(log2 x) (+1 if x >= 1.5*2^big);
// returns min(big,15), or 255 if big >= 16 (probability forced to zero), x >= 1 always.
__device__ __forceinline__ uint32_t log2_code(float sum_f, float exp_f) {
  float x = rintf(fdiv(sum_f, exp_f));
  uint32_t u = __float_as_uint(x);
  if (u >= 0x7f800000u) return 255u;                   // inf / nan (exp == 0)
  int big = int((u + 0x00400000u) >> 23) - 127;        // mantissa >= 1.5 carries into the exponent
  if (big < 0) big = 0;                                // x == 0 cannot happen (exp <= sum); keep defined
  return big >= 16 ? 255u : uint32_t(big);
}

// log_round(RNE(fl(tot / e))) of layers.py:376-381,422-427 without the division, as 2^(15-big) (0 when big >= 16).
//   x = RNE(q), q = fl(tot/e) >= 1;  big(x) = #{t in {2, 3, 6, 12, 24, ...} : x >= t}, and x >= t <=> q + 1/2 >= t up to the
//   tie rule.  g(q) = min(2q - 1, (4q + 2)/3) is increasing, equals 2 at q + 1/2 = 2 and 2^(j+2) at q + 1/2 = 3 * 2^j, so
//   big = floor(log2 g) = the exponent field of g.  g is computed from the table's reciprocal with two FFMAs (a few ulps off
//   the exact value), so the result can only differ from the reference when g is within 16 ulps of a power of two: with
//   pf = 2^(15-E) (E = exponent of g, an integer subtraction on the exponent field), g * pf lies in [2^15, 2^16) and the
//   guard is |g*pf - 1.5*2^15| >= 2^14 - 16 ulps; `gmax` collects it over a unit, which is then redone with the IEEE
//   division (log2_code).  pf + 2^23 leaves the integer 2^(15-E) (0 from E = 16 on: 0.5 rounds to even) in the low mantissa bits.
__device__ __forceinline__ uint32_t shr_clamp(uint32_t v, uint32_t n) {   // PTX shr: amounts > 31 give 0
  uint32_t r;
  asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
  return r;
}
// Reciprocal of a table entry for prob_bits_fast.  exp_int = 0 (the far tail of the table) must not become +inf: g = inf trips the
// guard band and sends the WHOLE WARP's unit to the IEEE-division redo - with raw observer scales (percentile / ema: codes spread
// over the full int8 range) most units hold such an entry.  2^60 keeps g finite (row sums are >= exp_int[0] >= 2^32 and < 2^56), puts
// its exponent far above 16, and the probability comes out 0 exactly as the reference's x = inf does.
__device__ __forceinline__ float prob_rcp(float e) { return e > 0.f ? fdiv(1.0f, e) : 1152921504606846976.f; }
constexpr float PROB_GUARD = 16384.f - 0.0625f;     // 16 ulps of [2^15, 2^16)
__device__ __forceinline__ uint32_t prob_bits_fast(float tot2, float tot43, float rcp, float& gmax) {
  const float g = fminf(__fmaf_rn(tot2, rcp, -1.0f), __fmaf_rn(tot43, rcp, 0.666666686534881591796875f));
  const float pf = __uint_as_float(0x86800000u - (__float_as_uint(g) & 0x7F800000u));     // 2^(15 - E), E = floor(log2 g)
  gmax = fmaxf(gmax, fabsf(__fmaf_rn(g, pf, -49152.f)));       // g * pf is exact (pf is a power of two), and so is the difference
  return __float_as_uint(fadd(pf, 8388608.f));      // low 16 bits: 2^(15-E)
}
// The same probability with no guard band: q = fl(sum / exp) exactly (div_rb from the table's reciprocal), x = RNE(q), and
// g' = (4x + 1) / 3 has floor(log2 g') = log_round(x) for every integer x >= 1 (x = 3 2^j - 1 gives 2^(j+2) - 1, x = 3 2^j gives
// 2^(j+2) + 1/3; the fp32 error of the FFMA stays below 0.2 up to the last threshold that matters, x = 49152).  Three FMA-pipe
// instructions more per score than prob_bits_fast, but no redo: for rows whose quotients keep landing exactly on the ties x.5.
__device__ __forceinline__ uint32_t prob_bits_div(float tot, float rcp, float e) {
  const float q = div_rb(tot, e, rcp);
  const float x = fsub(fadd(q, RMAGIC), RMAGIC);                     // RNE(q) below 2^22; beyond, any value that large gives 0
  const float g = __fmaf_rn(x, 1.33333337306976318359375f, 0.3333333432674407958984375f);
  const float pf = __uint_as_float(0x86800000u - (__float_as_uint(g) & 0x7F800000u));
  return __float_as_uint(fadd(pf, 8388608.f));
}
// exactly rounded (RNE) fp32 of the 128-bit integer hi*2^32 + lo  (hi, lo < 2^63)
__device__ __forceinline__ float u96_to_f32(unsigned long long hi, unsigned long long lo) {
  unsigned __int128 v = ((unsigned __int128)hi << 32) + lo;
  unsigned long long top = (unsigned long long)(v >> 64);
  unsigned long long bot = (unsigned long long)v;
  if (top == 0) return __ull2float_rn(bot);
  int lz = __clzll(top);                                // top != 0
  int sh = 64 - lz;                                     // bits of v above bit 63 -> shift right by sh keeps 64 bits
  unsigned long long kept = (unsigned long long)(v >> sh);
  unsigned long long lost = bot & ((1ull << sh) - 1ull);
  if (lost) kept |= 1ull;                               // sticky bit keeps RNE correct
  return fmul(__ull2float_rn(kept), pow2i(sh));
}

}  // namespace p2v
