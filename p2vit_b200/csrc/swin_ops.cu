// Swin-specific kernels (reference: models/swin_quant.py).  Window partition, cyclic shift, window reverse are not
// kernels: they are row maps applied by the LayerNorm store (p2v_layernorm_args.out_row_map) and by the GEMM epilogue
// (p2v_gemm_args.row_map).  Here: the per-window attention (quantized relative-position bias and SW-MSA mask inside),
// the 2x2 patch-merging gather, and the token average pool + QAct of the classifier tail.
#include <climits>
#include <cmath>
#include <algorithm>
#include "common.cuh"

namespace p2v {

__device__ __forceinline__ int dp4a_us_w(uint32_t a_u8x4, uint32_t b_s8x4, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
  return d;
}

constexpr int WA_WARPS = 4;
constexpr int WA_MAXT = 64;

// Persistent CTAs over (window, head) units: T = ws*ws <= 64 tokens, one warp per query row, a lane owns keys lane and lane + 32.
// Windows are tiny (49 x 32 per head), so this is an integer-ALU / latency kernel; dp4a for both matmuls.  What the first version
// paid per unit and this one does not: the 3 KB softmax table reloaded by every CTA (233 k CTAs at batch 256), three XU-pipe
// conversions, one IEEE division for qact2 and one for log_round per score, and twenty 64-bit shuffles per row for the exact sum:
//   * the table {exp_int hi, lo, fp32, reciprocal} is loaded once per CTA, one LDS.128 per score;
//   * roundings / saturations go through the 1.5 * 2^23 constant (exact, common.cuh), the division by a power-of-two s_attn2 is
//     a multiplication (POT2; other observers keep __fdiv_rn);
//   * the row sum is three REDUX.ADDs: the hi words and the two 16-bit halves of the lo words each stay below 2^32;
//   * 2^(15-code) from prob_bits_fast (exponent-field shortcut with its guard band; log2_code when the guard trips).
template <int DH, bool POT2>
__global__ void __launch_bounds__(WA_WARPS * 32) window_attention_kernel(p2v_window_attention_args a, uint32_t e_mask, int units) {
  constexpr int DW = DH / 4, KSTR = DW + 1, TW = WA_MAXT / 4, VSTR = TW + 1;
  __shared__ uint32_t sQ[WA_MAXT * DW], sK[WA_MAXT * KSTR], sVt[DH * VSTR], sP[WA_WARPS * 2 * TW];
  __shared__ uint4 sLut[256];          // hi, lo, bits(exp_f32), bits(1 / exp_f32)
  __shared__ int8_t sLab[WA_MAXT];
  const int T = a.T, H = a.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row_bytes = int64_t(3) * H * DH;
  for (int i = tid; i < 256; i += blockDim.x) {
    const float e = a.lut_dev->exp_f32[i];
    sLut[i] = make_uint4(a.lut_dev->hi[i], a.lut_dev->lo[i], __float_as_uint(e), __float_as_uint(prob_rcp(e)));
  }
  for (int i = tid; i < DH * VSTR; i += blockDim.x) sVt[i] = 0u;       // rows >= T of V and keys >= T of P stay zero for every unit
  for (int i = tid; i < WA_WARPS * 2 * TW; i += blockDim.x) sP[i] = 0u;
  const float e_mask_f = float(e_mask), r_mask = prob_rcp(e_mask_f);
  const float r2 = fdiv(1.0f, a.s_attn2);
  constexpr float LO = RMAGIC - 128.f, HI = RMAGIC + 127.f;
  uint32_t* pHi = sP + warp * 2 * TW;
  uint32_t* pLo = pHi + TW;
  uint8_t* bHi = reinterpret_cast<uint8_t*>(pHi);
  uint8_t* bLo = reinterpret_cast<uint8_t*>(pLo);
  constexpr int CH = DH / 16;
  pdl_wait();          // tables and zero fills above are independent of the previous kernel (common.cuh)
  pdl_trigger();

  for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int win = unit / H, h = unit % H;
    const int8_t* base = a.qkv + int64_t(win) * T * row_bytes + h * DH;
    __syncthreads();                    // the previous unit's rows are done with sQ / sK / sVt (first pass: the tables are written)
    if (tid < WA_MAXT) sLab[tid] = (a.labels && tid < T) ? a.labels[(win % a.windows_per_image) * T + tid] : int8_t(0);
    for (int idx = tid; idx < T * CH; idx += blockDim.x) {
      const int r = idx / CH, ch = idx % CH;
      const int8_t* p = base + int64_t(r) * row_bytes + ch * 16;
      const uint4 q4 = __ldg(reinterpret_cast<const uint4*>(p));
      const uint4 k4 = __ldg(reinterpret_cast<const uint4*>(p + int64_t(H) * DH));
      const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(p + int64_t(2) * H * DH));
      uint32_t* dq = sQ + r * DW + ch * 4;
      dq[0] = q4.x; dq[1] = q4.y; dq[2] = q4.z; dq[3] = q4.w;
      uint32_t* dk = sK + r * KSTR + ch * 4;
      dk[0] = k4.x; dk[1] = k4.y; dk[2] = k4.z; dk[3] = k4.w;
      const uint32_t vv[4] = {v4.x, v4.y, v4.z, v4.w};
      uint8_t* vt = reinterpret_cast<uint8_t*>(sVt);
#pragma unroll
      for (int e = 0; e < 16; ++e) vt[size_t(ch * 16 + e) * VSTR * 4 + r] = uint8_t(vv[e >> 2] >> ((e & 3) * 8));
    }
    __syncthreads();

    const float* bias_h = a.bias + size_t(h) * T * T;
    for (int i = warp; i < T; i += WA_WARPS) {
      int x[2];
      bool masked[2];
      int mx = INT_MIN;
      const uint32_t* qi = sQ + i * DW;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = lane + 32 * jj;
        x[jj] = INT_MIN;
        masked[jj] = false;
        if (j < T) {
          const uint32_t* kj = sK + j * KSTR;
          int s = 0;
#pragma unroll
          for (int w = 0; w < DW; ++w) s = __dp4a(int(qi[w]), int(kj[w]), s);
          // qact_attn1: float(sat_s8(s * m)) without leaving fp32 (the biased sum is monotone in its argument)
          const float c1 = fsub(fminf(fmaxf(fadd(fmul(__int2float_rn(s), a.score_mult), RMAGIC), LO), HI), RMAGIC);
          const float v = fadd(fmul(c1, a.s_attn1), __ldg(bias_h + i * T + j));                 // + relative position bias
          const float q2 = POT2 ? fmul(v, r2) : fdiv(v, a.s_attn2);                             // qact2
          const int c2 = __float_as_int(fminf(fmaxf(fadd(q2, RMAGIC), LO), HI)) - 0x4B400000;
          masked[jj] = sLab[i] != sLab[j];
          x[jj] = c2 + (masked[jj] ? a.mask_code : 0);                                          // + mask (after the quantizer)
          mx = max(mx, x[jj]);
        }
      }
      mx = __reduce_max_sync(0xffffffffu, mx);
      uint32_t hi = 0, lo0 = 0, lo1 = 0;
      float ef[2] = {1.f, 1.f}, rc[2] = {1.f, 1.f};
#pragma unroll
      for (int jj = 0; jj < 2; ++jj)
        if (lane + 32 * jj < T) {
          uint32_t lo;
          if (masked[jj] && mx - x[jj] > 255) {   // beyond the table: the clamped tail (host guarantees mask_code reaches it)
            lo = e_mask;
            ef[jj] = e_mask_f;
            rc[jj] = r_mask;
          } else {
            const uint4 t = sLut[mx - x[jj]];
            hi += t.x;
            lo = t.y;
            ef[jj] = __uint_as_float(t.z);
            rc[jj] = __uint_as_float(t.w);
          }
          lo0 += lo & 0xffffu;
          lo1 += lo >> 16;
        }
      // exact row sum: <= 64 entries below 2^55, so each of the three partial sums stays below 2^32
      const unsigned long long hs = __reduce_add_sync(0xffffffffu, hi);
      const unsigned long long ls = static_cast<unsigned long long>(__reduce_add_sync(0xffffffffu, lo0)) +
                                    (static_cast<unsigned long long>(__reduce_add_sync(0xffffffffu, lo1)) << 16);
      const float tot = __ull2float_rn((hs << 32) + ls);      // < 64 * 2^55: fits 64 bits, rounded once
      const float tot2 = fmul(tot, 2.0f), tot43 = fmul(tot, 1.33333337306976318359375f);
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = lane + 32 * jj;
        if (j < T) {
          float gmax = 0.f;
          uint32_t pv = prob_bits_fast(tot2, tot43, rc[jj], gmax) & 0xffffu;
          if (!(gmax < PROB_GUARD)) pv = shr_clamp(0x8000u, log2_code(tot, ef[jj]));      // next to a rounding / log2 boundary: IEEE division
          bHi[j] = uint8_t(pv >> 8);
          bLo[j] = uint8_t(pv & 0xffu);
        }
      }
      __syncwarp();
      const int64_t orow = a.out_row_map ? int64_t(__ldg(a.out_row_map + int64_t(win) * T + i)) : int64_t(win) * T + i;
#pragma unroll
      for (int cc = 0; cc < DH / 32; ++cc) {
        const int c = lane + 32 * cc;
        const uint32_t* vt = sVt + c * VSTR;
        int ah = 0, al = 0;
        for (int w = 0; w < (T + 3) / 4; ++w) {
          const uint32_t v = vt[w];
          ah = dp4a_us_w(pHi[w], v, ah);
          al = dp4a_us_w(pLo[w], v, al);
        }
        a.out[orow * (H * DH) + h * DH + c] = int8_t(sat_s8(fmul(float(ah * 256 + al), a.out_mult)));
      }
      __syncwarp();
    }
  }
}

int launch_window_attention(const p2v_window_attention_args& a, uint32_t e_mask, cudaStream_t stream) {
  const int units = a.n_windows * a.H;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = std::min(units, sms * 16);       // 16 CTAs of 128 threads per SM, each walks units grid apart
  int ex = 0;
  const bool pot2 = std::frexp(a.s_attn2, &ex) == 0.5f;
#define P2V_WA(DH_) \
  if (pot2) launch_pdl(window_attention_kernel<DH_, true>, dim3(grid), dim3(WA_WARPS * 32), 0, stream, a, e_mask, units); \
  else launch_pdl(window_attention_kernel<DH_, false>, dim3(grid), dim3(WA_WARPS * 32), 0, stream, a, e_mask, units);
  if (a.dh == 32) { P2V_WA(32) } else { P2V_WA(64) }
#undef P2V_WA
  count_launch();
  return check_launch("window_attention_i8");
}

// ------------------------------------------------------------------------------------------------ patch-merging gather
__global__ void __launch_bounds__(256) gather_rows_kernel(const int8_t* __restrict__ in, int8_t* __restrict__ out,
                                                          const int32_t* __restrict__ src, int64_t total16, int segs, int C16) {
  pdl_wait();
  pdl_trigger();
  // one 16-byte chunk per thread: chunk index -> (output row, segment, chunk in segment)
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total16; i += int64_t(gridDim.x) * blockDim.x) {
    const int ch = int(i % C16);
    const int64_t rs = i / C16;                 // output row * segs + segment
    const int32_t s = __ldg(src + rs);
    reinterpret_cast<uint4*>(out)[i] = __ldg(reinterpret_cast<const uint4*>(in + int64_t(s) * C16 * 16) + ch);
  }
}

int launch_gather_rows(const int8_t* in, int8_t* out, const int32_t* src, int rows_out, int segs, int C, cudaStream_t stream) {
  const int64_t total16 = int64_t(rows_out) * segs * (C / 16);
  const int blocks = int(std::min<int64_t>((total16 + 255) / 256, 148 * 16));
  launch_pdl(gather_rows_kernel, dim3(blocks), dim3(256), 0, stream, in, out, src, total16, segs, C / 16);
  count_launch();
  return check_launch("gather_rows_i8");
}

// ------------------------------------------------------------------------------------------------ token average pool + QAct
__global__ void __launch_bounds__(128) avgpool_quant_kernel(const int8_t* __restrict__ in, int8_t* __restrict__ out, int T, int C,
                                                            float s_in, float s_out) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    int s = 0;
    for (int t = 0; t < T; ++t) s += int(in[(int64_t(b) * T + t) * C + c]);
    out[int64_t(b) * C + c] = int8_t(sat_s8(fdiv(fdiv(fmul(float(s), s_in), float(T)), s_out)));
  }
}

int launch_avgpool_quant(const int8_t* in, int8_t* out, int B, int T, int C, float s_in, float s_out, cudaStream_t stream) {
  launch_pdl(avgpool_quant_kernel, dim3(B), dim3(128), 0, stream, in, out, T, C, s_in, s_out);
  count_launch();
  return check_launch("avgpool_quant_i8");
}

}  // namespace p2v
