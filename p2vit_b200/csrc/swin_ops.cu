// Swin-specific kernels (reference: models/swin_quant.py).  Window partition, cyclic shift, window reverse are not
// kernels: they are row maps applied by the LayerNorm store (p2v_layernorm_args.out_row_map) and by the GEMM epilogue
// (p2v_gemm_args.row_map).  Here: the per-window attention (quantized relative-position bias and SW-MSA mask inside),
// the 2x2 patch-merging gather, and the token average pool + QAct of the classifier tail.
#include <climits>
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

// One CTA per (window, head): T = ws*ws <= 64 tokens, one warp per query row, a lane owns keys lane and lane + 32.
// Windows are tiny (49 x 32 per head), so this is an integer-ALU / latency kernel; dp4a for both matmuls.
template <int DH>
__global__ void __launch_bounds__(WA_WARPS * 32) window_attention_kernel(p2v_window_attention_args a, uint32_t e_mask) {
  constexpr int DW = DH / 4, KSTR = DW + 1, TW = WA_MAXT / 4, VSTR = TW + 1;
  __shared__ uint32_t sQ[WA_MAXT * DW], sK[WA_MAXT * KSTR], sVt[DH * VSTR], sP[WA_WARPS * 2 * TW];
  __shared__ uint32_t sLh[256], sLl[256];
  __shared__ float sLe[256];
  __shared__ int8_t sLab[WA_MAXT];
  const int T = a.T, H = a.H;
  const int win = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row_bytes = int64_t(3) * H * DH;
  const int8_t* base = a.qkv + int64_t(win) * T * row_bytes + h * DH;
  for (int i = tid; i < 256; i += blockDim.x) { sLh[i] = a.lut_dev->hi[i]; sLl[i] = a.lut_dev->lo[i]; sLe[i] = a.lut_dev->exp_f32[i]; }
  for (int i = tid; i < DH * VSTR; i += blockDim.x) sVt[i] = 0u;
  for (int i = tid; i < WA_WARPS * 2 * TW; i += blockDim.x) sP[i] = 0u;
  if (tid < WA_MAXT) sLab[tid] = (a.labels && tid < T) ? a.labels[(win % a.windows_per_image) * T + tid] : int8_t(0);
  __syncthreads();
  constexpr int CH = DH / 16;
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

  uint32_t* pHi = sP + warp * 2 * TW;
  uint32_t* pLo = pHi + TW;
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
        const int c1 = sat_s8(fmul(float(s), a.score_mult));                                   // qact_attn1
        const float v = fadd(fmul(float(c1), a.s_attn1), __ldg(bias_h + i * T + j));            // + relative position bias
        const int c2 = sat_s8(fdiv(v, a.s_attn2));                                              // qact2
        masked[jj] = sLab[i] != sLab[j];
        x[jj] = c2 + (masked[jj] ? a.mask_code : 0);                                            // + mask (after the quantizer)
        mx = max(mx, x[jj]);
      }
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    unsigned long long hi = 0, lo = 0;
    float ef[2] = {0.f, 0.f};
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
      if (lane + 32 * jj < T) {
        if (masked[jj] && mx - x[jj] > 255) {   // beyond the table: the clamped tail (host guarantees mask_code reaches it)
          lo += e_mask;
          ef[jj] = float(e_mask);
        } else {
          const int d = mx - x[jj];
          hi += sLh[d]; lo += sLl[d];
          ef[jj] = sLe[d];
        }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { hi += __shfl_xor_sync(0xffffffffu, hi, o); lo += __shfl_xor_sync(0xffffffffu, lo, o); }
    const float tot = u96_to_f32(hi, lo);
    uint8_t* bHi = reinterpret_cast<uint8_t*>(pHi);
    uint8_t* bLo = reinterpret_cast<uint8_t*>(pLo);
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = lane + 32 * jj;
      if (j < T) {
        const uint32_t c = log2_code(tot, ef[jj]);
        const uint32_t pv = c == 255u ? 0u : (1u << (15 - c));
        bHi[j] = uint8_t(pv >> 8);
        bLo[j] = uint8_t(pv & 0xffu);
      }
    }
    __syncwarp();
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
      a.out[(int64_t(win) * T + i) * (H * DH) + h * DH + c] = int8_t(sat_s8(fmul(float(ah * 256 + al), a.out_mult)));
    }
    __syncwarp();
  }
}

int launch_window_attention(const p2v_window_attention_args& a, uint32_t e_mask, cudaStream_t stream) {
  const int grid = a.n_windows * a.H;
  if (a.dh == 32) window_attention_kernel<32><<<grid, WA_WARPS * 32, 0, stream>>>(a, e_mask);
  else window_attention_kernel<64><<<grid, WA_WARPS * 32, 0, stream>>>(a, e_mask);
  count_launch();
  return check_launch("window_attention_i8");
}

// ------------------------------------------------------------------------------------------------ patch-merging gather
__global__ void __launch_bounds__(256) gather_rows_kernel(const int8_t* __restrict__ in, int8_t* __restrict__ out,
                                                          const int32_t* __restrict__ src, int64_t total16, int segs, int C16) {
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
  gather_rows_kernel<<<blocks, 256, 0, stream>>>(in, out, src, total16, segs, C / 16);
  count_launch();
  return check_launch("gather_rows_i8");
}

// ------------------------------------------------------------------------------------------------ token average pool + QAct
__global__ void __launch_bounds__(128) avgpool_quant_kernel(const int8_t* __restrict__ in, int8_t* __restrict__ out, int T, int C,
                                                            float s_in, float s_out) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    int s = 0;
    for (int t = 0; t < T; ++t) s += int(in[(int64_t(b) * T + t) * C + c]);
    out[int64_t(b) * C + c] = int8_t(sat_s8(fdiv(fdiv(fmul(float(s), s_in), float(T)), s_out)));
  }
}

int launch_avgpool_quant(const int8_t* in, int8_t* out, int B, int T, int C, float s_in, float s_out, cudaStream_t stream) {
  avgpool_quant_kernel<<<B, 128, 0, stream>>>(in, out, T, C, s_in, s_out);
  count_launch();
  return check_launch("avgpool_quant_i8");
}

}  // namespace p2v
