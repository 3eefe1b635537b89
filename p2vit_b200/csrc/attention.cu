// Attention core between qact1 and qact2 (vit_fquant.py:373-389): S = q k^T -> qact_attn1 requant ->
// integer log2 softmax (layers.py:384-428) -> P v -> qact2 requant, all on integer codes.
//
// v1 kernel: one CTA per (image, head), K / V^T / Q resident in shared memory, dp4a for both matmuls,
// one warp per query row (row max / row sum are warp reductions).  The probabilities 2^(15-code) are
// kept as two u8 planes (hi, lo byte) so P.V is two u8 x s8 dot products: O = 256*acc_hi + acc_lo (exact).
#include "common.cuh"

namespace p2v {

__device__ __forceinline__ int dp4a_us(uint32_t a_u8x4, uint32_t b_s8x4, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
  return d;
}

constexpr int ATT_WARPS = 8;

template <int DH>
__global__ void __launch_bounds__(ATT_WARPS * 32) attention_simt_kernel(p2v_attention_args a) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int DW = DH / 4;             // words per head row
  constexpr int KSTR = DW + 1;           // padded K row stride (words), odd -> conflict-free column walks
  const int T = a.T, H = a.H;
  const int TW = (T + 3) / 4;            // words per probability row
  const int VSTR = (TW & 1) ? TW : TW + 1;
  uint32_t* sK = reinterpret_cast<uint32_t*>(smem);                  // [T][KSTR]
  uint32_t* sQ = sK + size_t(T) * KSTR;                              // [T][DW]
  uint32_t* sVt = sQ + size_t(T) * DW;                               // [DH][VSTR]  (bytes: V^T[c][j])
  uint32_t* sP = sVt + size_t(DH) * VSTR;                            // [ATT_WARPS][2][TW]
  uint32_t* sLh = sP + size_t(ATT_WARPS) * 2 * TW;                   // LUT hi / lo / exp
  uint32_t* sLl = sLh + 256;
  float* sLe = reinterpret_cast<float*>(sLl + 256);

  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row_bytes = int64_t(3) * H * DH;
  const int8_t* base = a.qkv + int64_t(b) * T * row_bytes + h * DH;

  for (int i = tid; i < 256; i += blockDim.x) { sLh[i] = a.lut_dev->hi[i]; sLl[i] = a.lut_dev->lo[i]; sLe[i] = a.lut_dev->exp_f32[i]; }
  for (int i = tid; i < DH * VSTR; i += blockDim.x) sVt[i] = 0u;
  for (int i = tid; i < ATT_WARPS * 2 * TW; i += blockDim.x) sP[i] = 0u;
  __syncthreads();
  // 16-byte chunks: T rows x (DH/16) chunks for each of q, k, v
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

  constexpr int NJ = 8;  // keys per lane: T <= 256
  uint32_t* pHi = sP + warp * 2 * TW;
  uint32_t* pLo = pHi + TW;
  for (int i = warp; i < T; i += ATT_WARPS) {
    // ---- S = q_i . k_j, requantised to qact_attn1 codes
    int code[NJ];
    int mx = -128;
    const uint32_t* qi = sQ + i * DW;
    // asymmetric qact1 (zp_qkv = z): sum (q - z)(k - z) = q.k - z (sum q + sum k) + DH z^2
    const int z = a.zp_qkv;
    int sq = 0;
    if (z != 0) {
#pragma unroll
      for (int w = 0; w < DW; ++w) sq = __dp4a(int(qi[w]), 0x01010101, sq);
    }
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      const int j = lane + 32 * jj;
      code[jj] = -1000;
      if (j < T) {
        const uint32_t* kj = sK + j * KSTR;
        int s = 0, sk = 0;
#pragma unroll
        for (int w = 0; w < DW; ++w) {
          s = __dp4a(int(qi[w]), int(kj[w]), s);
          if (z != 0) sk = __dp4a(int(kj[w]), 0x01010101, sk);
        }
        s += DH * z * z - z * (sq + sk);
        code[jj] = sat_s8(fadd(fmul(float(s), a.score_mult), a.zp_score));
        mx = max(mx, code[jj]);
      }
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    unsigned long long hi = 0, lo = 0;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj)
      if (lane + 32 * jj < T) { const int d = mx - code[jj]; hi += sLh[d]; lo += sLl[d]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { hi += __shfl_xor_sync(0xffffffffu, hi, o); lo += __shfl_xor_sync(0xffffffffu, lo, o); }
    const float tot = u96_to_f32(hi, lo);
    uint8_t* bHi = reinterpret_cast<uint8_t*>(pHi);
    uint8_t* bLo = reinterpret_cast<uint8_t*>(pLo);
    int psum = 0;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      const int j = lane + 32 * jj;
      if (j < T) {
        const uint32_t c = log2_code(tot, sLe[mx - code[jj]]);
        const uint32_t pv = c == 255u ? 0u : (1u << (15 - c));
        psum += int(pv);
        bHi[j] = uint8_t(pv >> 8);
        bLo[j] = uint8_t(pv & 0xffu);
        if (a.probs_or_null) a.probs_or_null[(int64_t(blockIdx.x) * T + i) * T + j] = uint8_t(c);
        if (a.scores_or_null) a.scores_or_null[(int64_t(blockIdx.x) * T + i) * T + j] = int8_t(code[jj]);
      }
    }
    psum = __reduce_add_sync(0xffffffffu, psum);
    __syncwarp();
    // ---- O = sum_j 2^(15-code_j) (v_j - z) ; lane owns channels lane, lane+32
#pragma unroll
    for (int cc = 0; cc < DH / 32; ++cc) {
      const int c = lane + 32 * cc;
      const uint32_t* vt = sVt + c * VSTR;
      int ah = 0, al = 0;
      for (int w = 0; w < TW; ++w) {
        const uint32_t v = vt[w];
        ah = dp4a_us(pHi[w], v, ah);
        al = dp4a_us(pLo[w], v, al);
      }
      const int O = ah * 256 + al - z * psum;
      a.out[(int64_t(b) * T + i) * (H * DH) + h * DH + c] = int8_t(sat_s8(fadd(fmul(float(O), a.out_mult), a.zp_out)));
    }
    __syncwarp();
  }
}

static size_t attention_smem_bytes(int T, int DH) {
  const int DW = DH / 4, KSTR = DW + 1, TW = (T + 3) / 4, VSTR = (TW & 1) ? TW : TW + 1;
  return sizeof(uint32_t) * (size_t(T) * KSTR + size_t(T) * DW + size_t(DH) * VSTR + size_t(ATT_WARPS) * 2 * TW + 768);
}

int launch_attention(const p2v_attention_args& a, cudaStream_t stream) {
  const size_t smem = attention_smem_bytes(a.T, a.dh);
  const int grid = a.B * a.H;
  if (a.dh == 64) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(attention_simt_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); attr = true; }
    attention_simt_kernel<64><<<grid, ATT_WARPS * 32, smem, stream>>>(a);
  } else {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(attention_simt_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); attr = true; }
    attention_simt_kernel<32><<<grid, ATT_WARPS * 32, smem, stream>>>(a);
  }
  count_launch();
  return check_launch("attention_i8");
}

}  // namespace p2v
