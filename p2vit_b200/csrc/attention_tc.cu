// Attention core between qact1 and qact2 (vit_fquant.py:373-389) on the 5th-generation tensor cores.
//
//   S = q k^T           tcgen05.mma kind::i8 (s8 x s8), q / k tiles TMA-loaded with 64-byte swizzle (K-major)
//   c = sat(RNE(S*m))   qact_attn1 codes, one thread per query row (TMEM lane == row): row max / row sum are
//   p = log2-softmax    thread-local, no shuffles.  exp_int comes from the 256-entry table (intmath.py), summed
//                       exactly; the log2 code uses an exact shortcut (see prob_code) with the IEEE division only
//                       next to a decision boundary.
//   O = P v             P = 2^(15-code) split into two u8 planes (hi, lo byte; exactly one is non-zero), written
//                       to shared memory in the K-major SWIZZLE_128B operand layout; v is consumed MN-major
//                       (token rows of 64 channels, exactly as TMA delivers them) -> two u8 x s8 MMA chains,
//                       O = 256*acc_hi + acc_lo (exact int32)
//   out = sat(RNE(O*m2))  qact2 codes, 64 contiguous bytes per row.
//
// One CTA = 4 softmax warps + 1 control warp (one lane issues TMA and MMA), two CTAs per SM (256 TMEM columns and
// ~112 KB shared memory each) so one CTA's MMA / TMA latency hides under the other's integer work.  A CTA walks
// heads blockIdx.x, blockIdx.x + gridDim.x, ...; each head is 1 or 2 query tiles of 128 rows.
#include <climits>
#include <type_traits>
#include "tc_common.cuh"

namespace p2v {

constexpr int AT_KV_ROWS = 224;                 // key / value rows staged per head (TMA box, >= T)
constexpr int AT_DH = 64;
constexpr int AT_THREADS = 160;
constexpr uint32_t AT_TMEM_COLS = 256;
constexpr uint32_t AT_OFF_Q = 0;                                  // 2 x [128 x 64]
constexpr uint32_t AT_OFF_K = 2 * 128 * AT_DH;                    // [224 x 64]
constexpr uint32_t AT_OFF_V = AT_OFF_K + AT_KV_ROWS * AT_DH;      // [224 x 64]
constexpr uint32_t AT_OFF_P = AT_OFF_V + AT_KV_ROWS * AT_DH;      // 2 planes x 2 chunks x [128 x 128]
constexpr uint32_t AT_P_CHUNK = 128 * 128, AT_P_PLANE = 2 * AT_P_CHUNK;
constexpr uint32_t AT_OFF_LUT = AT_OFF_P + 2 * AT_P_PLANE;        // uint2 [256] (hi, lo) + float [256] reciprocals
constexpr uint32_t AT_SMEM = AT_OFF_LUT + 256 * 8 + 256 * 4;
constexpr size_t AT_SMEM_ALLOC = AT_SMEM + 1024;                  // alignment slack
static_assert(AT_OFF_K % 1024 == 0 && AT_OFF_V % 1024 == 0 && AT_OFF_P % 1024 == 0, "swizzled tiles need 1024-byte alignment");

struct AttTcParams {
  int T, H, total_heads;
  int n_pad;      // key columns of S: T rounded up to 16
  int ksteps;     // 32-key MMA steps of P.V: ceil(T / 32)
  int mtiles;     // query tiles of 128 rows
  float score_mult, out_mult;
  const p2v_softmax_lut* lut;
  int8_t* out;
};

// log_round(RNE(fl(tot / e))) of layers.py:376-381,422-427 without the division.
//   x = RNE(q), q = fl(tot/e) >= 1;  big(x) = #{t in {2, 3, 6, 12, 24, ...} : x >= t}, and x >= t <=> q + 1/2 >= t up to
//   the tie rule.  With y = tot*rcp + 1/2 and w = y*(2/3):  big = [y >= 2] + max(0, floor(log2 w)).  tot*rcp, y and w
//   are each within a few ulps of the exact values, so the result can only differ from the reference when w is
//   within 16 ulps of a power of two >= 2 or y within 1e-5 of 2 (`near`): those take the exact path (log2_code).
// Returns 2^(15-big) (0 when big >= 16).
__device__ __forceinline__ uint32_t shr_clamp(uint32_t v, uint32_t n) {   // PTX shr: amounts > 31 give 0
  uint32_t r;
  asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
  return r;
}
__device__ __forceinline__ uint32_t prob_bits_fast(float tot, float rcp, bool& near) {
  const float y = fadd(fmul(tot, rcp), 0.5f);
  // 2/3 rounded up twice: y >= 1.5 - 1 ulp always (e <= tot), so w >= 1 and the exponent field needs no clamp; the
  // 1-ulp shift of the thresholds is inside the guard band
  const float w = fmul(y, 0.66666674613952636718750f);
  const uint32_t wb = __float_as_uint(w);
  const uint32_t nb = wb + 16u;
  near = ((nb & 0x007fffffu) < 32u && nb >= 0x40000000u) || fabsf(fsub(y, 2.0f)) < 1e-5f;
  return shr_clamp(y >= 2.0f ? 0x4000u : 0x8000u, (wb >> 23) - 127u);
}

__global__ void __launch_bounds__(AT_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, AttTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[6];
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_qk = smem_u32(&bars[0]), bar_v = smem_u32(&bars[1]), bar_s = smem_u32(&bars[2]);
  const uint32_t bar_p = smem_u32(&bars[3]), bar_o = smem_u32(&bars[4]), bar_free = smem_u32(&bars[5]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T, H = p.H;
  uint2* s_lut = reinterpret_cast<uint2*>(gbase + AT_OFF_LUT);
  float* s_rcp = reinterpret_cast<float*>(gbase + AT_OFF_LUT + 256 * 8);

  if (threadIdx.x == 0) {
    mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    mbar_init(bar_p, 4); mbar_init(bar_free, 4);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<AT_TMEM_COLS>(smem_u32(&tmem_slot));
  for (int i = threadIdx.x; i < 256; i += AT_THREADS) {
    s_lut[i] = make_uint2(p.lut->hi[i], p.lut->lo[i]);
    s_rcp[i] = fdiv(1.0f, p.lut->exp_f32[i]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 4) {
    // ================= control lane: TMA producer + MMA issuer =================
    if (lane == 0) {
      tma_prefetch_map(&tmQ);
      tma_prefetch_map(&tmKV);
      const int row_bytes_h = AT_DH;  // column offsets inside a token row: q at h*64, k at (H+h)*64, v at (2H+h)*64
      auto load_qk = [&](int hd) {
        const int b = hd / H, h = hd % H;
        mbar_expect_tx(bar_qk, uint32_t(p.mtiles) * 128u * AT_DH + AT_KV_ROWS * AT_DH);
        for (int mt = 0; mt < p.mtiles; ++mt) tma_load_3d(base + AT_OFF_Q + mt * 128 * AT_DH, &tmQ, bar_qk, h * row_bytes_h, mt * 128, b);
        tma_load_3d(base + AT_OFF_K, &tmKV, bar_qk, (H + h) * row_bytes_h, 0, b);
      };
      auto load_v = [&](int hd) {
        const int b = hd / H, h = hd % H;
        mbar_expect_tx(bar_v, AT_KV_ROWS * AT_DH);
        tma_load_3d(base + AT_OFF_V, &tmKV, bar_v, (2 * H + h) * row_bytes_h, 0, b);
      };
      const uint32_t idesc_qk = make_i8_idesc(128, p.n_pad, true, true);
      const uint32_t idesc_pv = make_i8_idesc(128, AT_DH, false, true, false, true);   // P u8 K-major, V s8 MN-major
      if (int(blockIdx.x) < p.total_heads) { load_qk(blockIdx.x); load_v(blockIdx.x); }
      uint32_t n = 0, hcount = 0;
      for (int hd = blockIdx.x; hd < p.total_heads; hd += gridDim.x, ++hcount) {
        const int next = hd + gridDim.x;
        mbar_wait(bar_qk, hcount & 1u);
        for (int mt = 0; mt < p.mtiles; ++mt, ++n) {
          mbar_wait(bar_free, (n & 1u) ^ 1u);          // previous tile's O has left TMEM
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < AT_DH / 32; ++k)
            umma_i8(tmem_base, make_smem_desc(base + AT_OFF_Q + mt * 128 * AT_DH + k * 32, 16, 512, UMMA_LAYOUT_SW64),
                    make_smem_desc(base + AT_OFF_K + k * 32, 16, 512, UMMA_LAYOUT_SW64), idesc_qk, uint32_t(k > 0));
          tc_commit(bar_s);
          if (mt == p.mtiles - 1 && next < p.total_heads) {
            mbar_wait(bar_s, n & 1u);                  // q / k tiles have been read: refill them for the next head
            load_qk(next);
          }
          mbar_wait(bar_p, n & 1u);                    // P planes written, S consumed
          if (mt == 0) mbar_wait(bar_v, hcount & 1u);
          tc_fence_after();
          for (int plane = 0; plane < 2; ++plane)
            for (int ks = 0; ks < p.ksteps; ++ks)
              umma_i8(tmem_base + plane * AT_DH,
                      make_kmajor_sw128_desc(base + AT_OFF_P + plane * AT_P_PLANE + (ks >> 2) * AT_P_CHUNK + (ks & 3) * 32),
                      make_smem_desc(base + AT_OFF_V + ks * 32 * AT_DH, AT_KV_ROWS * AT_DH, 512, UMMA_LAYOUT_SW64), idesc_pv,
                      uint32_t(ks > 0));
          tc_commit(bar_o);
          if (mt == p.mtiles - 1 && next < p.total_heads) {
            mbar_wait(bar_o, n & 1u);                  // v tile has been read
            load_v(next);
          }
        }
      }
    }
  } else {
    // ================= softmax warps: one thread per query row of the tile =================
    const int rloc = warp * 32 + lane;
    const uint32_t tlane = tmem_base + (uint32_t(warp * 32) << 16);
    const float mult = p.score_mult;
    uint32_t n = 0;
    for (int hd = blockIdx.x; hd < p.total_heads; hd += gridDim.x) {
      const int b = hd / H, h = hd % H;
      for (int mt = 0; mt < p.mtiles; ++mt, ++n) {
        const int row = mt * 128 + rloc;
        const bool warp_live = mt * 128 + warp * 32 < T;
        mbar_wait(bar_s, n & 1u);
        tc_fence_after();
        if (warp_live) {
          // S stays in TMEM and is re-read by every pass (TMEM reads are cheap; a rolled loop keeps the kernel in
          // the instruction cache, which a register-resident row of codes - a fully unrolled body - does not).
          const int nchunks = p.ksteps;                 // 32 key columns per chunk
          // ---- pass 1: row max of the raw scores; the requantisation is monotone, so max code = code(max S)
          int smax = INT_MIN;
#pragma unroll 1
          for (int c = 0; c < nchunks; ++c) {
            int acc[32];
            tmem_ld32(tlane + c * 32, acc);
            const int nv = T - c * 32;
#pragma unroll
            for (int e = 0; e < 32; ++e) smax = max(smax, e < nv ? acc[e] : INT_MIN);
          }
          const int mx = sat_s8(fmul(float(smax), mult));
          // ---- pass 2: exact row sum of exp_int(max - code); only the last chunk can hold columns >= T
          const int nfull = T >> 5;
          unsigned long long shi = 0, slo = 0;
          auto sum_chunk = [&](auto masked, int c) {
            int acc[32];
            tmem_ld32(tlane + c * 32, acc);
            const int nv = T - c * 32;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int d = mx - sat_s8(fmul(float(acc[e]), mult));
              const uint2 v = s_lut[d];
              if (!decltype(masked)::value || e < nv) { shi += v.x; slo += v.y; }
            }
          };
#pragma unroll 1
          for (int c = 0; c < nfull; ++c) sum_chunk(std::false_type{}, c);
          if (nfull < nchunks) sum_chunk(std::true_type{}, nfull);
          const float tot = u96_to_f32(shi, slo);
          // ---- pass 3: probabilities 2^(15-code) as hi / lo byte planes in the UMMA K-major SW128 layout
          uint8_t* prow = gbase + AT_OFF_P + rloc * 128;
          const uint32_t sw = uint32_t(rloc & 7);
          auto prob_chunk = [&](auto masked, int c) {
            int acc[32];
            tmem_ld32(tlane + c * 32, acc);
            const int nv = T - c * 32;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t pv[16];
              bool any_near = false;
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int d = mx - sat_s8(fmul(float(acc[half * 16 + e]), mult));
                acc[half * 16 + e] = d;
                bool near;
                pv[e] = prob_bits_fast(tot, s_rcp[d], near);
                any_near |= near;
              }
              if (any_near) {   // some element sits next to a rounding / log2 boundary: redo the unit with the IEEE division
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const uint2 v = s_lut[acc[half * 16 + e]];
                  const uint32_t big = log2_code(tot, __ull2float_rn((static_cast<unsigned long long>(v.x) << 32) | v.y));
                  pv[e] = shr_clamp(0x8000u, big);
                }
              }
              uint32_t lo[4], hi[4];
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                if (decltype(masked)::value) {
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    if (half * 16 + e4 * 4 + e >= nv) pv[e4 * 4 + e] = 0u;
                }
                const uint32_t p01 = pv[e4 * 4] | (pv[e4 * 4 + 1] << 16), p23 = pv[e4 * 4 + 2] | (pv[e4 * 4 + 3] << 16);
                lo[e4] = __byte_perm(p01, p23, 0x6420);
                hi[e4] = __byte_perm(p01, p23, 0x7531);
              }
              const int g = c * 2 + half;                       // 16-byte unit along the key axis
              uint8_t* dst = prow + (g >> 3) * AT_P_CHUNK + ((uint32_t(g & 7) ^ sw) << 4);
              *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(dst + AT_P_PLANE) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          };
#pragma unroll 1
          for (int c = 0; c < nfull; ++c) prob_chunk(std::false_type{}, c);
          if (nfull < nchunks) prob_chunk(std::true_type{}, nfull);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        // ---- O = 256*hi + lo -> qact2 codes
        mbar_wait(bar_o, n & 1u);
        tc_fence_after();
        if (warp_live) {
          int8_t* orow = p.out + (int64_t(b) * T + row) * (int64_t(H) * AT_DH) + h * AT_DH;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            int ah[32], al[32];
            tmem_ld32_async(tlane + half * 32, ah);
            tmem_ld32_async(tlane + AT_DH + half * 32, al);
            tmem_wait_ld();
            if (row < T) {
#pragma unroll
              for (int j = 0; j < 32; j += 16) {
                uint32_t w[4];
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                  int q[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const int O = ah[j + e4 * 4 + e] * 256 + al[j + e4 * 4 + e];
                    q[e] = sat_s8(fmul(float(O), p.out_mult));
                  }
                  w[e4] = pack4_s8(q[0], q[1], q[2], q[3]);
                }
                *reinterpret_cast<uint4*>(orow + half * 32 + j) = make_uint4(w[0], w[1], w[2], w[3]);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<AT_TMEM_COLS>(tmem_base);
  }
}

// int8 [B, T, W] tensor (W = 3*H*64 bytes per token), box = [64 bytes, box_rows tokens, 1 image], 64-byte swizzle;
// rows >= T are out of bounds and read as zero
static int make_tmap_qkv(CUtensorMap* m, const void* ptr, int B, int T, int W, int box_rows) {
  encode_tiled_fn enc = get_tensor_map_encoder();
  P2V_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {cuuint64_t(W), cuuint64_t(T), cuuint64_t(B)};
  cuuint64_t strides[2] = {cuuint64_t(W), cuuint64_t(W) * cuuint64_t(T)};
  cuuint32_t box[3] = {cuuint32_t(AT_DH), cuuint32_t(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  P2V_REQUIRE(r == CUDA_SUCCESS, "attention: cuTensorMapEncodeTiled failed (%d) B=%d T=%d W=%d", int(r), B, T, W);
  return 0;
}

bool attention_tc_supported(const p2v_attention_args& a) {
  return a.dh == AT_DH && a.T <= AT_KV_ROWS && a.probs_or_null == nullptr && a.scores_or_null == nullptr &&
         (reinterpret_cast<uintptr_t>(a.qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
}

int launch_attention_tc(const p2v_attention_args& a, cudaStream_t stream) {
  P2V_REQUIRE(attention_tc_supported(a), "attention_tc: needs head dim 64, T <= %d, 16-byte aligned tensors, no debug dumps", AT_KV_ROWS);
  CUtensorMap tmQ, tmKV;
  const int W = 3 * a.H * AT_DH;
  if (int r = make_tmap_qkv(&tmQ, a.qkv, a.B, a.T, W, 128)) return r;
  if (int r = make_tmap_qkv(&tmKV, a.qkv, a.B, a.T, W, AT_KV_ROWS)) return r;
  AttTcParams p;
  p.T = a.T; p.H = a.H; p.total_heads = a.B * a.H;
  p.n_pad = (a.T + 15) / 16 * 16;
  p.ksteps = (a.T + 31) / 32;
  p.mtiles = (a.T + 127) / 128;
  p.score_mult = a.score_mult; p.out_mult = a.out_mult;
  p.lut = a.lut_dev; p.out = a.out;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(AT_SMEM_ALLOC));
    P2V_REQUIRE(e == cudaSuccess, "attention_tc: cannot set %zu bytes of dynamic shared memory: %s", AT_SMEM_ALLOC, cudaGetErrorString(e));
  }
  const int grid = std::min(p.total_heads, 2 * sms);
  attention_tc_kernel<<<grid, AT_THREADS, AT_SMEM_ALLOC, stream>>>(tmQ, tmKV, p);
  count_launch();
  return check_launch("attention_tc");
}

}  // namespace p2v
