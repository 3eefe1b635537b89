// Swin window attention between attn.qact1 and attn.qact3 (swin_quant.py:211-249) on the 5th-generation tensor cores.
//
// A window is 49 tokens x 32 channels per head - a quarter of a 128-row tcgen05 tile - so a work unit is a PAIR of windows
// of one head: window A in rows / keys 0..63, window B in rows / keys 64..127 (rows >= 49 of each half are zero, filled by
// TMA's out-of-bounds rule).
//
//   S = Q K^T            ONE tcgen05.mma kind::i8 (M128 x N128 x K32, s8 x s8): the diagonal 64 x 64 blocks are the two
//                        windows' scores, the off-diagonal blocks are never read
//   c1 = sat(RNE(S*m))   qact_attn1;  c2 = sat(RNE((c1*s1 + bias[h,i,j]) / s2))   + quantized relative-position bias -> qact2
//   x  = c2 + mask_code*[label_i != label_j]      SW-MSA mask (one 64-bit word of mask bits per query row)
//   p  = log2-softmax(x) thread = query row (TMEM lane): the row's 64 scores live in registers, row max / exact 64-bit row sum are
//                        thread-local; 2^(15-code) through prob_bits_fast with the IEEE division behind its guard band
//   O = P V              P as two u8 planes (hi / lo byte) in the K-major SWIZZLE_128B operand layout, block diagonal (a row
//                        only ever writes its own window's 64 key bytes; the other half stays zero from the prologue); V is
//                        consumed MN-major exactly as TMA delivers it (SWIZZLE_32B): 2 planes x 4 k-steps of M128 x N32 x K32
//   out = sat(RNE(O*m2)) qact3, stored through out_row_map (window reverse + inverse cyclic shift, swin_quant.py:426-436)
//
// CTA = 4 softmax warps + 1 control warp (TMA producer, MMA issuer), three CTAs per SM; q / k / v tiles are double buffered so
// the loads of unit n+1 run under the softmax of unit n.  Same results, bit for bit, as the dp4a kernel in swin_ops.cu
// (tests/test_gpu_swin.py cross-checks the two and both against the oracle).
#include <climits>
#include <cmath>
#include <algorithm>
#include <type_traits>
#include "tc_common.cuh"

namespace p2v {

constexpr int WT_DH = 32;
constexpr int WT_THREADS = 160;
constexpr uint32_t WT_TILE = 128 * WT_DH;                    // one operand tile: 128 rows x 32 bytes
constexpr uint32_t WT_STAGE = 3 * WT_TILE;                   // q, k, v
constexpr uint32_t WT_OFF_P = 2 * WT_STAGE;                  // 2 planes x [128 x 128]
constexpr uint32_t WT_P_PLANE = 128 * 128;
constexpr uint32_t WT_OFF_LUT = WT_OFF_P + 2 * WT_P_PLANE;   // 256 x {hi, lo, exp_f32, 1 / exp_f32}
constexpr uint32_t WT_BIAS_PITCH = 80;                       // bias-code row pitch (p2v_window_attention_args.bias_codes): 16-byte row reads of 8 lanes hit 8 distinct bank groups
constexpr uint32_t WT_BIAS_SLOT = 64 * WT_BIAS_PITCH;        // int8 bias codes of one head (T <= 64 rows)
constexpr uint32_t WT_OFF_BIAS = WT_OFF_LUT + 4224;             // 257 table entries of 16 bytes, padded
constexpr uint32_t WT_OFF_BARS = WT_OFF_BIAS + 2 * WT_BIAS_SLOT;
constexpr uint32_t WT_SMEM = WT_OFF_BARS + 64;
constexpr size_t WT_SMEM_ALLOC = WT_SMEM + 1024;             // + alignment slack
constexpr uint32_t WT_TMEM_COLS = 128;                       // S [128 x 128]; O (2 planes x 32 columns) reuses its first 64 columns

struct WinTcParams {
  int T, H, n_windows, wpi, units;
  float score_mult, s_attn1, s_attn2, bias_scale, out_mult;
  int mask_code;
  uint32_t e_mask;
  const int8_t* bias_codes;                // [H, T, 80]
  const unsigned long long* mask_bits;     // [wpi, T] or NULL
  const p2v_softmax_lut* lut;
  const int32_t* out_row_map;
  int8_t* out;
};

template <bool POT2, bool MASK>
__global__ void __launch_bounds__(WT_THREADS, 3)
window_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, WinTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + WT_OFF_BARS;
  const uint32_t bar_load = bars, bar_s = bars + 16, bar_p = bars + 24, bar_o = bars + 32, bar_free = bars + 40;   // bar_load: 2 stages
  volatile uint32_t& tmem_slot = *reinterpret_cast<volatile uint32_t*>(gbase + WT_OFF_BARS + 48);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T, H = p.H;
  uint4* s_lut = reinterpret_cast<uint4*>(gbase + WT_OFF_LUT);

  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1); mbar_init(bar_load + 8, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    mbar_init(bar_p, 4); mbar_init(bar_free, 4);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<WT_TMEM_COLS>(base + WT_OFF_BARS + 48);
  // table entry d = max - x: {exp_int hi, lo, exp_int as fp32, its reciprocal}; entry 256 = the clamped tail of int_exp that
  // every masked score lands on (the host checks mask_code reaches it), so the lookup needs no branch: d = min(max - x, 256)
  for (int i = threadIdx.x; i < 257; i += WT_THREADS) {
    const bool tail = i == 256 && MASK;            // (without a mask entry 256 only serves the padding keys of a partial chunk: a copy of 255)
    const int k = min(i, 255);
    const float e = tail ? float(p.e_mask) : p.lut->exp_f32[k];
    s_lut[i] = make_uint4(tail ? 0u : p.lut->hi[k], tail ? p.e_mask : p.lut->lo[k], __float_as_uint(e), __float_as_uint(prob_rcp(e)));
  }
  for (uint32_t i = threadIdx.x; i < 2 * WT_P_PLANE / 16; i += WT_THREADS)       // off-diagonal blocks and key padding of P stay zero
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(base + WT_OFF_P + i * 16u), "r"(0u) : "memory");
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 4) {
    // ================= control warp: TMA producer + MMA issuer (warp-uniform, one elected lane issues) =================
    if (elect_one()) tma_prefetch_map(&tmQKV);
    auto load_unit = [&](int unit, uint32_t s) {
      const int pair = unit / H, h = unit % H;
      const uint32_t dst = base + s * WT_STAGE;
      mbar_expect_tx(bar_load + 8 * s, 6u * 64u * WT_DH);
      for (int w = 0; w < 2; ++w) {
        const int win = 2 * pair + w;      // win == n_windows (odd count): out of bounds, zero filled
        tma_load_3d(dst + w * 64 * WT_DH, &tmQKV, bar_load + 8 * s, h * WT_DH, 0, win);
        tma_load_3d(dst + WT_TILE + w * 64 * WT_DH, &tmQKV, bar_load + 8 * s, (H + h) * WT_DH, 0, win);
        tma_load_3d(dst + 2 * WT_TILE + w * 64 * WT_DH, &tmQKV, bar_load + 8 * s, (2 * H + h) * WT_DH, 0, win);
      }
    };
    const uint32_t idesc_qk = make_i8_idesc(128, 128, true, true);
    const uint32_t idesc_pv = make_i8_idesc(128, WT_DH, false, true, false, true);      // P u8 K-major, V s8 MN-major
    if (int(blockIdx.x) < p.units) {
      if (elect_one()) load_unit(blockIdx.x, 0);
      __syncwarp();
    }
    uint32_t n = 0;
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x, ++n) {
      const uint32_t s = n & 1u;
      const int next = unit + gridDim.x;
      mbar_wait(bar_free, (n & 1u) ^ 1u);           // O of the previous unit has left TMEM (so its P.V has read stage s^1 too)
      if (next < p.units) {
        if (elect_one()) load_unit(next, s ^ 1u);
        __syncwarp();
      }
      mbar_wait(bar_load + 8 * s, (n >> 1) & 1u);
      tc_fence_after();
      const uint32_t st = base + s * WT_STAGE;
      if (elect_one()) {
        umma_i8(tmem_base, make_smem_desc(st, 16, 256, UMMA_LAYOUT_SW32), make_smem_desc(st + WT_TILE, 16, 256, UMMA_LAYOUT_SW32), idesc_qk, 0u);
        tc_commit(bar_s);
      }
      __syncwarp();
      mbar_wait(bar_p, n & 1u);                     // P planes written, S consumed
      tc_fence_after();
      if (elect_one()) {
        for (int plane = 0; plane < 2; ++plane)
          for (int ks = 0; ks < 4; ++ks)
            umma_i8(tmem_base + plane * WT_DH, make_kmajor_sw128_desc(base + WT_OFF_P + plane * WT_P_PLANE + ks * 32),
                    make_smem_desc(st + 2 * WT_TILE + ks * 32 * WT_DH, WT_TILE, 256, UMMA_LAYOUT_SW32), idesc_pv, uint32_t(ks > 0));
        tc_commit(bar_o);
      }
      __syncwarp();
    }
  } else {
    // ================= softmax warps: thread = query row r = TMEM lane; r >> 6 = window of the pair, r & 63 = token =================
    const int r = warp * 32 + lane;
    const int w = r >> 6, i = r & 63;
    const uint32_t tlane = tmem_base + ((uint32_t(warp) * 32u) << 16);
    const uint32_t lut32 = base + WT_OFF_LUT;
    const float r2 = fdiv(1.0f, p.s_attn2);
    constexpr float LO = RMAGIC - 128.f, HI = RMAGIC + 127.f;
    const uint32_t prow = base + WT_OFF_P + uint32_t(r) * 128u;
    const uint32_t sw = uint32_t(r & 7);
    const int nfull = T >> 4, rem = T & 15;          // key chunks of 16: nfull complete ones, then one with `rem` valid keys
    uint32_t n = 0;
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x, ++n) {
      const int pair = unit / H, h = unit % H;
      const int win = 2 * pair + w;
      const bool live = i < T && win < p.n_windows;
      // ---- this head's bias codes into the slot of this unit, as code + 128 (unsigned bytes decode with one PRMT + FADD);
      //      slot n&1 was last read two units ago
      {
        uint4* slot = reinterpret_cast<uint4*>(gbase + WT_OFF_BIAS + (n & 1u) * WT_BIAS_SLOT);
        const uint4* src = reinterpret_cast<const uint4*>(p.bias_codes + size_t(h) * T * WT_BIAS_PITCH);
        for (int e = threadIdx.x; e < T * int(WT_BIAS_PITCH / 16); e += 128) {
          uint4 v = __ldg(src + e);
          v.x ^= 0x80808080u; v.y ^= 0x80808080u; v.z ^= 0x80808080u; v.w ^= 0x80808080u;
          slot[e] = v;
        }
      }
      uint32_t mlo = 0u, mhi = 0u;
      if (MASK && live) {
        const unsigned long long mb = __ldg(p.mask_bits + size_t(win % p.wpi) * T + i);
        mlo = uint32_t(mb); mhi = uint32_t(mb >> 32);
      }
      named_barrier(1, 128);                        // bias slot complete; every row is done with the unit before the last
      const uint8_t* brow = gbase + WT_OFF_BIAS + (n & 1u) * WT_BIAS_SLOT + i * WT_BIAS_PITCH;
      mbar_wait(bar_s, n & 1u);
      tc_fence_after();
      int x[4][16];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld16_async(tlane + uint32_t(w) * 64u + c * 16, x[c]);
      tmem_wait_ld();
      if (live) {
        // ---- pass A: scores -> qact_attn1 -> + bias -> qact2 -> + mask; row max.  Straight-line code per chunk: a complete chunk
        //      has no per-element tests at all, the last one replaces its padding keys after the fact.
        int mx = INT_MIN;
        auto pass_a = [&](auto partial, auto cc) {
          constexpr int c = decltype(cc)::value;
          const uint4 b16 = *reinterpret_cast<const uint4*>(brow + c * 16);
          const uint32_t bw[4] = {b16.x, b16.y, b16.z, b16.w};
          const uint32_t mword = c < 2 ? mlo : mhi;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            // float(sat_s8(s * m)) without leaving fp32 (the biased sum is monotone in its argument)
            const float c1 = fsub(fminf(fmaxf(fadd(fmul(__int2float_rn(x[c][e]), p.score_mult), RMAGIC), LO), HI), RMAGIC);
            // bias code: byte e of the row (stored + 128) -> 2^23 + (code + 128) -> code, exactly
            const float bc = fsub(__uint_as_float(__byte_perm(bw[e >> 2], 0x4B000000u, 0x7540 + (e & 3))), 8388736.f);
            const float v = fadd(fmul(c1, p.s_attn1), fmul(bc, p.bias_scale));              // + dequantized relative position bias
            const float q2 = POT2 ? fmul(v, r2) : fdiv(v, p.s_attn2);                        // qact2
            int xv = __float_as_int(fminf(fmaxf(fadd(q2, RMAGIC), LO), HI)) - 0x4B400000;
            if (MASK) xv += int(((mword >> ((c & 1) * 16 + e)) & 1u) * uint32_t(p.mask_code));   // + mask (after the quantizer)
            if (decltype(partial)::value) xv = e < rem ? xv : -(1 << 30);      // padding key: far below every score, lands on table entry 256
            x[c][e] = xv;
            mx = max(mx, xv);
          }
        };
        // ---- pass B: exact row sum of exp_int(max - x); x[] becomes the shared-memory address of the table entry
        unsigned long long sum = 0;
        auto pass_b = [&](auto partial, auto cc) {
          constexpr int c = decltype(cc)::value;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const uint32_t d = uint32_t(min(mx - x[c][e], 256));
            const uint32_t addr = lut32 + (d << 4);
            uint32_t vh, vl;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(vh), "=r"(vl) : "r"(addr));
            unsigned long long ev = (static_cast<unsigned long long>(vh) << 32) | vl;
            if (decltype(partial)::value) ev = e < rem ? ev : 0ull;
            sum += ev;
            x[c][e] = int(addr);
          }
        };
#define P2V_WT_CHUNKS(fn)                                                                                                   \
        if (nfull > 0) fn(std::false_type{}, std::integral_constant<int, 0>{}); else if (rem) fn(std::true_type{}, std::integral_constant<int, 0>{}); \
        if (nfull > 1) fn(std::false_type{}, std::integral_constant<int, 1>{}); else if (nfull == 1 && rem) fn(std::true_type{}, std::integral_constant<int, 1>{}); \
        if (nfull > 2) fn(std::false_type{}, std::integral_constant<int, 2>{}); else if (nfull == 2 && rem) fn(std::true_type{}, std::integral_constant<int, 2>{}); \
        if (nfull > 3) fn(std::false_type{}, std::integral_constant<int, 3>{}); else if (nfull == 3 && rem) fn(std::true_type{}, std::integral_constant<int, 3>{});
        P2V_WT_CHUNKS(pass_a)
        P2V_WT_CHUNKS(pass_b)
        const float tot = __ull2float_rn(sum);      // <= 64 entries below 2^55: exact in 64 bits, rounded once
        const float tot2 = fmul(tot, 2.0f), tot43 = fmul(tot, 1.33333337306976318359375f);
        // ---- pass C: 2^(15-code) as hi / lo byte planes, this row's 64 key bytes of each plane
        auto pass_c = [&](auto partial, auto cc) {
          constexpr int c = decltype(cc)::value;
          uint32_t pv[16];
          float gmax = 0.f;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float rcp;
            asm volatile("ld.shared.f32 %0, [%1+12];" : "=f"(rcp) : "r"(uint32_t(x[c][e])));
            pv[e] = prob_bits_fast(tot2, tot43, rcp, gmax);
          }
          if (!(gmax < PROB_GUARD)) {             // next to a rounding / log2 boundary (or not finite): IEEE division for the chunk
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              float ef;
              asm volatile("ld.shared.f32 %0, [%1+8];" : "=f"(ef) : "r"(uint32_t(x[c][e])));
              pv[e] = shr_clamp(0x8000u, log2_code(tot, ef));
            }
          }
          if (decltype(partial)::value) {
#pragma unroll
            for (int e = 0; e < 16; ++e) pv[e] = e < rem ? pv[e] : 0u;
          }
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const uint32_t p01 = __byte_perm(pv[e4 * 4], pv[e4 * 4 + 1], 0x5410), p23 = __byte_perm(pv[e4 * 4 + 2], pv[e4 * 4 + 3], 0x5410);
            lo[e4] = __byte_perm(p01, p23, 0x6420);
            hi[e4] = __byte_perm(p01, p23, 0x7531);
          }
          const uint32_t dst = prow + ((uint32_t(4 * w + c) ^ sw) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + WT_P_PLANE), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
        };
        P2V_WT_CHUNKS(pass_c)
#undef P2V_WT_CHUNKS
        // (key chunks beyond T are never written: they stay zero from the prologue)
      } else if (i < T) {
        // the missing partner of an odd window count: its P rows must not carry the previous unit's values (token-padding rows
        // are never written and stay zero from the prologue)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t dst = prow + ((uint32_t(4 * w + c) ^ sw) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst + WT_P_PLANE), "r"(0u) : "memory");
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      // ---- O = 256 * hi + lo -> qact3 codes, 32 contiguous bytes per row
      mbar_wait(bar_o, n & 1u);
      tc_fence_after();
      {
        int ah[32], al[32];
        tmem_ld32_async(tlane, ah);
        tmem_ld32_async(tlane + WT_DH, al);
        tmem_wait_ld();
        if (live) {
          const int64_t orow = p.out_row_map ? int64_t(__ldg(p.out_row_map + int64_t(win) * T + i)) : int64_t(win) * T + i;
          int8_t* o8 = p.out + orow * (int64_t(H) * WT_DH) + h * WT_DH;
#pragma unroll
          for (int j = 0; j < 32; j += 16) {
            uint32_t wd[4];
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              float rr[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) rr[e] = fadd(fmul(__int2float_rn(ah[j + e4 * 4 + e] * 256 + al[j + e4 * 4 + e]), p.out_mult), RMAGIC);
              wd[e4] = pack4_sat(rr[0], rr[1], rr[2], rr[3]);
            }
            *reinterpret_cast<uint4*>(o8 + j) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<WT_TMEM_COLS>(tmem_base);
  }
}

bool window_attention_tc_supported(const p2v_window_attention_args& a) {
  return a.dh == WT_DH && a.T <= 64 && a.bias_codes != nullptr && (a.labels == nullptr || a.mask_bits != nullptr) &&
         (reinterpret_cast<uintptr_t>(a.qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(a.bias_codes) & 15) == 0 && (a.H * a.dh) % 16 == 0;
}

int launch_window_attention_tc(const p2v_window_attention_args& a, uint32_t e_mask, cudaStream_t stream) {
  P2V_REQUIRE(window_attention_tc_supported(a), "window_attention_tc: needs head dim 32, T <= 64, bias codes (+ mask bits with labels)");
  encode_tiled_fn enc = get_tensor_map_encoder();
  P2V_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  // int8 [n_windows, T, W] (W = 3*H*32 bytes per token), box = [32 bytes, 64 tokens, 1 window]: rows >= T read as zero
  CUtensorMap tm;
  const cuuint64_t W = cuuint64_t(3) * a.H * WT_DH;
  cuuint64_t dims[3] = {W, cuuint64_t(a.T), cuuint64_t(a.n_windows)};
  cuuint64_t strides[2] = {W, W * cuuint64_t(a.T)};
  cuuint32_t box[3] = {cuuint32_t(WT_DH), 64, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(a.qkv), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  P2V_REQUIRE(r == CUDA_SUCCESS, "window_attention_tc: cuTensorMapEncodeTiled failed (%d) windows=%d T=%d W=%d", int(r), a.n_windows, a.T, int(W));
  WinTcParams p;
  p.T = a.T; p.H = a.H; p.n_windows = a.n_windows; p.wpi = a.windows_per_image;
  p.units = ((a.n_windows + 1) / 2) * a.H;
  p.score_mult = a.score_mult; p.s_attn1 = a.s_attn1; p.s_attn2 = a.s_attn2; p.bias_scale = a.bias_scale; p.out_mult = a.out_mult;
  p.mask_code = a.mask_code; p.e_mask = e_mask;
  p.bias_codes = a.bias_codes; p.mask_bits = a.labels ? reinterpret_cast<const unsigned long long*>(a.mask_bits) : nullptr;
  p.lut = a.lut_dev; p.out_row_map = a.out_row_map; p.out = a.out;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(window_attention_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(WT_SMEM_ALLOC));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(window_attention_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(WT_SMEM_ALLOC));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(window_attention_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(WT_SMEM_ALLOC));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(window_attention_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(WT_SMEM_ALLOC));
    P2V_REQUIRE(e == cudaSuccess, "window_attention_tc: cannot set %zu bytes of dynamic shared memory: %s", WT_SMEM_ALLOC, cudaGetErrorString(e));
  }
  const int grid = std::min(p.units, 3 * sms);
  int ex = 0;
  const bool pot2 = std::frexp(a.s_attn2, &ex) == 0.5f;
  pdl_next_kind(PDL_OTHER);
  const bool mask = p.mask_bits != nullptr;
  if (pot2 && mask) launch_pdl(window_attention_tc_kernel<true, true>, dim3(grid), dim3(WT_THREADS), WT_SMEM_ALLOC, stream, tm, p);
  else if (pot2) launch_pdl(window_attention_tc_kernel<true, false>, dim3(grid), dim3(WT_THREADS), WT_SMEM_ALLOC, stream, tm, p);
  else if (mask) launch_pdl(window_attention_tc_kernel<false, true>, dim3(grid), dim3(WT_THREADS), WT_SMEM_ALLOC, stream, tm, p);
  else launch_pdl(window_attention_tc_kernel<false, false>, dim3(grid), dim3(WT_THREADS), WT_SMEM_ALLOC, stream, tm, p);
  count_launch();
  return check_launch("window_attention_tc");
}

}  // namespace p2v
