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
// One CTA = 8 softmax warps + 1 control warp, two CTAs per SM (256 TMEM columns and ~112 KB shared memory each) so one CTA's
// MMA / TMA latency hides under the other's integer work.  Two softmax warps share each TMEM lane quarter (32 query rows) and
// split the key axis in 16-column units; their partial row max / row sum meet through spare TMEM columns (tcgen05.st, a
// 64-thread named barrier, tcgen05.ld) - the kernel is bound by the softmax warps' instruction issue, so 16 of them per SM
// instead of 8 is what hides the TMEM / shared-memory latencies.  Pass 2 writes the shared-memory address of the table
// entry of d = max - code back over S in TMEM so pass 3 does not requantise again.  A CTA walks heads blockIdx.x,
// blockIdx.x + gridDim.x, ...; each head is 1 or 2 query tiles of 128 rows.
//
// Both element loops are written for the pipe split of the SM sub-partition (ALU pipe: integer add / logic / min-max /
// shifts, FMA pipe: FFMA / FMUL / FADD, each one warp instruction per two cycles): the clamps of pass 2 are an FFMA.SAT,
// table indices leave an FFMA already scaled to byte offsets, pass 3 gets 2^(15-code) and its guard band from the
// exponent field with float multiplies - 4 ALU + 3 FMA (pass 2) and 4.5 ALU + 5 FMA (pass 3) instructions per score
// instead of 6 + 2 and 10 + 4.
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include "tc_common.cuh"

namespace p2v {

constexpr int AT_KV_ROWS = 224;                 // key / value rows staged per head (TMA box, >= T)
constexpr int AT_DH = 64;
constexpr int AT_SM_WARPS = 8;                  // softmax warps: lane quarter = warp & 3, key half = warp >> 2
constexpr int AT_THREADS = AT_SM_WARPS * 32 + 32;
constexpr uint32_t AT_XMAX = 224, AT_XSUM = 228;   // spare TMEM columns: partial row max (1 per half), partial row sum (2 per half)
constexpr uint32_t AT_TMEM_COLS = 256;
constexpr uint32_t AT_OFF_Q = 0;                                  // 2 x [128 x 64]
constexpr uint32_t AT_OFF_K = 2 * 128 * AT_DH;                    // [224 x 64]
constexpr uint32_t AT_OFF_V = AT_OFF_K + AT_KV_ROWS * AT_DH;      // [224 x 64]
constexpr uint32_t AT_OFF_P = AT_OFF_V + AT_KV_ROWS * AT_DH;      // 2 planes x 2 chunks x [128 x 128]
constexpr uint32_t AT_P_CHUNK = 128 * 128, AT_P_PLANE = 2 * AT_P_CHUNK;
// exp_int table: entries -1..255 of 8 bytes (hi, lo); entry -1 repeats entry 0 (a code above 127 saturates, and then the row
// max is 127).  Reciprocal table: the same 8-byte stride AT_RCP_OFF further on, so one address serves both lookups.
constexpr uint32_t AT_OFF_LUT = AT_OFF_P + 2 * AT_P_PLANE;
constexpr uint32_t AT_RCP_OFF = 2064;
constexpr uint32_t AT_OFF_BARS = AT_OFF_LUT + AT_RCP_OFF + 2064;  // 6 mbarriers + the TMEM slot
constexpr uint32_t AT_SMEM = AT_OFF_BARS + 64;
// The kernel has no static shared memory, so its dynamic window starts at the 1 KB the system reserves per CTA and is
// 1024-aligned already; 128 bytes of slack, checked at run time (two CTAs of 112 KB + tables have to fit one SM).
constexpr size_t AT_SMEM_ALLOC = AT_SMEM + 128;
// Asymmetric qact1 (zero point z, omse observers): a constant tile of the byte -z turns the correction terms into MMAs on the
// same accumulators,  (q - z)(k - z) = q k + (-z) k + q (-z) + dh z^2  and  P (v - z) = P v + P (-z),  so the softmax warps only
// add the scalar dh z^2.  Every byte of that tile is the same, so one 512-byte swizzle atom stands for all of it: the operand
// descriptors carry a stride of ZERO between the 8-row groups (K-major, the q / k side) and between the k atoms (MN-major, the
// v side), and the MMA reads the same atom for every group.  The K load writes only the n_pad rows S = q k^T reads, so for
// T <= 208 the atom sits in the unused tail of the K tile (rows 208..215) and the variant keeps two CTAs per SM like the others;
// beyond that it goes behind the barriers and one CTA fits (a 14-KB tile, r2 start, always meant one).
constexpr uint32_t AT_ZC_BYTES = 512;
constexpr uint32_t AT_OFF_ZC_K = AT_OFF_K + 208 * AT_DH;          // rows 208..215 of the K tile
constexpr uint32_t AT_OFF_ZC = (AT_SMEM + 511u) & ~511u;
constexpr size_t AT_SMEM_ALLOC_ZP = AT_OFF_ZC + AT_ZC_BYTES + 128;
static_assert(AT_OFF_ZC_K % 512 == 0 && AT_OFF_ZC % 512 == 0, "a SWIZZLE_64B atom needs 512-byte alignment");
static_assert(AT_OFF_K % 1024 == 0 && AT_OFF_V % 1024 == 0 && AT_OFF_P % 1024 == 0, "swizzled tiles need 1024-byte alignment");

struct AttTcParams {
  int T, H, total_heads;
  int n_pad;      // key columns of S: T rounded up to 16
  int ksteps;     // 32-key MMA steps of P.V: ceil(T / 32)
  int mtiles;     // query tiles of 128 rows
  float score_mult, out_mult;
  const p2v_softmax_lut* lut;
  int8_t* out;
  int exact_probs;          // pass 3 without the guard band (prob_bits_div)
  int k_rows;               // key rows the K load writes (= n_pad; the tile has room for AT_KV_ROWS)
  uint32_t zc_off, smem_alloc;   // ZP variant: shared-memory offset of the constant atom; bytes of dynamic shared memory of this launch
  int zp_qkv, zc2;          // ZP variant: zero point z of q / k / v and dh * z^2
  float zp_score, zp_out;   // ZP variant: zero points of qact_attn1 and qact2
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const int (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t& a, uint32_t& b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(a), "+r"(b) :: "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// partner exchange: both warps of a lane quarter have stored their words to TMEM; returns once each can load the other's
__device__ __forceinline__ void quarter_exchange_sync(uint32_t quarter) {
  tmem_wait_st();
  tc_fence_before();
  named_barrier(1 + quarter, 64);
  tc_fence_after();
}

// POTM: score_mult is a power of two (minmax observers), so S * mult is exact and one FFMA on the magic-biased integer does
// the conversion, the scaling and the RNE together; otherwise the reference's separately rounded product is kept.
// Measured and rejected on passes 2 / 3 (r2, all at 97-98 us for B = 256, H = 6, against 97.2 us as it stands): table entries split
// at 24 bits so that the row sum needs no IADD3.X carry chain; the reciprocal parked in TMEM over S by pass 2 (entry = {exp_int as
// fp32, 1 / exp_int}, sums through two exact float splits) so that pass 3 has no shared-memory lookup - 35 % fewer shared-memory
// wavefronts (the table lookups run at 2.3-3.1 wavefronts per ideal one, random indices in 32 lanes), three FADDs more per score.
// Neither the instruction count of pass 2 (146 -> 95 per 16-score unit with the subnormal-address FFMA below: 101 -> 97 us) nor
// the wavefront count is what bounds the kernel: with 4.5 warps per scheduler it issues one instruction per warp every 7 cycles,
// and no single stall reason exceeds 15 % of the samples (profiles/r2a_stalls.txt, profiles/r2_stalls.txt).
template <bool POTM, bool ZP>
__global__ void __launch_bounds__(AT_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    AttTcParams p) {
  static_assert(!(POTM && ZP), "zero points come with raw fp32 scales");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (base - smem_u32(smem_raw) + ((ZP && p.zc_off >= AT_SMEM) ? p.zc_off + AT_ZC_BYTES : AT_SMEM) > p.smem_alloc) __trap();     // dynamic window not aligned as assumed
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + AT_OFF_BARS;
  const uint32_t bar_qk = bars, bar_v = bars + 8, bar_s = bars + 16, bar_p = bars + 24, bar_o = bars + 32, bar_free = bars + 40;
  volatile uint32_t& tmem_slot = *reinterpret_cast<volatile uint32_t*>(gbase + AT_OFF_BARS + 48);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T, H = p.H;
  uint2* s_lut = reinterpret_cast<uint2*>(gbase + AT_OFF_LUT);                       // s_lut[1 + d]
  float2* s_rcp = reinterpret_cast<float2*>(gbase + AT_OFF_LUT + AT_RCP_OFF);        // s_rcp[1 + d].x

  if (threadIdx.x == 0) {
    mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    mbar_init(bar_p, AT_SM_WARPS); mbar_init(bar_free, AT_SM_WARPS);
    fence_mbar_init();
  }
  if (warp == AT_SM_WARPS) tmem_alloc<AT_TMEM_COLS>(base + AT_OFF_BARS + 48);
  for (int i = threadIdx.x; i < 257; i += AT_THREADS) {
    const int d = max(i - 1, 0);
    s_lut[i] = make_uint2(p.lut->hi[d], p.lut->lo[d]);
    s_rcp[i] = make_float2(prob_rcp(p.lut->exp_f32[d]), p.lut->exp_f32[d]);
  }
  // -z as an int8 byte; z = -128 has no int8 negative: the tile then holds 64 and the correction MMAs are issued twice
  const int zreps = ZP ? (p.zp_qkv == -128 ? 2 : 1) : 0;
  if (ZP) {
    const uint32_t zb = uint32_t(p.zp_qkv == -128 ? 64 : -p.zp_qkv) & 0xffu;
    const uint32_t zw = zb * 0x01010101u;
    for (int i = threadIdx.x; i < int(AT_ZC_BYTES) / 16; i += AT_THREADS)
      asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(base + p.zc_off + uint32_t(i) * 16u), "r"(zw) : "memory");
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = tmem_slot;

  if (warp == AT_SM_WARPS) {
    // ================= control warp: TMA producer + MMA issuer =================
    // The whole warp walks the loop (warp-uniform, so addresses and descriptors stay in uniform registers); one elected lane
    // issues each TMA / MMA / commit (tc_common.cuh: elect_one).
    {
      if (elect_one()) {
        tma_prefetch_map(&tmQ);
        tma_prefetch_map(&tmK);
        tma_prefetch_map(&tmV);
      }
      const int row_bytes_h = AT_DH;  // column offsets inside a token row: q at h*64, k at (H+h)*64, v at (2H+h)*64
      auto load_qk = [&](int hd) {
        const int b = hd / H, h = hd % H;
        mbar_expect_tx(bar_qk, uint32_t(p.mtiles) * 128u * AT_DH + uint32_t(p.k_rows) * AT_DH);
        for (int mt = 0; mt < p.mtiles; ++mt) tma_load_3d(base + AT_OFF_Q + mt * 128 * AT_DH, &tmQ, bar_qk, h * row_bytes_h, mt * 128, b);
        tma_load_3d(base + AT_OFF_K, &tmK, bar_qk, (H + h) * row_bytes_h, 0, b);
      };
      auto load_v = [&](int hd) {
        const int b = hd / H, h = hd % H;
        mbar_expect_tx(bar_v, AT_KV_ROWS * AT_DH);
        tma_load_3d(base + AT_OFF_V, &tmV, bar_v, (2 * H + h) * row_bytes_h, 0, b);
      };
      const uint32_t idesc_qk = make_i8_idesc(128, p.n_pad, true, true);
      const uint32_t idesc_pv = make_i8_idesc(128, AT_DH, false, true, false, true);   // P u8 K-major, V s8 MN-major
      if (int(blockIdx.x) < p.total_heads) {
        if (elect_one()) { load_qk(blockIdx.x); load_v(blockIdx.x); }
        __syncwarp();
      }
      uint32_t n = 0, hcount = 0;
      for (int hd = blockIdx.x; hd < p.total_heads; hd += gridDim.x, ++hcount) {
        const int next = hd + gridDim.x;
        mbar_wait(bar_qk, hcount & 1u);
        for (int mt = 0; mt < p.mtiles; ++mt, ++n) {
          mbar_wait(bar_free, (n & 1u) ^ 1u);          // previous tile's O has left TMEM
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < AT_DH / 32; ++k)
              umma_i8(tmem_base, make_smem_desc(base + AT_OFF_Q + mt * 128 * AT_DH + k * 32, 16, 512, UMMA_LAYOUT_SW64),
                      make_smem_desc(base + AT_OFF_K + k * 32, 16, 512, UMMA_LAYOUT_SW64), idesc_qk, uint32_t(k > 0));
            if (ZP) {       // + (-z) k + q (-z): the constant tile as the A operand, then as the B operand
              for (int rep = 0; rep < zreps; ++rep)
#pragma unroll
                for (int k = 0; k < AT_DH / 32; ++k) {
                  umma_i8(tmem_base, make_smem_desc(base + p.zc_off + k * 32, 16, 0, UMMA_LAYOUT_SW64),
                          make_smem_desc(base + AT_OFF_K + k * 32, 16, 512, UMMA_LAYOUT_SW64), idesc_qk, 1u);
                  umma_i8(tmem_base, make_smem_desc(base + AT_OFF_Q + mt * 128 * AT_DH + k * 32, 16, 512, UMMA_LAYOUT_SW64),
                          make_smem_desc(base + p.zc_off + k * 32, 16, 0, UMMA_LAYOUT_SW64), idesc_qk, 1u);
                }
            }
            tc_commit(bar_s);
          }
          __syncwarp();
          if (mt == p.mtiles - 1 && next < p.total_heads) {
            mbar_wait(bar_s, n & 1u);                  // q / k tiles have been read: refill them for the next head
            if (elect_one()) load_qk(next);
            __syncwarp();
          }
          mbar_wait(bar_p, n & 1u);                    // P planes written, S consumed
          if (mt == 0) mbar_wait(bar_v, hcount & 1u);
          tc_fence_after();
          if (elect_one()) {
            for (int plane = 0; plane < 2; ++plane) {
              for (int ks = 0; ks < p.ksteps; ++ks)
                umma_i8(tmem_base + plane * AT_DH,
                        make_kmajor_sw128_desc(base + AT_OFF_P + plane * AT_P_PLANE + (ks >> 2) * AT_P_CHUNK + (ks & 3) * 32),
                        make_smem_desc(base + AT_OFF_V + ks * 32 * AT_DH, AT_KV_ROWS * AT_DH, 512, UMMA_LAYOUT_SW64), idesc_pv,
                        uint32_t(ks > 0));
              if (ZP) {     // + P (-z)
                for (int rep = 0; rep < zreps; ++rep)
                  for (int ks = 0; ks < p.ksteps; ++ks)
                    umma_i8(tmem_base + plane * AT_DH,
                            make_kmajor_sw128_desc(base + AT_OFF_P + plane * AT_P_PLANE + (ks >> 2) * AT_P_CHUNK + (ks & 3) * 32),
                            make_smem_desc(base + p.zc_off, 0, 0, UMMA_LAYOUT_SW64), idesc_pv, 1u);
              }
            }
            tc_commit(bar_o);
          }
          __syncwarp();
          if (mt == p.mtiles - 1 && next < p.total_heads) {
            mbar_wait(bar_o, n & 1u);                  // v tile has been read
            if (elect_one()) load_v(next);
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ================= softmax warps: thread = one query row (TMEM lane) x one half of the key axis =================
    const uint32_t quarter = uint32_t(warp) & 3u, half = uint32_t(warp) >> 2;
    const bool exact_probs = p.exact_probs != 0;
    const int rloc = int(quarter) * 32 + lane;
    const uint32_t tlane = tmem_base + ((quarter * 32u) << 16);
    const float mult = p.score_mult;
    const uint32_t lut32 = base + AT_OFF_LUT + 8u;        // shared address of exp_int[0]
    // RMAGIC + S*mult in one rounding from t = RMAGIC + S (exact for |S| < 2^22; |S| <= 64 * 2^14 here)
    const float potm_c = fsub(RMAGIC, fmul(RMAGIC, mult));
    // key axis in 16-column units: units with a column < T hold scores, the P operand needs 2 * ksteps units (zero padded)
    const int units_p = 2 * p.ksteps, units_s = (T + 15) >> 4, u_mid = (units_p + 1) >> 1;
    const int u_begin = half ? u_mid : 0, u_end = half ? units_p : u_mid;
    const int u_send = min(u_end, units_s);               // my units that hold scores: [u_begin, u_send)
    uint32_t n = 0;
    for (int hd = blockIdx.x; hd < p.total_heads; hd += gridDim.x) {
      const int b = hd / H, h = hd % H;
      for (int mt = 0; mt < p.mtiles; ++mt, ++n) {
        const int row = mt * 128 + rloc;
        const bool warp_live = mt * 128 + int(quarter) * 32 < T;     // the same for both warps of a quarter
        mbar_wait(bar_s, n & 1u);
        tc_fence_after();
        if (warp_live) {
          // S stays in TMEM and is re-read by every pass (a rolled loop over units keeps the kernel in the instruction cache).
          // Only the unit that straddles column T needs per-element masking: [u_begin, u_full) run the unmasked bodies.
          const int u_full = min(u_send, T >> 4);
          // ---- pass 1: row max of the raw scores; the requantisation is monotone, so max code = code(max S)
          int smax = INT_MIN;
          auto max_unit = [&](auto masked, int u) {
            int acc[16];
            tmem_ld16(tlane + u * 16, acc);
            const int nv = T - u * 16;
#pragma unroll
            for (int e = 0; e < 16; ++e) smax = max(smax, (!decltype(masked)::value || e < nv) ? acc[e] : INT_MIN);
          };
#pragma unroll 1
          for (int u = u_begin; u < u_full; ++u) max_unit(std::false_type{}, u);
          if (u_full < u_send) max_unit(std::true_type{}, u_full);
          {
            uint32_t o0, o1;
            tmem_st2(tlane + AT_XMAX + 2 * half, uint32_t(smax), 0u);
            quarter_exchange_sync(quarter);
            tmem_ld2(tlane + AT_XMAX + 2 * (half ^ 1u), o0, o1);
            smax = max(smax, int(o0));
          }
          // ZP: the accumulator holds S - dh z^2 (the same offset in every column, so the row max is unaffected)
          const int mx = ZP ? sat_s8(fadd(fmul(float(smax + p.zc2), mult), p.zp_score)) : sat_s8(fmul(float(smax), mult));
          // ---- pass 2: exact row sum of exp_int(max - code) over my units; the table address of d = max - code goes back to
          //      TMEM over S.  u = RMAGIC + RNE(S * mult); v = sat((code + 128) / 256) clamps the code to [-128, 128] (FFMA.SAT,
          //      exact); 2048 v = 8 (code_sat + 128), so one FFMA in the subnormal range, rowk 2^-149 - v 2^-138, has the bit pattern
          //      &exp_int[mx + 128] - 8 (code_sat + 128) = &exp_int[mx - code_sat] (no flush-to-zero in this build).
          //      code_sat = 128 only when mx = 127: entry -1.
          const float rowk_f = __uint_as_float(lut32 + uint32_t(mx + 128) * 8u);      // &exp_int[mx + 128] (< 2^23: a subnormal)
          unsigned long long sum = 0;
          auto sum_unit = [&](auto masked, int u) {
            int acc[16];
            tmem_ld16(tlane + u * 16, acc);
            const int nv = T - u * 16;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float uq = POTM ? __fmaf_rn(__int_as_float(acc[e] + 0x4B400000), mult, potm_c)
                               : ZP ? fadd(fadd(fmul(__int2float_rn(acc[e] + p.zc2), mult), p.zp_score), RMAGIC)
                                    : fadd(fmul(__int2float_rn(acc[e]), mult), RMAGIC);
              const float v = fma_sat(uq, 0.00390625f, -49151.5f);
              // the table address as the bit pattern of a subnormal: (rowk - 2048 v) * 2^-149, exact (v is a multiple of 1/256)
              const uint32_t addr = __float_as_uint(__fmaf_rn(v, __uint_as_float(0x80000800u), rowk_f));
              acc[e] = int(addr);
              uint32_t vx, vy;
              asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(vx), "=r"(vy) : "r"(addr));
              if (!decltype(masked)::value || e < nv) sum += (static_cast<unsigned long long>(vx) << 32) | vy;
            }
            tmem_st16(tlane + u * 16, acc);
          };
#pragma unroll 1
          for (int u = u_begin; u < u_full; ++u) sum_unit(std::false_type{}, u);
          if (u_full < u_send) sum_unit(std::true_type{}, u_full);
          {
            uint32_t o0, o1;
            tmem_st2(tlane + AT_XSUM + 2 * half, uint32_t(sum), uint32_t(sum >> 32));
            quarter_exchange_sync(quarter);           // also orders the address write-back before pass 3's loads
            tmem_ld2(tlane + AT_XSUM + 2 * (half ^ 1u), o0, o1);
            sum += (static_cast<unsigned long long>(o1) << 32) | o0;
          }
          const float tot = __ull2float_rn(sum);      // the exact integer sum (< 2^64: intmath.build_softmax_lut bounds the table), rounded once
          const float tot2 = fmul(tot, 2.0f), tot43 = fmul(tot, 1.33333337306976318359375f);
          // ---- pass 3: probabilities 2^(15-code) as hi / lo byte planes in the UMMA K-major SW128 layout
          const uint32_t prow = base + AT_OFF_P + uint32_t(rloc) * 128u;
          const uint32_t sw = uint32_t(rloc & 7);
          auto store_unit = [&](int u, const uint32_t (&hi)[4], const uint32_t (&lo)[4]) {
            const uint32_t dst = prow + uint32_t(u >> 3) * AT_P_CHUNK + ((uint32_t(u & 7) ^ sw) << 4);    // unit = 16 bytes along the key axis
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + AT_P_PLANE), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
          };
          auto prob_unit = [&](auto masked, int u) {
            int aa[16];
            tmem_ld16(tlane + u * 16, aa);
            const int nv = T - u * 16;
            uint32_t pv[16];
            float gmax = 0.f;
            if (exact_probs) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                float rcp, ef;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(rcp), "=f"(ef) : "r"(uint32_t(aa[e])), "n"(AT_RCP_OFF));
                pv[e] = prob_bits_div(tot, rcp, ef);
              }
            } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              float rcp;
              asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(rcp) : "r"(uint32_t(aa[e])), "n"(AT_RCP_OFF));
              pv[e] = prob_bits_fast(tot2, tot43, rcp, gmax);
            }
            }
            if (!(gmax < PROB_GUARD)) {   // some element sits next to a rounding / log2 boundary (or is not finite): redo the unit with the IEEE division
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                uint32_t vx, vy;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(vx), "=r"(vy) : "r"(uint32_t(aa[e])));
                const uint32_t big = log2_code(tot, __ull2float_rn((static_cast<unsigned long long>(vx) << 32) | vy));
                pv[e] = shr_clamp(0x8000u, big);
              }
            }
            uint32_t lo[4], hi[4];
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              if (decltype(masked)::value) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (e4 * 4 + e >= nv) pv[e4 * 4 + e] = 0u;
              }
              const uint32_t p01 = __byte_perm(pv[e4 * 4], pv[e4 * 4 + 1], 0x5410), p23 = __byte_perm(pv[e4 * 4 + 2], pv[e4 * 4 + 3], 0x5410);
              lo[e4] = __byte_perm(p01, p23, 0x6420);
              hi[e4] = __byte_perm(p01, p23, 0x7531);
            }
            store_unit(u, hi, lo);
          };
#pragma unroll 1
          for (int u = u_begin; u < u_full; ++u) prob_unit(std::false_type{}, u);
          if (u_full < u_send) prob_unit(std::true_type{}, u_full);
          {
            const uint32_t zero[4] = {0u, 0u, 0u, 0u};
#pragma unroll 1
            for (int u = max(u_begin, units_s); u < u_end; ++u) store_unit(u, zero, zero);    // zero padding of the P operand
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        // ---- O = 256*hi + lo -> qact2 codes; this warp's half of the 64 head channels
        mbar_wait(bar_o, n & 1u);
        tc_fence_after();
        if (warp_live) {
          int8_t* orow = p.out + (int64_t(b) * T + row) * (int64_t(H) * AT_DH) + h * AT_DH + half * 32;
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
                float r[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float o = fmul(__int2float_rn(ah[j + e4 * 4 + e] * 256 + al[j + e4 * 4 + e]), p.out_mult);
                  r[e] = fadd(ZP ? fadd(o, p.zp_out) : o, RMAGIC);
                }
                w[e4] = pack4_sat(r[0], r[1], r[2], r[3]);
              }
              *reinterpret_cast<uint4*>(orow + j) = make_uint4(w[0], w[1], w[2], w[3]);
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
  if (warp == AT_SM_WARPS) {
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
  CUtensorMap tmQ, tmK, tmV;
  const int W = 3 * a.H * AT_DH;
  if (int r = make_tmap_qkv(&tmQ, a.qkv, a.B, a.T, W, 128)) return r;
  const int n_pad = (a.T + 15) / 16 * 16;
  if (int r = make_tmap_qkv(&tmK, a.qkv, a.B, a.T, W, n_pad)) return r;          // only the rows S = q k^T reads
  if (int r = make_tmap_qkv(&tmV, a.qkv, a.B, a.T, W, AT_KV_ROWS)) return r;
  AttTcParams p;
  p.T = a.T; p.H = a.H; p.total_heads = a.B * a.H;
  p.n_pad = n_pad;
  p.k_rows = n_pad;
  static const int force_mode = getenv("P2V_ATT_EXACT") ? atoi(getenv("P2V_ATT_EXACT")) : -1;      // triage: pin the mode
  p.exact_probs = force_mode >= 0 ? force_mode : a.prob_mode;
  p.zc_off = 0; p.smem_alloc = uint32_t(AT_SMEM_ALLOC);
  p.ksteps = (a.T + 31) / 32;
  p.mtiles = (a.T + 127) / 128;
  p.score_mult = a.score_mult; p.out_mult = a.out_mult;
  p.lut = a.lut_dev; p.out = a.out;
  p.zp_qkv = a.zp_qkv; p.zc2 = a.dh * a.zp_qkv * a.zp_qkv; p.zp_score = a.zp_score; p.zp_out = a.zp_out;
  const bool zp = a.zp_qkv != 0 || a.zp_score != 0.f || a.zp_out != 0.f;
  P2V_REQUIRE(a.zp_qkv >= -128 && a.zp_qkv <= 127, "attention: zp_qkv=%d outside the int8 range", a.zp_qkv);
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(AT_SMEM_ALLOC));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(AT_SMEM_ALLOC));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(AT_SMEM_ALLOC_ZP));
    P2V_REQUIRE(e == cudaSuccess, "attention_tc: cannot set %zu bytes of dynamic shared memory: %s", AT_SMEM_ALLOC_ZP, cudaGetErrorString(e));
  }
  if (zp) {      // asymmetric quantizers: constant-atom variant
    const bool in_k_tail = n_pad <= 208;
    p.zc_off = in_k_tail ? AT_OFF_ZC_K : AT_OFF_ZC;
    p.smem_alloc = uint32_t(in_k_tail ? AT_SMEM_ALLOC : AT_SMEM_ALLOC_ZP);
    pdl_next_kind(PDL_ATTENTION);
    launch_pdl(attention_tc_kernel<false, true>, dim3(std::min(p.total_heads, (in_k_tail ? 2 : 1) * sms)), dim3(AT_THREADS), p.smem_alloc, stream, tmQ, tmK, tmV, p);
    count_launch();
    return check_launch("attention_tc");
  }
  static const int ctas_per_sm = [] { const char* v = getenv("P2V_ATT_CTAS"); return v ? std::max(1, std::min(2, atoi(v))) : 2; }();   // perf triage only
  const int grid = std::min(p.total_heads, ctas_per_sm * sms);
  // power-of-two score multiplier in [2^-20, 2^8]: RMAGIC * (1 - mult) is then exact and so is the fused scaling
  int mexp = 0;
  const bool potm = std::frexp(a.score_mult, &mexp) == 0.5f && mexp >= -19 && mexp <= 9;
  pdl_next_kind(PDL_ATTENTION);
  if (potm) launch_pdl(attention_tc_kernel<true, false>, dim3(grid), dim3(AT_THREADS), AT_SMEM_ALLOC, stream, tmQ, tmK, tmV, p);
  else launch_pdl(attention_tc_kernel<false, false>, dim3(grid), dim3(AT_THREADS), AT_SMEM_ALLOC, stream, tmQ, tmK, tmV, p);
  count_launch();
  return check_launch("attention_tc");
}

}  // namespace p2v
