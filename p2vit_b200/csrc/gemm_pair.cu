// Block GEMMs (qkv / proj / fc1 / fc2) on CTA pairs: int8 x int8 -> int32 tcgen05.mma.cta_group::2 with the QAct
// epilogues fused, built for the case the old one-tile-per-group kernel (gemm_tc.cu) handles badly: K is short
// (384..1536), so the kernel is bound by the epilogue's instruction issue, by L2 -> SM operand traffic and by
// uncoalesced row-wise stores, not by the tensor pipe.
//
//   * cluster of 2 CTAs = one 256 x BN output tile (BN in {128, 192, 256}); each CTA loads its own 128 rows of A and
//     HALF of the W tile (TMA, SWIZZLE_128B); the leader CTA issues M256 x BN x K32 MMAs that read both halves, so a CTA
//     pulls (128 + BN/2) * K bytes per 128 x BN outputs - half the W traffic of cta_group::1;
//   * accumulators: 2 TMEM stages of 256 columns in each CTA; all 16 epilogue warps drain stage s while the MMAs of
//     the next tile fill stage s^1;
//   * epilogue warp = 32 rows (its TMEM lane quarter) x BN/4 columns, in chunks of 16 columns; the tcgen05.ld of chunk
//     c+1 is in flight while chunk c is computed; per-column constants live in a warp-private shared-memory table,
//     fetched one tile ahead;
//   * output codes are staged in shared memory (swizzled, conflict-free 16-byte writes) and leave with one TMA store
//     per warp and tile; the residual stream of the RESIDUAL epilogue arrives the same way (TMA load issued by the
//     producer warp one tile ahead) and is overwritten in place by the output;
//   * requantisation without F2I / FRND (both run on the 16-lane XU pipe): RNE through the 1.5*2^23 magic constant,
//     saturation by clamping the biased float, the code is its low byte.  Division by a non-power-of-two scale s:
//     RNE(y * r_lo) == RNE(y * r_hi) for r_lo < 1/s < r_hi (4 ulps apart) proves RNE(fl(y / s)) without dividing;
//     the rare disagreement re-runs the chunk with the IEEE division for the columns concerned.
//
// Same results, bit for bit, as gemm_tc.cu / gemm_simt.cu (tests/test_gpu_ops.py cross-checks the three).
#include <algorithm>
#include <cstdlib>
#include "tc_common.cuh"
#include "epilogue.cuh"

namespace p2v {

constexpr int PBM = 128;                 // rows per CTA (pair tile: 256)
constexpr int PBK = 128;                 // K bytes per pipeline stage (one 128-byte swizzle row)
constexpr int P_EPI_WARPS = 16;
constexpr int P_EPI_GROUPS = 4;          // column groups of 4 warps (one per TMEM lane quarter) sharing a constant table
constexpr int P_THREADS = 128 + P_EPI_WARPS * 32;   // warpgroup 0: TMA producer, MMA issuer, 2 idle warps; warpgroups 1-4: epilogue
constexpr int P_ACC_COLS = 256;          // TMEM columns per accumulator stage
constexpr uint32_t P_A_BYTES = PBM * PBK;
constexpr uint32_t P_B_BYTES = 128 * PBK;   // room for BN/2 <= 128 rows of W

constexpr float RMAGIC_LO = RMAGIC - 128.f, RMAGIC_HI = RMAGIC + 127.f;

struct PairGeom {
  int BN;        // columns per pair tile
  int W;         // columns per epilogue warp = BN / 4 (32, 48 or 64)
  int tiles_m, tiles_n, tiles;   // pair tiles
  int nkb;       // k blocks
  int swz;       // staging swizzle: 0 none (48-byte rows), 1 SWIZZLE_32B, 2 SWIZZLE_64B
  int nslot_log2;   // RESIDUAL: residual / output staging slots in flight (2 or 4 slots of 512 * W bytes)
  int nstages;      // operand ring depth (<= P_MAX_STAGES)
  int bres;         // 1: this CTA's half of the W column tile (all of K) stays resident in shared memory, only A streams;
                    //    tiles are then walked row-block fastest in one contiguous range per CTA pair, so W is reloaded only
                    //    when the range crosses into the next column tile.  0: A and W stream together, tiles column fastest.
  uint32_t stage_bytes, off_bres, off_stg, off_prm, off_gst;   // shared-memory layout relative to the 1024-aligned base
  uint32_t smem_bytes;   // dynamic shared memory of the launch
};
constexpr int P_MAX_STAGES = 8;

// -DPAIR_TRACE (tools/pair_trace.py): CTA pair 0 writes clock64 stamps of its producer / MMA / first epilogue warp per tile
// into the buffer passed as out_f32; not compiled into the product library.
#ifdef PAIR_TRACE
#define PTRACE(role, tile, slot_)                                                                                   \
  do {                                                                                                              \
    if (p.out_f32 != nullptr && pair == 0 && (tile) < 64)                                                           \
      reinterpret_cast<long long*>(p.out_f32)[((int(rank) * 19 + (role)) * 64 + int(tile)) * 4 + (slot_)] = clock64(); \
  } while (0)
#else
#define PTRACE(role, tile, slot_) do {} while (0)
#endif

// The pair's tiles in execution order.  Streaming mode: t = pair, pair + npairs, ... with the column block fastest (the pairs
// running at the same time share A row blocks through L2).  Resident-W mode: the contiguous range [pair*T/P, (pair+1)*T/P) of
// the row-block-fastest order, so the column block changes at most a few times per pair.
struct TileIter {
  int mt, nt, left;      // current tile, tiles left including it
  int dm, dn, tiles_m, tiles_n, bres;
  __device__ TileIter(const PairGeom& g, uint32_t pair, uint32_t npairs) : tiles_m(g.tiles_m), tiles_n(g.tiles_n), bres(g.bres) {
    if (bres) {
      const long long T = g.tiles;
      const int t0 = int(T * pair / npairs), t1 = int(T * (pair + 1) / npairs);
      left = t1 - t0;
      nt = t0 / tiles_m; mt = t0 % tiles_m; dm = dn = 0;
    } else {
      left = int(pair) < g.tiles ? (g.tiles - 1 - int(pair)) / int(npairs) + 1 : 0;
      mt = int(pair) / tiles_n; nt = int(pair) % tiles_n; dm = int(npairs) / tiles_n; dn = int(npairs) % tiles_n;
    }
  }
  __device__ void next() {
    --left;
    if (bres) {
      if (++mt == tiles_m) { mt = 0; ++nt; }
    } else {
      nt += dn; mt += dm;
      if (nt >= tiles_n) { nt -= tiles_n; ++mt; }
    }
  }
};

// ---- parameter rows (warp-private table: rows of 64 floats)
// ZP (asymmetric quantizers, omse; raw fp32 scales only): the accumulator loses zp_corr[n] = zp_in * sum_k W[n,k] first, and the
// output code of REQUANT / GELU is sat(RNE(fl(fl(y / scale) + out_zp))) - the exactly rounded quotient itself is needed, so it
// comes from the scale's reciprocal with two residual corrections (div_rb) instead of the reciprocal-bounds test.
template <int EPI, bool POT, bool ZP = false>
__host__ __device__ constexpr int prm_rows() {
  return EPI == P2V_EPI_RESIDUAL ? (ZP ? 10 : 9) : (POT ? (EPI == P2V_EPI_GELU ? 3 : 2) : 5);
}
enum { PR_S = 0, PR_B = 1, PR_RO = 2,                 // POT GELU: 1/out_scale
       PR_OLO = 2, PR_OHI = 3, PR_O = 4,              // general REQUANT / GELU
       PR_ZRO = 2, PR_ZC = 3,                         // ZP REQUANT / GELU: 1/out_scale, zp_corr (int32 bits); PR_O as above
       PR_MLO = 2, PR_MHI = 3, PR_M = 4, PR_RS = 5, PR_ROLO = 6, PR_ROHI = 7, PR_RO_ = 8, PR_RZC = 9 };   // RESIDUAL (+ zp_corr)

// ---- cluster / 2-SM primitives
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (release at CTA scope) as in CUTLASS' ClusterBarrier::arrive(cta_id): the hand-off it guards is
  // TMEM (ordered by tcgen05.fence), and a cluster-scope release costs a full memory barrier per tile and warp
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on a barrier of either CTA of the pair (cluster address)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc_pair(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_i8_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
// tcgen05.wait::ld with the destination registers as in/out operands: no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_wait_ld16(int (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}

// ---- exact requantisation helpers
// r_lo < 1/d < r_hi, 2..4 ulps either side of rs = RN(1/d): fl(y/d) lies between y*r_lo and y*r_hi
__device__ __forceinline__ float recip_lo(float rs) { return __fmaf_rn(rs, -0x1p-22f, rs); }
__device__ __forceinline__ float recip_hi(float rs) { return __fmaf_rn(rs, 0x1p-22f, rs); }

// clamp(RNE(fl(y / d)), -128, 127) + RMAGIC.  Fast pass: `flag` collects the columns whose two bounds round differently
// (or saturate far out); EXACT pass: those columns take the IEEE division.
template <bool EXACT, bool CLAMP = false>
__device__ __forceinline__ float quant_iv(float y, float rlo, float rhi, float d, uint32_t& flag) {
  const float tlo = __fmaf_rn(y, rlo, RMAGIC);
  float thi = __fmaf_rn(y, rhi, RMAGIC);
  const uint32_t diff = __float_as_uint(tlo) ^ __float_as_uint(thi);
  if (EXACT) {
    if (diff) thi = fadd(fminf(fmaxf(rintf(fdiv(y, d)), -128.f), 127.f), RMAGIC);
  } else {
    flag |= diff;
  }
  return CLAMP ? fminf(fmaxf(thi, RMAGIC_LO), RMAGIC_HI) : thi;     // final codes are saturated by pack4_sat
}
// RNE(x) + RMAGIC for x already on the output grid (power-of-two scales); saturated by pack4_sat
__device__ __forceinline__ float quant_pot(float x) { return fadd(x, RMAGIC); }

// One chunk: 16 accumulator columns of one row -> 16 output codes (4 packed words).  prm = this warp's table + the
// chunk's column offset.  `resx` = the row's 16 residual codes with the sign bits flipped (code + 128 as u8).
// Columns are processed four at a time; a group in which some column's reciprocal bounds disagree (a quotient within a few
// ulps of a rounding tie, ~1e-5 of the quotients) is redone on the spot with the IEEE division for those columns - the branch
// is short and keeps no extra state alive, so a hit costs a few hundred cycles of one warp instead of stalling the tile.
// 16-byte read of the warp's constant table through its 32-bit shared address.  The table used to be addressed through a generic
// pointer derived from the dynamic shared-memory base; under register pressure the compiler re-derived that pointer for every
// 4-column group (S2R SR_CgaCtaId / SR_TID.X, the 1024-byte alignment arithmetic: ~4 of the 37 instructions per 32 outputs of the
// RESIDUAL epilogue, ncu r1n).  An opaque 32-bit address cannot be rematerialised.
__device__ __forceinline__ float4 prm_ld4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// GST: 0 = no GELU step tables, 1 = step tables with the near-threshold distance test, 2 = clean tables (gelu_table.cu), no test
template <int EPI, bool POT, int GST, bool ZP>
__device__ __forceinline__ void pair_chunk(const float* __restrict__ prmg, const uint32_t prm, const int (&acc)[16], const uint4 resx, uint4& out,
                                           const GeluSteps& gst, const float out_zp) {
  uint32_t ow[4];
  const uint32_t rw[4] = {resx.x, resx.y, resx.z, resx.w};
  uint32_t gst_near = 0xffffffffu;
#pragma unroll
  for (int j4 = 0; j4 < 16; j4 += 4) {
    // RESIDUAL (9 rows, the tightest on registers) reads through the opaque shared address; the other epilogues keep the generic
    // pointer: their volatile loads would be ordered against the GELU table lookups and cost more than the re-derivation saves
    const float4 S4 = EPI == P2V_EPI_RESIDUAL ? prm_ld4(prm + (PR_S * 64 + j4) * 4) : *reinterpret_cast<const float4*>(prmg + PR_S * 64 + j4);
    const float4 B4 = EPI == P2V_EPI_RESIDUAL ? prm_ld4(prm + (PR_B * 64 + j4) * 4) : *reinterpret_cast<const float4*>(prmg + PR_B * 64 + j4);
    const float Sv[4] = {S4.x, S4.y, S4.z, S4.w}, Bv[4] = {B4.x, B4.y, B4.z, B4.w};
    float t[4];
    if (EPI == P2V_EPI_RESIDUAL) {
      const float4 a4 = prm_ld4(prm + (PR_MLO * 64 + j4) * 4);
      const float4 b4 = prm_ld4(prm + (PR_MHI * 64 + j4) * 4);
      const float4 c4 = prm_ld4(prm + (PR_M * 64 + j4) * 4);
      const float4 d4 = prm_ld4(prm + (PR_RS * 64 + j4) * 4);
      const float4 e4 = prm_ld4(prm + (PR_ROLO * 64 + j4) * 4);
      const float4 f4 = prm_ld4(prm + (PR_ROHI * 64 + j4) * 4);
      const float MLv[4] = {a4.x, a4.y, a4.z, a4.w}, MHv[4] = {b4.x, b4.y, b4.z, b4.w}, Mv[4] = {c4.x, c4.y, c4.z, c4.w};
      const float RSv[4] = {d4.x, d4.y, d4.z, d4.w}, OLv[4] = {e4.x, e4.y, e4.z, e4.w}, OHv[4] = {f4.x, f4.y, f4.z, f4.w};
      int zc[4] = {0, 0, 0, 0};
      if (ZP) {
        const float4 z4 = prm_ld4(prm + (PR_RZC * 64 + j4) * 4);
        zc[0] = __float_as_int(z4.x); zc[1] = __float_as_int(z4.y); zc[2] = __float_as_int(z4.z); zc[3] = __float_as_int(z4.w);
      }
      float y[4], r[4];
      uint32_t flag = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float af = __int2float_rn(acc[j4 + e] - zc[e]);
        y[e] = POT ? __fmaf_rn(af, Sv[e], Bv[e]) : fadd(fmul(af, Sv[e]), Bv[e]);
        const float k = fsub(quant_iv<false, true>(y[e], MLv[e], MHv[e], Mv[e], flag), RMAGIC);        // qact after the GEMM (code)
        // residual code: byte e of the word (sign already flipped) -> 2^23 + (code + 128) -> code, exactly
        r[e] = fmul(fsub(__uint_as_float(__byte_perm(rw[j4 >> 2], 0x4B000000u, 0x7540 + e)), 8388736.f), RSv[e]);
        t[e] = quant_iv<false>(fadd(r[e], fmul(k, Mv[e])), OLv[e], OHv[e], 1.f, flag);
      }
      if (flag) {
        const float4 g4 = prm_ld4(prm + (PR_RO_ * 64 + j4) * 4);
        const float Ov[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float k = fsub(quant_iv<true, true>(y[e], MLv[e], MHv[e], Mv[e], flag), RMAGIC);
          t[e] = quant_iv<true>(fadd(r[e], fmul(k, Mv[e])), OLv[e], OHv[e], Ov[e], flag);
        }
      }
    } else if (POT && EPI == P2V_EPI_REQUANT) {
#pragma unroll
      for (int e = 0; e < 4; ++e) t[e] = quant_pot(__fmaf_rn(__int2float_rn(acc[j4 + e]), Sv[e], Bv[e]));   // S, B pre-divided by out_scale
    } else if (EPI == P2V_EPI_GELU && GST) {
      // step tables (common.cuh: gelu_steps_code): ~19 instructions per column (9 ALU-pipe, 8 FMA-pipe) instead of ~40 for erff.  No branch inside the
      // chunk, so the 16 columns' lookup chains (two dependent shared-memory loads each) overlap; the near-threshold test is
      // taken once per chunk, after the loop.
      uint32_t q[4];
      int zc[4] = {0, 0, 0, 0};
      if (ZP) {
        const float4 z4 = *reinterpret_cast<const float4*>(prmg + PR_ZC * 64 + j4);
        zc[0] = __float_as_int(z4.x); zc[1] = __float_as_int(z4.y); zc[2] = __float_as_int(z4.z); zc[3] = __float_as_int(z4.w);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float af = __int2float_rn(acc[j4 + e] - zc[e]);
        q[e] = gelu_steps_code<GST == 1>(POT ? __fmaf_rn(af, Sv[e], Bv[e]) : fadd(fmul(af, Sv[e]), Bv[e]), gst, gst_near);
      }
      ow[j4 >> 2] = pack4_low_bytes(q[0], q[1], q[2], q[3]);
      continue;
    } else if (POT && EPI == P2V_EPI_GELU) {
      const float4 R4 = *reinterpret_cast<const float4*>(prmg + PR_RO * 64 + j4);
      const float Rv[4] = {R4.x, R4.y, R4.z, R4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float y = __fmaf_rn(__int2float_rn(acc[j4 + e]), Sv[e], Bv[e]);
        t[e] = __fmaf_rn(gelu_erf(y), Rv[e], RMAGIC);      // g * 2^k is exact
      }
    } else if (ZP) {
      const float4 r4 = *reinterpret_cast<const float4*>(prmg + PR_ZRO * 64 + j4);
      const float4 z4 = *reinterpret_cast<const float4*>(prmg + PR_ZC * 64 + j4);
      const float4 o4 = *reinterpret_cast<const float4*>(prmg + PR_O * 64 + j4);
      const float ROv[4] = {r4.x, r4.y, r4.z, r4.w}, Ov[4] = {o4.x, o4.y, o4.z, o4.w};
      const int zc[4] = {__float_as_int(z4.x), __float_as_int(z4.y), __float_as_int(z4.z), __float_as_int(z4.w)};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float y = fadd(fmul(__int2float_rn(acc[j4 + e] - zc[e]), Sv[e]), Bv[e]);
        if (EPI == P2V_EPI_GELU) y = gelu_erf(y);
        t[e] = fadd(fadd(div_rb(y, Ov[e], ROv[e]), out_zp), RMAGIC);
      }
    } else {
      const float4 a4 = *reinterpret_cast<const float4*>(prmg + PR_OLO * 64 + j4);
      const float4 b4 = *reinterpret_cast<const float4*>(prmg + PR_OHI * 64 + j4);
      const float OLv[4] = {a4.x, a4.y, a4.z, a4.w}, OHv[4] = {b4.x, b4.y, b4.z, b4.w};
      float y[4];
      uint32_t flag = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        y[e] = fadd(fmul(__int2float_rn(acc[j4 + e]), Sv[e]), Bv[e]);
        if (EPI == P2V_EPI_GELU) y[e] = gelu_erf(y[e]);
        t[e] = quant_iv<false>(y[e], OLv[e], OHv[e], 1.f, flag);
      }
      if (flag) {
        const float4 g4 = *reinterpret_cast<const float4*>(prmg + PR_O * 64 + j4);
        const float Ov[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) t[e] = quant_iv<true>(y[e], OLv[e], OHv[e], Ov[e], flag);
      }
    }
    ow[j4 >> 2] = pack4_sat(t[0], t[1], t[2], t[3]);
  }
  if (EPI == P2V_EPI_GELU && GST == 1) {
    if (gst_near <= 16u) {     // some y within 8 ulps of the threshold it consulted: the chunk takes the direct evaluation
#pragma unroll 1
      for (int j4 = 0; j4 < 16; j4 += 4) {
        const float4 S4 = *reinterpret_cast<const float4*>(prmg + PR_S * 64 + j4);
        const float4 B4 = *reinterpret_cast<const float4*>(prmg + PR_B * 64 + j4);
        const float4 R4 = *reinterpret_cast<const float4*>(prmg + (POT ? PR_RO : PR_O) * 64 + j4);     // POT: 1 / out_scale, else out_scale
        const int a0 = j4 == 0 ? acc[0] : j4 == 4 ? acc[4] : j4 == 8 ? acc[8] : acc[12];
        const int a1 = j4 == 0 ? acc[1] : j4 == 4 ? acc[5] : j4 == 8 ? acc[9] : acc[13];
        const int a2 = j4 == 0 ? acc[2] : j4 == 4 ? acc[6] : j4 == 8 ? acc[10] : acc[14];
        const int a3 = j4 == 0 ? acc[3] : j4 == 4 ? acc[7] : j4 == 8 ? acc[11] : acc[15];
        int z0 = 0, z1 = 0, z2 = 0, z3 = 0;
        if (ZP) {
          const float4 z4 = *reinterpret_cast<const float4*>(prmg + PR_ZC * 64 + j4);
          z0 = __float_as_int(z4.x); z1 = __float_as_int(z4.y); z2 = __float_as_int(z4.z); z3 = __float_as_int(z4.w);
        }
        auto code = [&](int a, float S, float B, float R) {
          return POT ? gelu_code_direct(__fmaf_rn(__int2float_rn(a), S, B), R) : gelu_code_div(fadd(fmul(__int2float_rn(a), S), B), R, out_zp);
        };
        const uint32_t w = pack4_sat_int(code(a0 - z0, S4.x, B4.x, R4.x), code(a1 - z1, S4.y, B4.y, R4.y), code(a2 - z2, S4.z, B4.z, R4.z),
                                         code(a3 - z3, S4.w, B4.w, R4.w));
        if (j4 == 0) ow[0] = w; else if (j4 == 4) ow[1] = w; else if (j4 == 8) ow[2] = w; else ow[3] = w;
      }
    }
  }
  out = make_uint4(ow[0], ow[1], ow[2], ow[3]);
}

// per-column constants of columns n (lane's first) and n + 32 (second, if < W): raw values, fetched one tile ahead
struct RawCol { float s, b, o, m, rs, zc; };
template <int EPI, bool ZP>
__device__ __forceinline__ RawCol load_raw_col(const EpiParams& p, int n, bool ok) {
  RawCol r;
  ok = ok && n < p.N;
  r.zc = (ZP && ok && p.zp_corr) ? __int_as_float(__ldg(p.zp_corr + n)) : 0.f;
  r.s = ok ? __ldg(p.acc_scale + n) : 0.f;
  r.b = (ok && p.bias) ? __ldg(p.bias + n) : 0.f;
  r.o = ok ? __ldg(p.out_scale + n) : 1.f;
  r.m = (EPI == P2V_EPI_RESIDUAL && ok) ? __ldg(p.mid_scale + n) : 1.f;
  r.rs = (EPI == P2V_EPI_RESIDUAL && ok) ? __ldg(p.res_scale + n) : 0.f;
  return r;
}
template <int EPI, bool POT, bool ZP>
__device__ __forceinline__ void store_col(float* prm, int c, const RawCol& r) {
  const float ro = __frcp_rn(r.o);     // == fdiv(1, o): both correctly rounded
  if (ZP && EPI != P2V_EPI_RESIDUAL) {
    prm[PR_S * 64 + c] = r.s; prm[PR_B * 64 + c] = r.b;
    prm[PR_ZRO * 64 + c] = ro; prm[PR_ZC * 64 + c] = r.zc; prm[PR_O * 64 + c] = r.o;
    return;
  }
  if (ZP) prm[PR_RZC * 64 + c] = r.zc;
  if (EPI == P2V_EPI_RESIDUAL) {
    const float rm = __frcp_rn(r.m);
    prm[PR_S * 64 + c] = r.s; prm[PR_B * 64 + c] = r.b;
    prm[PR_MLO * 64 + c] = recip_lo(rm); prm[PR_MHI * 64 + c] = recip_hi(rm); prm[PR_M * 64 + c] = r.m;
    prm[PR_RS * 64 + c] = r.rs;
    prm[PR_ROLO * 64 + c] = recip_lo(ro); prm[PR_ROHI * 64 + c] = recip_hi(ro); prm[PR_RO_ * 64 + c] = r.o;
  } else if (POT && EPI == P2V_EPI_REQUANT) {
    prm[PR_S * 64 + c] = fmul(r.s, ro); prm[PR_B * 64 + c] = fmul(r.b, ro);   // exact: ro is a power of two
  } else if (POT) {
    prm[PR_S * 64 + c] = r.s; prm[PR_B * 64 + c] = r.b; prm[PR_RO * 64 + c] = ro;
  } else {
    prm[PR_S * 64 + c] = r.s; prm[PR_B * 64 + c] = r.b;
    prm[PR_OLO * 64 + c] = recip_lo(ro); prm[PR_OHI * 64 + c] = recip_hi(ro); prm[PR_O * 64 + c] = r.o;
  }
}

template <int EPI, bool POT, int GST, bool ZP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
                 const __grid_constant__ CUtensorMap tmR, EpiParams p, PairGeom g) {
  constexpr bool RESID = EPI == P2V_EPI_RESIDUAL;
  constexpr int ROWS = prm_rows<EPI, POT, ZP>();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * P_MAX_STAGES + 13];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bres_base = ring + g.off_bres, stg = ring + g.off_stg;
  float* prm_all = reinterpret_cast<float*>(smem_raw + (ring - smem_u32(smem_raw)) + g.off_prm);
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[P_MAX_STAGES]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * P_MAX_STAGES]), bar_tempty = bar_tfull + 16;
  const uint32_t bar_rfull = bar_tfull + 32, bar_sempty = bar_tfull + 64;     // up to 4 residual slots each
  const uint32_t bar_bfull = bar_tfull + 96;                                   // resident W tile loaded
  const uint32_t slot_bytes = uint32_t(P_EPI_WARPS) * 32u * uint32_t(g.W), warp_stg = 32u * uint32_t(g.W);
  const uint32_t slot_mask = (1u << g.nslot_log2) - 1u;
  const uint32_t nstages = uint32_t(g.nstages), bhalf_bytes = uint32_t(g.BN / 2) * PBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P_MAX_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 2 * P_EPI_WARPS);     // every epilogue warp of both CTAs
    }
    for (int a = 0; a < 4; ++a) {
      mbar_init(bar_rfull + 8 * a, 1);
      mbar_init(bar_sempty + 8 * a, P_EPI_WARPS);
    }
    mbar_init(bar_bfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair(smem_u32(&tmem_slot), 512);
  if (GST) gelu_steps_fill_smem(p.gelu_table, ring + g.off_gst, int(threadIdx.x), P_THREADS);   // replicated GELU step tables
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive / TMA signal
  tc_fence_after();
  // Programmatic dependent launch (common.cuh): the dependency wait sits in the roles that touch activations - the producer after it
  // has started the first resident W tile (a constant), the epilogue warps before their first tile; the MMA warp reads shared memory only.
  const uint32_t tmem_base = tmem_slot;

  // Register file: 16 K registers per SM sub-partition hold 1 control warp + 4 epilogue warps.  The kernel starts with 96 per
  // lane (640 threads); the control warps keep 32 and the epilogue warps grow to 112 out of the CTA's own pool.
  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    {     // the whole warp walks the loop; one elected lane issues (see elect_one)
      if (elect_one()) {
        tma_prefetch_map(&tmA);
        tma_prefetch_map(&tmB);
        if (RESID) tma_prefetch_map(&tmR);
      }
      const uint32_t full0 = mapa_u32(bar_full, 0);       // the leader's full barriers collect both CTAs' bytes
      const uint32_t bfull0 = mapa_u32(bar_bfull, 0);
      uint32_t itk = 0, it = 0;
      int cur_nt = -1;
      TileIter ti(g, pair, npairs);
      if (g.bres && ti.left > 0) {      // first W tile before the dependency wait: it overlaps the previous kernel's tail
        const int nb0 = ti.nt * g.BN + int(rank) * (g.BN / 2);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(bar_bfull, 2u * uint32_t(g.nkb) * bhalf_bytes);
          for (int kb = 0; kb < g.nkb; ++kb) tma_load_2d_pair(bres_base + kb * bhalf_bytes, &tmB, bfull0, kb * PBK, nb0);
        }
        __syncwarp();
        cur_nt = ti.nt;
      }
      pdl_wait();
      pdl_trigger();
      for (; ti.left > 0; ti.next(), ++it) {
        const int mt = ti.mt, nt = ti.nt;
        const int m0 = mt * 256 + int(rank) * PBM, nb0 = nt * g.BN + int(rank) * (g.BN / 2);
        PTRACE(2, it, 0);
        if (g.bres && nt != cur_nt) {
          // new column tile: every MMA that reads the old W tile has completed once the stage filled last is free again
          if (itk > 0) mbar_wait(bar_empty + 8 * ((itk - 1) % nstages), ((itk - 1) / nstages) & 1u);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(bar_bfull, 2u * uint32_t(g.nkb) * bhalf_bytes);
            for (int kb = 0; kb < g.nkb; ++kb) tma_load_2d_pair(bres_base + kb * bhalf_bytes, &tmB, bfull0, kb * PBK, nb0);
          }
          __syncwarp();
          cur_nt = nt;
        }
        for (int kb = 0; kb < g.nkb; ++kb, ++itk) {
          const uint32_t s = itk % nstages, ph = (itk / nstages) & 1u;
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          const uint32_t sa = ring + s * g.stage_bytes;
          if (elect_one()) {
            if (g.bres) {
              if (rank == 0) mbar_expect_tx(bar_full + 8 * s, 2u * P_A_BYTES);
              tma_load_2d_pair(sa, &tmA, full0 + 8 * s, kb * PBK, m0);
            } else {
              if (rank == 0) mbar_expect_tx(bar_full + 8 * s, 2u * (P_A_BYTES + bhalf_bytes));
              tma_load_2d_pair(sa, &tmA, full0 + 8 * s, kb * PBK, m0);
              tma_load_2d_pair(sa + P_A_BYTES, &tmB, full0 + 8 * s, kb * PBK, nb0);
            }
          }
          __syncwarp();
        }
        PTRACE(2, it, 1);
        if (RESID) {   // residual codes of this tile, straight into the staging slot the epilogue will overwrite
          const uint32_t slot = it & slot_mask, u = it >> g.nslot_log2;
          mbar_wait(bar_sempty + 8 * slot, (u & 1u) ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(bar_rfull + 8 * slot, slot_bytes);
            for (int e = 0; e < P_EPI_WARPS; ++e) {
              const int q = e & 3, cg = e >> 2;
              tma_load_2d(stg + slot * slot_bytes + e * warp_stg, &tmR, bar_rfull + 8 * slot, nt * g.BN + cg * g.W, m0 + q * 32);
            }
          }
          __syncwarp();
          PTRACE(2, it, 2);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (rank == 0) {      // warp-uniform loop, one elected lane issues the MMAs and their commits
      const uint32_t idesc = make_i8_idesc(256, g.BN, true, true);
      uint32_t itk = 0, it = 0, reloads = 0;
      int cur_nt = -1;
      for (TileIter ti(g, pair, npairs); ti.left > 0; ti.next(), ++it) {
        const uint32_t a = it & 1u, use = it >> 1;
        PTRACE(0, it, 0);
        mbar_wait_poll(bar_tempty + 8 * a, (use & 1u) ^ 1u);     // arrivals come from both CTAs: poll, never suspend
        if (g.bres && ti.nt != cur_nt) {
          mbar_wait_poll(bar_bfull, reloads & 1u);
          ++reloads;
          cur_nt = ti.nt;
        }
        tc_fence_after();
        PTRACE(0, it, 1);
        const uint32_t d_tmem = tmem_base + a * P_ACC_COLS;
        for (int kb = 0; kb < g.nkb; ++kb, ++itk) {
          const uint32_t s = itk % nstages, ph = (itk / nstages) & 1u;
          mbar_wait_poll(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t sa = ring + s * g.stage_bytes, sb = g.bres ? bres_base + kb * bhalf_bytes : sa + P_A_BYTES;
          const int ksteps = min(PBK / 32, (p.K - kb * PBK + 31) / 32);
          if (elect_one()) {
            for (int k = 0; k < ksteps; ++k)
              umma_i8_pair(d_tmem, make_kmajor_sw128_desc(sa + k * 32), make_kmajor_sw128_desc(sb + k * 32), idesc, uint32_t(kb > 0 || k > 0));
            tc_commit_pair(bar_empty + 8 * s, 3);    // both CTAs' producers may refill the stage
            if (kb == g.nkb - 1) tc_commit_pair(bar_tfull + 8 * a, 3);      // accumulator complete in both CTAs' TMEM
          }
          __syncwarp();
        }
        PTRACE(0, it, 2);
      }
    }
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
  } else {
    // ================= epilogue: 16 warps, warp = 32 rows x W columns =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int e = warp - 4;
    const uint32_t quarter = uint32_t(warp) & 3u;         // TMEM lane quarter this warp may read
    const int cg = e >> 2;
    const int W = g.W, nch = W >> 4;
    // one constant table per column group, shared by its four warps (one per TMEM lane quarter): warp-private tables cost 36 KB of
    // shared memory for the RESIDUAL epilogue - the difference between a 3-stage and a 5-stage operand ring next to a resident W tile
    float* prm = prm_all + cg * ROWS * 64;
    uint32_t prm32 = ring + g.off_prm + uint32_t(cg) * uint32_t(ROWS) * 256u;    // the same table as a shared-window address (prm_ld4)
    asm volatile("" : "+r"(prm32));
    const uint32_t tempty0 = mapa_u32(bar_tempty, 0);
    // byte offset of this lane's 16-byte chunk c inside the warp's staging block (row pitch W bytes)
    const uint32_t row_off = uint32_t(lane) * uint32_t(W);
    const uint32_t xr = g.swz == 2 ? (uint32_t(lane) >> 1) & 3u : (g.swz == 1 ? (uint32_t(lane) >> 2) & 1u : 0u);
    GeluSteps gst = {};
    if (GST) gst = gelu_steps_view(p.gelu_table, ring + g.off_gst, lane);
    RawCol nx0, nx1;
    TileIter ti(g, pair, npairs);
    {
      const int n0 = ti.nt * g.BN + cg * W;
      nx0 = load_raw_col<EPI, ZP>(p, n0 + lane, ti.left > 0 && quarter == 0);          // the quarter-0 warp fills the group's table
      nx1 = load_raw_col<EPI, ZP>(p, n0 + 32 + lane, ti.left > 0 && quarter == 0 && 32 + lane < W);
    }
    pdl_wait();                // everything above read constants; the residual codes and the output buffer come after the dependency
    uint32_t it = 0;
    int prm_nt = -1;           // column tile whose constants the warp's table holds
    int pending_slot = -1;     // RESID: staging slot whose store has been issued but not yet released to the producer
    for (; ti.left > 0; ++it) {
      const int mt = ti.mt, nt = ti.nt;
      ti.next();
      const int m0 = mt * 256 + int(rank) * PBM + int(quarter) * 32, n0 = nt * g.BN + cg * W;
      const uint32_t a = it & 1u, use = it >> 1;
      const uint32_t slot = RESID ? (it & slot_mask) : 0u, ruse = it >> g.nslot_log2;
      const uint32_t my_stg = stg + slot * slot_bytes + uint32_t(e) * warp_stg;
      // this tile's column constants (fetched during the previous tile) into the table, then fetch the next tile's
      if (nt != prm_nt) {       // the four warps of the group walk the same tile sequence: they all take this branch together
        named_barrier(1 + uint32_t(cg), 128);       // every warp of the group is done reading the previous column tile's constants
        if (quarter == 0) {
          store_col<EPI, POT, ZP>(prm, lane, nx0);
          if (32 + lane < W) store_col<EPI, POT, ZP>(prm, 32 + lane, nx1);
        }
        named_barrier(1 + uint32_t(cg), 128);       // ... and sees the new ones
        prm_nt = nt;
      }
      if (ti.left > 0 && ti.nt != nt) {
        const int nn0 = ti.nt * g.BN + cg * W;
        nx0 = load_raw_col<EPI, ZP>(p, nn0 + lane, quarter == 0);
        nx1 = load_raw_col<EPI, ZP>(p, nn0 + 32 + lane, quarter == 0 && 32 + lane < W);
      }
      if (lane == 0) PTRACE(3 + e, it, 0);
      if (RESID) mbar_wait(bar_rfull + 8 * slot, ruse & 1u);
      if (lane == 0) PTRACE(3 + e, it, 1);
      mbar_wait(bar_tfull + 8 * a, use & 1u);
      tc_fence_after();
      if (lane == 0) PTRACE(3 + e, it, 2);
      const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + a * P_ACC_COLS + uint32_t(cg * W);
      int accA[16], accB[16];
      tmem_ld16_async(taddr, accA);
      // one 16-column chunk: wait for its accumulators, start the next chunk's tcgen05.ld (or hand the TMEM stage back), compute, stage
      auto do_chunk = [&](int c, int (&cur)[16], int (&nxt)[16]) {
        __syncwarp();
        tmem_wait_ld16(cur);
        if (c + 1 < nch) {
          tmem_ld16_async(taddr + (c + 1) * 16, nxt);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty0 + 8 * a);
        }
        const uint32_t caddr = my_stg + row_off + ((uint32_t(c) ^ xr) << 4);
        uint4 resx = make_uint4(0, 0, 0, 0);
        if (RESID) {
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(resx.x), "=r"(resx.y), "=r"(resx.z), "=r"(resx.w) : "r"(caddr));
          resx.x ^= 0x80808080u; resx.y ^= 0x80808080u; resx.z ^= 0x80808080u; resx.w ^= 0x80808080u;
        }
        uint4 o;
        pair_chunk<EPI, POT, GST, ZP>(prm + c * 16, prm32 + uint32_t(c) * 64u, cur, resx, o, gst, p.out_zp);
        if (c == 0) {
          if (!RESID) {                     // the previous tile's store must have finished reading this block
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
          } else if (pending_slot >= 0) {   // release the previous tile's slot to the producer (its store was issued a chunk ago)
            if (lane == 0) { bulk_wait_read0(); mbar_arrive(bar_sempty + 8 * uint32_t(pending_slot)); }
            pending_slot = -1;
          }
        }
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(caddr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
      };
#pragma unroll 1
      for (int c = 0; c < nch; c += 2) {
        do_chunk(c, accA, accB);
        if (c + 1 < nch) do_chunk(c + 1, accB, accA);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) {        // always lane 0 of the converged warp: the bulk group and its waits belong to that lane
        tma_store_2d(&tmO, my_stg, n0, m0);
        bulk_commit();
        PTRACE(3 + e, it, 3);
      }
      __syncwarp();
      if (RESID) pending_slot = int(slot);
    }
    if (lane == 0) bulk_wait_all();    // no arrive on bar_sempty: the producer has no further tile to load
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer may still be signalling this CTA's barriers / reading its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host
static int make_tmap_rows(CUtensorMap* m, const void* ptr, int rows, int cols, int box_cols, int box_rows, CUtensorMapSwizzle swz) {
  encode_tiled_fn enc = get_tensor_map_encoder();
  P2V_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(cols)};
  cuuint32_t box[2] = {cuuint32_t(box_cols), cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  P2V_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d box=%dx%d", int(r), rows, cols, box_rows, box_cols);
  return 0;
}

bool gemm_pair_supported(const p2v_gemm_args& a) {
  const int e = a.epilogue;
  if (e != P2V_EPI_REQUANT && e != P2V_EPI_GELU && e != P2V_EPI_RESIDUAL) return false;
  if (a.row_map || !a.out_i8) return false;
  if (a.mid_zp != 0.f || a.aux_zp != 0.f) return false;                          // EMBED's zero points: csrc/gemm_tc.cu
  if ((a.zp_corr || a.out_zp != 0.f) && a.pot_scales) return false;              // zero points come with raw fp32 scales (ZP variants)
  if (a.out_zp != 0.f && e == P2V_EPI_RESIDUAL) return false;
  if (a.N % 16 || a.K % 16) return false;
  if (reinterpret_cast<uintptr_t>(a.out_i8) & 15) return false;
  if (e == P2V_EPI_RESIDUAL && (reinterpret_cast<uintptr_t>(a.res) & 15)) return false;
  return true;
}

static int max_pairs() {
  static int cached = 0;
  if (!cached) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cached = std::max(1, sms / 2);
  }
  return cached;
}

constexpr size_t P_SMEM_BUDGET = 227 * 1024 - 1024;    // dynamic shared memory per CTA, static part and slack taken off

// Shared-memory plan of one candidate geometry: [operand ring][resident W half-tile][output / residual staging][16 warp-private
// constant tables (+ GELU step tables)].  rows = prm_rows<EPI, POT>().  false = no room for a two-stage ring.
static bool plan_pair(const p2v_gemm_args& a, int bn, int rows, bool gst, PairGeom& g) {
  static const int force_bres = getenv("P2V_PAIR_BRES") ? atoi(getenv("P2V_PAIR_BRES")) : -1;    // perf triage only
  g.BN = bn;
  g.W = bn / 4;
  g.tiles_m = (a.M + 255) / 256;
  g.tiles_n = (a.N + bn - 1) / bn;
  g.tiles = g.tiles_m * g.tiles_n;
  g.nkb = (a.K + PBK - 1) / PBK;
  g.swz = g.W == 64 ? 2 : (g.W == 32 ? 1 : 0);
  g.nslot_log2 = a.epilogue == P2V_EPI_RESIDUAL ? (g.W == 32 ? 2 : 1) : 0;
  const size_t prm_bytes = size_t(P_EPI_GROUPS) * rows * 64 * 4 + (gst ? P2V_GELU_STEPS_SMEM_MAX : 0);
  const size_t bhalf = size_t(g.BN / 2) * PBK;
  for (;;) {
    const size_t stg_bytes = (size_t(1) << g.nslot_log2) * P_EPI_WARPS * 32 * g.W;
    const size_t fixed = 1024 + prm_bytes + stg_bytes;
    const size_t bres_bytes = size_t(g.nkb) * bhalf;
    const bool fits = fixed + bres_bytes + 3 * P_A_BYTES <= P_SMEM_BUDGET;
    if (!fits && a.epilogue == P2V_EPI_RESIDUAL && g.nslot_log2 == 2 && fixed - stg_bytes / 2 + bres_bytes + 3 * P_A_BYTES <= P_SMEM_BUDGET) {
      g.nslot_log2 = 1;        // a resident W tile is worth more than the two extra residual slots
      continue;
    }
    g.bres = (fits && force_bres != 0) ? 1 : 0;
    g.stage_bytes = uint32_t(P_A_BYTES + (g.bres ? 0 : P_B_BYTES));
    if (fixed + (g.bres ? bres_bytes : 0) + 2 * g.stage_bytes > P_SMEM_BUDGET) return false;
    const size_t ring_room = P_SMEM_BUDGET - fixed - (g.bres ? bres_bytes : 0);
    g.nstages = int(std::min<size_t>(P_MAX_STAGES, ring_room / g.stage_bytes));
    g.off_bres = uint32_t(g.nstages) * g.stage_bytes;
    g.off_stg = g.off_bres + uint32_t(g.bres ? bres_bytes : 0);
    g.off_prm = g.off_stg + uint32_t(stg_bytes);
    g.off_gst = g.off_prm + uint32_t(P_EPI_GROUPS) * rows * 64 * 4;
    g.smem_bytes = uint32_t(1024 + g.off_prm + prm_bytes);
    return true;
  }
}

// Estimated cycles of the slowest CTA pair for a planned geometry (r1m measurements of 12 shapes x 3 tile widths, tools/gpu_exp.sh;
// DESIGN.md 3.1): per tile max(operand load, MMA, epilogue) + ~1500 cycles of hand-offs, times the tile waves over the pairs.
//   operand load: bytes per CTA at <= 28 B/clk (what L2 delivers per SM with every SM streaming) and at most the ring's bytes in
//                 flight per loaded round trip (2200 cycles while A fits L2, 2800 from DRAM): a resident W tile that leaves a
//                 3-stage ring is latency-bound (DeiT-S fc2 at BN = 128);
//   MMA:          2.46 cycles per column and k-block (256 x BN x 128 int8 MACs at 6.6 k MAC/clk/SM);
//   epilogue:     all BN columns of a tile cost (the warps of padded columns idle, the others are not faster), cycles per column
//                 by epilogue kind;
//   a narrower tile has to win by 8 % per step (per-tile costs the model does not see).
// With these constants the model picks the measured-fastest width on all 12 shapes.  The old rule (least waves x BN) ignored the
// fixed cost per tile and took 128-column tiles for short row counts: ViT-B qkv at batch 128 65 -> 42 us, ViT-L fc1 / fc2
// 163 -> 114 / 144 -> 97 us, DeiT-S fc2 55 -> 45 us.
static double pair_cost(const p2v_gemm_args& a, const PairGeom& g, bool gst) {
  const double waves = double((g.tiles + max_pairs() - 1) / max_pairs());
  const double bytes = (128.0 + (g.bres ? 0.0 : g.BN / 2.0)) * a.K;
  const double round_trip = double(a.M) * a.K <= 48e6 ? 2200.0 : 2800.0;
  const double rate = std::min(28.0, double(g.nstages) * g.stage_bytes / round_trip);
  const double load = bytes / rate;
  const double mma = double(g.nkb) * g.BN * 2.46;
  const bool pot = a.pot_scales != 0;
  const double per_col = a.epilogue == P2V_EPI_RESIDUAL ? 33.0 : a.epilogue == P2V_EPI_GELU ? (gst ? 31.0 : 50.0) : (pot ? 15.0 : 22.0);
  const double handicap = 1.0 + 0.08 * (g.BN == 256 ? 0 : g.BN == 192 ? 1 : 2);
  return waves * (std::max(load, std::max(mma, g.BN * per_col)) + 1500.0) * handicap;
}

template <int EPI, bool POT, int GST = 0, bool ZP = false>
static int launch_pair(const p2v_gemm_args& a, const PairGeom& g, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                       const CUtensorMap& tmR, cudaStream_t stream) {
  auto kern = gemm_pair_kernel<EPI, POT, GST, ZP>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(P_SMEM_BUDGET));
    P2V_REQUIRE(e == cudaSuccess, "gemm_pair: cannot set %zu bytes of dynamic shared memory: %s", P_SMEM_BUDGET, cudaGetErrorString(e));
    attr = true;
  }
  const int grid = 2 * std::min(g.tiles, max_pairs());
  EpiParams p = make_epi_params(a);
  pdl_next_kind(PDL_GEMM);
  launch_pdl(kern, dim3(grid), dim3(P_THREADS), g.smem_bytes, stream, tmA, tmB, tmO, tmR, p, g);
  count_launch();
  return check_launch("gemm_pair");
}

int launch_gemm_pair(const p2v_gemm_args& a, cudaStream_t stream) {
  P2V_REQUIRE(gemm_pair_supported(a), "gemm_pair: unsupported arguments");
  const bool pot = a.pot_scales != 0;
  const bool gst = a.epilogue == P2V_EPI_GELU && a.gelu_table;        // the table was built for this out_scale / out_zp (caller's contract)
  const bool zpv = !pot && (a.zp_corr != nullptr || a.out_zp != 0.f);
  const int rows = a.epilogue == P2V_EPI_RESIDUAL ? (zpv ? prm_rows<P2V_EPI_RESIDUAL, false, true>() : prm_rows<P2V_EPI_RESIDUAL, true>())
                   : a.epilogue == P2V_EPI_GELU   ? (pot ? prm_rows<P2V_EPI_GELU, true>() : prm_rows<P2V_EPI_GELU, false>())
                                                  : (pot ? prm_rows<P2V_EPI_REQUANT, true>() : prm_rows<P2V_EPI_REQUANT, false>());
  static const int force_bn = getenv("P2V_PAIR_BN") ? atoi(getenv("P2V_PAIR_BN")) : 0;    // perf triage only
  PairGeom g;
  double best = -1.0;
  for (int bn : {256, 192, 128}) {
    if (force_bn && bn != force_bn) continue;
    PairGeom c;
    if (!plan_pair(a, bn, rows, gst, c)) continue;
    const double cost = pair_cost(a, c, gst);
    if (best < 0.0 || cost < best) { best = cost; g = c; }
  }
  P2V_REQUIRE(best >= 0.0, "gemm_pair: no room for the operand ring (K=%d N=%d)", a.K, a.N);
  const CUtensorMapSwizzle oswz = g.W == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : (g.W == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE);
  CUtensorMap tmA, tmB, tmO, tmR;
  if (int r = make_tmap_rows(&tmA, a.A, a.M, a.K, PBK, PBM, CU_TENSOR_MAP_SWIZZLE_128B)) return r;
  if (int r = make_tmap_rows(&tmB, a.W, a.N, a.K, PBK, g.BN / 2, CU_TENSOR_MAP_SWIZZLE_128B)) return r;
  if (int r = make_tmap_rows(&tmO, a.out_i8, a.M, a.N, g.W, 32, oswz)) return r;
  tmR = tmO;
  if (a.epilogue == P2V_EPI_RESIDUAL)
    if (int r = make_tmap_rows(&tmR, a.res, a.M, a.N, g.W, 32, oswz)) return r;
  if (zpv) {
    switch (a.epilogue) {
      case P2V_EPI_REQUANT: return launch_pair<P2V_EPI_REQUANT, false, 0, true>(a, g, tmA, tmB, tmO, tmR, stream);
      case P2V_EPI_GELU:
        if (gst) return gelu_table_is_clean(a.gelu_table) ? launch_pair<P2V_EPI_GELU, false, 2, true>(a, g, tmA, tmB, tmO, tmR, stream)
                                                          : launch_pair<P2V_EPI_GELU, false, 1, true>(a, g, tmA, tmB, tmO, tmR, stream);
        return launch_pair<P2V_EPI_GELU, false, 0, true>(a, g, tmA, tmB, tmO, tmR, stream);
      default: return launch_pair<P2V_EPI_RESIDUAL, false, 0, true>(a, g, tmA, tmB, tmO, tmR, stream);
    }
  }
  switch (a.epilogue) {
    case P2V_EPI_REQUANT:
      return pot ? launch_pair<P2V_EPI_REQUANT, true>(a, g, tmA, tmB, tmO, tmR, stream)
                 : launch_pair<P2V_EPI_REQUANT, false>(a, g, tmA, tmB, tmO, tmR, stream);
    case P2V_EPI_GELU:
      if (gst && pot) return gelu_table_is_clean(a.gelu_table) ? launch_pair<P2V_EPI_GELU, true, 2>(a, g, tmA, tmB, tmO, tmR, stream)
                                                               : launch_pair<P2V_EPI_GELU, true, 1>(a, g, tmA, tmB, tmO, tmR, stream);
      if (gst) return gelu_table_is_clean(a.gelu_table) ? launch_pair<P2V_EPI_GELU, false, 2>(a, g, tmA, tmB, tmO, tmR, stream)
                                                        : launch_pair<P2V_EPI_GELU, false, 1>(a, g, tmA, tmB, tmO, tmR, stream);
      return pot ? launch_pair<P2V_EPI_GELU, true>(a, g, tmA, tmB, tmO, tmR, stream)
                 : launch_pair<P2V_EPI_GELU, false>(a, g, tmA, tmB, tmO, tmR, stream);
    default:
      return pot ? launch_pair<P2V_EPI_RESIDUAL, true>(a, g, tmA, tmB, tmO, tmR, stream)
                 : launch_pair<P2V_EPI_RESIDUAL, false>(a, g, tmA, tmB, tmO, tmR, stream);
  }
}

}  // namespace p2v
