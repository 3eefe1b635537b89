// int8 x int8 -> int32 GEMM on the 5th-generation tensor cores (sm_100a), fused QAct epilogues.
//
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared memory ring -> tcgen05.mma.cta_group::1.kind::i8
//   (A, W K-major, one elected thread) -> int32 accumulators in TMEM (2 stages) -> tcgen05.ld -> epilogue
//   in registers (epilogue.cuh) -> int8 / fp32 stores.
//
// Persistent, warp-specialised CTA of 18 warps (1 CTA per SM):
//   warp 0        TMA producer                (lane 0)
//   warp 1        TMEM allocator + MMA issuer (lane 0)
//   warps 2-17    four epilogue groups of 4 warps; group g owns TMEM accumulator stage g and this CTA's tiles
//                 g, g+4, g+8, ...  The epilogue is the ALU-bound part of these GEMMs (K is only 384..1536, every
//                 output element needs a requantisation), so 16 of the 18 warps work on it and the MMA of the next
//                 three tiles runs underneath.
// A warp may only read the TMEM lane quarter (warp_id % 4), so each group of 4 consecutive warps covers
// the 128 accumulator rows.  Per-column epilogue parameters are staged per tile in shared memory (double buffered
// per group, one named barrier per tile).  Tiles are enumerated n-fastest so the CTAs running at the same time share
// the A row-block through L2 and the (small) weight matrix stays L2 resident.
#include "tc_common.cuh"
#include "epilogue.cuh"

namespace p2v {

constexpr int BM = 128;
constexpr int BK = 128;           // one 128-byte swizzle row of int8
constexpr int UMMA_K = 32;        // K per tcgen05.mma for 8-bit operands
constexpr int TC_GROUPS = 4;       // epilogue groups == TMEM accumulator stages
constexpr int TC_THREADS = 64 + TC_GROUPS * 128;

__device__ __forceinline__ void group_barrier(uint32_t id) { named_barrier(id, 128); }

template <int BN, int STAGES, int EPI, bool POT, bool GTAB>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, EpiParams p, int tiles_m, int tiles_n) {
  constexpr uint32_t A_BYTES = BM * BK, B_BYTES = BN * BK, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = TC_GROUPS * BN;  // one accumulator stage per epilogue group (power of two)
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns");
  constexpr int NCH = (EPI == P2V_EPI_REQUANT || EPI == P2V_EPI_DEQUANT || EPI == P2V_EPI_F32) ? 32 : 16;  // columns per TMEM load
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 2 * TC_GROUPS];
  __shared__ __align__(16) float col_params[TC_GROUPS][2][CP_ROWS * BN];
  __shared__ uint32_t tmem_slot;

  const uint32_t tiles = uint32_t(tiles_m) * uint32_t(tiles_n);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t stage0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[STAGES]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * STAGES]), bar_tempty = smem_u32(&bars[2 * STAGES + TC_GROUPS]);
  const int nkb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int g = 0; g < TC_GROUPS; ++g) { mbar_init(bar_tfull + 8 * g, 1); mbar_init(bar_tempty + 8 * g, 4); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<TMEM_COLS>(smem_u32(&tmem_slot));
  }
  GeluTab gt{nullptr, 0, 0.f, 0.f};
  if (GTAB) {   // the step table of qact1(gelu(.)) moves to shared memory behind the operand ring
    const GeluTabHeader hd = *reinterpret_cast<const GeluTabHeader*>(p.gelu_table);
    uint2* s_tab = reinterpret_cast<uint2*>(smem_raw + (stage0 - smem_u32(smem_raw)) + STAGES * STAGE_BYTES);
    const uint2* src = reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(p.gelu_table) + sizeof(GeluTabHeader));
    for (int i = threadIdx.x; i < hd.n; i += TC_THREADS) s_tab[i] = __ldg(src + i);
    gt = GeluTab{s_tab, hd.n, hd.inv_w, -hd.y0 * hd.inv_w};
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: warp-uniform loop, one elected lane issues (tc_common.cuh: elect_one) =================
    if (elect_one()) {
      tma_prefetch_map(&tmA);
      tma_prefetch_map(&tmB);
    }
    uint32_t it = 0;
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int m0 = int(t / tiles_n) * BM, n0 = int(t % tiles_n) * BN;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(bar_full + 8 * s, STAGE_BYTES);
          const uint32_t sa = stage0 + s * STAGE_BYTES;
          tma_load_2d(sa, &tmA, bar_full + 8 * s, kb * BK, m0);
          tma_load_2d(sa + A_BYTES, &tmB, bar_full + 8 * s, kb * BK, n0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_i8_idesc(BM, BN, true, true);
    uint32_t it = 0, local = 0;
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x, ++local) {
      const uint32_t g = local % TC_GROUPS, use = local / TC_GROUPS;
      mbar_wait(bar_tempty + 8 * g, (use & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + g * BN;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
        mbar_wait(bar_full + 8 * s, ph);
        tc_fence_after();
        const uint32_t sa = stage0 + s * STAGE_BYTES, sb = sa + A_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_i8(d_tmem, make_kmajor_sw128_desc(sa + k * UMMA_K), make_kmajor_sw128_desc(sb + k * UMMA_K), idesc,
                    uint32_t(kb > 0 || k > 0));
          tc_commit(bar_empty + 8 * s);  // frees the smem stage when these MMAs have read it
          if (kb == nkb - 1) tc_commit(bar_tfull + 8 * g);    // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue groups =================
    const uint32_t g = uint32_t(warp - 2) >> 2;          // group == accumulator stage
    const uint32_t quarter = uint32_t(warp) & 3u;        // TMEM lane quarter this warp may access
    const int tg = int(threadIdx.x) - 64 - int(g) * 128;  // thread index inside the group
    uint32_t use = 0;
    for (uint32_t t = blockIdx.x + g * gridDim.x; t < tiles; t += TC_GROUPS * gridDim.x, ++use) {
      const int m0 = int(t / tiles_n) * BM, n0 = int(t % tiles_n) * BN;
      float* cp = col_params[g][use & 1u];
      stage_col_params<EPI, POT, BN>(p, cp, n0, tg);     // overlaps the MMA of this tile
      group_barrier(1 + g);
      const int row = m0 + int(quarter) * 32 + lane;
      const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + g * BN;
      uint32_t res_next[NCH / 4];
      if (EPI == P2V_EPI_RESIDUAL) load_residual<NCH>(p, row, n0, res_next);   // in flight while the MMA finishes
      mbar_wait(bar_tfull + 8 * g, use & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / NCH; ++c) {
        if (n0 + c * NCH >= p.N) break;
        int acc[NCH];
        uint32_t res_cur[NCH / 4];
#pragma unroll
        for (int j = 0; j < NCH / 4; ++j) res_cur[j] = res_next[j];
        __syncwarp();
        if (NCH == 32) tmem_ld32_async(taddr + c * NCH, reinterpret_cast<int(&)[32]>(acc[0]));
        else tmem_ld16_async(taddr + c * NCH, reinterpret_cast<int(&)[16]>(acc[0]));
        if (EPI == P2V_EPI_RESIDUAL && c + 1 < BN / NCH) load_residual<NCH>(p, row, n0 + (c + 1) * NCH, res_next);
        epilogue_row<EPI, POT, BN, NCH, true>(p, cp, row, n0, c * NCH, acc, res_cur, gt);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * g);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ host
encode_tiled_fn get_tensor_map_encoder() {
  static encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(f);
  }
  return fn;
}

// 2-D int8 row-major [rows, cols] tensor, box = [box_rows, 128 bytes], 128-byte swizzle, zero fill out of bounds
int make_tmap_i8(CUtensorMap* m, const void* ptr, int rows, int cols, int box_rows) {
  encode_tiled_fn enc = get_tensor_map_encoder();
  P2V_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(cols)};
  cuuint32_t box[2] = {cuuint32_t(BK), cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  P2V_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d box_rows=%d", int(r), rows, cols, box_rows);
  return 0;
}

template <int BN, int STAGES, bool GTAB>
static int launch_tc_bn(const p2v_gemm_args& a, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  if (int r = make_tmap_i8(&tmA, a.A, a.M, a.K, BM)) return r;
  if (int r = make_tmap_i8(&tmB, a.W, a.N, a.K, BN)) return r;
  EpiParams p = make_epi_params(a);
  const int tiles_m = (a.M + BM - 1) / BM, tiles_n = (a.N + BN - 1) / BN;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(tiles_m * tiles_n, sms);
  constexpr size_t smem = size_t(STAGES) * (BM * BK + BN * BK) + 1024 + (GTAB ? 8 * P2V_GELU_TABLE_MAX_ENTRIES : 0);
#define P2V_TC_LAUNCH(EPI_, POT_)                                                                                          \
  {                                                                                                                        \
    auto kern = gemm_tc_kernel<BN, STAGES, EPI_, POT_, GTAB>;                                                              \
    static bool attr = false;                                                                                              \
    if (!attr) {                                                                                                           \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));                  \
      P2V_REQUIRE(e == cudaSuccess, "gemm_tc: cannot set %zu bytes of dynamic shared memory: %s", smem, cudaGetErrorString(e)); \
      attr = true;                                                                                                         \
    }                                                                                                                      \
    launch_pdl(kern, dim3(grid), dim3(TC_THREADS), smem, stream, tmA, tmB, p, tiles_m, tiles_n);                                             \
  }
  if (GTAB) {
    P2V_TC_LAUNCH(P2V_EPI_GELU, true)
  } else {
    P2V_DISPATCH_EPI(a.epilogue, a.pot_scales != 0, P2V_TC_LAUNCH(EPI, POT));
  }
#undef P2V_TC_LAUNCH
  count_launch();
  return check_launch("gemm_tc");
}

bool gemm_pair_supported(const p2v_gemm_args& a);
int launch_gemm_pair(const p2v_gemm_args& a, cudaStream_t stream);
static int g_gemm_variant = 0;   // 0 auto, 1 one-tile-per-group kernel (this file), 2 CTA-pair kernel (gemm_pair.cu) whenever it applies
void set_gemm_variant(int v) { g_gemm_variant = v; }

int launch_gemm_tc(const p2v_gemm_args& a, cudaStream_t stream) {
  if (g_gemm_variant != 1 && gemm_pair_supported(a) && (g_gemm_variant == 2 || a.M >= 256)) return launch_gemm_pair(a, stream);
  // BN = 128: 32 KB / smem stage, 4 x 128 TMEM columns; 5 stages, or 4 stages + the 32 KB GELU step table
  if (a.epilogue == P2V_EPI_GELU && a.pot_scales && a.gelu_table) return launch_tc_bn<128, 4, true>(a, stream);
  return launch_tc_bn<128, 5, false>(a, stream);
}

}  // namespace p2v
