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
#include <cuda.h>
#include <unordered_map>
#include <mutex>
#include "epilogue.cuh"

namespace p2v {

constexpr int BM = 128;
constexpr int BK = 128;           // one 128-byte swizzle row of int8
constexpr int UMMA_K = 32;        // K per tcgen05.mma for 8-bit operands
constexpr int TC_GROUPS = 4;       // epilogue groups == TMEM accumulator stages
constexpr int TC_THREADS = 64 + TC_GROUPS * 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// try_wait suspends in hardware for a bounded time per call; the iteration cap turns a protocol bug into a
// trap (reported as a launch error by the next API call) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 operands, int32 accumulate
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void group_barrier(uint32_t id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor):
// start address>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major, 1) | SBO>>4 [32,46) = 1024 B between
// 8-row groups | version=1 [46,48) | layout type [61,64) = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3FFFu);
  d |= uint64_t(1) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

// cute/arch/mma_sm100_desc.hpp: InstrDescriptor.  c_format=S32(2) [4,6); a/b format [7,10)/[10,13): 1 = signed 8 bit,
// 0 = unsigned; a/b major bits 15/16 = 0 (K-major); N>>3 [17,23); M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_i8_idesc(int M, int N, bool a_signed, bool b_signed) {
  return (2u << 4) | (uint32_t(a_signed) << 7) | (uint32_t(b_signed) << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

template <int BN, int STAGES, int EPI, bool POT>
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
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      uint32_t it = 0;
      for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int m0 = int(t / tiles_n) * BM, n0 = int(t % tiles_n) * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          mbar_expect_tx(bar_full + 8 * s, STAGE_BYTES);
          const uint32_t sa = stage0 + s * STAGE_BYTES;
          tma_load_2d(sa, &tmA, bar_full + 8 * s, kb * BK, m0);
          tma_load_2d(sa + A_BYTES, &tmB, bar_full + 8 * s, kb * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
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
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_i8(d_tmem, make_kmajor_sw128_desc(sa + k * UMMA_K), make_kmajor_sw128_desc(sb + k * UMMA_K), idesc,
                    uint32_t(kb > 0 || k > 0));
          tc_commit(bar_empty + 8 * s);  // frees the smem stage when these MMAs have read it
        }
        tc_commit(bar_tfull + 8 * g);    // accumulator complete
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
      mbar_wait(bar_tfull + 8 * g, use & 1u);
      tc_fence_after();
      const int row = m0 + int(quarter) * 32 + lane;
      const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + g * BN;
#pragma unroll 1
      for (int c = 0; c < BN / NCH; ++c) {
        if (n0 + c * NCH >= p.N) break;
        int acc[NCH];
        __syncwarp();
        if (NCH == 32) tmem_ld32(taddr + c * NCH, reinterpret_cast<int(&)[32]>(acc[0]));
        else tmem_ld16(taddr + c * NCH, reinterpret_cast<int(&)[16]>(acc[0]));
        epilogue_row<EPI, POT, BN, NCH>(p, cp, row, n0, c * NCH, acc);
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
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
  encode_tiled_fn enc = get_encode();
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

template <int BN, int STAGES>
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
  constexpr size_t smem = size_t(STAGES) * (BM * BK + BN * BK) + 1024;
  P2V_DISPATCH_EPI(a.epilogue, a.pot_scales != 0, {
    auto kern = gemm_tc_kernel<BN, STAGES, EPI, POT>;
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      P2V_REQUIRE(e == cudaSuccess, "gemm_tc: cannot set %zu bytes of dynamic shared memory: %s", smem, cudaGetErrorString(e));
      attr = true;
    }
    kern<<<grid, TC_THREADS, smem, stream>>>(tmA, tmB, p, tiles_m, tiles_n);
  });
  count_launch();
  return check_launch("gemm_tc");
}

int launch_gemm_tc(const p2v_gemm_args& a, cudaStream_t stream) {
  // BN = 128: 32 KB / smem stage, 5 stages, 4 x 128 TMEM columns
  return launch_tc_bn<128, 5>(a, stream);
}

}  // namespace p2v
