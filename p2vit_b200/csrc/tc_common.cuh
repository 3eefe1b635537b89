// sm_100a building blocks shared by the tensor-core kernels (gemm_tc.cu, attention_tc.cu): mbarrier, TMA,
// tcgen05 (alloc / mma kind::i8 / commit / ld) wrappers, shared-memory and instruction descriptors, and the
// host-side tensor-map encoder.  Bit layouts follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor,
// InstrDescriptor) of the CUTLASS headers shipped with the image; nothing of CUTLASS is linked.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace p2v {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

// One leader lane of a converged warp (the lowest, deterministically for the full mask).  Control code runs warp-uniform and
// wraps only the TMA / tcgen05 / expect-tx instructions in `if (elect_one())`: their operands then stay in uniform registers.
// Under `if (lane == 0)` instead the compiler has to assume divergence and feeds every UTCIMMA / UTMALDG through an
// ELECT + R2UR.BROADCAST waterfall loop - about 300 cycles per MMA instruction issued, measured (tools/pair_trace.py).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// With the suspend-time hint ptxas emits TRYWAIT + NANOSLEEP.SYNCS: the warp sleeps until an mbarrier event (or the hint
// elapses) instead of spinning - spinning waiters took 18 % of the issue slots of the attention kernel (ncu, r1i).
constexpr uint32_t MBAR_SUSPEND_NS = 1000;
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_NS) : "memory");
  return ok != 0;
}
// try_wait suspends in hardware for a bounded time per call; a wait that lasts longer than 2 s of wall clock is a protocol
// bug and becomes a trap (reported as a launch error by the next API call) instead of a hung GPU.  No nanosleep in the loop:
// its granularity (~1 us) would add a microsecond to every wait that is not satisfied on the first try.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0u) {
      uint64_t t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t0 == 0) t0 = t1;
      if (t1 - t0 > 2000000000ull) __trap();
    }
  }
}
// Polling wait (test_wait never suspends).  For barriers whose last arrival can come from the peer CTA of a cluster
// (mbarrier.arrive.shared::cluster / a peer's TMA complete_tx): a thread suspended inside try_wait was observed to sleep
// on for thousands of cycles after such a remote arrival completed the phase (tools/pair_trace.py, proj GEMM).
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_poll(uint32_t bar, uint32_t parity) {
  if (mbar_test_wait(bar, parity)) return;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (!mbar_test_wait(bar, parity)) {
    if ((++spins & 4095u) == 0u) {
      uint64_t t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t0 == 0) t0 = t1;
      if (t1 - t0 > 2000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, 8-bit integer operands (signedness in idesc), int32 accumulate
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// asynchronous: the registers are valid after tmem_wait_ld()
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, int (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&r)[32]) {
  tmem_ld32_async(taddr, r);
  tmem_wait_ld();
}
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, int (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&r)[16]) {
  tmem_ld16_async(taddr, r);
  tmem_wait_ld();
}
__device__ __forceinline__ void named_barrier(uint32_t id, uint32_t threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Shared-memory matrix descriptor: start address>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) |
// layout type [61,64).  Canonical layouts (units of 16 bytes, cute/atom/mma_traits_sm100.hpp):
//   K-major  SW128: rows of 128 B, 8-row groups SBO = 1024 B apart           (LBO unused)
//   K-major  SW64 : rows of  64 B, 8-row groups SBO =  512 B apart           (LBO unused)
//   MN-major SW64 : 64 B of MN contiguous per K row, 8 K-rows per 512-B atom, atoms along K SBO apart,
//                   64-B column blocks along MN LBO apart
constexpr uint32_t UMMA_LAYOUT_SW128 = 2, UMMA_LAYOUT_SW64 = 4, UMMA_LAYOUT_SW32 = 6;
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3FFFu);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(layout) << 61;
  return d;
}
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) { return make_smem_desc(smem_addr, 16, 1024, UMMA_LAYOUT_SW128); }

// Instruction descriptor for kind::i8: c_format=S32(2) [4,6); a/b format [7,10)/[10,13): 1 = signed 8 bit, 0 = unsigned;
// a/b major bits 15/16 (0 = K-major, 1 = MN-major); N>>3 [17,23); M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_i8_idesc(int M, int N, bool a_signed, bool b_signed, bool a_mn_major = false,
                                                     bool b_mn_major = false) {
  return (2u << 4) | (uint32_t(a_signed) << 7) | (uint32_t(b_signed) << 10) | (uint32_t(a_mn_major) << 15) |
         (uint32_t(b_mn_major) << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn get_tensor_map_encoder();  // gemm_tc.cu; nullptr if the driver does not export cuTensorMapEncodeTiled

}  // namespace p2v
