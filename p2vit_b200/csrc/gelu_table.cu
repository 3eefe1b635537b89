// Builds the step table of  y -> sat(RNE(gelu_erf(y) / out_scale))  (include/p2vit_b200.h: p2v_build_gelu_table).
// One thread per segment: the segment's first / last fp32 values are found by bisection on gelu_segment() itself, the
// code change inside it (at most one for a segment of width out_scale/2) by bisection on gelu_code_direct(), and the
// neighbourhood of the threshold is probed so that a segment whose code is not a clean step is flagged `slow`
// (the epilogue then evaluates erf directly for every y that lands in it).
#include <cmath>
#include "common.cuh"

namespace p2v {

// order-preserving map fp32 <-> uint32 (negative floats reversed)
__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

__global__ void __launch_bounds__(128) build_gelu_table_kernel(GeluTabHeader hd, float ro, uint2* __restrict__ entries) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hd.n) return;
  const float off = -hd.y0 * hd.inv_w;
  const uint32_t kmin = f2key(-3.0e38f), kmax = f2key(3.0e38f);
  // first key whose segment is >= s (segments are monotone in y)
  auto first_key = [&](int s) {
    uint32_t lo = kmin, hi = kmax;
    if (gelu_segment(key2f(kmin), hd.inv_w, off, hd.n) >= s) return kmin;
    while (hi - lo > 1u) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (gelu_segment(key2f(mid), hd.inv_w, off, hd.n) >= s) hi = mid; else lo = mid;
    }
    return hi;
  };
  const uint32_t klo = first_key(i);
  const uint32_t khi = (i + 1 < hd.n) ? first_key(i + 1) - 1u : kmax;
  auto F = [&](uint32_t k) { return gelu_code_direct(key2f(k), ro); };
  const int cb = F(klo), ce = F(khi);
  uint32_t thr_bits = 0x7f800000u;   // +inf: no change inside the segment
  bool slow = false;
  int below = cb, above = cb;
  const uint32_t span = khi - klo;
  if (cb == ce) {
    for (int s = 1; s < 16; ++s) slow |= F(klo + uint32_t((uint64_t(span) * s) >> 4)) != cb;   // interior probes (dip near the minimum)
  } else if (ce - cb == 1 || cb - ce == 1) {
    uint32_t lo = klo, hi = khi;     // F(lo) == cb, F(hi) == ce
    while (hi - lo > 1u) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (F(mid) == ce) hi = mid; else lo = mid;
    }
    thr_bits = __float_as_uint(key2f(hi));
    above = ce;
    // clean step outside the +-8 ulp band the epilogue re-evaluates: probe up to 64 keys on both sides and 16 spread points
    for (uint32_t d = 9; d <= 64; ++d) {
      if (hi - klo >= d) slow |= F(hi - d) != cb;
      if (khi - hi >= d) slow |= F(hi + d) != ce;
    }
    for (int s = 1; s < 16; ++s) {
      const uint32_t k = klo + uint32_t((uint64_t(span) * s) >> 4);
      if (k + 8u < hi) slow |= F(k) != cb;
      if (k > hi + 8u) slow |= F(k) != ce;
    }
    // the threshold's sign must equal y's for the ulp-distance test: a threshold at +-0 / subnormal cannot be handled by it
    slow |= (thr_bits & 0x7fffffffu) < 0x00800000u;
  } else {
    slow = true;
  }
  entries[i] = make_uint2(thr_bits, (uint32_t(below) & 0xffu) | ((uint32_t(above) & 0xffu) << 8) | (slow ? 0x80000000u : 0u));
}

int launch_build_gelu_table(float out_scale, void* table_dev, cudaStream_t stream) {
  int ex = 0;
  const float m = frexpf(out_scale, &ex);
  if (!(out_scale > 0.f) || m != 0.5f) return 3;                     // power of two only
  const double inv_w = 2.0 / double(out_scale);                      // segment width out_scale / 2
  const double n_d = (8.5 + 128.0 * double(out_scale) + 0.5) * inv_w;
  if (n_d > double(P2V_GELU_TABLE_MAX_ENTRIES) || n_d < 4.0) return 3;
  GeluTabHeader hd;
  hd.y0 = -8.5f;
  hd.inv_w = float(inv_w);
  hd.n = int(n_d);
  hd.reserved = 0;
  cudaMemcpyAsync(table_dev, &hd, sizeof(hd), cudaMemcpyHostToDevice, stream);   // pageable source: staged before the call returns
  build_gelu_table_kernel<<<(hd.n + 127) / 128, 128, 0, stream>>>(hd, 1.0f / out_scale,
                                                                   reinterpret_cast<uint2*>(reinterpret_cast<char*>(table_dev) + sizeof(hd)));
  count_launch();
  return check_launch("build_gelu_table");
}

}  // namespace p2v
