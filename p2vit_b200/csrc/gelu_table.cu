// Builds the step table of  y -> sat(RNE(gelu_erf(y) / out_scale))  (include/p2vit_b200.h: p2v_build_gelu_table).
// One thread per segment: the segment's first / last fp32 values are found by bisection on gelu_segment() itself, the
// code change inside it (at most one for a segment of width out_scale/2) by bisection on gelu_code_direct(), and the
// neighbourhood of the threshold is probed so that a segment whose code is not a clean step is flagged `slow`
// (the epilogue then evaluates erf directly for every y that lands in it).
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>
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

// ------------------------------------------------------------------------------------------------
// second form (common.cuh: GeluStepsHeader): exact per-code thresholds on both branches + a segment-wise linear map
// ------------------------------------------------------------------------------------------------
// thread i: threshold of code boundary c = cr0 + i (right of y*: smallest y with code >= c) or, for i >= nr, c = cr0 + i - nr
// (left of y*: smallest y with code < c, the code being non-increasing in y there); +-inf when the boundary is never / always
// crossed on that branch.  A threshold whose neighbourhood is not a clean step outside the +-8 ulp band clears *ok.
__global__ void __launch_bounds__(128) build_gelu_steps_kernel(GeluStepsHeader h, float so, float zp, int cr0, float* __restrict__ thr, int* __restrict__ ok) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h.nr + h.nl) return;
  const bool left = i >= h.nr;
  const int c = cr0 + (left ? i - h.nr : i);
  auto F = [&](uint32_t k) { return gelu_code_div(key2f(k), so, zp); };
  const uint32_t k_star = f2key(h.ystar), k_lo = left ? f2key(h.ymin) : k_star, k_hi = left ? k_star : f2key(h.ymax);
  // on [k_lo, k_hi] the predicate Q(k) = (code >= c) on the right branch, (code < c) on the left one, goes from false to true
  auto Q = [&](uint32_t k) { return left ? F(k) < c : F(k) >= c; };
  const float inf = __int_as_float(0x7f800000);
  float out;
  bool bad = false;
  if (Q(k_lo)) {
    out = left ? -inf : -inf;           // true on the whole branch: right "y >= -inf", left "code < c everywhere" i.e. never code >= c
  } else if (!Q(k_hi)) {
    out = inf;                          // false on the whole branch
  } else {
    uint32_t lo = k_lo, hi = k_hi;
    while (hi - lo > 1u) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (Q(mid)) hi = mid; else lo = mid;
    }
    out = key2f(hi);
    for (uint32_t d = 9; d <= 64; ++d) {
      if (hi - k_lo >= d) bad |= Q(hi - d);
      if (k_hi - hi >= d) bad |= !Q(hi + d);
    }
    const uint32_t span = k_hi - k_lo;
    for (int s = 1; s < 32; ++s) {
      const uint32_t k = k_lo + uint32_t((uint64_t(span) * s) >> 5);
      if (k + 8u < hi) bad |= Q(k);
      if (k > hi + 8u) bad |= !Q(k);
    }
    bad |= fabsf(out) < 1e-30f;         // the ulp-distance test needs a threshold away from +-0
  }
  thr[i] = out;
  if (bad) atomicExch(ok, 0);
}

// Self-check with the epilogue's own code path (replicated shared-memory tables, gelu_steps_code): +-256 ulps around every
// threshold, a uniform grid over the active range and beyond, and a sweep of magnitudes; any y that is not sent to the direct
// evaluation (near_min > 16) must get exactly the direct code.  A mismatch clears *ok.  A y inside the band whose step code
// differs from the direct one clears *clean: the band around every threshold is scanned exhaustively, so a table that keeps
// `clean` needs no distance test in the epilogue.
__global__ void __launch_bounds__(256) verify_gelu_steps_kernel(const void* __restrict__ table, float so, float zp, int* __restrict__ ok, int* __restrict__ clean) {
  extern __shared__ uint8_t vsm[];
  const uint32_t base = (uint32_t(__cvta_generic_to_shared(vsm)) + 255u) & ~255u;
  gelu_steps_fill_smem(table, base, int(threadIdx.x), int(blockDim.x));
  __syncthreads();
  const GeluSteps t = gelu_steps_view(table, base, int(threadIdx.x) & 31);
  const GeluStepsHeader h = *reinterpret_cast<const GeluStepsHeader*>(reinterpret_cast<const char*>(table) + P2V_GELU_STEPS_OFFSET);
  const float* thr = reinterpret_cast<const float*>(reinterpret_cast<const char*>(table) + P2V_GELU_STEPS_OFFSET + sizeof(GeluStepsHeader) +
                                                    8 * P2V_GELU_STEPS_MAX_SEG);
  bool bad = false, unclean = false;
  auto check = [&](float y) {
    uint32_t nm = 0xffffffffu;
    const int c = int(int8_t(gelu_steps_code(y, t, nm) & 0xffu));
    if (c != gelu_code_div(y, so, zp)) {
      if (nm > 16u) bad = true; else unclean = true;
    }
  };
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
  const long long n1 = (long long)(h.nr + h.nl) * 513;
  for (long long j = gid; j < n1; j += gsz) {
    const float v = thr[j / 513];
    if (fabsf(v) < 1e30f) check(key2f(f2key(v) + uint32_t(int(j % 513) - 256)));
  }
  const long long n2 = 1ll << 21;
  const float a = h.ymin - 1.5f, w = (h.ymax - h.ymin + 3.0f) / float(n2);
  for (long long j = gid; j < n2; j += gsz) check(a + w * float(j));
  for (long long j = gid; j < 8192; j += gsz) {
    const float m = exp2f(-100.0f + 200.0f * float(j >> 1) / 4096.0f);
    check((j & 1) ? -m : m);
  }
  if (gid == 0) { check(0.0f); check(-0.0f); check(h.ymin); check(h.ymax); check(h.ystar); }
  if (bad) atomicExch(ok, 0);
  if (unclean) atomicExch(clean, 0);
}

// tables whose self-check found them clean, by device address (the table itself lives in device memory; the GEMM launcher
// picks the epilogue variant on the host)
static std::mutex g_clean_mu;
static std::unordered_map<const void*, bool> g_clean;
bool gelu_table_is_clean(const void* table_dev) {
  std::lock_guard<std::mutex> lk(g_clean_mu);
  auto it = g_clean.find(table_dev);
  return it != g_clean.end() && it->second;
}

static double gelu_f64(double y) { return 0.5 * y * (1.0 + erf(y * 0.70710678118654752440)); }

// host side of the second form: active range, segment map, table dimensions.  false = this scale is not tabulated.
static bool plan_gelu_steps(float out_scale, float out_zp, GeluStepsHeader& h, std::vector<float2>& seg, int& cr0) {
  const double so = out_scale, ro = 1.0 / so, zp = out_zp;
  if (!(zp >= -128.0 && zp <= 127.0) || zp != floor(zp)) return false;
  const float ystar = -0.7517916f;
  const double gmin = gelu_f64(ystar);
  if (-gmin * ro <= 0.45 || so > 0.25) return false;          // (almost) no negative codes: not worth a table
  double lo = 0.0, hi = 300.0 * so + 10.0;                     // ymax: gelu = 128.2 * so (every y above saturates to 127)
  for (int it = 0; it < 200; ++it) { const double mid = 0.5 * (lo + hi); if (gelu_f64(mid) < (128.2 - zp) * so) lo = mid; else hi = mid; }
  const double ymax = hi;
  lo = -40.0; hi = ystar;                                      // ymin: gelu = -0.4 * so left of the minimum (every y below has code 0)
  for (int it = 0; it < 200; ++it) { const double mid = 0.5 * (lo + hi); if (gelu_f64(mid) > -0.4 * so) lo = mid; else hi = mid; }
  const double ymin = lo;
  // lowest code - 2, but not below -128: with a zero point near -128 (post-GELU activations under an asymmetric observer) the low
  // codes saturate; f is then clamped at f0 = -129 and the thresholds of codes <= -128 are -inf, so the lookup returns -128 there
  cr0 = std::max(int(floor(gmin * ro + zp + 0.5)) - 2, -128);
  h.ymin = float(ymin); h.ymax = float(ymax); h.ystar = ystar;
  h.nr = 131 - cr0; h.nl = std::max(int(zp) + 3 - cr0, 1); h.k1 = 1 - cr0; h.ok = 1; h.clean = 1; h.zp = out_zp;
  h.pad[0] = 0;
  const double f0 = double(cr0 - 1);
  h.f_scale = float(126.0 - f0 + 0.49);
  const int entries = h.nr + h.nl;
  if (entries > P2V_GELU_STEPS_MAX_THR) return false;
  if (entries * 128 <= 26 * 1024) h.rep_log2 = 5;
  else if (entries * 64 <= 26 * 1024) h.rep_log2 = 4;
  else return false;
  for (int nseg = 16; nseg <= P2V_GELU_STEPS_MAX_SEG; nseg *= 2) {
    const double inv_w = double(nseg) / (double(h.ymax) - double(h.ymin)) * (1.0 - 1e-4);
    h.nseg = nseg; h.inv_w = float(inv_w); h.soff = float(-double(h.ymin) * double(h.inv_w) - 0.5 + 5e-4);
    h.seg_scale = float(double(nseg) - 1.0 + 0.49);
    seg.assign(nseg, make_float2(0.f, 0.f));
    double worst = 0.0;
    for (int s = 0; s < nseg; ++s) {      // y with RNE(y * inv_w + soff) == s, widened by 2 % of a segment, clipped to the active range
      double a = (double(s) - 0.52 - double(h.soff)) / double(h.inv_w), b = (double(s) + 0.52 - double(h.soff)) / double(h.inv_w);
      a = std::max(a, double(h.ymin)); b = std::min(b, double(h.ymax));
      if (!(b > a)) {
        seg[s] = make_float2(0.f, float((gelu_f64(std::min(std::max(a, double(h.ymin)), double(h.ymax))) * ro + zp - 0.5 - f0) / double(h.f_scale)));
        continue;
      }
      const double A = (gelu_f64(b) - gelu_f64(a)) * ro / (b - a);
      double dmin = 1e300, dmax = -1e300;
      for (int j = 0; j <= 128; ++j) {
        const double y = a + (b - a) * j / 128.0, d = gelu_f64(y) * ro + zp - A * y;
        dmin = std::min(dmin, d); dmax = std::max(dmax, d);
      }
      worst = std::max(worst, 0.5 * (dmax - dmin));
      seg[s] = make_float2(float(A / double(h.f_scale)), float((0.5 * (dmax + dmin) - 0.5 - f0) / double(h.f_scale)));
    }
    if (worst <= 0.30) return true;
  }
  return false;
}

// builds the second form behind the first in `table_dev`; synchronises `stream` to read the verdict back
static int build_gelu_steps(float out_scale, float out_zp, void* table_dev, cudaStream_t stream) {
  GeluStepsHeader h;
  std::vector<float2> seg;
  int cr0 = 0;
  char* base = reinterpret_cast<char*>(table_dev) + P2V_GELU_STEPS_OFFSET;
  if (!plan_gelu_steps(out_scale, out_zp, h, seg, cr0)) return 3;
  cudaMemcpyAsync(base, &h, sizeof(h), cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(base + sizeof(h), seg.data(), seg.size() * sizeof(float2), cudaMemcpyHostToDevice, stream);
  int* ok_dev = &reinterpret_cast<GeluStepsHeader*>(base)->ok;
  float* thr_dev = reinterpret_cast<float*>(base + sizeof(h) + 8 * P2V_GELU_STEPS_MAX_SEG);
  build_gelu_steps_kernel<<<(h.nr + h.nl + 127) / 128, 128, 0, stream>>>(h, out_scale, out_zp, cr0, thr_dev, ok_dev);
  const size_t smem = gelu_steps_smem_bytes(h) + 256;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(verify_gelu_steps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(P2V_GELU_STEPS_SMEM_MAX + 256));
    attr = true;
  }
  int* clean_dev = &reinterpret_cast<GeluStepsHeader*>(base)->clean;
  verify_gelu_steps_kernel<<<64, 256, smem, stream>>>(table_dev, out_scale, out_zp, ok_dev, clean_dev);
  count_launch(2);
  if (int r = check_launch("build_gelu_steps")) return r;
  GeluStepsHeader back;
  cudaMemcpyAsync(&back, base, sizeof(back), cudaMemcpyDeviceToHost, stream);
  if (cudaStreamSynchronize(stream) != cudaSuccess) { set_error("build_gelu_steps: %s", cudaGetErrorString(cudaGetLastError())); return 2; }
  static const bool no_clean = getenv("P2V_GELU_GUARD") && atoi(getenv("P2V_GELU_GUARD")) != 0;     // triage: keep the distance test
  {
    std::lock_guard<std::mutex> lk(g_clean_mu);
    g_clean[table_dev] = back.ok && back.clean && !no_clean;
  }
  return back.ok ? 0 : 3;
}

int launch_build_gelu_table(float out_scale, float out_zp, void* table_dev, cudaStream_t stream) {
  int ex = 0;
  const float m = frexpf(out_scale, &ex);
  if (!(out_scale > 0.f) || !(out_scale < 1e30f)) return 3;
  GeluTabHeader hd;
  hd.y0 = -8.5f;
  hd.reserved = 0;
  if (m != 0.5f || out_zp != 0.f) {
    // not a power of two (ema / percentile observers) or a zero point (omse): only the second form - its thresholds come from
    // the reference's own sequence fl(gelu(y) / out_scale) + zp, so it holds for any quantizer; the first form (one-tile and dp4a
    // kernels) stays empty
    hd.inv_w = 0.f;
    hd.n = 0;
    cudaMemcpyAsync(table_dev, &hd, sizeof(hd), cudaMemcpyHostToDevice, stream);
    return build_gelu_steps(out_scale, out_zp, table_dev, stream);
  }
  const double inv_w = 2.0 / double(out_scale);                      // segment width out_scale / 2
  const double n_d = (8.5 + 128.0 * double(out_scale) + 0.5) * inv_w;
  if (n_d > double(P2V_GELU_TABLE_MAX_ENTRIES) || n_d < 4.0) return 3;
  hd.inv_w = float(inv_w);
  hd.n = int(n_d);
  cudaMemcpyAsync(table_dev, &hd, sizeof(hd), cudaMemcpyHostToDevice, stream);   // pageable source: staged before the call returns
  build_gelu_table_kernel<<<(hd.n + 127) / 128, 128, 0, stream>>>(hd, 1.0f / out_scale,
                                                                   reinterpret_cast<uint2*>(reinterpret_cast<char*>(table_dev) + sizeof(hd)));
  count_launch();
  if (int r = check_launch("build_gelu_table")) return r;
  return build_gelu_steps(out_scale, 0.f, table_dev, stream);             // both forms or none: the caller passes one pointer to every kernel
}

}  // namespace p2v
