// HBM-bound element / row kernels: QAct quantize & fake-quant, patch gather, integer LayerNorm,
// stand-alone integer log2-softmax, calibration reductions.  Reference call sites are cited in
// include/p2vit_b200.h next to each entry point.
#include <cstdlib>
#include "common.cuh"

namespace p2v {

static inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ------------------------------------------------------------------------------------------------
// QAct: q = clamp(RNE(x/s + zp), lo, hi); y = (q - zp) * s           (uniform.py:83-86,125)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float quant_code(float x, float s, float zp, float lo, float hi) {
  float v = rintf(fadd(fdiv(x, s), zp));
  return fminf(fmaxf(v, lo), hi);
}

template <bool VEC4>
__global__ void __launch_bounds__(256) quantize_kernel(const float* __restrict__ x, int8_t* __restrict__ q, float* __restrict__ y,
                                                       int64_t n, int C, int64_t inner, const float* __restrict__ scale,
                                                       int n_scale, float zp, float lo, float hi) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  if (VEC4) {
    // 4 consecutive elements share... not necessarily a channel: inner % 4 == 0 (same channel) or inner == 1 && C % 4 == 0
    const int64_t n4 = n >> 2;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
      float s[4];
      if (n_scale == 1) {
        s[0] = s[1] = s[2] = s[3] = __ldg(scale);
      } else if (inner == 1) {
        const int c = int((i << 2) % C);
        const float4 sv = __ldg(reinterpret_cast<const float4*>(scale + c));
        s[0] = sv.x; s[1] = sv.y; s[2] = sv.z; s[3] = sv.w;
      } else {
        const int c = int(((i << 2) / inner) % C);
        s[0] = s[1] = s[2] = s[3] = __ldg(scale + c);
      }
      const float c0 = quant_code(v.x, s[0], zp, lo, hi), c1 = quant_code(v.y, s[1], zp, lo, hi);
      const float c2 = quant_code(v.z, s[2], zp, lo, hi), c3 = quant_code(v.w, s[3], zp, lo, hi);
      if (q) reinterpret_cast<uint32_t*>(q)[i] = pack4_s8(int(c0), int(c1), int(c2), int(c3));
      if (y) reinterpret_cast<float4*>(y)[i] = make_float4(fmul(fsub(c0, zp), s[0]), fmul(fsub(c1, zp), s[1]),
                                                            fmul(fsub(c2, zp), s[2]), fmul(fsub(c3, zp), s[3]));
    }
  } else {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float s = __ldg(scale + (n_scale == 1 ? 0 : int((i / inner) % C)));
      const float c = quant_code(__ldg(x + i), s, zp, lo, hi);
      if (q) q[i] = int8_t(int(c));
      if (y) y[i] = fmul(fsub(c, zp), s);
    }
  }
}

int launch_quantize(const float* x, int8_t* q, float* y, int64_t n, int C, int64_t inner, const float* scale, int n_scale,
                    float zp, int lo, int hi, cudaStream_t stream) {
  if (n == 0) return 0;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(q) & 3) == 0;
  const bool vec = aligned && (n % 4 == 0) && (n_scale == 1 || inner % 4 == 0 || (inner == 1 && C % 4 == 0));
  const int64_t work = vec ? n / 4 : n;
  const int blocks = int(std::min<int64_t>((work + 255) / 256, int64_t(num_sms()) * 16));
  if (vec) quantize_kernel<true><<<blocks, 256, 0, stream>>>(x, q, y, n, C, inner, scale, n_scale, zp, float(lo), float(hi));
  else quantize_kernel<false><<<blocks, 256, 0, stream>>>(x, q, y, n, C, inner, scale, n_scale, zp, float(lo), float(hi));
  count_launch();
  return check_launch("quantize");
}

__global__ void __launch_bounds__(256) dequantize_kernel(const int8_t* __restrict__ q, float* __restrict__ y, int64_t n, int C,
                                                         int64_t inner, const float* __restrict__ scale, int n_scale, float zp) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float s = __ldg(scale + (n_scale == 1 ? 0 : int((i / inner) % C)));
    y[i] = fmul(fsub(float(q[i]), zp), s);
  }
}

int launch_dequantize(const int8_t* q, float* y, int64_t n, int C, int64_t inner, const float* scale, int n_scale, float zp,
                      cudaStream_t stream) {
  if (n == 0) return 0;
  const int blocks = int(std::min<int64_t>((n + 255) / 256, int64_t(num_sms()) * 16));
  dequantize_kernel<<<blocks, 256, 0, stream>>>(q, y, n, C, inner, scale, n_scale, zp);
  count_launch();
  return check_launch("dequantize");
}

// ------------------------------------------------------------------------------------------------
// qact_input + patch gather: thread = one P-pixel row segment of one patch and channel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ img, int8_t* __restrict__ out, int B, int Cin,
                                                       int H, int W, int P, float s, float zp, float lo, float hi) {
  pdl_wait();
  pdl_trigger();
  const int gw = W / P, gh = H / P;
  const int64_t total = int64_t(B) * Cin * H * gw;  // segments
  const int K = Cin * P * P;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  // int8 codes ([-128, 127] bounds) at a power-of-two scale take the division-free path
  const bool pot = (__float_as_uint(s) & 0x007fffffu) == 0u && s > 0.f && lo >= -128.f && hi <= 127.f;
  const float rs = fdiv(1.f, s), mlo = fadd(RMAGIC, lo), mhi = fadd(RMAGIC, hi);
  const bool vec16 = P % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  for (int64_t seg = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; seg < total; seg += stride) {
    const int j = int(seg % gw);
    int64_t t = seg / gw;
    const int yy = int(t % H);
    t /= H;
    const int c = int(t % Cin);
    const int b = int(t / Cin);
    const float4* src = reinterpret_cast<const float4*>(img + ((int64_t(b) * Cin + c) * H + yy) * W + j * P);
    const int i = yy / P, py = yy % P;
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (int64_t(b) * gh * gw + int64_t(i) * gw + j) * K + (c * P + py) * P);
    if (pot) {
      // power-of-two scale: x * (1/s) == x / s exactly; RNE and the clamp on the 1.5 * 2^23-biased sum (monotone in its argument,
      // so the clamp is right for any magnitude); no division, FRND or F2I per pixel, one 16-byte store per four words
      uint32_t w4[4];
      for (int v0 = 0; v0 < P / 4; v0 += 4) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          if (v0 + v < P / 4) {
            const float4 f = __ldg(src + v0 + v);
            const float t0 = fminf(fmaxf(fadd(fadd(fmul(f.x, rs), zp), RMAGIC), mlo), mhi), t1 = fminf(fmaxf(fadd(fadd(fmul(f.y, rs), zp), RMAGIC), mlo), mhi);
            const float t2 = fminf(fmaxf(fadd(fadd(fmul(f.z, rs), zp), RMAGIC), mlo), mhi), t3 = fminf(fmaxf(fadd(fadd(fmul(f.w, rs), zp), RMAGIC), mlo), mhi);
            w4[v] = pack4_sat(t0, t1, t2, t3);
          }
        }
        if (vec16 && v0 + 4 <= P / 4) *reinterpret_cast<uint4*>(dst + v0) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        else {
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if (v0 + v < P / 4) dst[v0 + v] = w4[v];
        }
      }
    } else {
      for (int v = 0; v < P / 4; ++v) {
        const float4 f = __ldg(src + v);
        dst[v] = pack4_s8(int(quant_code(f.x, s, zp, lo, hi)), int(quant_code(f.y, s, zp, lo, hi)),
                          int(quant_code(f.z, s, zp, lo, hi)), int(quant_code(f.w, s, zp, lo, hi)));
      }
    }
  }
}

int launch_patchify(const float* img, int8_t* out, int B, int Cin, int H, int W, int P, float scale, float zp, int lo, int hi,
                    cudaStream_t stream) {
  const int64_t total = int64_t(B) * Cin * H * (W / P);
  const int blocks = int(std::min<int64_t>((total + 255) / 256, int64_t(num_sms()) * 16));
  launch_pdl(patchify_kernel, dim3(blocks), dim3(256), 0, stream, img, out, B, Cin, H, W, P, scale, zp, float(lo), float(hi));
  count_launch();
  return check_launch("patchify");
}

// ------------------------------------------------------------------------------------------------
// 8-bit pixels: ToTensor + Normalize + qact_input of a pixel depend only on (channel, byte), so the host tabulates the 256
// codes per channel (with the fp32 quantizer above) and this kernel is a table gather fused with the patch gather:
// 1 byte read + 1 byte written per pixel instead of 4 + 1.  Thread = one P-pixel row segment, like patchify_kernel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patchify_u8_kernel(const uint8_t* __restrict__ img, const int8_t* __restrict__ lut,
                                                          int8_t* __restrict__ out, int B, int Cin, int H, int W, int P) {
  extern __shared__ uint8_t lut_s[];
  pdl_wait();
  pdl_trigger();
  for (int i = threadIdx.x; i < Cin * 256; i += blockDim.x) lut_s[i] = uint8_t(lut[i]);
  __syncthreads();
  const int gw = W / P, gh = H / P;
  const int64_t total = int64_t(B) * Cin * H * gw;
  const int K = Cin * P * P;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const bool vec16 = P % 16 == 0 && W % 16 == 0 && ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(img)) & 15) == 0;
  auto code4 = [](const uint8_t* l, uint32_t w) {
    return uint32_t(l[w & 255u]) | (uint32_t(l[(w >> 8) & 255u]) << 8) | (uint32_t(l[(w >> 16) & 255u]) << 16) | (uint32_t(l[w >> 24]) << 24);
  };
  for (int64_t seg = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; seg < total; seg += stride) {
    const int j = int(seg % gw);
    int64_t t = seg / gw;
    const int yy = int(t % H);
    t /= H;
    const int c = int(t % Cin);
    const int b = int(t / Cin);
    const uint8_t* src = img + ((int64_t(b) * Cin + c) * H + yy) * W + j * P;
    const int i = yy / P, py = yy % P;
    int8_t* dst = out + (int64_t(b) * gh * gw + int64_t(i) * gw + j) * K + (c * P + py) * P;
    const uint8_t* l = lut_s + c * 256;
    if (vec16) {
      for (int v = 0; v < P / 16; ++v) {
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(src) + v);
        reinterpret_cast<uint4*>(dst)[v] = make_uint4(code4(l, w.x), code4(l, w.y), code4(l, w.z), code4(l, w.w));
      }
    } else {
      for (int v = 0; v < P / 4; ++v)
        reinterpret_cast<uint32_t*>(dst)[v] = code4(l, __ldg(reinterpret_cast<const uint32_t*>(src) + v));
    }
  }
}

int launch_patchify_u8(const uint8_t* img, const int8_t* lut, int8_t* out, int B, int Cin, int H, int W, int P, cudaStream_t stream) {
  const int64_t total = int64_t(B) * Cin * H * (W / P);
  const int blocks = int(std::min<int64_t>((total + 255) / 256, int64_t(num_sms()) * 8));
  launch_pdl(patchify_u8_kernel, dim3(blocks), dim3(256), size_t(Cin) * 256, stream, img, lut, out, B, Cin, H, W, P);
  count_launch();
  return check_launch("patchify_u8");
}

__global__ void fill_cls_kernel(int8_t* __restrict__ out, const int8_t* __restrict__ cls_row, int B, int T, int N) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x;
  for (int n = threadIdx.x; n < N; n += blockDim.x) out[size_t(b) * (T + 1) * N + n] = cls_row[n];
}

int launch_fill_cls(int8_t* out, const int8_t* cls_row, int B, int T, int N, cudaStream_t stream) {
  launch_pdl(fill_cls_kernel, dim3(B), dim3(128), 0, stream, out, cls_row, B, T, N);
  count_launch();
  return check_launch("fill_cls");
}

// ------------------------------------------------------------------------------------------------
// QIntLayerNorm (int mode) + following QAct.  One warp per row; lane owns words lane, lane+32, ...
// ------------------------------------------------------------------------------------------------
template <bool POT>
__device__ __forceinline__ float div_by(float x, float s) { return POT ? fmul(x, fdiv(1.f, s)) : fdiv(x, s); }

// word w (4 channels) of input row `row`: plain rows, or - Swin patch merging fused into the LayerNorm (p2v_layernorm_args.in_gather) -
// the concatenation of gather_segs source rows
__device__ __forceinline__ const uint32_t* ln_word_ptr(const p2v_layernorm_args& a, int row, int w) {
  if (a.in_gather == nullptr) return reinterpret_cast<const uint32_t*>(a.x + int64_t(row) * a.x_row_stride) + w;
  const int seg_words = a.C / (4 * a.gather_segs);
  const int seg = w / seg_words;
  return reinterpret_cast<const uint32_t*>(a.x + int64_t(__ldg(a.in_gather + int64_t(row) * a.gather_segs + seg)) * a.x_row_stride) + (w - seg * seg_words);
}

template <int WPL, bool POT>
__global__ void __launch_bounds__(256) layernorm_kernel(p2v_layernorm_args a) {
  pdl_wait();
  pdl_trigger();
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nwords = a.C >> 2;
  const float Cf = float(a.C);
  const float s1 = a.in_scale_min;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < a.rows; row += gridDim.x * warps_per_block) {
    int xv[WPL][4];
    int S1 = 0, S2 = 0;
#pragma unroll
    for (int i = 0; i < WPL; ++i) {
      const int w = lane + 32 * i;
      if (w < nwords) {
        const uint32_t u = __ldg(ln_word_ptr(a, row, w));
        const float4 m = __ldg(reinterpret_cast<const float4*>(a.in_mult) + w);
        xv[i][0] = int(int8_t(u & 0xff)) * int(m.x);
        xv[i][1] = int(int8_t((u >> 8) & 0xff)) * int(m.y);
        xv[i][2] = int(int8_t((u >> 16) & 0xff)) * int(m.z);
        xv[i][3] = int(int8_t(u >> 24)) * int(m.w);
#pragma unroll
        for (int e = 0; e < 4; ++e) { S1 += xv[i][e]; S2 += xv[i][e] * xv[i][e]; }
      }
    }
    S1 = __reduce_add_sync(0xffffffffu, S1);
    S2 = __reduce_add_sync(0xffffffffu, S2);
    // layers.py:315-318  (row sums are exact integers; every fp32 op below is one ATen op of the reference)
    const float S1f = float(S1), S2f = float(S2);
    const float mean = fmul(fdiv(S1f, Cf), s1);
    const float stdv = fmul(fdiv(s1, Cf), __fsqrt_rn(fsub(fmul(Cf, S2f), fmul(S1f, S1f))));
    const float t = fdiv(s1, stdv);
    const float mos = fdiv(mean, stdv);
#pragma unroll
    for (int i = 0; i < WPL; ++i) {
      const int w = lane + 32 * i;
      if (w < nwords) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.gamma) + w);
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.beta) + w);
        const float4 o4 = __ldg(reinterpret_cast<const float4*>(a.out_scale) + w);
        const float4 p4 = __ldg(reinterpret_cast<const float4*>(a.post_div) + w);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
        const float os[4] = {o4.x, o4.y, o4.z, o4.w}, pd[4] = {p4.x, p4.y, p4.z, p4.w};
        int q[4];
        float yf[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float A = div_by<POT>(fmul(t, g[e]), os[e]);                       // layers.py:320-324
          const float aA = fabsf(A);
          int N;                                                                    // get_MN, layers.py:270-274
          if (aA == 0.f) N = 31;
          else if (!(aA < __int_as_float(0x7f800000))) N = 0;
          else N = min(max(7 - floor_log2_as_fp32(aA), 0), 31);
          const float twoN = pow2i(N);
          const float M = fminf(fmaxf(floorf(fmul(aA, twoN)), 0.f), 255.f);
          const float sgn = A > 0.f ? 1.f : (A < 0.f ? -1.f : 0.f);
          const float Bv = rintf(fmul(div_by<POT>(fsub(be[e], fmul(mos, g[e])), os[e]), twoN));   // :327-334
          const float yq = rintf(fmul(fadd(fmul(fmul(sgn, M), float(xv[i][e])), Bv), pow2i(-N)));  // :336
          const float deq = fmul(yq, os[e]);                                                       // :337
          yf[e] = deq;
          const float mid = a.clamp_mid ? fmul(fminf(fmaxf(yq, -128.f), 127.f), os[e]) : deq;       // QAct at the LN's own scale
          q[e] = sat_s8(fadd(div_by<POT>(div_by<POT>(mid, pd[e]), a.next_scale), a.next_zp));   // next_zp: 0 unless the next QAct is asymmetric
        }
        if (a.out_i8) reinterpret_cast<uint32_t*>(a.out_i8 + int64_t(a.out_row_map ? __ldg(a.out_row_map + row) : row) * a.C)[w] = pack4_s8(q[0], q[1], q[2], q[3]);
        if (a.out_f32) reinterpret_cast<float4*>(a.out_f32 + int64_t(row) * a.C)[w] = make_float4(yf[0], yf[1], yf[2], yf[3]);
      }
    }
  }
}

template <bool SLOW>
__device__ __forceinline__ uint32_t ln_pot_word(float t, float mos, const float (&g)[4], const float (&bt)[4], const float (&f)[4],
                                                const int (&xv)[4], float clamp_hi) {
  int q[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float A = fmul(t, g[e]);
    const uint32_t ab = __float_as_uint(A) & 0x7fffffffu;
    int ex = int(ab >> 23) - 127;                  // 0 / subnormal -> N = 31, inf / nan -> N = 0 through the clamp
    if (SLOW) {
      if ((ab & 0x007fffffu) >= 0x007ffff0u && ab < 0x7f800000u) ex = floor_log2_as_fp32(__uint_as_float(ab));
    }
    const int N = min(max(7 - ex, 0), 31);
    const float twoN = __uint_as_float(uint32_t(N + 127) << 23), rtwoN = __uint_as_float(uint32_t(127 - N) << 23);
    const float M = fminf(floorf(fmul(__uint_as_float(ab), twoN)), 255.f);
    const float sM = __uint_as_float(__float_as_uint(M) | (__float_as_uint(A) & 0x80000000u));
    const float Bv = rintf(fmul(fsub(bt[e], fmul(mos, g[e])), twoN));
    float yq = rintf(fmul(fadd(fmul(sM, float(xv[e])), Bv), rtwoN));
    yq = fminf(fmaxf(yq, -clamp_hi - 1.f), clamp_hi);       // clamp_mid: [-128,127]; otherwise +-inf bounds (no-op)
    q[e] = sat_s8(fmul(yq, f[e]));
  }
  return pack4_s8(q[0], q[1], q[2], q[3]);
}

// one lane's words of a row through ln_pot_word<true>, every constant re-derived from global memory as the kernel prologue does
__device__ __noinline__ void ln_pot_row_slow(const p2v_layernorm_args& a, int row, uint32_t* __restrict__ orow, int sub,
                                             int lpr, int wpln, float t, float mos, float clamp_hi) {
  const float rnext = fdiv(1.f, a.next_scale);
  for (int i = 0; i < wpln; ++i) {
    const int w = sub + lpr * i;
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.gamma) + w), b4 = __ldg(reinterpret_cast<const float4*>(a.beta) + w);
    const float4 o4 = __ldg(reinterpret_cast<const float4*>(a.out_scale) + w), p4 = __ldg(reinterpret_cast<const float4*>(a.post_div) + w);
    const float4 m4 = __ldg(reinterpret_cast<const float4*>(a.in_mult) + w);
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w}, oo[4] = {o4.x, o4.y, o4.z, o4.w};
    const float pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w};
    const uint32_t u = __ldg(ln_word_ptr(a, row, w));
    const int cx[4] = {int(int8_t(u & 0xff)), int(int8_t((u >> 8) & 0xff)), int(int8_t((u >> 16) & 0xff)), int(int8_t(u >> 24))};
    float g[4], bt[4], f[4];
    int xv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ros = fdiv(1.f, oo[e]);
      g[e] = fmul(gg[e], ros);
      bt[e] = fmul(bb[e], ros);
      f[e] = fmul(fmul(oo[e], fdiv(1.f, pp[e])), rnext);
      xv[e] = cx[e] * int(mm[e]);
    }
    orow[w] = ln_pot_word<true>(t, mos, g, bt, f, xv, clamp_hi);
  }
}

// The same arithmetic for the common case 2^-24 <= |A| < 2^8 with integer / magic-constant tricks (results identical):
//   N = 7 - (E - 127) with E the biased exponent of A, so 2^N and 2^-N are exponent-field subtractions;
//   M = floor(|A| * 2^N) = the top 8 bits of A's significand, and sign(A) * M as a float is A with its low 16 mantissa bits
//   cleared and its exponent set to 7:  (bits(A) & 0x807f0000) | 0x43000000;
//   RNE(v) for |v| < 2^22 = (v + 1.5*2^23) - 1.5*2^23, fused with the preceding power-of-two scaling into one FFMA;
//   the final saturation to int8 is pack4_sat's.  `mant_max` collects max(mantissa(A)) for the caller's corner-case test.
template <bool CLAMP_MID>
__device__ __forceinline__ uint32_t ln_pot_fast_word(float t, float mos, const float (&g)[4], const float (&bt)[4], const float (&f)[4],
                                                     const int (&xv)[4], uint32_t& mant_max) {
  float r[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const uint32_t Ab = __float_as_uint(fmul(t, g[e]));
    const uint32_t Ef = Ab & 0x7f800000u;
    mant_max = max(mant_max, Ab & 0x007fffffu);
    const float twoN = __uint_as_float(0x82800000u - Ef);            // 2^(134 - E)
    const float rtwoN = __uint_as_float(Ef - 0x03800000u);           // 2^(E - 134)
    const float sM = __uint_as_float((Ab & 0x807f0000u) | 0x43000000u);
    const float Bv = rintf(fmul(fsub(bt[e], fmul(mos, g[e])), twoN));
    const float sum = __fmaf_rn(sM, __int2float_rn(xv[e]), Bv);     // sM * x is exact (8 x 11 bits): one rounding, as fadd(fmul(..), Bv)
    float yq = fsub(__fmaf_rn(sum, rtwoN, RMAGIC), RMAGIC);          // RNE(sum / 2^N)
    if (CLAMP_MID) yq = fminf(fmaxf(yq, -128.f), 127.f);      // compile-time: two FMNMX per element that the ViT LayerNorms do not need
    r[e] = __fmaf_rn(yq, f[e], RMAGIC);                              // RNE(yq * f) + RMAGIC, saturated below
  }
  return pack4_sat(r[0], r[1], r[2], r[3]);
}

// Fast path for power-of-two output scales (every minmax-calibrated model): LPR lanes share a row (8, 16 or 32, so a
// lane owns >= 12 channels and the per-row scalar work - three IEEE divisions and a square root - is amortised), the
// per-channel constants live in registers for the whole persistent loop, and every division by a scale is folded into
// them: with ros = 1/out_scale, f = out_scale/(post_div*next_scale) exact powers of two,
//   A  = fl(fl(t*g)*ros)            == fl(t*g'),            g' = g*ros
//   Bv = RNE(fl(fl(b - fl(m*g))*ros)*2^N) == RNE(fl(b' - fl(m*g'))*2^N),  b' = b*ros
//   q  = sat(RNE(((yq*os)/pd)/next))  == sat(RNE(yq*f))
// (scaling by a power of two commutes with rounding), so the codes equal the generic kernel's bit for bit.
// Sums of two integers over each group of LPR lanes, every lane of the warp taking part.  A full warp is one REDUX each; for
// groups of 8 / 16 lanes __reduce_add_sync with the group's own mask compiles to a uniformity test and, the masks differing
// between the groups, a WARPSYNC.COLLECTIVE loop over them - C = 96 ran at 342 Gelement/s against 877 at C = 384 (r2) - so they
// take xor butterflies, which never leave the group.
template <int LPR>
__device__ __forceinline__ void group_sum2(int& a, int& b) {
  if (LPR == 32) {
    a = __reduce_add_sync(0xffffffffu, a);
    b = __reduce_add_sync(0xffffffffu, b);
  } else {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
  }
}
template <int LPR, int WPLN, bool CLAMP_MID, bool GATHER>
__global__ void __launch_bounds__(128, (WPLN <= 3 && !GATHER) ? 4 : 3) layernorm_pot_kernel(p2v_layernorm_args a) {
  constexpr int GPW = 32 / LPR;                       // rows per warp iteration
  const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int row_stride = gridDim.x * (blockDim.x >> 5) * GPW;
  float g[WPLN][4], bt[WPLN][4], f[WPLN][4];
  int sh[WPLN][4];
  const float rnext = fdiv(1.f, a.next_scale);
#pragma unroll
  for (int i = 0; i < WPLN; ++i) {
    const int w = sub + LPR * i;
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.gamma) + w), b4 = __ldg(reinterpret_cast<const float4*>(a.beta) + w);
    const float4 o4 = __ldg(reinterpret_cast<const float4*>(a.out_scale) + w), p4 = __ldg(reinterpret_cast<const float4*>(a.post_div) + w);
    const float4 m4 = __ldg(reinterpret_cast<const float4*>(a.in_mult) + w);
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w}, oo[4] = {o4.x, o4.y, o4.z, o4.w};
    const float pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ros = fdiv(1.f, oo[e]);
      g[i][e] = fmul(gg[e], ros);
      bt[i][e] = fmul(bb[e], ros);
      f[i][e] = fmul(fmul(oo[e], fdiv(1.f, pp[e])), rnext);
      sh[i][e] = int(mm[e]);
    }
  }
  const float Cf = float(a.C), s1 = a.in_scale_min, s1c = fdiv(s1, Cf);
  const float clamp_hi = a.clamp_mid ? 127.f : __int_as_float(0x7f800000);
  float gmin = __int_as_float(0x7f800000), gmax = 0.f;     // bounds of |g'| over the row's channels (NaN-free: fminf / fmaxf drop NaN)
#pragma unroll
  for (int i = 0; i < WPLN; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) { gmin = fminf(gmin, fabsf(g[i][e])); gmax = fmaxf(gmax, fabsf(g[i][e])); }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    gmin = fminf(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
    gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
  }
  pdl_wait();          // the channel constants above are plan-time data; the rows are the previous kernel's output (common.cuh)
  pdl_trigger();
  // gathered input (patch merging; its own instantiation - the plain kernel carries none of this): segment and offset of each
  // of the lane's words are row independent
  int gseg[GATHER ? WPLN : 1], goff[GATHER ? WPLN : 1];
  if (GATHER) {
    const int seg_words = a.C / (4 * a.gather_segs);
#pragma unroll
    for (int i = 0; i < WPLN; ++i) { gseg[i] = (sub + LPR * i) / seg_words; goff[i] = (sub + LPR * i) - gseg[i] * seg_words; }
  }
  // The row's words are loaded one iteration ahead: with three warps per scheduler the wait for them was a third of the
  // kernel's stall samples (ncu r2: long scoreboard on the first unpack).
  auto load_row = [&](int row, uint32_t (&u)[WPLN]) {
    const uint32_t* xr = reinterpret_cast<const uint32_t*>(a.x + int64_t(row) * a.x_row_stride);
#pragma unroll
    for (int i = 0; i < WPLN; ++i)
      u[i] = GATHER ? __ldg(reinterpret_cast<const uint32_t*>(a.x + int64_t(__ldg(a.in_gather + int64_t(row) * a.gather_segs + gseg[GATHER ? i : 0])) * a.x_row_stride) + goff[GATHER ? i : 0])
                    : __ldg(xr + sub + LPR * i);
  };
  // unpack (x * in_mult), exact integer sums over the row, the row scalars t = s1 / std and mos = mean / std (layers.py:316-323)
  auto row_stats = [&](const uint32_t (&u)[WPLN], int (&xv)[WPLN][4], float& t, float& mos) {
    int S1 = 0, S2 = 0;
#pragma unroll
    for (int i = 0; i < WPLN; ++i) {
      xv[i][0] = int(int8_t(u[i] & 0xff)) * sh[i][0];
      xv[i][1] = int(int8_t((u[i] >> 8) & 0xff)) * sh[i][1];
      xv[i][2] = int(int8_t((u[i] >> 16) & 0xff)) * sh[i][2];
      xv[i][3] = int(int8_t(u[i] >> 24)) * sh[i][3];
#pragma unroll
      for (int e = 0; e < 4; ++e) { S1 += xv[i][e]; S2 += xv[i][e] * xv[i][e]; }
    }
    group_sum2<LPR>(S1, S2);
    const float S1f = float(S1), S2f = float(S2);
    const float mean = fmul(fdiv(S1f, Cf), s1);
    const float stdv = fmul(s1c, __fsqrt_rn(fsub(fmul(Cf, S2f), fmul(S1f, S1f))));
    t = fdiv(s1, stdv);
    mos = fdiv(mean, stdv);
  };
  // Fast element loop (ln_pot_fast_word): no conversion / rounding instruction except one FRND.  It needs every |A| = |t*g'| in
  // [2^-24, 2^8) (N = 7 - floor(log2|A|) unclamped) - checked per row against the row-independent bounds of |g'| - and no
  // mantissa of A within 16 ulps below a power of two (floor_log2_as_fp32's corner) - collected by the loop itself.  Rows that
  // fail either test (or have std == 0 / non-finite statistics) are redone with the reference-order code.
  auto finish_row = [&](int row, float t, float mos, const uint32_t (&qw)[WPLN], uint32_t mant_max) {
    uint32_t* orow = reinterpret_cast<uint32_t*>(a.out_i8 + int64_t(a.out_row_map ? __ldg(a.out_row_map + row) : row) * a.C);
    const bool in_range = fmul(t, gmax) < 256.f && fmul(t, gmin) >= 0x1p-24f;
    if (!in_range || mant_max >= 0x007ffff0u) {
      ln_pot_row_slow(a, row, orow, sub, LPR, WPLN, t, mos, clamp_hi);     // rare: out of line, constants re-read from memory
    } else {
#pragma unroll
      for (int i = 0; i < WPLN; ++i) orow[sub + LPR * i] = qw[i];
    }
  };
  // All 32 lanes stay in the loop (the lane groups of a warp's last iteration that have no row left redo the last row and skip
  // the store), so the group sums are full-mask butterflies - see group_sum2.
  const int rb0 = warp_global * GPW;
  if (rb0 >= a.rows) return;
  const int last = a.rows - 1;
  // (Measured and rejected, r2: two rows in flight - the statistics chain of row n + 1 in the same straight-line block as the
  // element loop of row n, 143 registers - 23.3 vs 22.5 us at C = 384 stand-alone, 0.74 vs 0.71 ms per DeiT-S step.)
  uint32_t ucur[WPLN], unext[WPLN];
  load_row(min(rb0 + grp, last), ucur);
  for (int rb = rb0; rb < a.rows; rb += row_stride) {
    const int row = rb + grp;
    if (rb + row_stride < a.rows) load_row(min(rb + row_stride + grp, last), unext);
    int xv[WPLN][4];
    float t, mos;
    row_stats(ucur, xv, t, mos);
#pragma unroll
    for (int i = 0; i < WPLN; ++i) ucur[i] = unext[i];
    uint32_t qw[WPLN];
    uint32_t mant_max = 0;
    if (fmul(t, gmax) < 256.f && fmul(t, gmin) >= 0x1p-24f) {
#pragma unroll
      for (int i = 0; i < WPLN; ++i) qw[i] = ln_pot_fast_word<CLAMP_MID>(t, mos, g[i], bt[i], f[i], xv[i], mant_max);
    }
    if (row <= last) finish_row(row, t, mos, qw, mant_max);
  }
}

// ------------------------------------------------------------------------------------------------
// The same row-persistent kernel for arbitrary fp32 scales (ema / percentile / omse observers): the reference divides by the
// LayerNorm's output scale twice and by the next QAct's scale once per element (layers.py:320-337 and the following QAct,
// uniform.py:83-86).  Each of them is an exactly rounded IEEE quotient here too, but from the divisor's reciprocal, computed once
// per channel: q0 = a * rb, two residual corrections q <- q + (a - b q) * rb with the residual exact in one FFMA.  The first
// correction leaves q within half an ulp (+ 2^-40) of a / b, the second then rounds correctly (Markstein's theorem: rb = RN(1 / b),
// b's significand not all ones; the operands here are far from the exponent range's ends, zero operands give zero).  A channel whose
// scale has an all-ones significand, a post-divisor that is not a power of two, a row outside the fast loop's range (see
// layernorm_pot_kernel) take the reference-order code.  Constants sit in shared memory ([6][C] floats, one LDS.128 per word and
// row): the register-resident form of the power-of-two kernel would need 6 x 24 registers at C = 768.
// ViT-B percentile: 6.8 ms of LayerNorm per 256 images with the generic kernel (r2).
// ------------------------------------------------------------------------------------------------
// one lane's words of a row in the reference's operation order (the generic kernel's element code)
__device__ __noinline__ void ln_np_row_slow(const p2v_layernorm_args& a, int row, uint32_t* __restrict__ orow, int sub, int lpr, int wpln,
                                            float t, float mos) {
  for (int i = 0; i < wpln; ++i) {
    const int w = sub + lpr * i;
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.gamma) + w), b4 = __ldg(reinterpret_cast<const float4*>(a.beta) + w);
    const float4 o4 = __ldg(reinterpret_cast<const float4*>(a.out_scale) + w), p4 = __ldg(reinterpret_cast<const float4*>(a.post_div) + w);
    const float4 m4 = __ldg(reinterpret_cast<const float4*>(a.in_mult) + w);
    const float g[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w}, os[4] = {o4.x, o4.y, o4.z, o4.w};
    const float pd[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w};
    const uint32_t u = __ldg(ln_word_ptr(a, row, w));
    const int cx[4] = {int(int8_t(u & 0xff)), int(int8_t((u >> 8) & 0xff)), int(int8_t((u >> 16) & 0xff)), int(int8_t(u >> 24))};
    int q[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float A = fdiv(fmul(t, g[e]), os[e]);
      const float aA = fabsf(A);
      int N;
      if (aA == 0.f) N = 31;
      else if (!(aA < __int_as_float(0x7f800000))) N = 0;
      else N = min(max(7 - floor_log2_as_fp32(aA), 0), 31);
      const float twoN = pow2i(N);
      const float M = fminf(fmaxf(floorf(fmul(aA, twoN)), 0.f), 255.f);
      const float sgn = A > 0.f ? 1.f : (A < 0.f ? -1.f : 0.f);
      const float Bv = rintf(fmul(fdiv(fsub(be[e], fmul(mos, g[e])), os[e]), twoN));
      const float yq = rintf(fmul(fadd(fmul(fmul(sgn, M), float(cx[e] * int(mm[e]))), Bv), pow2i(-N)));
      const float mid = a.clamp_mid ? fmul(fminf(fmaxf(yq, -128.f), 127.f), os[e]) : fmul(yq, os[e]);
      q[e] = sat_s8(fadd(fdiv(fdiv(mid, pd[e]), a.next_scale), a.next_zp));
    }
    orow[w] = pack4_s8(q[0], q[1], q[2], q[3]);
  }
}

template <int LPR, int WPLN, bool CLAMP_MID>
__global__ void __launch_bounds__(128, 4) layernorm_np_kernel(p2v_layernorm_args a) {
  extern __shared__ float4 ln_np_sm[];                 // [6][C / 4]: gamma, beta, out_scale, 1 / out_scale, 1 / post_div, in_mult
  constexpr int GPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int row_stride = gridDim.x * (blockDim.x >> 5) * GPW;
  const int nw = a.C >> 2;
  bool bad = false;
  for (int w = threadIdx.x; w < nw; w += blockDim.x) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.gamma) + w), o4 = __ldg(reinterpret_cast<const float4*>(a.out_scale) + w);
    const float4 p4 = __ldg(reinterpret_cast<const float4*>(a.post_div) + w);
    const float oo[4] = {o4.x, o4.y, o4.z, o4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
    float ro[4], rp[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ro[e] = fdiv(1.f, oo[e]);
      rp[e] = fdiv(1.f, pp[e]);
      bad |= (__float_as_uint(oo[e]) & 0x007fffffu) == 0x007fffffu || !(oo[e] > 0x1p-60f && oo[e] < 0x1p60f);
      bad |= (__float_as_uint(pp[e]) & 0x007fffffu) != 0u || !(pp[e] > 0x1p-60f && pp[e] < 0x1p60f);
    }
    ln_np_sm[w] = g4;
    ln_np_sm[nw + w] = __ldg(reinterpret_cast<const float4*>(a.beta) + w);
    ln_np_sm[2 * nw + w] = o4;
    ln_np_sm[3 * nw + w] = make_float4(ro[0], ro[1], ro[2], ro[3]);
    ln_np_sm[4 * nw + w] = make_float4(rp[0], rp[1], rp[2], rp[3]);
    ln_np_sm[5 * nw + w] = __ldg(reinterpret_cast<const float4*>(a.in_mult) + w);
  }
  const float next = a.next_scale, rnext = fdiv(1.f, next), zp = a.next_zp;
  bad |= (__float_as_uint(next) & 0x007fffffu) == 0x007fffffu || !(next > 0x1p-60f && next < 0x1p60f);
  bad = __syncthreads_or(bad);
  // row-independent bounds of |g / out_scale| over the lane group's channels, and the integer input multipliers in registers
  int sh[WPLN][4];
  float gmin = __int_as_float(0x7f800000), gmax = 0.f;
#pragma unroll
  for (int i = 0; i < WPLN; ++i) {
    const float4 g4 = ln_np_sm[sub + LPR * i], r4 = ln_np_sm[3 * nw + sub + LPR * i], m4 = ln_np_sm[5 * nw + sub + LPR * i];
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, rr[4] = {r4.x, r4.y, r4.z, r4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float q = fabsf(fmul(gg[e], rr[e]));
      gmin = fminf(gmin, q); gmax = fmaxf(gmax, q);
      sh[i][e] = int(mm[e]);
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    gmin = fminf(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
    gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
  }
  const float Cf = float(a.C), s1 = a.in_scale_min, s1c = fdiv(s1, Cf);
  pdl_wait();
  pdl_trigger();
  const int rb0 = warp_global * GPW;
  if (rb0 >= a.rows) return;
  const int last = a.rows - 1;
  auto load_row = [&](int row, uint32_t (&u)[WPLN]) {
    const uint32_t* xr = reinterpret_cast<const uint32_t*>(a.x + int64_t(row) * a.x_row_stride);
#pragma unroll
    for (int i = 0; i < WPLN; ++i) u[i] = __ldg(xr + sub + LPR * i);
  };
  uint32_t ucur[WPLN], unext[WPLN];
  load_row(min(rb0 + grp, last), ucur);
  for (int rb = rb0; rb < a.rows; rb += row_stride) {      // every lane stays in the loop: see layernorm_pot_kernel
    const int row = min(rb + grp, last);
    const bool live = rb + grp <= last;
    if (rb + row_stride < a.rows) load_row(min(rb + row_stride + grp, last), unext);
    int xv[WPLN][4];
    int S1 = 0, S2 = 0;
#pragma unroll
    for (int i = 0; i < WPLN; ++i) {
      xv[i][0] = int(int8_t(ucur[i] & 0xff)) * sh[i][0];
      xv[i][1] = int(int8_t((ucur[i] >> 8) & 0xff)) * sh[i][1];
      xv[i][2] = int(int8_t((ucur[i] >> 16) & 0xff)) * sh[i][2];
      xv[i][3] = int(int8_t(ucur[i] >> 24)) * sh[i][3];
#pragma unroll
      for (int e = 0; e < 4; ++e) { S1 += xv[i][e]; S2 += xv[i][e] * xv[i][e]; }
      ucur[i] = unext[i];
    }
    group_sum2<LPR>(S1, S2);
    const float S1f = float(S1), S2f = float(S2);
    const float mean = fmul(fdiv(S1f, Cf), s1);
    const float stdv = fmul(s1c, __fsqrt_rn(fsub(fmul(Cf, S2f), fmul(S1f, S1f))));
    const float t = fdiv(s1, stdv);
    const float mos = fdiv(mean, stdv);
    uint32_t* orow = reinterpret_cast<uint32_t*>(a.out_i8 + int64_t(a.out_row_map ? __ldg(a.out_row_map + row) : row) * a.C);
    // |A| = |fl(fl(t g) / os)| within [2^-24, 2^8) for every channel, with a margin for the two roundings
    const bool in_range = !bad && fmul(t, gmax) < 255.99f && fmul(t, gmin) >= 0x1.0002p-24f;
    uint32_t qw[WPLN];
    uint32_t mant_max = 0;
    if (in_range) {
#pragma unroll
      for (int i = 0; i < WPLN; ++i) {
        const int w = sub + LPR * i;
        const float4 g4 = ln_np_sm[w], b4 = ln_np_sm[nw + w], o4 = ln_np_sm[2 * nw + w], r4 = ln_np_sm[3 * nw + w], p4 = ln_np_sm[4 * nw + w];
        const float g[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w}, os[4] = {o4.x, o4.y, o4.z, o4.w};
        const float ro[4] = {r4.x, r4.y, r4.z, r4.w}, rp[4] = {p4.x, p4.y, p4.z, p4.w};
        float r[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t Ab = __float_as_uint(div_rb(fmul(t, g[e]), os[e], ro[e]));
          const uint32_t Ef = Ab & 0x7f800000u;
          mant_max = max(mant_max, Ab & 0x007fffffu);
          const float twoN = __uint_as_float(0x82800000u - Ef);            // 2^(134 - E)
          const float rtwoN = __uint_as_float(Ef - 0x03800000u);           // 2^(E - 134)
          const float sM = __uint_as_float((Ab & 0x807f0000u) | 0x43000000u);
          const float Bv = rintf(fmul(div_rb(fsub(be[e], fmul(mos, g[e])), os[e], ro[e]), twoN));
          const float sum = __fmaf_rn(sM, __int2float_rn(xv[i][e]), Bv);  // sM * x is exact (8 x 11 bits)
          float yq = fsub(__fmaf_rn(sum, rtwoN, RMAGIC), RMAGIC);          // RNE(sum / 2^N)
          if (CLAMP_MID) yq = fminf(fmaxf(yq, -128.f), 127.f);
          const float v = fmul(fmul(yq, os[e]), rp[e]);                    // (yq * os) / post_div, the second factor a power of two
          r[e] = fadd(fadd(div_rb(v, next, rnext), zp), RMAGIC);           // RNE(v / next + zp) + RMAGIC, saturated below
        }
        qw[i] = pack4_sat(r[0], r[1], r[2], r[3]);
      }
    }
    if (!live) continue;
    if (!in_range || mant_max >= 0x007ffff0u) {
      ln_np_row_slow(a, row, orow, sub, LPR, WPLN, t, mos);
    } else {
#pragma unroll
      for (int i = 0; i < WPLN; ++i) orow[sub + LPR * i] = qw[i];
    }
  }
}
template <int LPR, int WPLN>
static void launch_ln_np(const p2v_layernorm_args& a, cudaStream_t stream) {
  constexpr int GPW = 32 / LPR;
  const int rows_per_block = 4 * GPW;
  const size_t smem = size_t(a.C) * 24;
  const int blocks = std::max(1, std::min((a.rows + rows_per_block - 1) / rows_per_block, num_sms() * 4));
  pdl_next_kind(PDL_LAYERNORM);
  if (a.clamp_mid) launch_pdl(layernorm_np_kernel<LPR, WPLN, true>, dim3(blocks), dim3(128), smem, stream, a);
  else launch_pdl(layernorm_np_kernel<LPR, WPLN, false>, dim3(blocks), dim3(128), smem, stream, a);
}

// The power-of-two kernel for rows of 6 to 8 words per lane (C = 768 .. 1024: ViT-B, ViT-L).  Its register-resident channel
// constants (4 x 4 x WPL floats per lane) fill the 170 registers of three blocks per SM at C = 768 and spill at C = 1024 (144 bytes),
// so this form keeps them in shared memory ([4][C / 4] float4: g', b', f, in_mult; three conflict-free LDS.128 per word and row) and
// runs four blocks per SM: C = 1024 262 -> 191 us per 200 k rows, C = 768 40.3 -> 38.0 us per 50 k rows; below that the registers win
// (C = 512 30.2 vs 34.9 us, C = 384 22.3 vs 27.0 us).  Same arithmetic, same helper functions, same slow path as layernorm_pot_kernel.
template <int WPLN, bool CLAMP_MID, bool GATHER = false>
__global__ void __launch_bounds__(128, WPLN <= 8 ? 4 : 3) layernorm_pot_smem_kernel(p2v_layernorm_args a) {
  extern __shared__ float4 ln_ps_sm[];
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int row_stride = gridDim.x * (blockDim.x >> 5);
  const int nw = a.C >> 2;
  const float rnext = fdiv(1.f, a.next_scale);
  for (int w = threadIdx.x; w < nw; w += blockDim.x) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.gamma) + w), b4 = __ldg(reinterpret_cast<const float4*>(a.beta) + w);
    const float4 o4 = __ldg(reinterpret_cast<const float4*>(a.out_scale) + w), p4 = __ldg(reinterpret_cast<const float4*>(a.post_div) + w);
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w}, oo[4] = {o4.x, o4.y, o4.z, o4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
    float g[4], bt[4], f[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ros = fdiv(1.f, oo[e]);
      g[e] = fmul(gg[e], ros);
      bt[e] = fmul(bb[e], ros);
      f[e] = fmul(fmul(oo[e], fdiv(1.f, pp[e])), rnext);
    }
    ln_ps_sm[w] = make_float4(g[0], g[1], g[2], g[3]);
    ln_ps_sm[nw + w] = make_float4(bt[0], bt[1], bt[2], bt[3]);
    ln_ps_sm[2 * nw + w] = make_float4(f[0], f[1], f[2], f[3]);
    ln_ps_sm[3 * nw + w] = __ldg(reinterpret_cast<const float4*>(a.in_mult) + w);
  }
  __syncthreads();
  int sh[WPLN][4];
  float gmin = __int_as_float(0x7f800000), gmax = 0.f;
#pragma unroll
  for (int i = 0; i < WPLN; ++i) {
    const float4 g4 = ln_ps_sm[lane + 32 * i], m4 = ln_ps_sm[3 * nw + lane + 32 * i];
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      gmin = fminf(gmin, fabsf(gg[e])); gmax = fmaxf(gmax, fabsf(gg[e]));
      sh[i][e] = int(mm[e]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gmin = fminf(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
    gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
  }
  const float Cf = float(a.C), s1 = a.in_scale_min, s1c = fdiv(s1, Cf);
  const float clamp_hi = a.clamp_mid ? 127.f : __int_as_float(0x7f800000);
  pdl_wait();
  pdl_trigger();
  if (warp_global >= a.rows) return;
  // gathered input (patch merging): segment and offset of each of the lane's words are row independent
  int gseg[GATHER ? WPLN : 1], goff[GATHER ? WPLN : 1];
  if (GATHER) {
    const int seg_words = a.C / (4 * a.gather_segs);
#pragma unroll
    for (int i = 0; i < WPLN; ++i) { gseg[i] = (lane + 32 * i) / seg_words; goff[i] = (lane + 32 * i) - gseg[i] * seg_words; }
  }
  auto load_row = [&](int row, uint32_t (&u)[WPLN]) {
    const uint32_t* xr = reinterpret_cast<const uint32_t*>(a.x + int64_t(row) * a.x_row_stride);
#pragma unroll
    for (int i = 0; i < WPLN; ++i)
      u[i] = GATHER ? __ldg(reinterpret_cast<const uint32_t*>(a.x + int64_t(__ldg(a.in_gather + int64_t(row) * a.gather_segs + gseg[GATHER ? i : 0])) * a.x_row_stride) + goff[GATHER ? i : 0])
                    : __ldg(xr + lane + 32 * i);
  };
  uint32_t ucur[WPLN], unext[WPLN];
  load_row(warp_global, ucur);
  for (int row = warp_global; row < a.rows; row += row_stride) {
    if (row + row_stride < a.rows) load_row(row + row_stride, unext);
    int xv[WPLN][4];
    int S1 = 0, S2 = 0;
#pragma unroll
    for (int i = 0; i < WPLN; ++i) {
      xv[i][0] = int(int8_t(ucur[i] & 0xff)) * sh[i][0];
      xv[i][1] = int(int8_t((ucur[i] >> 8) & 0xff)) * sh[i][1];
      xv[i][2] = int(int8_t((ucur[i] >> 16) & 0xff)) * sh[i][2];
      xv[i][3] = int(int8_t(ucur[i] >> 24)) * sh[i][3];
#pragma unroll
      for (int e = 0; e < 4; ++e) { S1 += xv[i][e]; S2 += xv[i][e] * xv[i][e]; }
      ucur[i] = unext[i];
    }
    group_sum2<32>(S1, S2);
    const float S1f = float(S1), S2f = float(S2);
    const float mean = fmul(fdiv(S1f, Cf), s1);
    const float stdv = fmul(s1c, __fsqrt_rn(fsub(fmul(Cf, S2f), fmul(S1f, S1f))));
    const float t = fdiv(s1, stdv);
    const float mos = fdiv(mean, stdv);
    uint32_t* orow = reinterpret_cast<uint32_t*>(a.out_i8 + int64_t(a.out_row_map ? __ldg(a.out_row_map + row) : row) * a.C);
    const bool in_range = fmul(t, gmax) < 256.f && fmul(t, gmin) >= 0x1p-24f;
    uint32_t qw[WPLN];
    uint32_t mant_max = 0;
    if (in_range) {
#pragma unroll
      for (int i = 0; i < WPLN; ++i) {
        const int w = lane + 32 * i;
        const float4 g4 = ln_ps_sm[w], b4 = ln_ps_sm[nw + w], f4 = ln_ps_sm[2 * nw + w];
        const float g[4] = {g4.x, g4.y, g4.z, g4.w}, bt[4] = {b4.x, b4.y, b4.z, b4.w}, f[4] = {f4.x, f4.y, f4.z, f4.w};
        qw[i] = ln_pot_fast_word<CLAMP_MID>(t, mos, g, bt, f, xv[i], mant_max);
      }
    }
    if (!in_range || mant_max >= 0x007ffff0u) {
      ln_pot_row_slow(a, row, orow, lane, 32, WPLN, t, mos, clamp_hi);
    } else {
#pragma unroll
      for (int i = 0; i < WPLN; ++i) orow[lane + 32 * i] = qw[i];
    }
  }
}
template <int WPLN>
static void launch_ln_pot_smem(const p2v_layernorm_args& a, cudaStream_t stream) {
  const size_t smem = size_t(a.C) * 16;
  const int blocks = std::max(1, std::min((a.rows + 3) / 4, num_sms() * 4));
  pdl_next_kind(PDL_LAYERNORM);
  if (a.clamp_mid) launch_pdl(layernorm_pot_smem_kernel<WPLN, true>, dim3(blocks), dim3(128), smem, stream, a);
  else launch_pdl(layernorm_pot_smem_kernel<WPLN, false>, dim3(blocks), dim3(128), smem, stream, a);
}
// patch-merging form (gathered rows, no clamp): the wide merges (4C = 768 .. 1536)
template <int WPLN>
static void launch_ln_pot_smem_gather(const p2v_layernorm_args& a, cudaStream_t stream) {
  const size_t smem = size_t(a.C) * 16;
  const int blocks = std::max(1, std::min((a.rows + 3) / 4, num_sms() * (WPLN <= 8 ? 4 : 3)));
  pdl_next_kind(PDL_LAYERNORM);
  launch_pdl(layernorm_pot_smem_kernel<WPLN, false, true>, dim3(blocks), dim3(128), smem, stream, a);
}

template <int LPR, int WPLN, bool CLAMP_MID, bool GATHER = false>
static void launch_ln_pot_c(const p2v_layernorm_args& a, cudaStream_t stream) {
  constexpr int GPW = 32 / LPR;
  const int rows_per_block = 4 * GPW;
  // persistent: several rows per lane group so the register-resident channel constants are amortised; one wave of resident blocks
  static int occ = 0;
  if (!occ) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, layernorm_pot_kernel<LPR, WPLN, CLAMP_MID, GATHER>, 128, 0) != cudaSuccess || occ < 1) occ = 3;
  }
  const int blocks = std::max(1, std::min((a.rows + rows_per_block - 1) / rows_per_block, num_sms() * occ));
  pdl_next_kind(PDL_LAYERNORM);
  launch_pdl(layernorm_pot_kernel<LPR, WPLN, CLAMP_MID, GATHER>, dim3(blocks), dim3(128), 0, stream, a);
}
template <int LPR, int WPLN>
static void launch_ln_pot(const p2v_layernorm_args& a, cudaStream_t stream) {
  if (a.clamp_mid) launch_ln_pot_c<LPR, WPLN, true>(a, stream);
  else launch_ln_pot_c<LPR, WPLN, false>(a, stream);
}

int launch_layernorm(const p2v_layernorm_args& a, cudaStream_t stream) {
  const int nwords = a.C / 4;
  if (a.in_gather && a.pot_scales && a.out_i8 && !a.clamp_mid && nwords % 32 == 0 && nwords / 32 >= 3 && nwords / 32 <= 12) {
    // patch-merging LayerNorm (4C = 384 / 768 / 1536 for Swin-T/S, 512 / 1024 / 2048 for Swin-B): gathered instantiations
    bool done = true;
    static const bool gather_smem = !(getenv("P2V_LN_SMEM_MIN") && atoi(getenv("P2V_LN_SMEM_MIN")) > 8);      // triage: 9 = register form everywhere
    switch (nwords / 32) {
      case 3: launch_ln_pot_c<32, 3, false, true>(a, stream); break;
      case 4: launch_ln_pot_c<32, 4, false, true>(a, stream); break;
      case 6: if (gather_smem) launch_ln_pot_smem_gather<6>(a, stream); else launch_ln_pot_c<32, 6, false, true>(a, stream); break;
      case 8: if (gather_smem) launch_ln_pot_smem_gather<8>(a, stream); else launch_ln_pot_c<32, 8, false, true>(a, stream); break;
      case 12: if (gather_smem) launch_ln_pot_smem_gather<12>(a, stream); else launch_ln_pot_c<32, 12, false, true>(a, stream); break;
      default: done = false;
    }
    if (done) {
      count_launch();
      return check_launch("layernorm_int");
    }
  }
  if (a.pot_scales && a.out_i8 && !a.out_f32 && !a.in_gather) {
    bool done = true;
    // lanes per row: as many as leave a lane >= 12 channels (3 words) - more rows in flight per SM beat the amortisation of
    // the per-row scalar work (C = 384: 32 lanes x 3 words 26.8 us vs 16 x 6 32.9 us for 50 k rows, tools/ln_bench.py)
#define P2V_LN_POT(LPR_, N_) case N_: launch_ln_pot<LPR_, N_>(a, stream); break;
    static const int smem_min = getenv("P2V_LN_SMEM_MIN") ? atoi(getenv("P2V_LN_SMEM_MIN")) : 6;       // triage; measured: 768 38.0 vs 40.3 us, 640 equal, 512 / 384 slower
    if (nwords % 32 == 0 && nwords / 32 >= smem_min && nwords / 32 >= 3 && nwords / 32 <= 8) {
      switch (nwords / 32) {
        case 3: launch_ln_pot_smem<3>(a, stream); break;
        case 4: launch_ln_pot_smem<4>(a, stream); break;
        case 5: launch_ln_pot_smem<5>(a, stream); break;
        case 6: launch_ln_pot_smem<6>(a, stream); break;
        case 7: launch_ln_pot_smem<7>(a, stream); break;
        default: launch_ln_pot_smem<8>(a, stream); break;
      }
    } else if (nwords % 32 == 0 && nwords / 32 >= 3 && nwords / 32 <= 8) {
      switch (nwords / 32) { P2V_LN_POT(32, 3) P2V_LN_POT(32, 4) P2V_LN_POT(32, 5) P2V_LN_POT(32, 6) P2V_LN_POT(32, 7) P2V_LN_POT(32, 8) }
    } else if (nwords % 16 == 0 && nwords / 16 >= 3 && nwords / 16 <= 6 && a.rows >= 2) {
      switch (nwords / 16) { P2V_LN_POT(16, 3) P2V_LN_POT(16, 4) P2V_LN_POT(16, 5) P2V_LN_POT(16, 6) }
    } else if (nwords % 8 == 0 && nwords / 8 >= 3 && nwords / 8 <= 6 && a.rows >= 4) {
      switch (nwords / 8) { P2V_LN_POT(8, 3) P2V_LN_POT(8, 4) P2V_LN_POT(8, 5) P2V_LN_POT(8, 6) }
    } else if (nwords % 32 == 0 && nwords / 32 <= 2) {
      switch (nwords / 32) { P2V_LN_POT(32, 1) P2V_LN_POT(32, 2) }
#undef P2V_LN_POT
    } else {
      done = false;
    }
    if (done) {
      count_launch();
      return check_launch("layernorm_int");
    }
  }
  static const bool np_off = getenv("P2V_LN_NP") && atoi(getenv("P2V_LN_NP")) == 0;     // triage: generic kernel for non-power-of-two scales
  if (!a.pot_scales && a.out_i8 && !a.out_f32 && !a.in_gather && a.C <= 2048 && !np_off) {
    bool done = true;
#define P2V_LN_NP(LPR_, N_) case N_: launch_ln_np<LPR_, N_>(a, stream); break;
    if (nwords % 32 == 0 && nwords / 32 >= 3 && nwords / 32 <= 8) {
      switch (nwords / 32) { P2V_LN_NP(32, 3) P2V_LN_NP(32, 4) P2V_LN_NP(32, 5) P2V_LN_NP(32, 6) P2V_LN_NP(32, 7) P2V_LN_NP(32, 8) }
    } else if (nwords % 16 == 0 && nwords / 16 >= 3 && nwords / 16 <= 6 && a.rows >= 2) {
      switch (nwords / 16) { P2V_LN_NP(16, 3) P2V_LN_NP(16, 4) P2V_LN_NP(16, 5) P2V_LN_NP(16, 6) }
    } else if (nwords % 8 == 0 && nwords / 8 >= 3 && nwords / 8 <= 6 && a.rows >= 4) {
      switch (nwords / 8) { P2V_LN_NP(8, 3) P2V_LN_NP(8, 4) P2V_LN_NP(8, 5) P2V_LN_NP(8, 6) }
    } else if (nwords % 32 == 0 && nwords / 32 <= 2) {
      switch (nwords / 32) { P2V_LN_NP(32, 1) P2V_LN_NP(32, 2) }
#undef P2V_LN_NP
    } else {
      done = false;
    }
    if (done) {
      count_launch();
      return check_launch("layernorm_int");
    }
  }
  const int wpl = (a.C / 4 + 31) / 32;
  const int blocks = std::min((a.rows + 7) / 8, num_sms() * 8);
#define P2V_LN(W)                                                                   \
  if (a.pot_scales) launch_pdl(layernorm_kernel<W, true>, dim3(blocks), dim3(256), 0, stream, a);       \
  else launch_pdl(layernorm_kernel<W, false>, dim3(blocks), dim3(256), 0, stream, a);
  if (wpl <= 1) { P2V_LN(1) } else if (wpl <= 2) { P2V_LN(2) } else if (wpl <= 3) { P2V_LN(3) } else if (wpl <= 4) { P2V_LN(4) }
  else if (wpl <= 6) { P2V_LN(6) } else if (wpl <= 8) { P2V_LN(8) } else if (wpl <= 12) { P2V_LN(12) }
  else if (wpl <= 16) { P2V_LN(16) } else { P2V_LN(32) }
#undef P2V_LN
  count_launch();
  return check_launch("layernorm_int");
}

// ------------------------------------------------------------------------------------------------
// Stand-alone QIntSoftmax on int8 score codes (one warp per row)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_kernel(const int8_t* __restrict__ scores, uint8_t* __restrict__ out, int64_t rows,
                                                      int n, const p2v_softmax_lut* __restrict__ lut) {
  __shared__ uint32_t s_hi[256], s_lo[256];
  __shared__ float s_e[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) { s_hi[i] = lut->hi[i]; s_lo[i] = lut->lo[i]; s_e[i] = lut->exp_f32[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t row = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); row < rows; row += int64_t(gridDim.x) * wpb) {
    const int8_t* r = scores + row * n;
    int mx = -128;
    for (int j = lane; j < n; j += 32) mx = max(mx, int(r[j]));
    mx = __reduce_max_sync(0xffffffffu, mx);
    unsigned long long hi = 0, lo = 0;
    for (int j = lane; j < n; j += 32) { const int d = mx - int(r[j]); hi += s_hi[d]; lo += s_lo[d]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { hi += __shfl_xor_sync(0xffffffffu, hi, o); lo += __shfl_xor_sync(0xffffffffu, lo, o); }
    const float tot = u96_to_f32(hi, lo);
    for (int j = lane; j < n; j += 32) out[row * n + j] = uint8_t(log2_code(tot, s_e[mx - int(r[j])]));
  }
}

int launch_softmax(const int8_t* scores, uint8_t* out, int64_t rows, int n, const p2v_softmax_lut* lut, cudaStream_t stream) {
  const int blocks = int(std::min<int64_t>((rows + 7) / 8, int64_t(num_sms()) * 8));
  softmax_kernel<<<blocks, 256, 0, stream>>>(scores, out, rows, n, lut);
  count_launch();
  return check_launch("int_softmax_log2");
}

// ------------------------------------------------------------------------------------------------
// Calibration reductions
// ------------------------------------------------------------------------------------------------
// x viewed as [outer, C, inner]; block b handles a slab of `outer` (inner == 1: rows) and writes partial
// min/max per channel; the second kernel folds the partials.
__global__ void __launch_bounds__(256) minmax_partial_kernel(const float* __restrict__ x, float* __restrict__ part, int64_t outer,
                                                             int C, int64_t inner, int64_t outer_per_block) {
  const int64_t o0 = int64_t(blockIdx.x) * outer_per_block;
  const int64_t o1 = min(outer, o0 + outer_per_block);
  float* pmin = part + size_t(blockIdx.x) * 2 * C;
  float* pmax = pmin + C;
  if (inner == 1) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float mn = __int_as_float(0x7f800000), mx = -mn;
      for (int64_t o = o0; o < o1; ++o) { const float v = __ldg(x + o * C + c); mn = fminf(mn, v); mx = fmaxf(mx, v); }
      pmin[c] = mn; pmax[c] = mx;
    }
  } else {
    __shared__ float smn[256], smx[256];
    for (int c = 0; c < C; ++c) {
      float mn = __int_as_float(0x7f800000), mx = -mn;
      for (int64_t o = o0; o < o1; ++o) {
        const float* p = x + (o * C + c) * inner;
        for (int64_t i = threadIdx.x; i < inner; i += blockDim.x) { const float v = __ldg(p + i); mn = fminf(mn, v); mx = fmaxf(mx, v); }
      }
      smn[threadIdx.x] = mn; smx[threadIdx.x] = mx;
      __syncthreads();
      for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { smn[threadIdx.x] = fminf(smn[threadIdx.x], smn[threadIdx.x + s]); smx[threadIdx.x] = fmaxf(smx[threadIdx.x], smx[threadIdx.x + s]); }
        __syncthreads();
      }
      if (threadIdx.x == 0) { pmin[c] = smn[0]; pmax[c] = smx[0]; }
      __syncthreads();
    }
  }
}

__global__ void minmax_final_kernel(const float* __restrict__ part, float* __restrict__ out, int nblocks, int C) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    float mn = __int_as_float(0x7f800000), mx = -mn;
    for (int b = 0; b < nblocks; ++b) { mn = fminf(mn, part[size_t(b) * 2 * C + c]); mx = fmaxf(mx, part[size_t(b) * 2 * C + C + c]); }
    out[c] = mn; out[C + c] = mx;
  }
}

static int minmax_blocks(int64_t outer, int64_t& opb) {
  int nblocks = int(std::min<int64_t>(outer, int64_t(num_sms()) * 4));
  if (nblocks < 1) nblocks = 1;
  opb = (outer + nblocks - 1) / nblocks;
  return int((outer + opb - 1) / opb);
}
// bytes of caller-owned scratch (block partials): the library keeps no device memory of its own, so concurrent streams /
// devices cannot share a buffer by accident
int64_t minmax_scratch_bytes(int64_t n, int C, int64_t inner) {
  int64_t opb;
  return int64_t(minmax_blocks(n / (int64_t(C) * inner), opb)) * 2 * C * int64_t(sizeof(float));
}

int launch_minmax(const float* x, float* minmax, int64_t n, int C, int64_t inner, float* part, cudaStream_t stream) {
  const int64_t outer = n / (int64_t(C) * inner);
  int64_t opb;
  const int nblocks = minmax_blocks(outer, opb);
  minmax_partial_kernel<<<nblocks, 256, 0, stream>>>(x, part, outer, C, inner, opb);
  minmax_final_kernel<<<(C + 255) / 256, 256, 0, stream>>>(part, minmax, nblocks, C);
  count_launch(2);
  return check_launch("minmax_per_channel");
}

// sum (x - fq_k(x))^2 for K candidate scales.  Thread t of a block owns channel positions t, t+256, ...
// of a slab of rows (inner == 1), so per-channel sums need no cross-thread reduction.
constexpr int MSE_KG = 8;
__global__ void __launch_bounds__(256) mse_scores_kernel(const float* __restrict__ x, int64_t outer, int C, int64_t inner,
                                                         const float* __restrict__ scales, const float* __restrict__ zps, int K,
                                                         int n_scale, int per_channel_out, float lo, float hi,
                                                         double* __restrict__ part, int64_t outer_per_block) {
  // part: [gridDim.x, K, n_out] block partials, folded in block order by mse_final_kernel: the scores (and so the argmin of
  // near-tied candidates) are bit-reproducible run to run, unlike floating-point atomics
  const int k0 = blockIdx.y * MSE_KG;
  const int kn = min(MSE_KG, K - k0);
  const int64_t o0 = int64_t(blockIdx.x) * outer_per_block;
  const int64_t o1 = min(outer, o0 + outer_per_block);
  const int n_out = per_channel_out ? C : 1;
  double tot[MSE_KG];
#pragma unroll
  for (int k = 0; k < MSE_KG; ++k) tot[k] = 0.0;
  const int64_t per_row = int64_t(C) * inner;
  for (int64_t e = threadIdx.x; e < per_row; e += blockDim.x) {
    const int c = int(e / inner);
    float acc[MSE_KG];
#pragma unroll
    for (int k = 0; k < MSE_KG; ++k) acc[k] = 0.f;
    for (int64_t o = o0; o < o1; ++o) {
      const float v = __ldg(x + o * per_row + e);
#pragma unroll
      for (int k = 0; k < MSE_KG; ++k) {
        if (k < kn) {
          const float s = __ldg(scales + size_t(k0 + k) * n_scale + (n_scale == 1 ? 0 : c));
          const float zp = zps ? __ldg(zps + size_t(k0 + k) * n_scale + (n_scale == 1 ? 0 : c)) : 0.f;
          const float qv = fminf(fmaxf(rintf(fadd(fdiv(v, s), zp)), lo), hi);
          const float d = fsub(v, fmul(fsub(qv, zp), s));
          acc[k] = fadd(acc[k], fmul(d, d));
        }
      }
    }
    if (per_channel_out) {
#pragma unroll
      for (int k = 0; k < MSE_KG; ++k) if (k < kn) part[(size_t(blockIdx.x) * K + k0 + k) * n_out + c] = double(acc[k]);   // inner == 1: e == c, every (block, k, c) is written exactly once
    } else {
#pragma unroll
      for (int k = 0; k < MSE_KG; ++k) tot[k] += double(acc[k]);
    }
  }
  if (!per_channel_out) {
    __shared__ double sh[256];
    for (int k = 0; k < kn; ++k) {
      sh[threadIdx.x] = tot[k];
      __syncthreads();
      for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s]; __syncthreads(); }
      if (threadIdx.x == 0) part[size_t(blockIdx.x) * K + k0 + k] = sh[0];
      __syncthreads();
    }
  }
}

__global__ void mse_final_kernel(const double* __restrict__ part, double* __restrict__ out, int nblocks, int64_t per_block) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < per_block; i += int64_t(gridDim.x) * blockDim.x) {
    double t = 0.0;
    for (int b = 0; b < nblocks; ++b) t += part[size_t(b) * per_block + i];
    out[i] = t;
  }
}

static int mse_blocks(int64_t outer, int64_t& opb) {
  int nblocks = int(std::min<int64_t>(outer, int64_t(num_sms()) * 2));
  if (nblocks < 1) nblocks = 1;
  opb = (outer + nblocks - 1) / nblocks;
  return int((outer + opb - 1) / opb);
}
int64_t mse_scratch_bytes(int64_t n, int C, int64_t inner, int K, int per_channel_out) {
  int64_t opb;
  return int64_t(mse_blocks(n / (int64_t(C) * inner), opb)) * K * (per_channel_out ? C : 1) * int64_t(sizeof(double));
}

int launch_mse_scores(const float* x, int64_t n, int C, int64_t inner, const float* scales, const float* zps, int K, int n_scale,
                      int per_channel_out, int lo, int hi, double* out, double* part, cudaStream_t stream) {
  const int64_t outer = n / (int64_t(C) * inner);
  const int n_out = per_channel_out ? C : 1;
  int64_t opb;
  const int nblocks = mse_blocks(outer, opb);
  P2V_REQUIRE(!(per_channel_out && inner != 1), "mse_scores: per-channel scores need channels-last data (inner == 1)");
  dim3 grid(nblocks, (K + MSE_KG - 1) / MSE_KG);
  mse_scores_kernel<<<grid, 256, 0, stream>>>(x, outer, C, inner, scales, zps, K, n_scale, per_channel_out, float(lo), float(hi), part, opb);
  const int64_t per_block = int64_t(K) * n_out;
  mse_final_kernel<<<int(std::min<int64_t>((per_block + 255) / 256, 1024)), 256, 0, stream>>>(part, out, nblocks, per_block);
  count_launch(2);
  return check_launch("quant_mse_scores");
}

// ------------------------------------------------------------------------------------------------
// Order statistics for the percentile observer (observer/percentile.py:26-55): one pass of a most-significant-digit radix
// select.  Keys are the floats' bit patterns mapped monotonically to unsigned integers; the pass counts, among the elements
// whose key matches `prefix_value` under `prefix_mask`, the digit (key >> shift) & (2^nbits - 1).  Counts are integers, so the
// histograms of the ranks of a data-parallel calibration add up exactly (all-reduce SUM) and three passes (12 + 12 + 8 bits)
// pin down the global k-th smallest element - the same value a single process would find on the concatenated batch.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_order_key(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void __launch_bounds__(512) radix_hist_kernel(const float* __restrict__ x, int64_t n, uint32_t prefix_mask, uint32_t prefix_value,
                                                         int shift, int nbits, unsigned long long* __restrict__ hist) {
  extern __shared__ uint32_t sh_hist[];
  const int nb = 1 << nbits;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) sh_hist[i] = 0u;
  __syncthreads();
  const uint32_t dmask = uint32_t(nb - 1);
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const int64_t n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += stride) {       // 16-byte loads
    const float4 v = __ldg(x4 + i);
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t k = float_order_key(vv[e]);
      if ((k & prefix_mask) == prefix_value) atomicAdd(&sh_hist[(k >> shift) & dmask], 1u);
    }
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += stride) {
    const uint32_t k = float_order_key(__ldg(x + i));
    if ((k & prefix_mask) == prefix_value) atomicAdd(&sh_hist[(k >> shift) & dmask], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x)
    if (sh_hist[i]) atomicAdd(hist + i, static_cast<unsigned long long>(sh_hist[i]));
}

int launch_radix_hist(const float* x, int64_t n, uint32_t prefix_mask, uint32_t prefix_value, int shift, int nbits,
                      unsigned long long* hist, cudaStream_t stream) {
  const int64_t per_block = 512 * 4 * 8;        // a block's smem counters are 32 bit: 16 K elements per block and grid-stride pass stay far below 2^32
  int grid = int(std::min<int64_t>((n + per_block - 1) / per_block, int64_t(num_sms()) * 4));
  if (grid < 1) grid = 1;
  radix_hist_kernel<<<grid, 512, size_t(4) << nbits, stream>>>(x, n, prefix_mask, prefix_value, shift, nbits, hist);
  count_launch();
  return check_launch("radix_hist_f32");
}

}  // namespace p2v
