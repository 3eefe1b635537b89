// fp32 GEMM on the CUDA cores for the two places of the path whose operands are genuinely fp32 (nothing to put on the int8
// tensor pipe, and TF32 would change the reference's arithmetic):
//
//   * calibration - the weight power-of-two search of MinmaxObserver (observer/minmax.py:145-207).  The reference re-runs the
//     layer per output channel and candidate; its score is  sum_rows (layer(x; W)[., j] - layer(x; fq_k(W))[., j])^2 =
//     sum_rows (x . (W - fq_k(W))[j, :])^2.  p2v_linear_sqerr_scores takes the stacked difference rows D [n, K] (all candidates
//     of a layer in one launch) and returns, per row of D, the column sum of squares of x D^T - the [rows, n] product never
//     exists in memory: a block reduces its 128 x 128 tile to 128 partial sums, a second kernel folds the row blocks in a fixed
//     order (bit-reproducible scores);
//   * the ViT-Large stem (vit_fquant.py:1063: no input quantizer, so fp32 pixels meet fake-quantized weights - an fp32 GEMM in
//     the reference too): p2v_embed_f32 = patch gather + GEMM + the EMBED epilogue chain of include/p2vit_b200.h, int8 out.
//
// One kernel body: 128 x 128 x 16 tiles, 256 threads, 8 x 8 accumulators per thread, operands staged transposed in shared
// memory (conflict-free 16-byte reads), next tile's global loads in flight under the current tile's FMAs.  The k loop runs in
// ascending order with one FMA chain per accumulator, so the result is deterministic (no split-k, no atomics).
#include <algorithm>
#include "epilogue.cuh"

namespace p2v {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16, SG_THREADS = 256;

// A: row r of the activation matrix.  PATCH > 0: rows are patches of an NCHW image (k = c * P * P + py * P + px), gathered on the fly.
struct SgemmA {
  const float* x;
  int M, K;
  int patch, Cin, H, W;      // patch == 0: plain [M, K] row-major
};
__device__ __forceinline__ float4 sg_load_a(const SgemmA& a, int row, int k) {
  if (row >= a.M || k >= a.K) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.patch == 0) return __ldg(reinterpret_cast<const float4*>(a.x + size_t(row) * a.K + k));
  const int P = a.patch, gw = a.W / P, gh = a.H / P;
  const int b = row / (gw * gh), t = row % (gw * gh), ty = t / gw, tx = t % gw;
  const int c = k / (P * P), rem = k % (P * P), py = rem / P, px = rem % P;      // px % 4 == 0 (k % 4 == 0, P % 4 == 0)
  return __ldg(reinterpret_cast<const float4*>(a.x + ((size_t(b) * a.Cin + c) * a.H + ty * P + py) * a.W + tx * P + px));
}

struct SgemmEmbed {      // EMBED epilogue on an fp32 accumulator (include/p2vit_b200.h: P2V_EPI_EMBED), zero points included
  const float* bias;         // (MODE 2: bias or NULL, out_f32 the [M, N] result)
  float* out_f32;
  const float* pos;          // [(T+1), N]
  const float* out_scale;    // [N]
  float mid_scale, mid_zp, aux_scale, aux_zp;
  int tokens_per_image;
  int8_t* out;               // [B*(T+1), N]
};

template <int MODE>      // 0: column sums of squares -> part[blockIdx.y][n];  1: EMBED epilogue;  2: out_f32 = acc + bias
__global__ void __launch_bounds__(SG_THREADS) sgemm_kernel(SgemmA a, const float* __restrict__ Wt, int N, double* __restrict__ part, SgemmEmbed ep) {
  // operand tiles, and - after the k loop - the MODE 0 reduction buffer over the same bytes
  __shared__ __align__(16) unsigned char sg_smem[2 * 2 * SG_BK * (SG_BM + 4) * 4];
  float (*As)[SG_BK][SG_BM + 4] = reinterpret_cast<float (*)[SG_BK][SG_BM + 4]>(sg_smem);
  float (*Bs)[SG_BK][SG_BN + 4] = reinterpret_cast<float (*)[SG_BK][SG_BN + 4]>(sg_smem + 2 * SG_BK * (SG_BM + 4) * 4);
  static_assert(SG_BM == SG_BN && sizeof(double) * 16 * SG_BN <= sizeof(sg_smem), "reduction buffer must fit the operand tiles");
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  // global -> shared: 128 rows x 16 k per operand = 512 float4; thread t loads rows (t >> 2) and (t >> 2) + 64, k offset (t & 3) * 4
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      ra[h] = sg_load_a(a, m0 + lr + 64 * h, k0 + lk);
      const int n = n0 + lr + 64 * h;
      rb[h] = (n < N && k0 + lk < a.K) ? __ldg(reinterpret_cast<const float4*>(Wt + size_t(n) * a.K + k0 + lk)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + 64 * h;
      As[buf][lk][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
      Bs[buf][lk][r] = rb[h].x; Bs[buf][lk + 1][r] = rb[h].y; Bs[buf][lk + 2][r] = rb[h].z; Bs[buf][lk + 3][r] = rb[h].w;
    }
  };
  const int nk = (a.K + SG_BK - 1) / SG_BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) gload((kb + 1) * SG_BK);
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      // thread (ty, tx): rows ty*4..+3 and 64+ty*4..+3, columns tx*4..+3 and 64+tx*4..+3 (two float4 per operand, conflict free)
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]), a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]), b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
    }
    if (kb + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }
  if (MODE == 0) {
    // column sums of squares over the tile's rows (rows >= M contributed zeros): thread partials -> shared -> one double per column
    double (*red)[SG_BN] = reinterpret_cast<double (*)[SG_BN]>(sg_smem);       // the k loop ended with a __syncthreads
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += double(acc[i][j]) * double(acc[i][j]);
      red[ty][(j < 4 ? 0 : 64) + tx * 4 + (j & 3)] = s;
    }
    __syncthreads();
    if (tid < SG_BN) {
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < 16; ++r) s += red[r][tid];
      if (n0 + tid < N) part[size_t(blockIdx.y) * N + n0 + tid] = s;
    }
  } else if (MODE == 2) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = m0 + (i < 4 ? 0 : 64) + ty * 4 + (i & 3);
      if (row >= a.M) continue;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        const int n = n0 + jh * 64 + tx * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (n + e < N) ep.out_f32[size_t(row) * N + n + e] = ep.bias ? fadd(acc[i][jh * 4 + e], __ldg(ep.bias + n + e)) : acc[i][jh * 4 + e];
      }
    }
  } else {
    const float e_rsm = fdiv(1.f, ep.mid_scale), e_raux = fdiv(1.f, ep.aux_scale);
    const int T = ep.tokens_per_image;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = m0 + (i < 4 ? 0 : 64) + ty * 4 + (i & 3);
      if (row >= a.M) continue;
      const int tok = row % T;
      const size_t orow = size_t(row / T) * (T + 1) + tok + 1;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        const int n = n0 + jh * 64 + tx * 4;
        if (n >= N) continue;       // N % 4 == 0 (host check)
        int q[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float y = fadd(acc[i][jh * 4 + e], __ldg(ep.bias + n + e));
          bool slow = false;       // EXACT = true below: the IEEE divisions, as the reference performs them
          const float c = quant_div<true>(y, ep.mid_scale, e_rsm, slow, ep.mid_zp);
          const float ecode = quant_div<true>(fmul(fsub(c, ep.mid_zp), ep.mid_scale), ep.aux_scale, e_raux, slow, ep.aux_zp);
          const float v = fadd(fmul(fsub(ecode, ep.aux_zp), ep.aux_scale), __ldg(ep.pos + size_t(tok + 1) * N + n + e));
          const float o = __ldg(ep.out_scale + n + e);
          q[e] = quant_div_s8<true>(v, o, 0.f, slow);
        }
        *reinterpret_cast<uint32_t*>(ep.out + orow * N + n) = pack4_s8(q[0], q[1], q[2], q[3]);
      }
    }
  }
}

__global__ void sg_final_kernel(const double* __restrict__ part, double* __restrict__ out, int nblocks, int N) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    double t = 0.0;
    for (int b = 0; b < nblocks; ++b) t += part[size_t(b) * N + i];
    out[i] = t;
  }
}

int64_t linear_sqerr_scratch_bytes(int M, int n) { return int64_t((M + SG_BM - 1) / SG_BM) * n * int64_t(sizeof(double)); }

int launch_linear_sqerr(const float* x, int M, int K, int patch, int Cin, int H, int W, const float* D, int n, double* out, double* scratch,
                        cudaStream_t stream) {
  SgemmA a{x, M, K, patch, Cin, H, W};
  SgemmEmbed ep{};
  dim3 grid((n + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM);
  sgemm_kernel<0><<<grid, SG_THREADS, 0, stream>>>(a, D, n, scratch, ep);
  sg_final_kernel<<<std::min((n + 255) / 256, 1024), 256, 0, stream>>>(scratch, out, int(grid.y), n);
  count_launch(2);
  return check_launch("linear_sqerr_scores");
}

int launch_linear_f32(const float* x, int M, int K, int patch, int Cin, int H, int W, const float* Wt, const float* bias, int N, float* out,
                      cudaStream_t stream) {
  SgemmA a{x, M, K, patch, Cin, H, W};
  SgemmEmbed ep{};
  ep.bias = bias;
  ep.out_f32 = out;
  dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM);
  sgemm_kernel<2><<<grid, SG_THREADS, 0, stream>>>(a, Wt, N, nullptr, ep);
  count_launch();
  return check_launch("linear_f32");
}

int launch_embed_f32(const float* img, int B, int Cin, int H, int W, int P, const float* w_hat, int N, const SgemmEmbed& ep, cudaStream_t stream) {
  const int M = B * (H / P) * (W / P), K = Cin * P * P;
  SgemmA a{img, M, K, P, Cin, H, W};
  dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM);
  sgemm_kernel<1><<<grid, SG_THREADS, 0, stream>>>(a, w_hat, N, nullptr, ep);
  count_launch();
  return check_launch("embed_f32");
}

}  // namespace p2v
