// CUDA-core (dp4a) int8 GEMM with the same fused epilogues as the tcgen05 kernel.
// NOT the product path: tests use it to cross-check gemm_tc.cu and to localise a failure to
// "tensor-core plumbing" vs "epilogue arithmetic" (include/p2vit_b200.h: p2v_gemm_i8_simt).
#include "epilogue.cuh"

namespace p2v {

constexpr int SIMT_NC = 8;        // columns per thread
constexpr int SIMT_THREADS = 128;

template <int EPI, bool POT>
__global__ void __launch_bounds__(SIMT_THREADS) gemm_simt_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ W, EpiParams p) {
  // block = 16 rows x 8 column-groups (64 cols); thread = 1 row x 8 cols
  __shared__ __align__(16) float cp[CP_ROWS * 64];
  const int n0 = blockIdx.x * 64;
  stage_col_params<EPI, POT, 64>(p, cp, n0, int(threadIdx.x));
  __syncthreads();
  const int row = blockIdx.y * 16 + threadIdx.x / 8;
  const int c0 = (threadIdx.x % 8) * SIMT_NC;
  const int col0 = n0 + c0;
  if (row >= p.M || col0 >= p.N) return;
  int acc[SIMT_NC];
#pragma unroll
  for (int j = 0; j < SIMT_NC; ++j) acc[j] = 0;
  const int K = p.K;  // multiple of 16
  const int4* a4 = reinterpret_cast<const int4*>(A + size_t(row) * K);
  for (int k = 0; k < K / 16; ++k) {
    const int4 av = __ldg(a4 + k);
#pragma unroll
    for (int j = 0; j < SIMT_NC; ++j) {
      if (col0 + j < p.N) {
        const int4 wv = __ldg(reinterpret_cast<const int4*>(W + size_t(col0 + j) * K) + k);
        acc[j] = __dp4a(av.x, wv.x, acc[j]);
        acc[j] = __dp4a(av.y, wv.y, acc[j]);
        acc[j] = __dp4a(av.z, wv.z, acc[j]);
        acc[j] = __dp4a(av.w, wv.w, acc[j]);
      }
    }
  }
  uint32_t resw[SIMT_NC / 4];
  if (EPI == P2V_EPI_RESIDUAL) load_residual<SIMT_NC>(p, row, col0, resw);
  GeluTab gt{nullptr, 0, 0.f, 0.f};
  if (EPI == P2V_EPI_GELU && POT && p.gelu_table) {
    const GeluTabHeader hd = *reinterpret_cast<const GeluTabHeader*>(p.gelu_table);
    gt = GeluTab{reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(p.gelu_table) + sizeof(GeluTabHeader)), hd.n, hd.inv_w,
                 -hd.y0 * hd.inv_w};
  }
  epilogue_row<EPI, POT, 64, SIMT_NC>(p, cp, row, n0, c0, acc, resw, gt);
}

int launch_gemm_simt(const p2v_gemm_args& a, cudaStream_t stream) {
  EpiParams p = make_epi_params(a);
  dim3 grid((a.N + 63) / 64, (a.M + 15) / 16);
  P2V_DISPATCH_EPI(a.epilogue, a.pot_scales != 0,
                   gemm_simt_kernel<EPI, POT><<<grid, SIMT_THREADS, 0, stream>>>(a.A, a.W, p););
  count_launch();
  return check_launch("gemm_simt");
}

}  // namespace p2v
