// fp32 GEMM with the same layouts and fused epilogue as the tcgen05 kernel: exact fp32 products, fp32 FFMA
// accumulation.  It serves the fp32 parity path (1e-5 vs the float64 oracle) and the small, launch-bound
// fp32 shapes (BASELINE config 1: BatchEnsemble MLP 784-512-512-10 at batch 128), where the problem is far
// below one tensor-core wave and the step is latency- rather than math-bound.
#include "gemm.h"
#include "status.h"

#include <cuda_bf16.h>

namespace ed2 {
namespace {

constexpr int TK = 16;

__device__ __forceinline__ float ld_elem(const void* p, int dt, long long idx) {
  return dt == kBF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx])
                     : reinterpret_cast<const float*>(p)[idx];
}
__device__ __forceinline__ void st_elem(void* p, int dt, long long idx, float v) {
  if (dt == kBF16) reinterpret_cast<__nv_bfloat16*>(p)[idx] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[idx] = v;
}

// Tile TM x TN per 256-thread block, (TM/16) x (TN/16) outputs per thread.  64x64 for large problems; 32x32 for the
// small, latency-bound shapes (BASELINE config 1) so that the grid still covers the SMs.
template <int MA, int MB, int TM, int TN>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, int lda,
                                                       const float* __restrict__ B, int ldb, int M, int N, int K,
                                                       EpiParams ep) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int ty = t / 16, tx = t % 16;
  constexpr int RM = TM / 16, RN = TN / 16;
  float acc[RM][RN] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
    for (int i = 0; i < (TM * TK) / 256; ++i) {
      const int idx = t + i * 256;
      int m, k;
      if (MA == kMajorK) { k = idx % TK; m = idx / TK; } else { m = idx % TM; k = idx / TM; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K) v = MA == kMajorK ? A[(long long)gm * lda + gk] : A[(long long)gk * lda + gm];
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < (TN * TK) / 256; ++i) {
      const int idx = t + i * 256;
      int n, kk;
      if (MB == kMajorK) { kk = idx % TK; n = idx / TK; } else { n = idx % TN; kk = idx / TN; }
      const int gn = n0 + n, gk2 = k0 + kk;
      float w = 0.f;
      if (gn < N && gk2 < K) w = MB == kMajorK ? B[(long long)gn * ldb + gk2] : B[(long long)gk2 * ldb + gn];
      Bs[kk][n] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[RM], b[RN];
#pragma unroll
      for (int i = 0; i < RM; ++i) a[i] = As[k][ty * RM + i];
#pragma unroll
      for (int j = 0; j < RN; ++j) b[j] = Bs[k][tx * RN + j];
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int m = m0 + ty * RM + i;
    if (m >= M) continue;
    const int grp = ep.group_rows > 0 ? m / ep.group_rows : 0;
    float rs = 0.f;
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      const int n = n0 + tx * RN + j;
      if (n >= N) continue;
      float v = acc[i][j] * ep.alpha;
      if (ep.raw) st_elem(ep.raw, ep.raw_dt, (long long)m * ep.raw_ld + n, v);
      if (ep.elem_scale) v *= ld_elem(ep.elem_scale, ep.elem_scale_dt, (long long)m * ep.elem_scale_ld + n);
      if (ep.col_scale) v *= ep.col_scale[(long long)grp * ep.col_scale_ld + n];
      if (ep.addend) v = fmaf(ep.beta, ld_elem(ep.addend, ep.addend_dt, (long long)m * ep.addend_ld + n), v);
      if (ep.col_bias) v += ep.col_bias[(long long)grp * ep.col_bias_ld + n];
      if (ep.act == kActRelu) v = fmaxf(v, 0.f);
      else if (ep.act == kActCos) v = cosf(v) * ep.act_scale;
      if (ep.mask_src && !(ld_elem(ep.mask_src, ep.mask_dt, (long long)m * ep.mask_ld + n) > 0.f)) v = 0.f;
      if (ep.out) st_elem(ep.out, ep.out_dt, (long long)m * ep.out_ld + n, v);
      if (ep.col_sum) atomicAdd(ep.col_sum + (long long)grp * ep.col_sum_ld + n, v);
      rs += v;
    }
    if (ep.row_sum) atomicAdd(ep.row_sum + m, rs);
  }
}

template <int TM, int TN>
void launch_f32(const GemmDesc& g, cudaStream_t stream) {
  dim3 grid((g.N + TN - 1) / TN, (g.M + TM - 1) / TM);
  const float* A = reinterpret_cast<const float*>(g.A);
  const float* B = reinterpret_cast<const float*>(g.B);
  if (g.major_a == kMajorK && g.major_b == kMajorK)
    gemm_f32_kernel<kMajorK, kMajorK, TM, TN><<<grid, 256, 0, stream>>>(A, g.lda, B, g.ldb, g.M, g.N, g.K, g.ep);
  else if (g.major_a == kMajorK && g.major_b == kMajorMN)
    gemm_f32_kernel<kMajorK, kMajorMN, TM, TN><<<grid, 256, 0, stream>>>(A, g.lda, B, g.ldb, g.M, g.N, g.K, g.ep);
  else if (g.major_a == kMajorMN && g.major_b == kMajorK)
    gemm_f32_kernel<kMajorMN, kMajorK, TM, TN><<<grid, 256, 0, stream>>>(A, g.lda, B, g.ldb, g.M, g.N, g.K, g.ep);
  else
    gemm_f32_kernel<kMajorMN, kMajorMN, TM, TN><<<grid, 256, 0, stream>>>(A, g.lda, B, g.ldb, g.M, g.N, g.K, g.ep);
}

}  // namespace

int gemm_f32(const GemmDesc& g, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return set_error(ED2_ERR_BAD_SHAPE, "gemm_f32: empty shape");
  const long long tiles64 = (long long)((g.M + 63) / 64) * ((g.N + 63) / 64);
  if (tiles64 >= 148) launch_f32<64, 64>(g, stream);
  else launch_f32<32, 32>(g, stream);
  count_launch();
  return check_launch("gemm_f32");
}

int gemm(int dtype, const GemmDesc& g, cudaStream_t stream) {
  if (dtype == kBF16) return gemm_bf16(g, stream);
  if (dtype == kF32) return gemm_f32(g, stream);
  return set_error(ED2_ERR_BAD_DTYPE, "gemm: unknown dtype %d", dtype);
}

}  // namespace ed2
