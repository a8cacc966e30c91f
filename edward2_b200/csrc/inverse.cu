// Sigma = A^-1 for the symmetric positive definite A = ridge I + precision of the Laplace random-feature covariance
// (edward2/tensorflow/layers/random_feature.py:447-459 tf.linalg.inv; edward2/jax/nn/random_feature.py:393-404), on the
// GPU without a library call: in-place block Gauss-Jordan (Jordan exchange) with 64-wide pivot blocks.
//
//   pivot block P = A_jj:   A_jj <- P^-1;  A_jk <- P^-1 A_jk;  A_ij <- -A_ij P^-1;  A_ik <- A_ik - A_ij P^-1 A_jk  (i, k != j)
//
// After the last block A holds the inverse.  Every pivot block is a Schur complement of an SPD matrix, hence SPD: no
// pivoting is needed.  The O(n^2 * 64) work of a step is three fp32 FFMA GEMMs (gemm_f32: exact fp32 products — the bf16
// tensor-core path would put ~1e-3 relative error into an inverse); the pivot block itself is inverted by one CTA in
// shared memory.  2 n^3 FLOP in total, fp32 throughout: the result matches a float64 inverse to ~cond(A) * 1e-7.
#include <cuda_runtime.h>

#include <algorithm>

#include "gemm.h"
#include "kernels.h"
#include "status.h"

namespace ed2 {
namespace {

constexpr int NB = 64;

// w[bs][NB] <- (a[j0:j0+bs, j0:j0+bs])^-1 by Gauss-Jordan without pivoting (SPD block), one 256-thread CTA
__global__ void __launch_bounds__(256) invert_pivot_kernel(const float* __restrict__ a, long long lda, int bs,
                                                          float* __restrict__ w, int* __restrict__ status) {
  __shared__ float s[NB][NB + 1];
  __shared__ float rowp[NB], colp[NB];
  const int t = threadIdx.x;
  for (int i = t; i < bs * bs; i += 256) s[i / bs][i % bs] = a[static_cast<long long>(i / bs) * lda + (i % bs)];
  __syncthreads();
  for (int p = 0; p < bs; ++p) {
    const float piv = s[p][p];
    if (t == 0 && !(piv > 0.f)) *status = 1;  // not positive definite (or NaN)
    const float inv = 1.f / piv;
    if (t < bs) {
      rowp[t] = (t == p ? 1.f : s[p][t]) * inv;
      colp[t] = s[t][p];
    }
    __syncthreads();
    for (int i = t; i < bs * bs; i += 256) {
      const int r = i / bs, c = i - r * bs;
      if (r == p) s[r][c] = rowp[c];
      else s[r][c] = (c == p ? 0.f : s[r][c]) - colp[r] * rowp[c];
    }
    __syncthreads();
  }
  for (int i = t; i < bs * bs; i += 256) w[(i / bs) * NB + (i % bs)] = s[i / bs][i % bs];
}

int f32_gemm(const float* A, int lda, int major_a, const float* B, int ldb, int major_b, int M, int N, int K, float alpha,
             const float* addend, int addend_ld, float* out, int out_ld, cudaStream_t s) {
  GemmDesc g{};
  g.A = A; g.lda = lda; g.major_a = major_a;
  g.B = B; g.ldb = ldb; g.major_b = major_b;
  g.M = M; g.N = N; g.K = K;
  g.ep = EpiParams{};
  g.ep.alpha = alpha; g.ep.beta = 1.f; g.ep.act_scale = 1.f;
  g.ep.addend = addend; g.ep.addend_ld = addend_ld; g.ep.addend_dt = kF32;
  g.ep.out = out; g.ep.out_ld = out_ld; g.ep.out_dt = kF32;
  return gemm_f32(g, s);
}

}  // namespace

long long k_spd_inverse_workspace_floats(int n) { return 2LL * n * NB + NB * NB + 4; }

int k_spd_inverse(float* a, int n, float* ws, cudaStream_t s) {
  if (n <= 0) return set_error(ED2_ERR_BAD_SHAPE, "spd_inverse: empty matrix");
  float* w = ws;                           // [NB][NB]   inverse of the pivot block
  float* c = w + NB * NB;                  // [n][NB]    the pivot column panel before the step
  float* r = c + static_cast<long long>(n) * NB;  // [NB][n]  P^-1 times the pivot row panel
  int* status = reinterpret_cast<int*>(r + static_cast<long long>(n) * NB);
  cudaMemsetAsync(status, 0, sizeof(int), s);
  for (int j0 = 0; j0 < n; j0 += NB) {
    const int bs = std::min(NB, n - j0);
    float* ajj = a + static_cast<long long>(j0) * n + j0;
    invert_pivot_kernel<<<1, 256, 0, s>>>(ajj, n, bs, w, status);
    count_launch();
    int st = check_launch("spd_inverse(pivot)");
    if (st != ED2_OK) return st;
    if (n == bs) {  // single block: done
      cudaMemcpy2DAsync(a, sizeof(float) * n, w, sizeof(float) * NB, sizeof(float) * bs, bs, cudaMemcpyDeviceToDevice, s);
      break;
    }
    // C <- A[:, j]  (n x bs)
    cudaMemcpy2DAsync(c, sizeof(float) * NB, a + j0, sizeof(float) * n, sizeof(float) * bs, n, cudaMemcpyDeviceToDevice, s);
    // R <- P^-1 A[j, :]   (bs x n)
    st = f32_gemm(w, NB, kMajorK, a + static_cast<long long>(j0) * n, n, kMajorMN, bs, n, bs, 1.f, nullptr, 0, r, n, s);
    if (st != ED2_OK) return st;
    // A <- A - C R  (all rows and columns; the pivot row / column blocks are overwritten below)
    st = f32_gemm(c, NB, kMajorK, r, n, kMajorMN, n, n, bs, -1.f, a, n, a, n, s);
    if (st != ED2_OK) return st;
    // A[:, j] <- -C P^-1
    st = f32_gemm(c, NB, kMajorK, w, NB, kMajorMN, n, bs, bs, -1.f, nullptr, 0, a + j0, n, s);
    if (st != ED2_OK) return st;
    // A[j, :] <- R, then A_jj <- P^-1
    cudaMemcpyAsync(a + static_cast<long long>(j0) * n, r, sizeof(float) * static_cast<size_t>(bs) * n, cudaMemcpyDeviceToDevice, s);
    cudaMemcpy2DAsync(ajj, sizeof(float) * n, w, sizeof(float) * NB, sizeof(float) * bs, bs, cudaMemcpyDeviceToDevice, s);
  }
  return ED2_OK;
}

}  // namespace ed2
