// Pointwise part of the Bayesian LSTM cells (edward2/tensorflow/layers/recurrent.py:30 LSTMCellReparameterization;
// Keras LSTMCell._compute_carry_and_output, gate order [i | f | c | o] along the 4*units axis, :163-184):
//   i = sigmoid(z_i), f = sigmoid(z_f), g = tanh(z_c), o = sigmoid(z_o);  c = f c_prev + i g;  h = o tanh(c)
// z = x·W + h_prev·U + b comes from two tcgen05 GEMMs (the second accumulates onto the first through the addend).
// One thread per (row, unit): four strided reads of z, fp32 cell state, h in the compute dtype.
#include <cuda_bf16.h>

#include <algorithm>

#include "gemm_epilogue.h"
#include "kernels.h"
#include "status.h"

namespace ed2 {
namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ float ldz(const void* p, int dt, long long i) {
  return dt == kBF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void stz(void* p, int dt, long long i, float v) {
  if (dt == kBF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[i] = v;
}
__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(kBlock) lstm_pointwise_fwd_kernel(const void* __restrict__ z, int dt, long long z_ld,
                                                                   const float* __restrict__ c_prev, int B, int H,
                                                                   float* __restrict__ c, void* __restrict__ h, long long h_ld,
                                                                   const void* __restrict__ z2, long long z2_ld) {
  const long long total = static_cast<long long>(B) * H;
  for (long long idx = blockIdx.x * (long long)kBlock + threadIdx.x; idx < total; idx += (long long)gridDim.x * kBlock) {
    const long long b = idx / H;
    const int j = static_cast<int>(idx - b * H);
    const long long zr = b * z_ld, zr2 = b * z2_ld;
    auto pre = [&](int gate) {  // pre-activation of one gate: z (+ z2: the recurrent branch of the Flipout / rank-1 cells)
      const float v = ldz(z, dt, zr + gate * H + j);
      return z2 != nullptr ? v + ldz(z2, dt, zr2 + gate * H + j) : v;
    };
    const float i = sigm(pre(0)), f = sigm(pre(1));
    const float g = tanhf(pre(2)), o = sigm(pre(3));
    const float cn = fmaf(f, c_prev[idx], i * g);
    c[idx] = cn;
    stz(h, dt, b * h_ld + j, o * tanhf(cn));
  }
}

// dz = [di | df | dg | do] (pre-activation gradients), dc_prev, dbias[4H] += column sums of dz (fp32 atomics; zeroed by
// the launcher).  The gates are recomputed from z.
__global__ void __launch_bounds__(kBlock) lstm_pointwise_bwd_kernel(const void* __restrict__ z, int dt, long long z_ld,
                                                                   const float* __restrict__ c_prev, const float* __restrict__ c,
                                                                   const void* __restrict__ dh, long long dh_ld,
                                                                   const float* __restrict__ dc, int B, int H,
                                                                   void* __restrict__ dz, long long dz_ld,
                                                                   float* __restrict__ dc_prev, float* __restrict__ dbias,
                                                                   const void* __restrict__ z2, long long z2_ld) {
  const long long total = static_cast<long long>(B) * H;
  for (long long idx = blockIdx.x * (long long)kBlock + threadIdx.x; idx < total; idx += (long long)gridDim.x * kBlock) {
    const long long b = idx / H;
    const int j = static_cast<int>(idx - b * H);
    const long long zr = b * z_ld, zr2 = b * z2_ld;
    auto pre = [&](int gate) {
      const float v = ldz(z, dt, zr + gate * H + j);
      return z2 != nullptr ? v + ldz(z2, dt, zr2 + gate * H + j) : v;
    };
    const float i = sigm(pre(0)), f = sigm(pre(1));
    const float g = tanhf(pre(2)), o = sigm(pre(3));
    const float tc = tanhf(c[idx]);
    const float gh = dh != nullptr ? ldz(dh, dt, b * dh_ld + j) : 0.f;
    const float dct = (dc != nullptr ? dc[idx] : 0.f) + gh * o * (1.f - tc * tc);
    const float d_i = dct * g * i * (1.f - i), d_f = dct * c_prev[idx] * f * (1.f - f);
    const float d_g = dct * i * (1.f - g * g), d_o = gh * tc * o * (1.f - o);
    const long long dr = b * dz_ld;
    stz(dz, dt, dr + j, d_i); stz(dz, dt, dr + H + j, d_f); stz(dz, dt, dr + 2 * H + j, d_g); stz(dz, dt, dr + 3 * H + j, d_o);
    dc_prev[idx] = dct * f;
    if (dbias != nullptr) {
      // the values the weight-gradient GEMMs read (rounded to the compute dtype)
      atomicAdd(dbias + j, ldz(dz, dt, dr + j));
      atomicAdd(dbias + H + j, ldz(dz, dt, dr + H + j));
      atomicAdd(dbias + 2 * H + j, ldz(dz, dt, dr + 2 * H + j));
      atomicAdd(dbias + 3 * H + j, ldz(dz, dt, dr + 3 * H + j));
    }
  }
}

int grid_for(long long work) {
  return static_cast<int>(std::max<long long>(1, std::min<long long>((work + kBlock - 1) / kBlock, 148 * 8)));
}

}  // namespace

int k_lstm_pointwise_fwd(const void* z, int dt, long long z_ld, const float* c_prev, int B, int H, float* c, void* h,
                         long long h_ld, cudaStream_t s, const void* z2, long long z2_ld) {
  if (B <= 0 || H <= 0) return set_error(ED2_ERR_BAD_SHAPE, "lstm_pointwise: empty");
  lstm_pointwise_fwd_kernel<<<grid_for(static_cast<long long>(B) * H), kBlock, 0, s>>>(z, dt, z_ld, c_prev, B, H, c, h, h_ld, z2, z2_ld);
  count_launch();
  return check_launch("lstm_pointwise_fwd");
}

int k_lstm_pointwise_bwd(const void* z, int dt, long long z_ld, const float* c_prev, const float* c, const void* dh,
                         long long dh_ld, const float* dc, int B, int H, void* dz, long long dz_ld, float* dc_prev,
                         float* dbias, cudaStream_t s, const void* z2, long long z2_ld) {
  if (B <= 0 || H <= 0) return set_error(ED2_ERR_BAD_SHAPE, "lstm_pointwise: empty");
  if (dbias != nullptr) cudaMemsetAsync(dbias, 0, sizeof(float) * 4 * static_cast<size_t>(H), s);
  lstm_pointwise_bwd_kernel<<<grid_for(static_cast<long long>(B) * H), kBlock, 0, s>>>(z, dt, z_ld, c_prev, c, dh, dh_ld, dc, B,
                                                                                        H, dz, dz_ld, dc_prev, dbias, z2, z2_ld);
  count_launch();
  return check_launch("lstm_pointwise_bwd");
}

}  // namespace ed2
