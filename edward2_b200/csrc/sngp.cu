// SNGP-head kernels that are not contractions: the spectral-norm power iteration (coalesced, bandwidth-bound
// matvecs over the weight matrix), input LayerNorm, the precision-matrix momentum blend and mean-field logits.
#include "kernels.h"

#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "gemm_epilogue.h"
#include "status.h"

namespace ed2 {
namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[kBlock / 32];
  __shared__ float total;
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < kBlock / 32 ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) total = t;
  }
  __syncthreads();
  const float r = total;
  __syncthreads();
  return r;
}

// t[i] = sum_o W[i][o] * u[o]   — one warp per row, float4 loads along the contiguous axis.
__global__ void __launch_bounds__(kBlock) matvec_rows_kernel(const float* __restrict__ w, int rows, int cols,
                                                             const float* __restrict__ u, float* __restrict__ t) {
  const int warp = (blockIdx.x * kBlock + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * kBlock) >> 5;
  const bool vec = (cols & 3) == 0;
  for (int r = warp; r < rows; r += nwarps) {
    const float* wr = w + (long long)r * cols;
    float acc = 0.f;
    if (vec) {
      for (int c = lane * 4; c < cols; c += 128) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(wr + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(u + c));
        acc += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
      }
    } else {
      for (int c = lane; c < cols; c += 32) acc += wr[c] * u[c];
    }
    acc = warp_sum(acc);
    if (lane == 0) t[r] = acc;
  }
}

// t[o] += sum_{i in slab} v[i] * W[i][o]   — block = 128 columns x a slab of rows; coalesced along o.
constexpr int kSlabRows = 128;
__global__ void __launch_bounds__(kBlock) matvec_cols_kernel(const float* __restrict__ w, int rows, int cols,
                                                             const float* __restrict__ v, float* __restrict__ t) {
  __shared__ float red[kBlock / 32][128 + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 128 + lane * 4;
  const int r0 = blockIdx.y * kSlabRows;
  const int r1 = min(rows, r0 + kSlabRows);
  float acc[4] = {0, 0, 0, 0};
  if (c < cols) {
    const bool vec = (cols & 3) == 0;
    for (int r = r0 + warp; r < r1; r += kBlock / 32) {
      const float vv = __ldg(v + r);
      const float* wr = w + (long long)r * cols + c;
      if (vec) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(wr));
        acc[0] += vv * a.x; acc[1] += vv * a.y; acc[2] += vv * a.z; acc[3] += vv * a.w;
      } else {
        for (int j = 0; j < 4; ++j)
          if (c + j < cols) acc[j] += vv * wr[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[warp][lane * 4 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 128) {
    const int cc = blockIdx.x * 128 + threadIdx.x;
    if (cc < cols) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < kBlock / 32; ++k) s += red[k][threadIdx.x];
      atomicAdd(t + cc, s);
    }
  }
}

// dst = src * rsqrt(max(sum(src^2), 1e-12))   (tf.nn.l2_normalize / edward2/jax/nn/normalization.py:50-51)
__global__ void __launch_bounds__(kBlock) l2_normalize_kernel(const float* __restrict__ src, int n, float* dst) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += kBlock) acc += src[i] * src[i];
  const float tot = block_sum(acc);
  const float inv = rsqrtf(fmaxf(tot, 1e-12f));
  for (int i = threadIdx.x; i < n; i += kBlock) dst[i] = src[i] * inv;
}

// scal[0] = sigma = <u, u_raw>;  scal[1] = multiplicative factor applied to W
__global__ void __launch_bounds__(kBlock) sigma_kernel(const float* __restrict__ u, const float* __restrict__ u_raw,
                                                       int n, float c, int mode, float* scal, float* sigma_out) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += kBlock) acc += u[i] * u_raw[i];
  const float sigma = block_sum(acc);
  if (threadIdx.x == 0) {
    float factor;
    if (mode == 0) factor = (c / sigma) < 1.f ? (c / sigma) : 1.f;  // normalization.py:387-390
    else factor = 1.f / fmaxf(sigma / c, 1.f);                      // jax/nn/normalization.py:175-178
    scal[0] = sigma;
    scal[1] = factor;
    if (sigma_out != nullptr) *sigma_out = sigma;
  }
}

__global__ void scale_w_kernel(const float* __restrict__ w, long long rows, long long cols, const float* scal,
                               int mode, float c, float* w_out, __nv_bfloat16* w_lp, long long lp_ld) {
  const float sigma = scal[0], factor = scal[1];
  const float div = fmaxf(sigma / c, 1.f);
  const long long n = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = w[i];
    const float y = mode == 0 ? (factor < 1.f ? factor * x : x) : x / div;
    if (w_out != nullptr) w_out[i] = y;
    if (w_lp != nullptr) {
      const long long r = i / cols, cc = i - r * cols;
      w_lp[r * lp_ld + cc] = __float2bfloat16_rn(y);
    }
  }
}

__device__ __forceinline__ float ld1(const void* p, int dt, long long i) {
  return dt == kBF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                     : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st1(void* p, int dt, long long i, float v) {
  if (dt == kBF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[i] = v;
}

// one warp per row; two-pass mean/variance in registers-from-L1 (row is re-read from L1/L2, not HBM)
__global__ void __launch_bounds__(kBlock) input_normalize_kernel(const void* x, int xdt, long long x_ld, void* y,
                                                                 int ydt, long long y_ld, long long rows, int cols,
                                                                 const float* gamma, const float* beta, float eps,
                                                                 int scale_only, float scale) {
  const long long warp = (blockIdx.x * (long long)kBlock + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * kBlock) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    if (scale_only) {
      for (int c = lane; c < cols; c += 32) st1(y, ydt, r * y_ld + c, ld1(x, xdt, r * x_ld + c) * scale);
      continue;
    }
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += ld1(x, xdt, r * x_ld + c);
    const float mean = warp_sum(s) / cols;
    float q = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float d = ld1(x, xdt, r * x_ld + c) - mean;
      q += d * d;
    }
    const float inv = rsqrtf(warp_sum(q) / cols + eps);
    for (int c = lane; c < cols; c += 32) {
      float v = (ld1(x, xdt, r * x_ld + c) - mean) * inv;
      if (gamma != nullptr) v *= gamma[c];
      if (beta != nullptr) v += beta[c];
      st1(y, ydt, r * y_ld + c, v);
    }
  }
}


// LayerNorm backward, one warp per row; column sums for dgamma / dbeta are accumulated in shared memory per block
// and flushed with one global atomic per column per block.
//   xhat = (x - mean) * inv;  g = dy * gamma;  dx = inv * (g - mean(g) - xhat * mean(g * xhat))
__global__ void __launch_bounds__(kBlock) layernorm_bwd_kernel(const void* x, int xdt, long long x_ld, const void* dy,
                                                               int dydt, long long dy_ld, long long rows, int cols,
                                                               const float* gamma, float eps, void* dx, int dxdt,
                                                               long long dx_ld, float* dgamma, float* dbeta) {
  extern __shared__ float sh[];
  float* sg = sh;
  float* sb = sh + cols;
  const bool want_param = dgamma != nullptr || dbeta != nullptr;
  if (want_param) {
    for (int c = threadIdx.x; c < cols; c += kBlock) { sg[c] = 0.f; sb[c] = 0.f; }
    __syncthreads();
  }
  const long long warp = (blockIdx.x * (long long)kBlock + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * kBlock) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += ld1(x, xdt, r * x_ld + c);
    const float mean = warp_sum(s) / cols;
    float q = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float d = ld1(x, xdt, r * x_ld + c) - mean;
      q += d * d;
    }
    const float inv = rsqrtf(warp_sum(q) / cols + eps);
    float sum_g = 0.f, sum_gx = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float xh = (ld1(x, xdt, r * x_ld + c) - mean) * inv;
      const float d = ld1(dy, dydt, r * dy_ld + c);
      const float g = d * (gamma != nullptr ? gamma[c] : 1.f);
      sum_g += g;
      sum_gx += g * xh;
      if (want_param) {
        atomicAdd(&sg[c], d * xh);
        atomicAdd(&sb[c], d);
      }
    }
    const float mg = warp_sum(sum_g) / cols, mgx = warp_sum(sum_gx) / cols;
    if (dx != nullptr) {
      for (int c = lane; c < cols; c += 32) {
        const float xh = (ld1(x, xdt, r * x_ld + c) - mean) * inv;
        const float g = ld1(dy, dydt, r * dy_ld + c) * (gamma != nullptr ? gamma[c] : 1.f);
        st1(dx, dxdt, r * dx_ld + c, inv * (g - mg - xh * mgx));
      }
    }
  }
  if (want_param) {
    __syncthreads();
    for (int c = threadIdx.x; c < cols; c += kBlock) {
      if (dgamma != nullptr) atomicAdd(dgamma + c, sg[c]);
      if (dbeta != nullptr) atomicAdd(dbeta + c, sb[c]);
    }
  }
}

// The same backward for cols == NJ * 32 <= 1024 with the row held in registers: x and dy are read ONCE (the kernel above
// re-reads x four times and issues two shared-memory atomics per element), the dgamma / dbeta column sums of the rows
// a warp owns accumulate in its registers and reach shared memory once per warp, global memory once per block.
template <int NJ>
__global__ void __launch_bounds__(kBlock) layernorm_bwd_rows_kernel(const void* x, int xdt, long long x_ld, const void* dy,
                                                                    int dydt, long long dy_ld, long long rows,
                                                                    const float* gamma, float eps, void* dx, int dxdt,
                                                                    long long dx_ld, float* dgamma, float* dbeta) {
  constexpr int cols = NJ * 32;
  __shared__ float sg[cols], sb[cols];
  const bool want_param = dgamma != nullptr || dbeta != nullptr;
  if (want_param) {
    for (int c = threadIdx.x; c < cols; c += kBlock) { sg[c] = 0.f; sb[c] = 0.f; }
    __syncthreads();
  }
  const long long warp = (blockIdx.x * (long long)kBlock + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * kBlock) >> 5;
  float ag[NJ], ab[NJ], gam[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    ag[j] = 0.f; ab[j] = 0.f;
    gam[j] = gamma != nullptr ? gamma[j * 32 + lane] : 1.f;
  }
  for (long long r = warp; r < rows; r += nwarps) {
    float xv[NJ], dv[NJ];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      xv[j] = ld1(x, xdt, r * x_ld + j * 32 + lane);
      dv[j] = ld1(dy, dydt, r * dy_ld + j * 32 + lane);
      s += xv[j];
    }
    const float mean = warp_sum(s) / cols;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float d = xv[j] - mean;
      q += d * d;
    }
    const float inv = rsqrtf(warp_sum(q) / cols + eps);
    float sum_g = 0.f, sum_gx = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float xh = (xv[j] - mean) * inv;
      const float g = dv[j] * gam[j];
      sum_g += g;
      sum_gx += g * xh;
      ag[j] = fmaf(dv[j], xh, ag[j]);
      ab[j] += dv[j];
      xv[j] = xh;
      dv[j] = g;
    }
    const float mg = warp_sum(sum_g) / cols, mgx = warp_sum(sum_gx) / cols;
    if (dx != nullptr) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) st1(dx, dxdt, r * dx_ld + j * 32 + lane, inv * (dv[j] - mg - xv[j] * mgx));
    }
  }
  if (want_param) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      atomicAdd(&sg[j * 32 + lane], ag[j]);
      atomicAdd(&sb[j * 32 + lane], ab[j]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cols; c += kBlock) {
      if (dgamma != nullptr) atomicAdd(dgamma + c, sg[c]);
      if (dbeta != nullptr) atomicAdd(dbeta + c, sb[c]);
    }
  }
}

__global__ void precision_blend_kernel(float* p, const float* __restrict__ summed, long long n, float momentum,
                                       float inv_batch) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (momentum > 0.f) p[i] = momentum * p[i] + (1.f - momentum) * summed[i] * inv_batch;
    else p[i] = p[i] + summed[i];
  }
}

__global__ void mean_field_kernel(const float* __restrict__ logits, const float* __restrict__ var, long long batch,
                                  int classes, float factor, int likelihood, float var_scale, float* out) {
  const long long n = batch * classes;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / classes;
    float l = logits[i];
    if (factor >= 0.f) {  // utils.py:400-402: a negative factor returns the logits unchanged
      const float v = var[b] * var_scale;
      const float scale = likelihood == 1 ? expf(-v * factor * 0.5f) : sqrtf(1.f + v * factor);
      l = l / scale;
    }
    out[i] = l;
  }
}

__global__ void rff_bwd_prep_kernel(const void* dphi, int ddt, long long dphi_ld, const float* __restrict__ pre,
                                    long long pre_ld, const float* __restrict__ bias, float scale, long long rows, int cols, void* gpre, int gdt,
                                    long long gpre_ld) {
  const long long n = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    const float g = ld1(dphi, ddt, r * dphi_ld + c);
    st1(gpre, gdt, r * gpre_ld + c, -scale * sinf(pre[r * pre_ld + c] + (bias != nullptr ? bias[c] : 0.f)) * g);
  }
}

// LayerNorm forward that writes the bf16 SPLIT A-operand of the fp32-grade random-feature GEMM directly:
//   y = (x - mean) * rsqrt(var + eps) * gamma + beta;   h = bf16(y), m = bf16(y - h);   out row = [h | h | m]
// (the 2-term split of ed2_split_bf16, is_b = 0) — one pass over x instead of LayerNorm -> fp32 tensor -> split.
// One warp per row, 8 columns per lane per trip (cols % 8 == 0), the row is re-read from L1 / L2 for the three passes.
__device__ __forceinline__ void ld8(const void* p, int dt, long long off, float (&f)[8]) {
  if (dt == kBF16) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + off));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 v = __bfloat1622float2(h[t]);
      f[2 * t] = v.x; f[2 * t + 1] = v.y;
    }
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + off));
    const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + off) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
}
__device__ __forceinline__ uint4 pack8_bf16(const float (&f)[8]) {
  uint4 q;
  __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
  q.x = *reinterpret_cast<uint32_t*>(&h0); q.y = *reinterpret_cast<uint32_t*>(&h1);
  q.z = *reinterpret_cast<uint32_t*>(&h2); q.w = *reinterpret_cast<uint32_t*>(&h3);
  return q;
}

__global__ void __launch_bounds__(kBlock) layernorm_split_kernel(const void* __restrict__ x, int xdt, long long x_ld,
                                                                 __nv_bfloat16* __restrict__ out, long long out_ld,
                                                                 long long rows, int cols, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float eps) {
  const long long warp = (blockIdx.x * (long long)kBlock + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * kBlock) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    float s = 0.f;
    for (int c = lane * 8; c < cols; c += 256) {
      float f[8];
      ld8(x, xdt, r * x_ld + c, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[j];
    }
    const float mean = warp_sum(s) / cols;
    float q = 0.f;
    for (int c = lane * 8; c < cols; c += 256) {
      float f[8];
      ld8(x, xdt, r * x_ld + c, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = f[j] - mean; q += d * d; }
    }
    const float inv = rsqrtf(warp_sum(q) / cols + eps);
    __nv_bfloat16* o = out + r * out_ld;
    for (int c = lane * 8; c < cols; c += 256) {
      float f[8], g[8], b[8], hi[8], lo[8];
      ld8(x, xdt, r * x_ld + c, f);
      if (gamma != nullptr) ld8(gamma, kF32, c, g);
      if (beta != nullptr) ld8(beta, kF32, c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = (f[j] - mean) * inv;
        if (gamma != nullptr) v *= g[j];
        if (beta != nullptr) v += b[j];
        hi[j] = __bfloat162float(__float2bfloat16_rn(v));
        lo[j] = v - hi[j];
      }
      const uint4 ph = pack8_bf16(hi), pl = pack8_bf16(lo);
      *reinterpret_cast<uint4*>(o + c) = ph;
      *reinterpret_cast<uint4*>(o + cols + c) = ph;
      *reinterpret_cast<uint4*>(o + 2LL * cols + c) = pl;
    }
  }
}

// out = a ∘ b, bf16, 8 elements per thread-trip (the backward of the random-feature map: dphi ∘ (-sin * scale))
__global__ void __launch_bounds__(kBlock) mul_bf16_kernel(const __nv_bfloat16* __restrict__ a, long long a_ld,
                                                          const __nv_bfloat16* __restrict__ b, long long b_ld,
                                                          __nv_bfloat16* __restrict__ out, long long out_ld, long long rows,
                                                          int cols) {
  const long long vpr = cols >> 3, n = rows * vpr;
  for (long long i = blockIdx.x * (long long)kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
    const long long r = i / vpr, c = (i - r * vpr) << 3;
    float fa[8], fb[8];
    ld8(a, kBF16, r * a_ld + c, fa);
    ld8(b, kBF16, r * b_ld + c, fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) fa[j] *= fb[j];
    *reinterpret_cast<uint4*>(out + r * out_ld + c) = pack8_bf16(fa);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Whole spectral-norm update of a small kernel in ONE launch: a cluster of 8 CTAs keeps W resident in distributed shared
// memory (CTA c owns in_dim / 8 rows), so W is read from HBM once and written once; both matvecs, both l2-normalisations,
// sigma and the scaling pass run out of shared memory, with cluster barriers + DSMEM reads for the cross-CTA sums.
// Replaces 7 launches (matvec_rows, l2_normalize, memset, matvec_cols, l2_normalize, sigma, scale_w) that are pure
// launch latency at the SNGP backbone's sizes (512 x 512 = 1 MB).  Same arithmetic order per element as those kernels
// up to the summation order of the dot products.
constexpr int kSnCluster = 8;
constexpr int kSnThreads = 512;

__device__ __forceinline__ float sn_block_sum(float v, float* red) {  // red: kSnThreads / 32 + 1 floats
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < kSnThreads / 32 ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[kSnThreads / 32] = t;
  }
  __syncthreads();
  const float r = red[kSnThreads / 32];
  __syncthreads();
  return r;
}

__global__ void __cluster_dims__(kSnCluster, 1, 1) __launch_bounds__(kSnThreads)
spectral_norm_cluster_kernel(const float* __restrict__ w, int in_dim, int out_dim, float* __restrict__ u,
                             float* __restrict__ v, int iterations, float c, int mode, int update_uv, float* w_out,
                             __nv_bfloat16* w_lp, long long lp_ld, float* sigma_out, float* scal) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float sm[];
  const int rank = static_cast<int>(cluster.block_rank());
  const int rows = in_dim / kSnCluster, r0 = rank * rows;
  float* ws = sm;                                  // [rows][out_dim]   this CTA's rows of W
  float* su = ws + (size_t)rows * out_dim;         // [out_dim]         current u
  float* part = su + out_dim;                      // [out_dim]         this CTA's partial of v W
  float* ucopy = part + out_dim;                   // [out_dim]         final u (kept while su receives v W)
  float* sv = ucopy + out_dim;                     // [rows]            this CTA's slice of v
  float* red = sv + rows;                          // block-sum scratch [17]
  float* xch = red + 32;                           // [8] cluster-visible scalars (partial sums of squares)
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

  for (int i = t * 4; i < rows * out_dim; i += kSnThreads * 4)
    *reinterpret_cast<float4*>(ws + i) = __ldg(reinterpret_cast<const float4*>(w + (size_t)r0 * out_dim + i));
  for (int o = t; o < out_dim; o += kSnThreads) su[o] = u[o];
  for (int i = t; i < rows; i += kSnThreads) sv[i] = v[r0 + i];
  __syncthreads();

  auto cluster_sum = [&](float local, int slot) -> float {  // sum of `local` (one value per CTA) over the cluster
    if (t == 0) xch[slot] = local;
    cluster.sync();
    float tot = 0.f;
    for (int q = 0; q < kSnCluster; ++q) tot += *cluster.map_shared_rank(xch + slot, q);
    cluster.sync();  // everyone has read before the slot is reused
    return tot;
  };
  // u_raw = v W for the current v: partial over this CTA's rows, then every CTA sums the partials of all CTAs (DSMEM)
  auto vw = [&]() {
    for (int o = t; o < out_dim; o += kSnThreads) {
      float acc = 0.f;
      for (int i = 0; i < rows; ++i) acc = fmaf(sv[i], ws[(size_t)i * out_dim + o], acc);
      part[o] = acc;
    }
    cluster.sync();
    for (int o = t; o < out_dim; o += kSnThreads) {
      float acc = 0.f;
      for (int q = 0; q < kSnCluster; ++q) acc += cluster.map_shared_rank(part, q)[o];
      su[o] = acc;  // u_raw (replicated in every CTA)
    }
    cluster.sync();  // partials consumed
  };

  const int iters = update_uv ? iterations : 0;
  for (int it = 0; it < iters; ++it) {
    // v_raw = u W^T (own rows), v = l2n(v_raw)
    float sq = 0.f;
    for (int i = warp; i < rows; i += kSnThreads / 32) {
      float acc = 0.f;
      for (int o = lane * 4; o < out_dim; o += 128) {
        const float4 a = *reinterpret_cast<const float4*>(ws + (size_t)i * out_dim + o);
        const float4 b = *reinterpret_cast<const float4*>(su + o);
        acc += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
      }
      acc = warp_sum(acc);
      if (lane == 0) { sv[i] = acc; sq += acc * acc; }
    }
    sq = sn_block_sum(sq, red);
    const float inv_v = rsqrtf(fmaxf(cluster_sum(sq, 0), 1e-12f));
    for (int i = t; i < rows; i += kSnThreads) sv[i] *= inv_v;
    __syncthreads();
    // u_raw = v W, u = l2n(u_raw)
    vw();
    float squ = 0.f;
    for (int o = t; o < out_dim; o += kSnThreads) squ += su[o] * su[o];
    squ = sn_block_sum(squ, red);  // su is replicated: every CTA computes the same total
    const float inv_u = rsqrtf(fmaxf(squ, 1e-12f));
    for (int o = t; o < out_dim; o += kSnThreads) su[o] *= inv_u;
    __syncthreads();
  }
  // sigma = v W u^T with the final u, v (normalization.py:381): keep u, recompute v W
  if (iters > 0) {
    if (rank == 0)
      for (int o = t; o < out_dim; o += kSnThreads) u[o] = su[o];
    for (int i = t; i < rows; i += kSnThreads) v[r0 + i] = sv[i];
  }
  for (int o = t; o < out_dim; o += kSnThreads) ucopy[o] = su[o];
  __syncthreads();
  vw();
  float acc_s = 0.f;
  for (int o = t; o < out_dim; o += kSnThreads) acc_s += ucopy[o] * su[o];
  const float sg = sn_block_sum(acc_s, red);
  float factor;
  if (mode == 0) factor = (c / sg) < 1.f ? (c / sg) : 1.f;   // normalization.py:387-390
  else factor = 1.f / fmaxf(sg / c, 1.f);                    // jax/nn/normalization.py:175-178
  if (rank == 0 && t == 0) {
    scal[0] = sg;
    scal[1] = factor;
    if (sigma_out != nullptr) *sigma_out = sg;
  }
  if (w_out != nullptr || w_lp != nullptr) {
    const float div = fmaxf(sg / c, 1.f);
    for (int i = t * 4; i < rows * out_dim; i += kSnThreads * 4) {
      float4 x = *reinterpret_cast<const float4*>(ws + i);
      float y[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] = mode == 0 ? (factor < 1.f ? factor * y[j] : y[j]) : y[j] / div;
      if (w_out != nullptr) *reinterpret_cast<float4*>(w_out + (size_t)r0 * out_dim + i) = make_float4(y[0], y[1], y[2], y[3]);
      if (w_lp != nullptr) {
        const int r = i / out_dim, cc = i - r * out_dim;
        __nv_bfloat16* p = w_lp + (size_t)(r0 + r) * lp_ld + cc;
        p[0] = __float2bfloat16_rn(y[0]); p[1] = __float2bfloat16_rn(y[1]);
        p[2] = __float2bfloat16_rn(y[2]); p[3] = __float2bfloat16_rn(y[3]);
      }
    }
  }
  cluster.sync();  // no CTA exits while a peer may still read its shared memory
}

size_t sn_cluster_smem(int in_dim, int out_dim) {
  const size_t rows = in_dim / kSnCluster;
  return sizeof(float) * (rows * out_dim + 3 * (size_t)out_dim + rows + 32 + 8);
}

int grid_for(long long n) {
  long long g = (n + kBlock - 1) / kBlock;
  if (g < 1) g = 1;
  if (g > 148 * 8) g = 148 * 8;
  return static_cast<int>(g);
}

}  // namespace

int k_spectral_norm(const float* w, int in_dim, int out_dim, float* u, float* v, int iterations, float c, int mode,
                    int update_uv, float* w_out, void* w_out_lp, long long w_lp_ld, float* sigma_out, float* ws,
                    cudaStream_t s) {
  // workspace: v_raw [in_dim] | u_raw [out_dim] | scal [8]
  float* v_raw = ws;
  float* u_raw = ws + in_dim;
  float* scal = u_raw + out_dim;
  // small kernels: one cluster launch with W resident in distributed shared memory
  static const bool fused_ok = [] { const char* e = std::getenv("ED2_SN_FUSED"); return e == nullptr || std::atoi(e) != 0; }();
  if (fused_ok && in_dim % kSnCluster == 0 && out_dim % 4 == 0 && sn_cluster_smem(in_dim, out_dim) <= 200 * 1024 &&
      (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (w_out == nullptr || (reinterpret_cast<uintptr_t>(w_out) & 15) == 0) &&
      (w_out_lp == nullptr || w_lp_ld % 4 == 0)) {
    const size_t smem = sn_cluster_smem(in_dim, out_dim);
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(spectral_norm_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr = true;
    }
    spectral_norm_cluster_kernel<<<kSnCluster, kSnThreads, smem, s>>>(w, in_dim, out_dim, u, v, iterations, c, mode, update_uv,
                                                                      w_out, reinterpret_cast<__nv_bfloat16*>(w_out_lp),
                                                                      w_lp_ld, sigma_out, scal);
    count_launch();
    return check_launch("spectral_norm(cluster)");
  }
  const int row_grid = std::min(148 * 4, (in_dim + 7) / 8);
  dim3 col_grid((out_dim + 127) / 128, (in_dim + kSlabRows - 1) / kSlabRows);
  int launches = 0;
  const int iters = update_uv ? iterations : 0;
  for (int it = 0; it < iters; ++it) {
    matvec_rows_kernel<<<row_grid, kBlock, 0, s>>>(w, in_dim, out_dim, u, v_raw);       // u Wᵀ
    l2_normalize_kernel<<<1, kBlock, 0, s>>>(v_raw, in_dim, v);
    cudaMemsetAsync(u_raw, 0, sizeof(float) * out_dim, s);
    matvec_cols_kernel<<<col_grid, kBlock, 0, s>>>(w, in_dim, out_dim, v, u_raw);       // v W
    l2_normalize_kernel<<<1, kBlock, 0, s>>>(u_raw, out_dim, u);
    launches += 4;
  }
  if (iters == 0) {
    cudaMemsetAsync(u_raw, 0, sizeof(float) * out_dim, s);
    matvec_cols_kernel<<<col_grid, kBlock, 0, s>>>(w, in_dim, out_dim, v, u_raw);
    ++launches;
  }
  sigma_kernel<<<1, kBlock, 0, s>>>(u, u_raw, out_dim, c, mode, scal, sigma_out);
  ++launches;
  if (w_out != nullptr || w_out_lp != nullptr) {
    scale_w_kernel<<<grid_for((long long)in_dim * out_dim), kBlock, 0, s>>>(
        w, in_dim, out_dim, scal, mode, c, w_out, reinterpret_cast<__nv_bfloat16*>(w_out_lp), w_lp_ld);
    ++launches;
  }
  count_launch(launches);
  return check_launch("spectral_norm");
}

int k_input_normalize(const void* x, int xdt, long long x_ld, void* y, int ydt, long long y_ld, long long rows,
                      int cols, const float* gamma, const float* beta, float eps, int scale_only, float scale,
                      cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return ED2_OK;
  const int g = static_cast<int>(std::min<long long>(148 * 8, (rows + 7) / 8));
  input_normalize_kernel<<<g, kBlock, 0, s>>>(x, xdt, x_ld, y, ydt, y_ld, rows, cols, gamma, beta, eps, scale_only, scale);
  count_launch();
  return check_launch("input_normalize");
}


int k_layernorm_split(const void* x, int xdt, long long x_ld, void* out, long long out_ld, long long rows, int cols,
                      const float* gamma, const float* beta, float eps, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return ED2_OK;
  if (cols % 8 != 0 || out_ld % 8 != 0 || x_ld % 8 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(ED2_ERR_BAD_ARG, "layernorm_split: needs cols %% 8 == 0 and 16-byte aligned rows");
  const int g = static_cast<int>(std::min<long long>(148 * 8, (rows + 7) / 8));
  layernorm_split_kernel<<<g, kBlock, 0, s>>>(x, xdt, x_ld, reinterpret_cast<__nv_bfloat16*>(out), out_ld, rows, cols, gamma, beta, eps);
  count_launch();
  return check_launch("layernorm_split");
}

int k_mul_bf16(const void* a, long long a_ld, const void* b, long long b_ld, void* out, long long out_ld, long long rows,
               int cols, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return ED2_OK;
  if (cols % 8 != 0 || a_ld % 8 != 0 || b_ld % 8 != 0 || out_ld % 8 != 0)
    return set_error(ED2_ERR_BAD_ARG, "mul_bf16: needs 8-element aligned rows");
  mul_bf16_kernel<<<grid_for(rows * (cols / 8)), kBlock, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(a), a_ld,
                                                                  reinterpret_cast<const __nv_bfloat16*>(b), b_ld,
                                                                  reinterpret_cast<__nv_bfloat16*>(out), out_ld, rows, cols);
  count_launch();
  return check_launch("mul_bf16");
}

int k_layernorm_bwd(const void* x, int xdt, long long x_ld, const void* dy, int dydt, long long dy_ld, long long rows,
                    int cols, const float* gamma, float eps, void* dx, int dxdt, long long dx_ld, float* dgamma,
                    float* dbeta, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return ED2_OK;
  if (cols > 8192) return set_error(ED2_ERR_BAD_SHAPE, "layernorm_bwd: feature dimension %d > 8192", cols);
  if (dgamma) cudaMemsetAsync(dgamma, 0, sizeof(float) * cols, s);
  if (dbeta) cudaMemsetAsync(dbeta, 0, sizeof(float) * cols, s);
  if (cols == 128 || cols == 256 || cols == 512 || cols == 1024) {
    // register-resident rows; a few rows per warp so that the per-block flush of the column sums stays small
    const int g = static_cast<int>(std::min<long long>(148LL * 2, (rows + 7) / 8));
#define ED2_LN_ROWS(NJ) layernorm_bwd_rows_kernel<NJ><<<g, kBlock, 0, s>>>(x, xdt, x_ld, dy, dydt, dy_ld, rows, gamma, eps, dx, dxdt, dx_ld, dgamma, dbeta)
    if (cols == 128) ED2_LN_ROWS(4);
    else if (cols == 256) ED2_LN_ROWS(8);
    else if (cols == 512) ED2_LN_ROWS(16);
    else ED2_LN_ROWS(32);
#undef ED2_LN_ROWS
    count_launch();
    return check_launch("layernorm_bwd");
  }
  const size_t smem = 2 * sizeof(float) * (size_t)cols;
  // one warp per row and latency-bound: as many blocks per SM as the per-block column accumulators leave room for
  const int per_sm = static_cast<int>(std::max<size_t>(2, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(smem, 1024))));
  const int g = static_cast<int>(std::min<long long>(148LL * per_sm, (rows + 7) / 8));
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(layernorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4 * 8192);
    attr = true;
  }
  layernorm_bwd_kernel<<<g, kBlock, smem, s>>>(x, xdt, x_ld, dy, dydt, dy_ld, rows, cols, gamma, eps, dx, dxdt, dx_ld, dgamma, dbeta);
  count_launch();
  return check_launch("layernorm_bwd");
}

// P[i][j] = P[j][i] for i < j: completes a matrix whose lower triangle was computed (SYRK); 32 x 32 tiles via smem
__global__ void mirror_lower_kernel(float* __restrict__ p, int n) {
  __shared__ float tile[32][33];
  const int bj = blockIdx.x, bi = blockIdx.y;  // source tile (rows bi*32.., cols bj*32..), only bj <= bi
  if (bj > bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    tile[r][tx] = (i < n && j < n) ? p[(long long)i * n + j] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bj * 32 + r, j = bi * 32 + tx;  // destination: transposed position
    if (i < n && j < n && i < j) p[(long long)i * n + j] = tile[tx][r];
  }
}

int k_mirror_lower(float* p, int n, cudaStream_t s) {
  const int nb = (n + 31) / 32;
  mirror_lower_kernel<<<dim3(nb, nb), kBlock, 0, s>>>(p, n);
  count_launch();
  return check_launch("mirror_lower");
}

int k_precision_blend(float* precision, const float* summed, long long n, float momentum, long long batch_total,
                      cudaStream_t s) {
  precision_blend_kernel<<<grid_for(n), kBlock, 0, s>>>(precision, summed, n, momentum, 1.f / (float)batch_total);
  count_launch();
  return check_launch("precision_blend");
}

int k_mean_field(const float* logits, const float* var, long long batch, int classes, float factor, int likelihood,
                 float var_scale, float* out, cudaStream_t s) {
  if (batch <= 0) return ED2_OK;
  mean_field_kernel<<<grid_for(batch * classes), kBlock, 0, s>>>(logits, var, batch, classes, factor, likelihood, var_scale, out);
  count_launch();
  return check_launch("mean_field");
}

int k_rff_bwd_prep(const void* dphi, int ddt, long long dphi_ld, const float* pre, long long pre_ld, const float* bias, float scale,
                   long long rows, int cols, void* gpre, int gdt, long long gpre_ld, cudaStream_t s) {
  if (rows <= 0) return ED2_OK;
  rff_bwd_prep_kernel<<<grid_for(rows * cols), kBlock, 0, s>>>(dphi, ddt, dphi_ld, pre, pre_ld, bias, scale, rows, cols, gpre, gdt, gpre_ld);
  count_launch();
  return check_launch("rff_bwd_prep");
}

}  // namespace ed2
