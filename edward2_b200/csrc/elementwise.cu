// Bandwidth-bound kernels of the stochastic-weight layers: Philox fills, casts, TrainableNormal sampling with
// the analytic KL fused in, KL forward/backward, fast-weight scaling and the column reductions of the backward
// passes.  All are coalesced, 16-byte vectorised where alignment allows, and use warp-shuffle + shared-memory
// reductions with one atomic per block (fp64 for the KL scalar).
#include "kernels.h"

#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "gemm_epilogue.h"
#include "peer.cuh"
#include "status.h"

namespace ed2 {
namespace {

constexpr int kBlock = 256;
constexpr int kMaxGrid = 148 * 8;

// Grid cap of the issue-bound sampling / finishing kernels.  They are meant to run on side streams BESIDE a GEMM
// (tensor cores), measured on B200: true co-residency with the persistent GEMM does not pay (the GEMM slows by as much as the side kernel gains), so the cap defaults to full occupancy; ED2_SIDE_BLOCKS_PER_SM lowers it for experiments.
int side_grid_cap() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("ED2_SIDE_BLOCKS_PER_SM");
    int b = e ? std::atoi(e) : 8;
    if (b < 1 || b > 8) b = 8;
    v = 148 * b;
  }
  return v;
}
}  // namespace

// ED2_KL_FUSED=0 restores the separate, libm-accurate KL pass of the bf16 layers (api.cu forks it beside the GEMM).
bool fused_kl_enabled() {
  static const bool v = [] {
    const char* e = std::getenv("ED2_KL_FUSED");
    return e == nullptr || std::atoi(e) != 0;
  }();
  return v;
}

namespace {
constexpr float kSoftplusEps = 1e-7f;  // keras.backend.epsilon(), edward2/tensorflow/constraints.py:56-63

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float ld1(const void* p, int dt, long long i) {
  return dt == kBF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                     : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st1(void* p, int dt, long long i, float v) {
  if (dt == kBF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[i] = v;
}

// 4 consecutive elements starting at element offset i; vectorised when the address is 4-element aligned.
__device__ __forceinline__ void ld4(const void* p, int dt, long long i, int valid, float (&f)[4]) {
  if (dt == kBF16) {
    const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(p) + i;
    if (valid == 4 && (reinterpret_cast<uintptr_t>(q) & 7) == 0) {
      uint2 r = *reinterpret_cast<const uint2*>(q);
      float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.x));
      float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.y));
      f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) f[j] = j < valid ? __bfloat162float(q[j]) : 0.f;
    }
  } else {
    const float* q = reinterpret_cast<const float*>(p) + i;
    if (valid == 4 && (reinterpret_cast<uintptr_t>(q) & 15) == 0) {
      float4 r = *reinterpret_cast<const float4*>(q);
      f[0] = r.x; f[1] = r.y; f[2] = r.z; f[3] = r.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) f[j] = j < valid ? q[j] : 0.f;
    }
  }
}
__device__ __forceinline__ void st4(void* p, int dt, long long i, int valid, const float (&f)[4]) {
  if (dt == kBF16) {
    __nv_bfloat16* q = reinterpret_cast<__nv_bfloat16*>(p) + i;
    if (valid == 4 && (reinterpret_cast<uintptr_t>(q) & 7) == 0) {
      __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
      uint2 r;
      r.x = *reinterpret_cast<uint32_t*>(&a);
      r.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(q) = r;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < valid) q[j] = __float2bfloat16_rn(f[j]);
    }
  } else {
    float* q = reinterpret_cast<float*>(p) + i;
    if (valid == 4 && (reinterpret_cast<uintptr_t>(q) & 15) == 0) {
      *reinterpret_cast<float4*>(q) = make_float4(f[0], f[1], f[2], f[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < valid) q[j] = f[j];
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum -> one fp64 atomic
__device__ __forceinline__ void block_atomic_add(double* out, float v, float scale) {
  __shared__ float red[kBlock / 32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < kBlock / 32 ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(out, static_cast<double>(t) * static_cast<double>(scale));
  }
}

int grid_for(long long work_items) {
  long long g = (work_items + kBlock - 1) / kBlock;
  if (g < 1) g = 1;
  if (g > kMaxGrid) g = kMaxGrid;
  return static_cast<int>(g);
}

// ------------------------------------------------------------------------------------------- Philox fills
template <int kKind>
__global__ void philox_fill_kernel(float* out, long long n, uint64_t seed, uint64_t stream) {
  const long long nblk = (n + 3) >> 2;
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < nblk; b += (long long)gridDim.x * blockDim.x) {
    float e[4];
    if (kKind == 0) philox_normal4(seed, stream, b, e);
    else if (kKind == 3) philox_normal4<true>(seed, stream, b, e);  // hardware-approximate Box-Muller (bf16 paths)
    else if (kKind == 1) philox_uniform4(seed, stream, b, e);
    else philox_sign4(seed, stream, static_cast<uint64_t>(b) << 2, e);
    const long long i = b << 2;
    const int valid = static_cast<int>(min(4LL, n - i));
    st4(out, kF32, i, valid, e);
  }
}

__global__ void philox_raw_kernel(uint32_t* out, long long nblk, uint64_t seed, uint64_t stream) {
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < nblk; b += (long long)gridDim.x * blockDim.x) {
    uint32_t x[4];
    philox_block(seed, stream, b, x);
    reinterpret_cast<uint4*>(out)[b] = make_uint4(x[0], x[1], x[2], x[3]);
  }
}

// ------------------------------------------------------------------------------------------- cast
__global__ void cast2d_kernel(const void* src, int sdt, long long sld, void* dst, int ddt, long long dld,
                              long long rows, long long cols) {
  const long long gpr = (cols + 3) >> 2;  // groups per row
  const long long total = rows * gpr;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long r = g / gpr, c = (g - r * gpr) << 2;
    const int valid = static_cast<int>(min(4LL, cols - c));
    float f[4];
    ld4(src, sdt, r * sld + c, valid, f);
    st4(dst, ddt, r * dld + c, valid, f);
  }
}


// fp32 -> split-bf16 operand for an fp32-grade tensor-core product: x = h + m + l with h = bf16(x), m = bf16(x - h),
// l = bf16(x - h - m) (24 mantissa bits).  The three-term product sum_{i+j<=2} a_i b_j is evaluated by ONE bf16 GEMM
// whose contraction axis is the concatenation   A' = [h|h|h|m|m|l],  B' = [h|m|l|h|m|h]   (K' = 6K; nsplit = 2 keeps
// the 3 leading terms: A' = [h|h|m], B' = [h|m|h], 16 mantissa bits).  fp32 accumulation happens in TMEM.
__global__ void split_bf16_kernel(const void* __restrict__ src, int sdt, long long sld, __nv_bfloat16* __restrict__ dst,
                                  long long dld, long long rows, int cols, int is_b, int nsplit) {
  const long long n = rows * cols;
  const int terms = nsplit == 3 ? 6 : 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    const float x = ld1(src, sdt, r * sld + c);
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const __nv_bfloat16 l = __float2bfloat16_rn(r1 - __bfloat162float(m));
    __nv_bfloat16* d = dst + r * dld + c;
    if (nsplit == 3) {
      const __nv_bfloat16 a6[6] = {h, h, h, m, m, l}, b6[6] = {h, m, l, h, m, h};
#pragma unroll
      for (int t = 0; t < 6; ++t) d[(long long)t * cols] = is_b ? b6[t] : a6[t];
    } else {
      const __nv_bfloat16 a3[3] = {h, h, m}, b3[3] = {h, m, h};
#pragma unroll
      for (int t = 0; t < 3; ++t) d[(long long)t * cols] = is_b ? b3[t] : a3[t];
    }
    (void)terms;
  }
}

// ------------------------------------------------------------------------------------------- sampling + KL
__device__ __forceinline__ float kl_elem(float mu, float sigma, float pm, float ps) {
  const float d = mu - pm;
  return logf(ps / sigma) + (sigma * sigma + d * d) / (2.f * ps * ps) - 0.5f;
}

// Lean per-element math of the sampling / gradient passes (they are issue-bound, not HBM-bound: Philox + Box-Muller
// + softplus + KL is ~100 instructions per weight):  e = exp(rho), sigma = log1p(e) + eps, and from the same e
// sigmoid(rho) = e / (1 + e).  KL uses log(ps) - log(sigma) with the hardware lg2 (abs err 2^-21 near 1, 2 ulp away).
__device__ __forceinline__ float sigma_from_rho(float rho, float& e) {
  e = __expf(fminf(rho, 80.f));
  return (rho > 20.f ? rho : log1pf(e)) + kSoftplusEps;
}
// bf16-output variant: sigma only needs ~3 significant digits, so softplus is lg2(1 + ex2(rho)) on the SFU.
__device__ __forceinline__ float sigma_from_rho_fast(float rho, float& e) {
  e = __expf(fminf(rho, 80.f));
  const float sp = e < 2.44140625e-4f ? e : __logf(1.f + e);
  return (rho > 20.f ? rho : sp) + kSoftplusEps;
}
__device__ __forceinline__ float kl_fast(float mu, float sigma, float pm, float log_ps, float inv_2s2) {
  const float d = mu - pm;
  return log_ps - __logf(sigma) + (sigma * sigma + d * d) * inv_2s2 - 0.5f;
}

// Thread = 4 consecutive columns of one row (= one Philox block when cols % 4 == 0); blockIdx.y strides over rows,
// so no index division is needed for the pitched output.
template <bool kWantKl, bool kFast>
__global__ void __launch_bounds__(kBlock) sample_weights_kernel(const float* __restrict__ mu,
                                                                const float* __restrict__ rho, Noise eps, int rows,
                                                                int cols, int mode, void* w, int wdt, long long wld,
                                                                void* mean_out, long long mean_ld, double* kl_out,
                                                                float kl_scale, Prior p) {
  const int c = (blockIdx.x * kBlock + threadIdx.x) * 4;
  const int valid = min(4, cols - c);
  const float log_ps = __logf(p.stddev), inv_2s2 = 0.5f / (p.stddev * p.stddev);
  const long long n = (long long)rows * cols;
  float kl = 0.f;
  if (valid > 0) {
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
      const long long i = (long long)r * cols + c;
      float m[4], rh[4], e[4], o[4];
      ld4(mu, kF32, i, valid, m);
      ld4(rho, kF32, i, valid, rh);
      if ((i & 3) == 0) noise4<0, kFast>(eps, i, n, e);
      else
        for (int j = 0; j < 4; ++j) e[j] = j < valid ? noise1<0>(eps, i + j) : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float ex;
        const float sigma = kFast ? sigma_from_rho_fast(rh[j], ex) : sigma_from_rho(rh[j], ex);
        o[j] = mode == 0 ? fmaf(sigma, e[j], m[j]) : sigma * e[j];
        if (kWantKl && j < valid) kl += kl_fast(m[j], sigma, p.mean, log_ps, inv_2s2);
      }
      st4(w, wdt, (long long)r * wld + c, valid, o);
      if (mean_out != nullptr) st4(mean_out, wdt, (long long)r * mean_ld + c, valid, m);
    }
  }
  if (kWantKl) block_atomic_add(kl_out, kl, kl_scale);
}


// ---- lean hot-path kernels (bf16 compute mode, Philox noise, 8-element-aligned rows): straight-line code, 8 weights
// per thread (two Philox blocks), no alignment / tail / injected-noise branches.  ~40 instructions and 4 MUFU ops per
// weight instead of ~125: the sampling pass sits on the forward critical path (the GEMM needs W).
__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int kMode, bool kMean, bool kKl>
__global__ void __launch_bounds__(kBlock) sample_bf16_lean_kernel(const float* __restrict__ mu,
                                                                  const float* __restrict__ rho, uint64_t seed,
                                                                  uint64_t stream, const uint64_t* step, int rows,
                                                                  int cols, __nv_bfloat16* __restrict__ w, int wld,
                                                                  __nv_bfloat16* __restrict__ mean_out, int mean_ld,
                                                                  double* kl_out, float kl_scale, Prior p) {
  const int c = (blockIdx.x * kBlock + threadIdx.x) * 8;
  const bool active = c < cols;
  if (!kKl && !active) return;
  const uint64_t key = step != nullptr ? seed + __ldg(reinterpret_cast<const unsigned long long*>(step)) * 0x9E3779B97F4A7C15ull : seed;
  const float log_ps = __logf(p.stddev), inv_2s2 = 0.5f / (p.stddev * p.stddev);
  float kl = 0.f;
  for (int r = blockIdx.y; active && r < rows; r += gridDim.y) {
    const long long i = (long long)r * cols + c;
    const float4 m0 = __ldg(reinterpret_cast<const float4*>(mu + i)), m1 = __ldg(reinterpret_cast<const float4*>(mu + i) + 1);
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(rho + i)), r1 = __ldg(reinterpret_cast<const float4*>(rho + i) + 1);
    float e[8];
    philox_normal8<true>(key, stream, static_cast<uint64_t>(i >> 3), e);  // i % 8 == 0: one Philox block
    const float m[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    const float rh[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float ex;
      const float sigma = sigma_from_rho_fast(rh[j], ex);
      o[j] = kMode == 0 ? fmaf(sigma, e[j], m[j]) : sigma * e[j];
      if (kKl) kl += kl_fast(m[j], sigma, p.mean, log_ps, inv_2s2);  // same mu / sigma registers: one lg2 + 4 FMA more
    }
    uint4 q;
    q.x = pack2_bf16(o[0], o[1]); q.y = pack2_bf16(o[2], o[3]); q.z = pack2_bf16(o[4], o[5]); q.w = pack2_bf16(o[6], o[7]);
    *reinterpret_cast<uint4*>(w + (long long)r * wld + c) = q;
    if (kMean) {
      uint4 pk;
      pk.x = pack2_bf16(m[0], m[1]); pk.y = pack2_bf16(m[2], m[3]); pk.z = pack2_bf16(m[4], m[5]); pk.w = pack2_bf16(m[6], m[7]);
      *reinterpret_cast<uint4*>(mean_out + (long long)r * mean_ld + c) = pk;
    }
  }
  if (kKl) block_atomic_add(kl_out, kl, kl_scale);
}

// dmu = d_mean + klg (mu - m) / s^2 ;  drho = (d_pert * eps + klg (sigma / s^2 - 1 / sigma)) * sigmoid(rho)
__global__ void __launch_bounds__(kBlock) grad_finish_lean_kernel(const float* d_mean, const float* d_pert,
                                                                  const float* __restrict__ mu,
                                                                  const float* __restrict__ rho, uint64_t seed,
                                                                  uint64_t stream, const uint64_t* step, long long n,
                                                                  float klg_in, const double* kl_upstream, Prior p,
                                                                  float* dmu, float* drho, __nv_bfloat16* dmu_lp,
                                                                  __nv_bfloat16* drho_lp) {
  const float klg = klg_in * (kl_upstream != nullptr ? static_cast<float>(*kl_upstream) : 1.f);
  const float inv_s2 = 1.f / (p.stddev * p.stddev);
  const uint64_t key = step != nullptr ? seed + __ldg(reinterpret_cast<const unsigned long long*>(step)) * 0x9E3779B97F4A7C15ull : seed;
  const long long nblk = n >> 2;  // n % 4 == 0 (checked by the launcher)
  for (long long b = blockIdx.x * (long long)kBlock + threadIdx.x; b < nblk; b += (long long)gridDim.x * kBlock) {
    const float4 m = __ldg(reinterpret_cast<const float4*>(mu) + b), r = __ldg(reinterpret_cast<const float4*>(rho) + b);
    const float4 gm = reinterpret_cast<const float4*>(d_mean)[b], gp = reinterpret_cast<const float4*>(d_pert)[b];
    float e[4];
    philox_normal4<true>(key, stream, static_cast<uint64_t>(b), e);
    const float mm[4] = {m.x, m.y, m.z, m.w}, rr[4] = {r.x, r.y, r.z, r.w};
    const float g1[4] = {gm.x, gm.y, gm.z, gm.w}, g2[4] = {gp.x, gp.y, gp.z, gp.w};
    float om[4], orr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float ex;
      const float sigma = sigma_from_rho_fast(rr[j], ex);
      const float sgm = __fdividef(ex, 1.f + ex);
      om[j] = g1[j] + klg * (mm[j] - p.mean) * inv_s2;
      orr[j] = (g2[j] * e[j] + klg * (sigma * inv_s2 - __fdividef(1.f, sigma))) * sgm;
    }
    reinterpret_cast<float4*>(dmu)[b] = make_float4(om[0], om[1], om[2], om[3]);
    reinterpret_cast<float4*>(drho)[b] = make_float4(orr[0], orr[1], orr[2], orr[3]);
    if (dmu_lp != nullptr) {  // bf16 copies for the all-reduce buckets, written while the values are in registers
      reinterpret_cast<uint2*>(dmu_lp)[b] = make_uint2(pack2_bf16(om[0], om[1]), pack2_bf16(om[2], om[3]));
      reinterpret_cast<uint2*>(drho_lp)[b] = make_uint2(pack2_bf16(orr[0], orr[1]), pack2_bf16(orr[2], orr[3]));
    }
  }
}

// Finishing pass fed by a bf16 weight gradient (the all-reduced bucket of a data-parallel step).
template <bool kFast>
__global__ void __launch_bounds__(kBlock) grad_finish_bf16_kernel(const __nv_bfloat16* __restrict__ dw,
                                                                  const float* __restrict__ mu,
                                                                  const float* __restrict__ rho, Noise eps, long long n,
                                                                  float klg_in, const double* kl_upstream, Prior p,
                                                                  float* dmu, float* drho) {
  const float klg = klg_in * (kl_upstream != nullptr ? static_cast<float>(*kl_upstream) : 1.f);
  const float inv_s2 = 1.f / (p.stddev * p.stddev);
  const long long nblk = (n + 3) >> 2;
  for (long long b = blockIdx.x * (long long)kBlock + threadIdx.x; b < nblk; b += (long long)gridDim.x * kBlock) {
    const long long i = b << 2;
    const int valid = static_cast<int>(min(4LL, n - i));
    float m[4], r[4], g[4], e[4], om[4], orr[4];
    ld4(mu, kF32, i, valid, m);
    ld4(rho, kF32, i, valid, r);
    ld4(dw, kBF16, i, valid, g);
    noise4<0, kFast>(eps, i, n, e);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float ex;
      const float sigma = kFast ? sigma_from_rho_fast(r[j], ex) : sigma_from_rho(r[j], ex);
      const float sgm = __fdividef(ex, 1.f + ex);
      om[j] = g[j] + klg * (m[j] - p.mean) * inv_s2;
      orr[j] = (g[j] * e[j] + klg * (sigma * inv_s2 - __fdividef(1.f, sigma))) * sgm;
    }
    st4(dmu, kF32, i, valid, om);
    st4(drho, kF32, i, valid, orr);
  }
}

// Finishing pass of the peer-memory exchange (peer.cuh): the reduced bf16 gradient is read slice by slice from the
// memory of the rank that owns (reduced) it — the all-gather half of the all-reduce never exists as a separate pass.
template <bool kFast>
__global__ void __launch_bounds__(kBlock, 6) grad_finish_peer_kernel(PeerExchange x, const float* __restrict__ mu,
                                                                     const float* __restrict__ rho, Noise eps,
                                                                     float klg_in, const double* kl_upstream, Prior p,
                                                                     float* dmu, float* drho, float* small_out) {
  unsigned long long* lw = peer::local_words(x);
  const unsigned long long epoch = lw[0];
  if (threadIdx.x < x.world)
    peer::wait_flag(peer::reduced_flags(x, x.rank) + threadIdx.x, epoch, lw + 2, 0x200 + threadIdx.x);
  __syncthreads();
  const float klg = klg_in * (kl_upstream != nullptr ? static_cast<float>(*kl_upstream) : 1.f);
  const float inv_s2 = 1.f / (p.stddev * p.stddev);
  const uint64_t key = noise_key(eps);
  const long long nblk = x.n >> 2;  // n is a multiple of 8; thread = 4 weights = one Philox block = one 8-byte load
  const unsigned chunk4 = static_cast<unsigned>(x.chunk >> 2);
  // half or more of the gradient loads cross NVLink (2-3 us): two of them stay in flight per thread ahead of the math
  const long long stride = (long long)gridDim.x * kBlock;
  auto load_dw = [&](long long b) -> uint2 {
    if (b >= nblk) return make_uint2(0u, 0u);
    const int owner = static_cast<int>(static_cast<unsigned>(b) / chunk4);
    return peer::ld_sys_v2(reinterpret_cast<const __nv_bfloat16*>(x.reduced[owner]) + (b << 2));
  };
  const long long b0 = blockIdx.x * (long long)kBlock + threadIdx.x;
  uint2 raw = load_dw(b0), raw1 = load_dw(b0 + stride);
  for (long long b = b0; b < nblk; b += stride) {
    const uint2 raw2 = load_dw(b + 2 * stride);
    const float4 m4 = __ldg(reinterpret_cast<const float4*>(mu) + b), r4 = __ldg(reinterpret_cast<const float4*>(rho) + b);
    float e[4];
    philox_normal4<kFast>(key, eps.stream, static_cast<uint64_t>(b), e);
    const float2 ga = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
    const float2 gb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
    const float m[4] = {m4.x, m4.y, m4.z, m4.w}, r[4] = {r4.x, r4.y, r4.z, r4.w}, g[4] = {ga.x, ga.y, gb.x, gb.y};
    float om[4], orr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float ex;
      const float sigma = kFast ? sigma_from_rho_fast(r[j], ex) : sigma_from_rho(r[j], ex);
      const float sgm = __fdividef(ex, 1.f + ex);
      om[j] = g[j] + klg * (m[j] - p.mean) * inv_s2;
      orr[j] = (g[j] * e[j] + klg * (sigma * inv_s2 - __fdividef(1.f, sigma))) * sgm;
    }
    reinterpret_cast<float4*>(dmu)[b] = make_float4(om[0], om[1], om[2], om[3]);
    reinterpret_cast<float4*>(drho)[b] = make_float4(orr[0], orr[1], orr[2], orr[3]);
    raw = raw1;
    raw1 = raw2;
  }
  // companion vector (bias gradient): one-shot sum over ranks, same order on every rank
  if (small_out != nullptr) {
    for (long long j = blockIdx.x * (long long)kBlock + threadIdx.x; j < x.small_n; j += (long long)gridDim.x * kBlock) {
      float s = 0.f;
      for (int q = 0; q < x.world; ++q) s += peer::ld_sys_f32(x.small[q] + (epoch & 1ull) * x.small_n + j);
      small_out[j] = s;
    }
  }
}

__global__ void normal_kl_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ rho, long long n, Prior p,
                                     float scale, double* out) {
  const long long nblk = (n + 3) >> 2;
  float kl = 0.f;
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < nblk; b += (long long)gridDim.x * blockDim.x) {
    const long long i = b << 2;
    const int valid = static_cast<int>(min(4LL, n - i));
    float m[4], r[4];
    ld4(mu, kF32, i, valid, m);
    ld4(rho, kF32, i, valid, r);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < valid) kl += kl_elem(m[j], softplus_f(r[j]) + kSoftplusEps, p.mean, p.stddev);
  }
  block_atomic_add(out, kl, scale);
}

__global__ void normal_kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ rho, long long n, Prior p,
                                     float scale, const float* upstream, float* dmu, float* drho) {
  const float g = (upstream != nullptr ? *upstream : 1.f) * scale;
  const float inv_s2 = 1.f / (p.stddev * p.stddev);
  const long long nblk = (n + 3) >> 2;
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < nblk; b += (long long)gridDim.x * blockDim.x) {
    const long long i = b << 2;
    const int valid = static_cast<int>(min(4LL, n - i));
    float m[4], r[4], a[4], c[4];
    ld4(mu, kF32, i, valid, m);
    ld4(rho, kF32, i, valid, r);
    ld4(dmu, kF32, i, valid, a);
    ld4(drho, kF32, i, valid, c);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float sigma = softplus_f(r[j]) + kSoftplusEps;
      a[j] += g * (m[j] - p.mean) * inv_s2;
      c[j] += g * (sigma * inv_s2 - 1.f / sigma) * sigmoid_f(r[j]);
    }
    st4(dmu, kF32, i, valid, a);
    st4(drho, kF32, i, valid, c);
  }
}

template <bool kFast>
__global__ void __launch_bounds__(kBlock) normal_grad_finish_kernel(const float* d_mean, const float* d_pert,
                                                                    const float* __restrict__ mu,
                                                                    const float* __restrict__ rho, Noise eps,
                                                                    long long n, float klg_in, const double* kl_upstream,
                                                                    Prior p, float* dmu, float* drho) {
  const float klg = klg_in * (kl_upstream != nullptr ? static_cast<float>(*kl_upstream) : 1.f);
  const float inv_s2 = 1.f / (p.stddev * p.stddev);
  const long long nblk = (n + 3) >> 2;
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < nblk; b += (long long)gridDim.x * blockDim.x) {
    const long long i = b << 2;
    const int valid = static_cast<int>(min(4LL, n - i));
    float m[4], r[4], gm[4], gp[4], e[4], om[4], orr[4];
    ld4(mu, kF32, i, valid, m);
    ld4(rho, kF32, i, valid, r);
    ld4(d_mean, kF32, i, valid, gm);
    ld4(d_pert, kF32, i, valid, gp);
    noise4<0, kFast>(eps, i, n, e);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float ex;
      const float sigma = kFast ? sigma_from_rho_fast(r[j], ex) : sigma_from_rho(r[j], ex);
      const float sgm = __fdividef(ex, 1.f + ex);  // sigmoid(rho)
      om[j] = gm[j] + klg * (m[j] - p.mean) * inv_s2;
      orr[j] = (gp[j] * e[j] + klg * (sigma * inv_s2 - __fdividef(1.f, sigma))) * sgm;
    }
    st4(dmu, kF32, i, valid, om);
    st4(drho, kF32, i, valid, orr);
  }
}

// ------------------------------------------------------------------------------------------- fast weights
constexpr float kClip = 10.f;  // reference default, edward2/tensorflow/layers/dense.py:1025-1026 (lean kernels)

// factor for 4 consecutive columns c..c+3 of row `row` (member e)
template <int kMode>
__device__ __forceinline__ void fast_factor4(const float* fw_mu, const float* fw_rho, const Noise& nz, long long row,
                                             int e, int c, int cols, int valid, float (&f)[4], float (&raw)[4],
                                             float (&en)[4], bool fast = false, float lo = -kClip, float hi = kClip) {
  if (kMode == 2) {
    const long long idx = noise_idx(nz, row, cols, c);
    if ((idx & 3) == 0) noise4<1>(nz, idx, (long long)1 << 62, f);
    else
      for (int j = 0; j < 4; ++j) f[j] = j < valid ? noise1<1>(nz, idx + j) : 0.f;
    return;
  }
  float m[4];
  ld4(fw_mu, kF32, (long long)e * cols + c, valid, m);
  if (kMode == 0 || fw_rho == nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[j] = m[j]; raw[j] = m[j]; en[j] = 0.f; }
    return;
  }
  float r[4];
  ld4(fw_rho, kF32, (long long)e * cols + c, valid, r);
  const long long idx = noise_idx(nz, row, cols, c);
  if ((idx & 3) == 0) {
    // bf16 compute mode: hardware-approximate Box-Muller (the factor is rounded to bf16 before use)
    if (fast) noise4<0, true>(nz, idx, (long long)1 << 62, en);
    else noise4<0>(nz, idx, (long long)1 << 62, en);
  }
  else
    for (int j = 0; j < 4; ++j) en[j] = j < valid ? noise1<0>(nz, idx + j) : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    raw[j] = fmaf(softplus_f(r[j]) + kSoftplusEps, en[j], m[j]);
    f[j] = fminf(fmaxf(raw[j], lo), hi);
  }
}

template <int kMode>
__global__ void scale_rows_kernel(const void* x, int dt, long long x_ld, long long rows, int cols, int rpm,
                                  const float* fw_mu, const float* fw_rho, Noise nz, void* xo, long long xo_ld,
                                  void* f_out, long long f_ld, FastOpts fo) {
  const long long gpr = (cols + 3) >> 2;
  const long long total = rows * gpr;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long r = g / gpr;
    const int c = static_cast<int>(g - r * gpr) << 2;
    const int valid = min(4, cols - c);
    const int e = static_cast<int>(r / rpm);
    float f[4], raw[4], en[4];
    fast_factor4<kMode>(fw_mu, fw_rho, nz, r, e, c, cols, valid, f, raw, en, dt == kBF16, fo.clip_lo, fo.clip_hi);
    if (x != nullptr) {
      float xv[4];
      ld4(x, dt, r * x_ld + c, valid, xv);
#pragma unroll
      for (int j = 0; j < 4; ++j) xv[j] = fo.additive ? xv[j] + f[j] : xv[j] * f[j];
      st4(xo, dt, r * xo_ld + c, valid, xv);
    }
    if (f_out != nullptr) st4(f_out, dt, r * f_ld + c, valid, f);
  }
}


// ---- lean rank-1 kernels (bf16, Philox noise, 8-aligned): one thread owns 8 consecutive columns and walks down a
// chunk of rows of ONE member, so the per-(member, column) quantities (mu, sigma, sigmoid(rho)) are computed once,
// the noise is two Philox blocks per row, every access is a 16-byte vector, and the column reductions of the backward
// passes are per-thread accumulators flushed with one atomic per column per block (no shared memory).
constexpr int kLeanRows = 64;  // upper bound; the host shrinks it until the grid fills the GPU (lean_rows_for)
inline int lean_rows_for(long long rows, int cols) {
  const long long col_blocks = (cols / 8 + kBlock - 1) / kBlock;
  int r = kLeanRows;
  // one column block (<= 2048 columns, e.g. the 1000-unit head): every row chunk ends in one atomic per column per
  // accumulator on the SAME few thousand addresses, so fewer, longer chunks win (one block per SM is enough)
  const long long want = col_blocks == 1 ? 148 : 148 * 6;
  while (r > 4 && (rows / r) * col_blocks < want) r >>= 1;
  return r;
}

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 v = __bfloat1622float2(h[t]);
    f[2 * t] = v.x;
    f[2 * t + 1] = v.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 q;
  q.x = pack2_bf16(f[0], f[1]); q.y = pack2_bf16(f[2], f[3]); q.z = pack2_bf16(f[4], f[5]); q.w = pack2_bf16(f[6], f[7]);
  return q;
}
__device__ __forceinline__ void normal8_fast(uint64_t key, uint64_t stream, long long idx, float (&e)[8]) {
  if ((idx & 7) == 0) {
    philox_normal8<true>(key, stream, static_cast<uint64_t>(idx >> 3), e);
  } else {  // idx % 4 == 0: halves of two blocks
    float t[4];
    philox_normal4<true>(key, stream, static_cast<uint64_t>(idx >> 2), t);
    e[0] = t[0]; e[1] = t[1]; e[2] = t[2]; e[3] = t[3];
    philox_normal4<true>(key, stream, static_cast<uint64_t>(idx >> 2) + 1, t);
    e[4] = t[0]; e[5] = t[1]; e[6] = t[2]; e[7] = t[3];
  }
}
__device__ __forceinline__ uint64_t lean_key(uint64_t seed, const uint64_t* step) {
  return step != nullptr ? seed + __ldg(reinterpret_cast<const unsigned long long*>(step)) * 0x9E3779B97F4A7C15ull : seed;
}
struct LeanCols {
  float m[8], sg[8], sgm[8];
};
__device__ __forceinline__ void load_cols(const float* mu, const float* rho, int e, int c, int cols, LeanCols& k) {
  const float4 a0 = __ldg(reinterpret_cast<const float4*>(mu + (long long)e * cols + c));
  const float4 a1 = __ldg(reinterpret_cast<const float4*>(mu + (long long)e * cols + c) + 1);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(rho + (long long)e * cols + c));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(rho + (long long)e * cols + c) + 1);
  const float mm[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  const float rr[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    k.m[j] = mm[j];
    k.sg[j] = softplus_f(rr[j]) + kSoftplusEps;
    k.sgm[j] = sigmoid_f(rr[j]);
  }
}

// xo = x ∘ clip(mu[e] + sigma[e] eps, -10, 10)   (x == nullptr: write the factor itself)
template <bool kHasX>
__global__ void __launch_bounds__(kBlock) rank1_scale_lean_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int rpm,
                                                                  int cols, const float* __restrict__ fw_mu,
                                                                  const float* __restrict__ fw_rho, Noise nz,
                                                                  __nv_bfloat16* __restrict__ out, int out_ld,
                                                                  int lean_rows) {
  const int c = (blockIdx.x * kBlock + threadIdx.x) * 8;
  if (c >= cols) return;
  const int chunks = (rpm + lean_rows - 1) / lean_rows;
  const int e = blockIdx.y / chunks;
  const long long row0 = (long long)e * rpm + (long long)(blockIdx.y - e * chunks) * lean_rows;
  const long long row1 = min((long long)(e + 1) * rpm, row0 + lean_rows);
  const uint64_t key = lean_key(nz.seed, nz.step);
  LeanCols k;
  load_cols(fw_mu, fw_rho, e, c, cols, k);
  for (long long r = row0; r < row1; ++r) {
    float en[8], f[8];
    normal8_fast(key, nz.stream, noise_row(nz, r) * cols + c, en);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fminf(fmaxf(fmaf(k.sg[j], en[j], k.m[j]), -kClip), kClip);
    if (kHasX) {
      float xv[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + r * x_ld + c)), xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= xv[j];
    }
    *reinterpret_cast<uint4*>(out + r * out_ld + c) = pack8(f);
  }
}

// rank-1 output side: gp = gy ∘ relu'(y); dz = gp ∘ s; ds = gp ∘ z ∘ 1[unclipped]; dmu_s, drho_s, dbias column sums
__global__ void __launch_bounds__(kBlock) rank1_bwd_out_lean_kernel(BwdOutArgs a) {
  const int c = (blockIdx.x * kBlock + threadIdx.x) * 8;
  if (c >= a.cols) return;
  const int chunks = (a.rows_per_member + a.lean_rows - 1) / a.lean_rows;
  const int e = blockIdx.y / chunks;
  const long long row0 = (long long)e * a.rows_per_member + (long long)(blockIdx.y - e * chunks) * a.lean_rows;
  const long long row1 = min((long long)(e + 1) * a.rows_per_member, row0 + a.lean_rows);
  const uint64_t key = lean_key(a.eps_s.seed, a.eps_s.step);
  LeanCols k;
  load_cols(a.mu_s, a.rho_s, e, c, a.cols, k);
  const __nv_bfloat16* gy = reinterpret_cast<const __nv_bfloat16*>(a.gy);
  const __nv_bfloat16* y = reinterpret_cast<const __nv_bfloat16*>(a.y);
  const __nv_bfloat16* z = reinterpret_cast<const __nv_bfloat16*>(a.z);
  const __nv_bfloat16* sb = reinterpret_cast<const __nv_bfloat16*>(a.s_buf);
  __nv_bfloat16* dz = reinterpret_cast<__nv_bfloat16*>(a.dz);
  float acc_mu[8] = {}, acc_rho[8] = {}, acc_b[8] = {};
  for (long long r = row0; r < row1; ++r) {
    float g[8], yv[8], zv[8], sv[8], en[8], o[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(gy + r * a.gy_ld + c)), g);
    unpack8(__ldg(reinterpret_cast<const uint4*>(z + r * a.z_ld + c)), zv);
    if (sb != nullptr) unpack8(__ldg(reinterpret_cast<const uint4*>(sb + r * a.s_ld + c)), sv);
    if (a.act == kActRelu) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(y + r * a.y_ld + c)), yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.f ? g[j] : 0.f;
    }
    normal8_fast(key, a.eps_s.stream, noise_row(a.eps_s, r) * a.cols + c, en);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float raw = fmaf(k.sg[j], en[j], k.m[j]);
      const float ds = (raw < -kClip || raw > kClip) ? 0.f : g[j] * zv[j];
      acc_b[j] += g[j];
      acc_mu[j] += ds;
      acc_rho[j] += ds * en[j] * k.sgm[j];
      // s was not saved (fused forward): it is the clipped draw itself
      o[j] = g[j] * (sb != nullptr ? sv[j] : fminf(fmaxf(raw, -kClip), kClip));
    }
    *reinterpret_cast<uint4*>(dz + r * a.dz_ld + c) = pack8(o);
  }
  const long long base = (long long)e * a.cols + c;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(a.d_fast_mu + base + j, acc_mu[j]);
    if (a.d_fast_rho != nullptr) atomicAdd(a.d_fast_rho + base + j, acc_rho[j]);
    if (a.dbias != nullptr) atomicAdd(a.dbias + base + j, acc_b[j]);
  }
}

// rank-1 input side: dx = dxr ∘ r; dr = dxr ∘ x ∘ 1[unclipped]; dmu_r, drho_r column sums
__global__ void __launch_bounds__(kBlock) rank1_bwd_in_lean_kernel(BwdInArgs a) {
  const int c = (blockIdx.x * kBlock + threadIdx.x) * 8;
  if (c >= a.cols) return;
  const int chunks = (a.rows_per_member + a.lean_rows - 1) / a.lean_rows;
  const int e = blockIdx.y / chunks;
  const long long row0 = (long long)e * a.rows_per_member + (long long)(blockIdx.y - e * chunks) * a.lean_rows;
  const long long row1 = min((long long)(e + 1) * a.rows_per_member, row0 + a.lean_rows);
  const uint64_t key = lean_key(a.eps.seed, a.eps.step);
  LeanCols k;
  load_cols(a.fw_mu, a.fw_rho, e, c, a.cols, k);
  const __nv_bfloat16* dxf = reinterpret_cast<const __nv_bfloat16*>(a.dxf);
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(a.x);
  __nv_bfloat16* dx = reinterpret_cast<__nv_bfloat16*>(a.dx);
  float acc_mu[8] = {}, acc_rho[8] = {};
  for (long long r = row0; r < row1; ++r) {
    float d[8], xv[8], en[8], o[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dxf + r * a.dxf_ld + c)), d);
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + r * a.x_ld + c)), xv);
    normal8_fast(key, a.eps.stream, noise_row(a.eps, r) * a.cols + c, en);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float raw = fmaf(k.sg[j], en[j], k.m[j]);
      const bool clipped = raw < -kClip || raw > kClip;
      const float df = clipped ? 0.f : d[j] * xv[j];
      acc_mu[j] += df;
      acc_rho[j] += df * en[j] * k.sgm[j];
      o[j] = d[j] * fminf(fmaxf(raw, -kClip), kClip);
    }
    if (dx != nullptr) *reinterpret_cast<uint4*>(dx + r * a.dx_ld + c) = pack8(o);
  }
  const long long base = (long long)e * a.cols + c;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(a.d_fast_mu + base + j, acc_mu[j]);
    if (a.d_fast_rho != nullptr) atomicAdd(a.d_fast_rho + base + j, acc_rho[j]);
  }
}

__host__ inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Column-reduction tiles: a block owns 128 columns x up to kRowsPerBlock rows that all belong to one member.
constexpr int kRowsPerBlock = 64;

__device__ __forceinline__ void reduce_cols_and_add(float (&a0)[4], float (&a1)[4], float (&a2)[4], bool use0, bool use1,
                                                    bool use2, float* o0, float* o1, float* o2, long long base, int c,
                                                    int cols, const float* a3 = nullptr, float* o3 = nullptr) {
  __shared__ float red[4][kBlock / 32][128 + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[0][warp][lane * 4 + j] = a0[j];
    red[1][warp][lane * 4 + j] = a1[j];
    red[2][warp][lane * 4 + j] = a2[j];
    red[3][warp][lane * 4 + j] = a3 != nullptr ? a3[j] : 0.f;
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int cc = c - lane * 4 + threadIdx.x;  // block's first column + threadIdx.x
    if (cc < cols) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int w = 0; w < kBlock / 32; ++w) {
        s0 += red[0][w][threadIdx.x];
        s1 += red[1][w][threadIdx.x];
        s2 += red[2][w][threadIdx.x];
        s3 += red[3][w][threadIdx.x];
      }
      if (use0) atomicAdd(o0 + base + cc, s0);
      if (use1) atomicAdd(o1 + base + cc, s1);
      if (use2) atomicAdd(o2 + base + cc, s2);
      if (o3 != nullptr) atomicAdd(o3 + base + cc, s3);
    }
  }
}

template <int kFast>
__global__ void __launch_bounds__(kBlock) bwd_out_kernel(BwdOutArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks_per_member = (a.rows_per_member + kRowsPerBlock - 1) / kRowsPerBlock;
  const int e = blockIdx.y / chunks_per_member;
  const int chunk = blockIdx.y - e * chunks_per_member;
  const long long row0 = (long long)e * a.rows_per_member + (long long)chunk * kRowsPerBlock;
  const long long row_end = min((long long)(e + 1) * a.rows_per_member, row0 + kRowsPerBlock);
  const int c = blockIdx.x * 128 + lane * 4;
  const int valid = max(0, min(4, a.cols - c));
  float acc_mu[4] = {0, 0, 0, 0}, acc_rho[4] = {0, 0, 0, 0}, acc_b[4] = {0, 0, 0, 0}, acc_brho[4] = {0, 0, 0, 0};
  float gam[4] = {1, 1, 1, 1}, mus[4], rhos[4], sig_s[4], sgm_s[4], sgm_b[4] = {0, 0, 0, 0};
  const bool sampled_bias = kFast == 2 && a.rho_b != nullptr && a.d_bias_rho != nullptr;
  if (valid > 0 && sampled_bias) {
    float rb[4];
    ld4(a.rho_b, kF32, (long long)e * a.cols + c, valid, rb);
#pragma unroll
    for (int j = 0; j < 4; ++j) sgm_b[j] = sigmoid_f(rb[j]);
  }
  if (valid > 0) {
    if (kFast == 1) ld4(a.gamma, kF32, (long long)e * a.cols + c, valid, gam);
    if (kFast == 2) {
      ld4(a.mu_s, kF32, (long long)e * a.cols + c, valid, mus);
      if (a.rho_s != nullptr) {
        ld4(a.rho_s, kF32, (long long)e * a.cols + c, valid, rhos);
#pragma unroll
        for (int j = 0; j < 4; ++j) { sig_s[j] = softplus_f(rhos[j]) + kSoftplusEps; sgm_s[j] = sigmoid_f(rhos[j]); }
      }
    }
  }
  if (kFast == 0 && valid > 0) {
    // plain layer: 4 rows per iteration so that 8 independent 8/16-byte loads are in flight per thread
    constexpr int kStep = kBlock / 32;
    for (long long r = row0 + warp; r < row_end; r += 4 * kStep) {
      float g[4][4], y[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long rr = r + u * kStep;
        if (rr < row_end) {
          ld4(a.gy, a.dt, rr * a.gy_ld + c, valid, g[u]);
          if (a.act == kActRelu) ld4(a.y, a.dt, rr * a.y_ld + c, valid, y[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long rr = r + u * kStep;
        if (rr < row_end) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (a.act == kActRelu) g[u][j] = y[u][j] > 0.f ? g[u][j] : 0.f;
            acc_b[j] += g[u][j];
          }
          st4(a.dz, a.dt, rr * a.dz_ld + c, valid, g[u]);
        }
      }
    }
  } else if (valid > 0) {
    for (long long r = row0 + warp; r < row_end; r += kBlock / 32) {
      float g[4], y[4];
      ld4(a.gy, a.dt, r * a.gy_ld + c, valid, g);
      if (a.act == kActRelu) {
        ld4(a.y, a.dt, r * a.y_ld + c, valid, y);
#pragma unroll
        for (int j = 0; j < 4; ++j) g[j] = y[j] > 0.f ? g[j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) acc_b[j] += g[j];
      float dz[4];
      if (kFast == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dz[j] = g[j];
      } else if (kFast == 1) {
        float z[4];
        ld4(a.z, a.dt, r * a.z_ld + c, valid, z);
#pragma unroll
        for (int j = 0; j < 4; ++j) { dz[j] = g[j] * gam[j]; acc_mu[j] += g[j] * z[j]; }
      } else if (kFast == 2) {
        float z[4], sv[4];
        if (a.opts.additive) {  // y = z + s + b: dz = gp, ds = gp
#pragma unroll
          for (int j = 0; j < 4; ++j) { z[j] = 1.f; sv[j] = 1.f; }
        } else {
          ld4(a.z, a.dt, r * a.z_ld + c, valid, z);
          ld4(a.s_buf, a.dt, r * a.s_ld + c, valid, sv);
        }
        if (sampled_bias) {
          float eb[4];
          const long long bidx = noise_idx(a.eps_b, r, a.cols, c);
          if ((bidx & 3) == 0) {
            if (a.dt == kBF16) noise4<0, true>(a.eps_b, bidx, (long long)1 << 62, eb);
            else noise4<0>(a.eps_b, bidx, (long long)1 << 62, eb);
          } else {
            for (int j = 0; j < 4; ++j) eb[j] = j < valid ? noise1<0>(a.eps_b, bidx + j) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) acc_brho[j] += g[j] * eb[j] * sgm_b[j];
        }
        if (a.rho_s != nullptr) {
          float en[4];
          const long long idx = noise_idx(a.eps_s, r, a.cols, c);
          if ((idx & 3) == 0) {
            if (a.dt == kBF16) noise4<0, true>(a.eps_s, idx, (long long)1 << 62, en);
            else noise4<0>(a.eps_s, idx, (long long)1 << 62, en);
          }
          else
            for (int j = 0; j < 4; ++j) en[j] = j < valid ? noise1<0>(a.eps_s, idx + j) : 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float raw = fmaf(sig_s[j], en[j], mus[j]);
            const float pass = (raw < a.opts.clip_lo || raw > a.opts.clip_hi) ? 0.f : 1.f;
            const float ds = g[j] * z[j] * pass;
            acc_mu[j] += ds;
            acc_rho[j] += ds * en[j] * sgm_s[j];
            dz[j] = g[j] * sv[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) { acc_mu[j] += g[j] * z[j]; dz[j] = g[j] * sv[j]; }
        }
      } else {  // flipout: dz = gp, dz2 = gp ∘ T
        float tt[4];
        const long long idx = noise_idx(a.sign_out, r, a.cols, c);
        if ((idx & 3) == 0) noise4<1>(a.sign_out, idx, (long long)1 << 62, tt);
        else
          for (int j = 0; j < 4; ++j) tt[j] = j < valid ? noise1<1>(a.sign_out, idx + j) : 0.f;
        float d2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { dz[j] = g[j]; d2[j] = g[j] * tt[j]; }
        st4(a.dz2, a.dt, r * a.dz2_ld + c, valid, d2);
      }
      st4(a.dz, a.dt, r * a.dz_ld + c, valid, dz);
    }
  }
  reduce_cols_and_add(acc_mu, acc_rho, acc_b, a.d_fast_mu != nullptr, a.d_fast_rho != nullptr, a.dbias != nullptr,
                      a.d_fast_mu, a.d_fast_rho, a.dbias, (long long)e * a.cols, c, a.cols,
                      sampled_bias ? acc_brho : nullptr, sampled_bias ? a.d_bias_rho : nullptr);
}

__global__ void __launch_bounds__(kBlock) bwd_in_kernel(BwdInArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks_per_member = (a.rows_per_member + kRowsPerBlock - 1) / kRowsPerBlock;
  const int e = blockIdx.y / chunks_per_member;
  const int chunk = blockIdx.y - e * chunks_per_member;
  const long long row0 = (long long)e * a.rows_per_member + (long long)chunk * kRowsPerBlock;
  const long long row_end = min((long long)(e + 1) * a.rows_per_member, row0 + kRowsPerBlock);
  const int c = blockIdx.x * 128 + lane * 4;
  const int valid = max(0, min(4, a.cols - c));
  float acc_mu[4] = {0, 0, 0, 0}, acc_rho[4] = {0, 0, 0, 0}, unused[4] = {0, 0, 0, 0};
  float sgm[4] = {0, 0, 0, 0};
  if (valid > 0 && a.fast == 2 && a.fw_rho != nullptr) {
    float rr[4];
    ld4(a.fw_rho, kF32, (long long)e * a.cols + c, valid, rr);
#pragma unroll
    for (int j = 0; j < 4; ++j) sgm[j] = sigmoid_f(rr[j]);
  }
  if (valid > 0) {
    for (long long r = row0 + warp; r < row_end; r += kBlock / 32) {
      float d[4], xv[4], f[4], raw[4], en[4];
      ld4(a.dxf, a.dt, r * a.dxf_ld + c, valid, d);
      ld4(a.x, a.dt, r * a.x_ld + c, valid, xv);
      if (a.fast == 1) fast_factor4<0>(a.fw_mu, nullptr, a.eps, r, e, c, a.cols, valid, f, raw, en);
      else fast_factor4<1>(a.fw_mu, a.fw_rho, a.eps, r, e, c, a.cols, valid, f, raw, en, a.dt == kBF16,
                           a.opts.clip_lo, a.opts.clip_hi);
      const bool add = a.fast == 2 && a.opts.additive;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pass = (a.fast == 2 && a.fw_rho != nullptr && (raw[j] < a.opts.clip_lo || raw[j] > a.opts.clip_hi)) ? 0.f : 1.f;
        const float df = (add ? d[j] : d[j] * xv[j]) * pass;
        acc_mu[j] += df;
        acc_rho[j] += df * en[j] * sgm[j];
        if (!add) d[j] *= f[j];
      }
      if (a.dx != nullptr) st4(a.dx, a.dt, r * a.dx_ld + c, valid, d);
    }
  }
  reduce_cols_and_add(acc_mu, acc_rho, unused, a.d_fast_mu != nullptr, a.d_fast_rho != nullptr, false, a.d_fast_mu,
                      a.d_fast_rho, nullptr, (long long)e * a.cols, c, a.cols);
}

template <int kKind>
__global__ void fill_noise_kernel(Noise nz, void* out, int dt, long long ld, long long rows, long long cols) {
  const long long gpr = (cols + 3) >> 2;
  const long long total = rows * gpr;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long r = g / gpr, c = (g - r * gpr) << 2;
    const int valid = static_cast<int>(min(4LL, cols - c));
    const long long idx = noise_idx(nz, r, cols, c);
    float e[4];
    if ((idx & 3) == 0) noise4<kKind>(nz, idx, nz.injected != nullptr ? rows * cols : (long long)1 << 62, e);
    else
      for (int j = 0; j < 4; ++j) e[j] = j < valid ? noise1<kKind>(nz, idx + j) : 0.f;
    st4(out, dt, r * ld + c, valid, e);
  }
}

__global__ void scale_inplace_kernel(float* x, long long n, float a) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] *= a;
}


// Mean softmax cross-entropy, forward and gradient in one pass (one warp per row): the synthetic loss of the
// benchmark harness.  loss += sum_rows (logsumexp - z[label]) * loss_scale;
// dlogits = (softmax - onehot) * grad_scale.
__global__ void __launch_bounds__(kBlock) softmax_xent_kernel(const void* logits, int dt, long long ld,
                                                              const long long* labels, long long rows, int cols,
                                                              float loss_scale, float grad_scale, double* loss,
                                                              void* dlogits, int ddt, long long dld, XentExtras ex) {
  const long long warp = (blockIdx.x * (long long)kBlock + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * kBlock) >> 5;
  float local = 0.f;
  for (long long r = warp; r < rows; r += nwarps) {
    float mx = -INFINITY;
    for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, ld1(logits, dt, r * ld + c));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int c = lane; c < cols; c += 32) se += expf(ld1(logits, dt, r * ld + c) - mx);
    se = warp_sum(se);
    const int lab = static_cast<int>(labels[r]);
    const float lse = logf(se) + mx;
    if (lane == 0) local += lse - ld1(logits, dt, r * ld + lab);
    if (dlogits != nullptr) {
      const float inv = 1.f / se;
      for (int c = lane; c < cols; c += 32) {
        float p = expf(ld1(logits, dt, r * ld + c) - mx) * inv;
        if (c == lab) p -= 1.f;
        st1(dlogits, ddt, r * dld + c, p * grad_scale);
      }
    }
  }
  block_atomic_add(loss, local, loss_scale);
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // regulariser terms (e.g. the layers' KL scalars) join the same scalar
    double extra = 0.0;
    for (int i = 0; i < ex.n; ++i) extra += *ex.terms[i];
    if (ex.n > 0) atomicAdd(loss, extra);
  }
}

}  // namespace

// =============================================================================================== launchers
int k_philox_fill(int kind, float* out, long long n, uint64_t seed, uint64_t stream, cudaStream_t s) {
  if (n <= 0) return ED2_OK;
  const int g = grid_for((n + 3) / 4);
  if (kind == 0) philox_fill_kernel<0><<<g, kBlock, 0, s>>>(out, n, seed, stream);
  else if (kind == 1) philox_fill_kernel<1><<<g, kBlock, 0, s>>>(out, n, seed, stream);
  else if (kind == 3) philox_fill_kernel<3><<<g, kBlock, 0, s>>>(out, n, seed, stream);
  else philox_fill_kernel<2><<<g, kBlock, 0, s>>>(out, n, seed, stream);
  count_launch();
  return check_launch("philox_fill");
}

int k_philox_raw(uint32_t* out, long long nblk, uint64_t seed, uint64_t stream, cudaStream_t s) {
  if (nblk <= 0) return ED2_OK;
  philox_raw_kernel<<<grid_for(nblk), kBlock, 0, s>>>(out, nblk, seed, stream);
  count_launch();
  return check_launch("philox_raw");
}

int k_cast2d(const void* src, int sdt, long long sld, void* dst, int ddt, long long dld, long long rows,
             long long cols, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return ED2_OK;
  cast2d_kernel<<<grid_for(rows * ((cols + 3) / 4)), kBlock, 0, s>>>(src, sdt, sld, dst, ddt, dld, rows, cols);
  count_launch();
  return check_launch("cast2d");
}

int k_split_bf16(const void* src, int sdt, long long sld, void* dst, long long dld, long long rows, int cols, int is_b,
                 int nsplit, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return ED2_OK;
  split_bf16_kernel<<<grid_for(rows * cols), kBlock, 0, s>>>(src, sdt, sld, reinterpret_cast<__nv_bfloat16*>(dst), dld, rows,
                                                             cols, is_b, nsplit);
  count_launch();
  return check_launch("split_bf16");
}

// Kernels that are forked beside the persistent GEMM ask for the GEMM's shared-memory carveout: an SM cannot change its
// L1 / shared split while blocks of the other configuration are resident, so a mismatch serialises the two kernels.
template <typename K>
void match_gemm_carveout(K kernel) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}
void set_side_carveouts() {
  static const bool once = [] {
    match_gemm_carveout(sample_bf16_lean_kernel<0, false, false>);
    match_gemm_carveout(sample_bf16_lean_kernel<0, false, true>);
    match_gemm_carveout(sample_bf16_lean_kernel<0, true, false>);
    match_gemm_carveout(sample_bf16_lean_kernel<0, true, true>);
    match_gemm_carveout(sample_bf16_lean_kernel<1, false, false>);
    match_gemm_carveout(sample_bf16_lean_kernel<1, false, true>);
    match_gemm_carveout(sample_bf16_lean_kernel<1, true, false>);
    match_gemm_carveout(sample_bf16_lean_kernel<1, true, true>);
    match_gemm_carveout(grad_finish_lean_kernel);
    match_gemm_carveout(normal_grad_finish_kernel<true>);
    match_gemm_carveout(normal_grad_finish_kernel<false>);
    match_gemm_carveout(normal_kl_fwd_kernel);
    return true;
  }();
  (void)once;
}

int k_sample_weights(const float* mu, const float* rho, Noise eps, long long rows, long long cols, int mode, void* w,
                     int wdt, long long wld, void* mean_out, long long mean_ld, double* kl_out, float kl_scale,
                     Prior p, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return ED2_OK;
  set_side_carveouts();
  const int gx = static_cast<int>((cols + 4 * kBlock - 1) / (4 * kBlock));
  const int gy = static_cast<int>(std::min<long long>(rows, std::max(1, side_grid_cap() / gx)));
  dim3 grid(gx, gy);
  // bf16 output: hardware-approximate transcendentals (the result is rounded to 8 mantissa bits).  The KL is either
  // left to a separate accurate pass (kl_out == nullptr here) or, by default, reduced in the same lean kernel from the
  // mu / sigma it already holds (fused_kl_enabled(): lg2-based, ~1e-6 relative on the sum).
  const bool fast = wdt == kBF16 && (kl_out == nullptr || fused_kl_enabled());
  const bool lean = fast && eps.injected == nullptr && cols % 8 == 0 && wld % 8 == 0 && (mean_out == nullptr || mean_ld % 8 == 0) &&
                    (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(mu) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(rho) & 15) == 0 && (reinterpret_cast<uintptr_t>(mean_out) & 15) == 0;
  if (lean) {
    const int lgx = static_cast<int>((cols + 8 * kBlock - 1) / (8 * kBlock));
    const int lgy = static_cast<int>(std::min<long long>(rows, std::max(1, side_grid_cap() / lgx)));
    dim3 lgrid(lgx, lgy);
    auto* wp = reinterpret_cast<__nv_bfloat16*>(w);
    auto* mp = reinterpret_cast<__nv_bfloat16*>(mean_out);
#define ED2_LEAN(MODE, MEAN, KL)                                                                                      \
  sample_bf16_lean_kernel<MODE, MEAN, KL><<<lgrid, kBlock, 0, s>>>(mu, rho, eps.seed, eps.stream, eps.step, (int)rows, \
                                                                   (int)cols, wp, (int)wld, mp, (int)mean_ld, kl_out,  \
                                                                   kl_scale, p)
    const bool kl = kl_out != nullptr;
    if (mode == 0 && mp == nullptr) { if (kl) ED2_LEAN(0, false, true); else ED2_LEAN(0, false, false); }
    else if (mode == 0) { if (kl) ED2_LEAN(0, true, true); else ED2_LEAN(0, true, false); }
    else if (mp == nullptr) { if (kl) ED2_LEAN(1, false, true); else ED2_LEAN(1, false, false); }
    else { if (kl) ED2_LEAN(1, true, true); else ED2_LEAN(1, true, false); }
#undef ED2_LEAN
    count_launch();
    return check_launch("sample_weights(lean)");
  }
  if (kl_out == nullptr && fast) {
    sample_weights_kernel<false, true><<<grid, kBlock, 0, s>>>(mu, rho, eps, (int)rows, (int)cols, mode, w, wdt, wld,
                                                               mean_out, mean_ld, kl_out, kl_scale, p);
    count_launch();
    return check_launch("sample_weights");
  }
  if (kl_out != nullptr)
    sample_weights_kernel<true, false><<<grid, kBlock, 0, s>>>(mu, rho, eps, (int)rows, (int)cols, mode, w, wdt, wld,
                                                               mean_out, mean_ld, kl_out, kl_scale, p);
  else
    sample_weights_kernel<false, false><<<grid, kBlock, 0, s>>>(mu, rho, eps, (int)rows, (int)cols, mode, w, wdt, wld,
                                                                mean_out, mean_ld, kl_out, kl_scale, p);
  count_launch();
  return check_launch("sample_weights");
}

int k_normal_kl_fwd(const float* mu, const float* rho, long long n, Prior p, float scale, double* out, cudaStream_t s) {
  if (n <= 0) return ED2_OK;
  normal_kl_fwd_kernel<<<grid_for((n + 3) / 4), kBlock, 0, s>>>(mu, rho, n, p, scale, out);
  count_launch();
  return check_launch("normal_kl_fwd");
}

int k_normal_kl_bwd(const float* mu, const float* rho, long long n, Prior p, float scale, const float* upstream,
                    float* dmu, float* drho, cudaStream_t s) {
  if (n <= 0) return ED2_OK;
  normal_kl_bwd_kernel<<<grid_for((n + 3) / 4), kBlock, 0, s>>>(mu, rho, n, p, scale, upstream, dmu, drho);
  count_launch();
  return check_launch("normal_kl_bwd");
}

int k_normal_grad_finish(const float* d_mean, const float* d_pert, const float* mu, const float* rho, Noise eps,
                         long long n, float klg, const double* kl_upstream, Prior p, float* dmu, float* drho, bool fast,
                         cudaStream_t s, void* dmu_lp, void* drho_lp) {
  if (n <= 0) return ED2_OK;
  set_side_carveouts();
  const int g = std::min(grid_for((n + 3) / 4), side_grid_cap());
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (fast && eps.injected == nullptr && n % 4 == 0 && al16(d_mean) && al16(d_pert) && al16(mu) && al16(rho) && al16(dmu) && al16(drho)) {
    grad_finish_lean_kernel<<<g, kBlock, 0, s>>>(d_mean, d_pert, mu, rho, eps.seed, eps.stream, eps.step, n, klg, kl_upstream, p, dmu, drho,
                                                 reinterpret_cast<__nv_bfloat16*>(dmu_lp), reinterpret_cast<__nv_bfloat16*>(drho_lp));
    count_launch();
    return check_launch("normal_grad_finish(lean)");
  }
  if (dmu_lp != nullptr) {  // general path: the bf16 copies are made by the cast kernel afterwards
    int st = k_normal_grad_finish(d_mean, d_pert, mu, rho, eps, n, klg, kl_upstream, p, dmu, drho, fast, s, nullptr, nullptr);
    if (st != ED2_OK) return st;
    st = k_cast2d(dmu, kF32, n, dmu_lp, kBF16, n, 1, n, s);
    if (st != ED2_OK) return st;
    return k_cast2d(drho, kF32, n, drho_lp, kBF16, n, 1, n, s);
  }
  if (fast) normal_grad_finish_kernel<true><<<g, kBlock, 0, s>>>(d_mean, d_pert, mu, rho, eps, n, klg, kl_upstream, p, dmu, drho);
  else normal_grad_finish_kernel<false><<<g, kBlock, 0, s>>>(d_mean, d_pert, mu, rho, eps, n, klg, kl_upstream, p, dmu, drho);
  count_launch();
  return check_launch("normal_grad_finish");
}

int k_normal_grad_finish_any(const void* dw, int dw_dt, const float* mu, const float* rho, Noise eps, long long n,
                             float klg, const double* kl_upstream, Prior p, float* dmu, float* drho, bool fast,
                             cudaStream_t s) {
  if (dw_dt == kF32) {
    const float* d = reinterpret_cast<const float*>(dw);
    return k_normal_grad_finish(d, d, mu, rho, eps, n, klg, kl_upstream, p, dmu, drho, fast, s, nullptr, nullptr);
  }
  if (n <= 0) return ED2_OK;
  const int g = std::min(grid_for((n + 3) / 4), side_grid_cap());
  const __nv_bfloat16* d = reinterpret_cast<const __nv_bfloat16*>(dw);
  if (fast) grad_finish_bf16_kernel<true><<<g, kBlock, 0, s>>>(d, mu, rho, eps, n, klg, kl_upstream, p, dmu, drho);
  else grad_finish_bf16_kernel<false><<<g, kBlock, 0, s>>>(d, mu, rho, eps, n, klg, kl_upstream, p, dmu, drho);
  count_launch();
  return check_launch("normal_grad_finish(bf16 dW)");
}

int k_peer_grad_finish(const PeerExchange& x, const float* mu, const float* rho, Noise eps, float klg,
                       const double* kl_upstream, Prior p, float* dmu, float* drho, float* small_out, bool fast,
                       int blocks_per_sm, cudaStream_t s) {
  if (x.world < 2 || x.world > kMaxPeers || x.n <= 0 || x.n % 8 != 0 || x.chunk % 8 != 0 || x.n >= (1LL << 33))
    return set_error(ED2_ERR_BAD_ARG, "peer grad finish: bad exchange (world %d, n %lld, chunk %lld)", x.world, x.n, x.chunk);
  if (eps.injected != nullptr) return set_error(ED2_ERR_BAD_ARG, "peer grad finish: needs in-kernel (Philox) noise");
  // Persistent grid, 40 registers per thread (6 blocks per SM would fit an idle SM).  The caller says how many blocks
  // per SM to launch: 3 when the pass runs on a side stream beside the persistent GEMM (29 K of the 64 K registers are
  // taken), 5 when it runs alone (room for a concurrent owner pass of another exchange).
  if (blocks_per_sm < 1 || blocks_per_sm > 6) blocks_per_sm = 5;
  const int g = std::min(grid_for(x.n / 4), 148 * blocks_per_sm);
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al16(mu) || !al16(rho) || !al16(dmu) || !al16(drho))
    return set_error(ED2_ERR_BAD_ARG, "peer grad finish: mu / rho / dmu / drho must be 16-byte aligned");
  static const bool carve = [] {  // same shared-memory carveout as the GEMM: no SM reconfiguration between them
    cudaFuncSetAttribute(grad_finish_peer_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(grad_finish_peer_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    return true;
  }();
  (void)carve;
  if (fast) grad_finish_peer_kernel<true><<<g, kBlock, 0, s>>>(x, mu, rho, eps, klg, kl_upstream, p, dmu, drho, small_out);
  else grad_finish_peer_kernel<false><<<g, kBlock, 0, s>>>(x, mu, rho, eps, klg, kl_upstream, p, dmu, drho, small_out);
  count_launch();
  return check_launch("peer_grad_finish");
}

__global__ void fast_prep_kernel(const float* __restrict__ rho, long long n, float* __restrict__ sg, float* __restrict__ sgm,
                                 float* __restrict__ z0, float* __restrict__ z1, float* __restrict__ z2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (rho != nullptr) {
      const float r = __ldg(rho + i);
      sg[i] = softplus_f(r) + kSoftplusEps;
      sgm[i] = sigmoid_f(r);
    }
    if (z0 != nullptr) z0[i] = 0.f;
    if (z1 != nullptr) z1[i] = 0.f;
    if (z2 != nullptr) z2[i] = 0.f;
  }
}

int k_fast_prep(const float* rho, long long n, float* sg, float* sgm, cudaStream_t s, float* z0, float* z1, float* z2) {
  if (n <= 0) return ED2_OK;
  if (rho == nullptr && z0 == nullptr && z1 == nullptr && z2 == nullptr) return ED2_OK;
  fast_prep_kernel<<<static_cast<int>(std::min<long long>((n + kBlock - 1) / kBlock, 1024)), kBlock, 0, s>>>(rho, n, sg, sgm, z0, z1, z2);
  count_launch();
  return check_launch("fast_prep");
}

int k_scale_rows(const void* x, int dt, long long x_ld, long long rows, int cols, int rpm, int mode,
                 const float* fw_mu, const float* fw_rho, Noise nz, void* xo, long long xo_ld, void* f_out,
                 long long f_ld, cudaStream_t s, FastOpts fo) {
  if (rows <= 0 || cols <= 0) return ED2_OK;
  if (rpm <= 0) rpm = static_cast<int>(rows);
  const bool default_opts = fo.clip_lo == -kClip && fo.clip_hi == kClip && !fo.additive;
  if (default_opts && mode == 1 && dt == kBF16 && fw_rho != nullptr && nz.injected == nullptr && cols % 8 == 0 && rows % rpm == 0 &&
      (x == nullptr || (x_ld % 8 == 0 && al16(x))) && al16(fw_mu) && al16(fw_rho) && (cols % 4 == 0) &&
      ((x != nullptr && f_out == nullptr && xo_ld % 8 == 0 && al16(xo)) ||
       (x == nullptr && f_out != nullptr && f_ld % 8 == 0 && al16(f_out)))) {
    const int lean_rows = lean_rows_for(rows, cols);
    const int chunks = (rpm + lean_rows - 1) / lean_rows;
    dim3 grid((cols / 8 + kBlock - 1) / kBlock, static_cast<unsigned>((rows / rpm) * chunks));
    if (x != nullptr)
      rank1_scale_lean_kernel<true><<<grid, kBlock, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), (int)x_ld, rpm, cols,
                                                            fw_mu, fw_rho, nz,
                                                            reinterpret_cast<__nv_bfloat16*>(xo), (int)xo_ld, lean_rows);
    else
      rank1_scale_lean_kernel<false><<<grid, kBlock, 0, s>>>(nullptr, 0, rpm, cols, fw_mu, fw_rho, nz,
                                                             reinterpret_cast<__nv_bfloat16*>(f_out), (int)f_ld, lean_rows);
    count_launch();
    return check_launch("scale_rows(lean rank-1)");
  }
  const int g = grid_for(rows * ((cols + 3) / 4));
  if (mode == 0) scale_rows_kernel<0><<<g, kBlock, 0, s>>>(x, dt, x_ld, rows, cols, rpm, fw_mu, fw_rho, nz, xo, xo_ld, f_out, f_ld, fo);
  else if (mode == 1) scale_rows_kernel<1><<<g, kBlock, 0, s>>>(x, dt, x_ld, rows, cols, rpm, fw_mu, fw_rho, nz, xo, xo_ld, f_out, f_ld, fo);
  else scale_rows_kernel<2><<<g, kBlock, 0, s>>>(x, dt, x_ld, rows, cols, rpm, fw_mu, fw_rho, nz, xo, xo_ld, f_out, f_ld, fo);
  count_launch();
  return check_launch("scale_rows");
}

int k_bwd_out(const BwdOutArgs& a_in, cudaStream_t s) {
  BwdOutArgs a = a_in;
  if (a.rows <= 0 || a.cols <= 0) return ED2_OK;
  if (a.rows_per_member <= 0) a.rows_per_member = static_cast<int>(a.rows);
  const int members = static_cast<int>(a.rows / a.rows_per_member);
  const size_t bytes = sizeof(float) * (size_t)members * a.cols;
  if (a.d_fast_mu) cudaMemsetAsync(a.d_fast_mu, 0, bytes, s);
  if (a.d_fast_rho) cudaMemsetAsync(a.d_fast_rho, 0, bytes, s);
  if (a.dbias) cudaMemsetAsync(a.dbias, 0, bytes, s);
  if (a.d_bias_rho) cudaMemsetAsync(a.d_bias_rho, 0, bytes, s);
  const bool default_opts = a.opts.clip_lo == -kClip && a.opts.clip_hi == kClip && !a.opts.additive && a.rho_b == nullptr;
  if (default_opts && a.fast == 2 && a.dt == kBF16 && a.rho_s != nullptr && a.eps_s.injected == nullptr && a.cols % 8 == 0 &&
      a.gy_ld % 8 == 0 && a.y_ld % 8 == 0 && a.z_ld % 8 == 0 && (a.s_buf == nullptr || (a.s_ld % 8 == 0 && al16(a.s_buf))) &&
      a.dz_ld % 8 == 0 && al16(a.gy) && al16(a.y) && al16(a.z) && al16(a.dz) && al16(a.mu_s) && al16(a.rho_s) && a.d_fast_mu != nullptr) {
    a.lean_rows = lean_rows_for(a.rows, a.cols);
    const int lchunks = (a.rows_per_member + a.lean_rows - 1) / a.lean_rows;
    dim3 lgrid((a.cols / 8 + kBlock - 1) / kBlock, members * lchunks);
    rank1_bwd_out_lean_kernel<<<lgrid, kBlock, 0, s>>>(a);
    count_launch();
    return check_launch("bwd_out(lean rank-1)");
  }
  if (a.fast == 2 && a.s_buf == nullptr && !a.opts.additive)
    return set_error(ED2_ERR_BAD_ARG, "rank-1 backward: s_buf is required on the general (non-lean) path");
  const int chunks = (a.rows_per_member + kRowsPerBlock - 1) / kRowsPerBlock;
  dim3 grid((a.cols + 127) / 128, members * chunks);
  switch (a.fast) {
    case 0: bwd_out_kernel<0><<<grid, kBlock, 0, s>>>(a); break;
    case 1: bwd_out_kernel<1><<<grid, kBlock, 0, s>>>(a); break;
    case 2: bwd_out_kernel<2><<<grid, kBlock, 0, s>>>(a); break;
    default: bwd_out_kernel<3><<<grid, kBlock, 0, s>>>(a); break;
  }
  count_launch();
  return check_launch("bwd_out");
}

int k_bwd_in(const BwdInArgs& a_in, cudaStream_t s) {
  BwdInArgs a = a_in;
  if (a.rows <= 0 || a.cols <= 0) return ED2_OK;
  if (a.rows_per_member <= 0) a.rows_per_member = static_cast<int>(a.rows);
  const int members = static_cast<int>(a.rows / a.rows_per_member);
  const size_t bytes = sizeof(float) * (size_t)members * a.cols;
  if (a.d_fast_mu) cudaMemsetAsync(a.d_fast_mu, 0, bytes, s);
  if (a.d_fast_rho) cudaMemsetAsync(a.d_fast_rho, 0, bytes, s);
  const bool default_opts = a.opts.clip_lo == -kClip && a.opts.clip_hi == kClip && !a.opts.additive;
  if (default_opts && a.fast == 2 && a.dt == kBF16 && a.fw_rho != nullptr && a.eps.injected == nullptr && a.cols % 8 == 0 &&
      a.dxf_ld % 8 == 0 && a.x_ld % 8 == 0 && (a.dx == nullptr || (a.dx_ld % 8 == 0 && al16(a.dx))) && al16(a.dxf) &&
      al16(a.x) && al16(a.fw_mu) && al16(a.fw_rho) && a.d_fast_mu != nullptr) {
    a.lean_rows = lean_rows_for(a.rows, a.cols);
    const int lchunks = (a.rows_per_member + a.lean_rows - 1) / a.lean_rows;
    dim3 lgrid((a.cols / 8 + kBlock - 1) / kBlock, members * lchunks);
    rank1_bwd_in_lean_kernel<<<lgrid, kBlock, 0, s>>>(a);
    count_launch();
    return check_launch("bwd_in(lean rank-1)");
  }
  const int chunks = (a.rows_per_member + kRowsPerBlock - 1) / kRowsPerBlock;
  dim3 grid((a.cols + 127) / 128, members * chunks);
  bwd_in_kernel<<<grid, kBlock, 0, s>>>(a);
  count_launch();
  return check_launch("bwd_in");
}

int k_fill_noise(int kind, Noise nz, void* out, int dt, long long ld, long long rows, long long cols, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return ED2_OK;
  const int g = grid_for(rows * ((cols + 3) / 4));
  if (kind == 0) fill_noise_kernel<0><<<g, kBlock, 0, s>>>(nz, out, dt, ld, rows, cols);
  else fill_noise_kernel<1><<<g, kBlock, 0, s>>>(nz, out, dt, ld, rows, cols);
  count_launch();
  return check_launch("fill_noise");
}

int k_softmax_xent(const void* logits, int dt, long long ld, const long long* labels, long long rows, int cols,
                   float loss_scale, float grad_scale, double* loss, void* dlogits, int ddt, long long dld,
                   const XentExtras& ex, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return ED2_OK;
  const int g = static_cast<int>(std::min<long long>(kMaxGrid, (rows + 7) / 8));
  softmax_xent_kernel<<<g, kBlock, 0, s>>>(logits, dt, ld, labels, rows, cols, loss_scale, grad_scale, loss, dlogits, ddt, dld, ex);
  count_launch();
  return check_launch("softmax_xent");
}

int k_scale_inplace(float* x, long long n, float a, cudaStream_t s) {
  if (n <= 0) return ED2_OK;
  scale_inplace_kernel<<<grid_for(n), kBlock, 0, s>>>(x, n, a);
  count_launch();
  return check_launch("scale_inplace");
}

}  // namespace ed2
