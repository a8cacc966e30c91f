// Conv2D stochastic-weight layers (edward2/tensorflow/layers/convolutional.py:32,167,560,2056) on the dense hot path.
//
// In every variant the per-example factors are vectors over CHANNELS, broadcast over the spatial axes
// (Flipout signs [B,1,1,C] :235-249, BatchEnsemble alpha/gamma :681-699, rank-1 r/s :1976-1989).  A convolution
// restated on the patch matrix  cols[(b,ho,wo), (i,j,c)] = x[b, ho*sh+i-pt, wo*sw+j-pl, c]  is therefore a DENSE
// BatchEnsemble-style product with one "member" per example:  y = act((cols ∘ tile(f_in[b])) W ∘ f_out[b] + bias[b])
// with W = kernel reshaped [kh*kw*Cin, Cout] — exactly what ed2_be_dense_fwd/_bwd, ed2_reparam_dense_* compute on the
// tcgen05 GEMM with rows-per-member = Ho*Wo.  This file holds what the dense path lacks: the patch gather (im2col), its
// adjoint (col2im, gather form: no atomics), the expansion of per-member / per-example channel tables to the patch
// layout and its adjoint, and the rank-1 factor table with its gradient.  NHWC ("channels_last"), dilation 1.
#include <cuda_bf16.h>

#include <algorithm>

#include "gemm_epilogue.h"
#include "kernels.h"
#include "philox.cuh"
#include "status.h"

namespace ed2 {
namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ float ldv(const void* p, int dt, long long i) {
  return dt == kBF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void stv(void* p, int dt, long long i, float v) {
  if (dt == kBF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[i] = v;
}

struct ConvGeom {
  int B, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo;
};

// One thread per (patch row m, kernel tap t, channel vector): kVec channels (16 bytes) moved as one vector, zero-filled
// outside the image.  Consecutive threads walk the channels of one tap, then the taps of one patch: reads and writes
// are both contiguous runs of C elements.
// x ∘ fac[b][c..c+7] for a vector of 8 bf16 / 4 fp32 channels (fac fp32)
__device__ __forceinline__ uint4 scale_vec(uint4 v, const float* __restrict__ f) {
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 a = __bfloat1622float2(h[t]);
    h[t] = __floats2bfloat162_rn(a.x * __ldg(f + 2 * t), a.y * __ldg(f + 2 * t + 1));
  }
  return v;
}
__device__ __forceinline__ float4 scale_vec(float4 v, const float* __restrict__ f) {
  return make_float4(v.x * __ldg(f), v.y * __ldg(f + 1), v.z * __ldg(f + 2), v.w * __ldg(f + 3));
}

template <int kVec, typename V>
__global__ void __launch_bounds__(kBlock) im2col_vec_kernel(const V* __restrict__ x, V* __restrict__ cols, long long cols_ldv,
                                                           ConvGeom g, const float* __restrict__ fac = nullptr) {
  const int cv = g.C / kVec;
  const long long per_row = static_cast<long long>(g.kh) * g.kw * cv;
  const long long total = static_cast<long long>(g.B) * g.Ho * g.Wo * per_row;
  for (long long idx = blockIdx.x * (long long)kBlock + threadIdx.x; idx < total; idx += (long long)gridDim.x * kBlock) {
    const long long m = idx / per_row;
    const int rem = static_cast<int>(idx - m * per_row);
    const int t = rem / cv, c = rem - t * cv;
    const int i = t / g.kw, j = t - i * g.kw;
    const int wo = static_cast<int>(m % g.Wo);
    const long long bh = m / g.Wo;
    const int ho = static_cast<int>(bh % g.Ho), b = static_cast<int>(bh / g.Ho);
    const int h = ho * g.sh + i - g.pt, w = wo * g.sw + j - g.pl;
    V v{};
    if (h >= 0 && h < g.H && w >= 0 && w < g.W) {
      v = __ldg(x + ((static_cast<long long>(b) * g.H + h) * g.W + w) * cv + c);
      if (fac != nullptr) v = scale_vec(v, fac + static_cast<long long>(b) * g.C + c * kVec);
    }
    cols[m * cols_ldv + static_cast<long long>(t) * cv + c] = v;
  }
}

// The same copy with ONE WARP PER PATCH ROW (the kh*kw*C elements of a row are contiguous in `cols`): the row's (b, ho, wo)
// is decoded once per warp, the tap offsets come from a per-block table and a lane tracks its (tap, channel vector)
// incrementally — no division per element (the flat kernel above spends ~5 integer divisions per 16-byte store and is
// instruction-bound at 1.8 TB/s).  Needs kh*kw <= kMaxTaps and fewer than 2^31 rows.
constexpr int kMaxTaps = 64;
template <int kVec, typename V>
__global__ void __launch_bounds__(kBlock) im2col_rows_kernel(const V* __restrict__ x, V* __restrict__ cols, long long cols_ldv,
                                                            ConvGeom g, const float* __restrict__ fac, unsigned rows) {
  __shared__ int s_dh[kMaxTaps], s_dw[kMaxTaps];
  const int taps = g.kh * g.kw;
  if (threadIdx.x < taps) {
    const int i = threadIdx.x / g.kw;
    s_dh[threadIdx.x] = i - g.pt;
    s_dw[threadIdx.x] = threadIdx.x - i * g.kw - g.pl;
  }
  __syncthreads();
  const int cv = g.C / kVec, per_row = taps * cv;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t0 = lane / cv, c0 = lane - t0 * cv;    // (tap, channel vector) of element `lane` of a row
  const int dt_ = 32 / cv, dc_ = 32 - dt_ * cv;     // ... and of a step of 32 elements
  const unsigned wpb = kBlock / 32;
  for (unsigned m = blockIdx.x * wpb + warp; m < rows; m += gridDim.x * wpb) {
    const unsigned wo = m % static_cast<unsigned>(g.Wo), bh = m / static_cast<unsigned>(g.Wo);
    const unsigned ho = bh % static_cast<unsigned>(g.Ho), b = bh / static_cast<unsigned>(g.Ho);
    const int h0 = static_cast<int>(ho) * g.sh, w0 = static_cast<int>(wo) * g.sw;
    const V* __restrict__ img = x + static_cast<long long>(b) * g.H * g.W * cv;
    const float* __restrict__ fb = fac != nullptr ? fac + static_cast<long long>(b) * g.C : nullptr;
    V* __restrict__ dst = cols + static_cast<long long>(m) * cols_ldv;
    int t = t0, c = c0;
    // four elements per lane and pass: the four loads are issued before the first store (a row is only kh*kw*C/kVec
    // vectors long, so without this a warp has one 512-byte request in flight at a time)
    for (int rem = lane; rem < per_row; rem += 128) {
      V v[4];
      int cu[4];
      bool in[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = V{};
        cu[u] = c;
        in[u] = false;
        if (rem + 32 * u < per_row) {
          const int h = h0 + s_dh[t], w = w0 + s_dw[t];
          in[u] = h >= 0 && h < g.H && w >= 0 && w < g.W;
          if (in[u]) v[u] = __ldg(img + (h * g.W + w) * cv + c);
        }
        t += dt_; c += dc_;
        if (c >= cv) { c -= cv; ++t; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (rem + 32 * u < per_row) {
          if (fb != nullptr && in[u]) v[u] = scale_vec(v[u], fb + cu[u] * kVec);
          dst[rem + 32 * u] = v[u];
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kBlock) im2col_kernel(const void* __restrict__ x, int dt, void* __restrict__ cols,
                                                       long long cols_ld, ConvGeom g, const float* __restrict__ fac = nullptr) {
  const long long per_row = static_cast<long long>(g.kh) * g.kw * g.C;
  const long long total = static_cast<long long>(g.B) * g.Ho * g.Wo * per_row;
  for (long long idx = blockIdx.x * (long long)kBlock + threadIdx.x; idx < total; idx += (long long)gridDim.x * kBlock) {
    const long long m = idx / per_row;
    const int rem = static_cast<int>(idx - m * per_row);
    const int t = rem / g.C, c = rem - t * g.C;
    const int i = t / g.kw, j = t - i * g.kw;
    const int wo = static_cast<int>(m % g.Wo);
    const long long bh = m / g.Wo;
    const int ho = static_cast<int>(bh % g.Ho), b = static_cast<int>(bh / g.Ho);
    const int h = ho * g.sh + i - g.pt, w = wo * g.sw + j - g.pl;
    float v = 0.f;
    if (h >= 0 && h < g.H && w >= 0 && w < g.W) {
      v = ldv(x, dt, ((static_cast<long long>(b) * g.H + h) * g.W + w) * g.C + c);
      if (fac != nullptr) v *= __ldg(fac + static_cast<long long>(b) * g.C + c);
    }
    stv(cols, dt, m * cols_ld + rem, v);
  }
}

// Adjoint of im2col in gather form: dx[b,h,w,c] = sum over the taps (i,j) whose patch (ho,wo) covers (h,w) of
// dcols[(b,ho,wo), (i,j,c)] — fp32 accumulation, every output element written once, no atomics.
__global__ void __launch_bounds__(kBlock) col2im_kernel(const void* __restrict__ dcols, int dt, long long cols_ld,
                                                       void* __restrict__ dx, int dx_dt, ConvGeom g) {
  const long long total = static_cast<long long>(g.B) * g.H * g.W * g.C;
  for (long long idx = blockIdx.x * (long long)kBlock + threadIdx.x; idx < total; idx += (long long)gridDim.x * kBlock) {
    const int c = static_cast<int>(idx % g.C);
    long long r = idx / g.C;
    const int w = static_cast<int>(r % g.W);
    r /= g.W;
    const int h = static_cast<int>(r % g.H), b = static_cast<int>(r / g.H);
    float acc = 0.f;
    for (int i = 0; i < g.kh; ++i) {
      const int hh = h + g.pt - i;
      if (hh < 0 || hh % g.sh != 0) continue;
      const int ho = hh / g.sh;
      if (ho >= g.Ho) continue;
      for (int j = 0; j < g.kw; ++j) {
        const int ww = w + g.pl - j;
        if (ww < 0 || ww % g.sw != 0) continue;
        const int wo = ww / g.sw;
        if (wo >= g.Wo) continue;
        const long long m = (static_cast<long long>(b) * g.Ho + ho) * g.Wo + wo;
        acc += ldv(dcols, dt, m * cols_ld + (static_cast<long long>(i) * g.kw + j) * g.C + c);
      }
    }
    stv(dx, dx_dt, idx, acc);
  }
}

// bf16 patch gradients, C % 8 == 0: one thread owns 8 consecutive channels of one pixel (16-byte loads per tap)
// kUnit: strides 1 x 1 (no division per tap); the element index is decoded in 32 bits (the launcher checks the range)
template <bool kUnit>
__global__ void __launch_bounds__(kBlock) col2im_bf16x8_kernel(const uint4* __restrict__ dcols, long long cols_ldv,
                                                              void* __restrict__ dx, int dx_dt, ConvGeom g) {
  const unsigned cv = g.C / 8;
  const unsigned total = static_cast<unsigned>(g.B) * g.H * g.W * cv;
  for (unsigned idx = blockIdx.x * kBlock + threadIdx.x; idx < total; idx += gridDim.x * kBlock) {
    const int c = static_cast<int>(idx % cv);
    unsigned r = idx / cv;
    const int w = static_cast<int>(r % static_cast<unsigned>(g.W));
    r /= static_cast<unsigned>(g.W);
    const int h = static_cast<int>(r % static_cast<unsigned>(g.H)), b = static_cast<int>(r / static_cast<unsigned>(g.H));
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < g.kh; ++i) {
      const int hh = h + g.pt - i;
      if (hh < 0 || (!kUnit && hh % g.sh != 0)) continue;
      const int ho = kUnit ? hh : hh / g.sh;
      if (ho >= g.Ho) continue;
      for (int j = 0; j < g.kw; ++j) {
        const int ww = w + g.pl - j;
        if (ww < 0 || (!kUnit && ww % g.sw != 0)) continue;
        const int wo = kUnit ? ww : ww / g.sw;
        if (wo >= g.Wo) continue;
        const long long m = (static_cast<long long>(b) * g.Ho + ho) * g.Wo + wo;
        const uint4 q = __ldg(dcols + m * cols_ldv + (static_cast<long long>(i) * g.kw + j) * cv + c);
        const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __bfloat1622float2(hp[t]);
          acc[2 * t] += f.x;
          acc[2 * t + 1] += f.y;
        }
      }
    }
    if (dx_dt == kBF16) {
      uint4 o;
      __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int t = 0; t < 4; ++t) op[t] = __floats2bfloat162_rn(acc[2 * t], acc[2 * t + 1]);
      reinterpret_cast<uint4*>(dx)[idx] = o;
    } else {
      float4* o = reinterpret_cast<float4*>(dx) + idx * 2;
      o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

// Image-space tail of a factorised convolution's backward: g = gradient w.r.t. the SCALED image x∘fac (from col2im).
//   dx[b,h,w,c] = g * fac[b,c];   dfac[b,c] += sum_{h,w} g * x      (dfac zeroed by the caller)
// Block = one image b and a slab of pixels; thread = one channel (coalesced rows of C), one atomic per thread.
__global__ void __launch_bounds__(kBlock) conv_scale_reduce_kernel(const void* __restrict__ gimg, const void* __restrict__ x,
                                                                  int dt, const float* __restrict__ fac, void* __restrict__ dx,
                                                                  int dx_dt, float* __restrict__ dfac, int HW, int C,
                                                                  int pixels_per_block) {
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * pixels_per_block, p1 = min(HW, p0 + pixels_per_block);
  for (int c = threadIdx.x; c < C; c += kBlock) {
    const float f = fac != nullptr ? __ldg(fac + static_cast<long long>(b) * C + c) : 1.f;
    float acc = 0.f;
    for (int p = p0; p < p1; ++p) {
      const long long i = (static_cast<long long>(b) * HW + p) * C + c;
      const float gv = ldv(gimg, dt, i);
      acc += gv * ldv(x, dt, i);
      if (dx != nullptr) stv(dx, dx_dt, i, gv * f);
    }
    if (dfac != nullptr) atomicAdd(dfac + static_cast<long long>(b) * C + c, acc);
  }
}

// out[b][t*C + c] = src[b / rows_per_entry][c], t < reps   (per-member or per-example channel table -> patch layout)
__global__ void expand_table_kernel(const float* __restrict__ src, float* __restrict__ out, int B, int C, int rows_per_entry,
                                    int reps) {
  const long long total = static_cast<long long>(B) * reps * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    const int b = static_cast<int>(idx / (static_cast<long long>(reps) * C));
    out[idx] = __ldg(src + static_cast<long long>(b / rows_per_entry) * C + c);
  }
}
// adjoint: dsrc[e][c] = sum_{b in entry e} sum_t dout[b][t*C + c]   (one thread per (e, c); the tables are tiny)
__global__ void expand_table_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dsrc, int entries, int C,
                                        int rows_per_entry, int reps) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= entries * C) return;
  const int e = idx / C, c = idx - e * C;
  float acc = 0.f;
  for (int b = e * rows_per_entry; b < (e + 1) * rows_per_entry; ++b)
    for (int t = 0; t < reps; ++t) acc += dout[(static_cast<long long>(b) * reps + t) * C + c];
  dsrc[idx] = acc;
}

// gradient of the rank-1 factor table f[b][c] = clip(mu[e][c] + sigma[e][c] * eps[b][c], lo, hi), e = b / n:
//   dmu[e][c] = sum_b df * 1[unclipped];  drho[e][c] = sum_b df * 1[unclipped] * eps * sigmoid(rho)
__global__ void rank1_factor_bwd_kernel(const float* __restrict__ df, const float* __restrict__ mu, const float* __restrict__ rho,
                                        Noise eps, int E, int n, int C, float lo, float hi, float* __restrict__ dmu,
                                        float* __restrict__ drho) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= E * C) return;
  const int e = idx / C, c = idx - e * C;
  const float m = mu[idx];
  const float rh = rho != nullptr ? rho[idx] : 0.f;
  const float sigma = rho != nullptr ? (rh > 20.f ? rh : log1pf(expf(rh))) + 1e-7f : 0.f;
  const float sgm = 1.f / (1.f + expf(-rh));
  float am = 0.f, ar = 0.f;
  for (int i = 0; i < n; ++i) {
    const long long b = static_cast<long long>(e) * n + i;
    const float ev = rho != nullptr ? noise1<0>(eps, noise_idx(eps, b, C, c)) : 0.f;
    const float raw = fmaf(sigma, ev, m);
    const float pass = (raw < lo || raw > hi) ? 0.f : 1.f;
    const float d = df[b * C + c] * pass;
    am += d;
    ar += d * ev;
  }
  dmu[idx] = am;
  if (drho != nullptr) drho[idx] = ar * sgm;
}

int grid_for(long long work) {
  long long g = (work + kBlock - 1) / kBlock;
  return static_cast<int>(std::max<long long>(1, std::min<long long>(g, 148 * 16)));
}

int geom(const ConvArgs& a, ConvGeom* g) {
  if (a.batch <= 0 || a.height <= 0 || a.width <= 0 || a.channels <= 0 || a.kh <= 0 || a.kw <= 0 || a.sh <= 0 || a.sw <= 0)
    return set_error(ED2_ERR_BAD_SHAPE, "conv: non-positive dimension");
  auto out = [](int n, int k, int s, int same, int* o, int* p0) {
    if (!same) { *o = (n - k) / s + 1; *p0 = 0; return; }
    *o = (n + s - 1) / s;  // TF 'SAME': extra padding goes to the bottom / right
    const int total = std::max((*o - 1) * s + k - n, 0);
    *p0 = total / 2;
  };
  g->B = a.batch; g->H = a.height; g->W = a.width; g->C = a.channels; g->kh = a.kh; g->kw = a.kw; g->sh = a.sh; g->sw = a.sw;
  out(a.height, a.kh, a.sh, a.same, &g->Ho, &g->pt);
  out(a.width, a.kw, a.sw, a.same, &g->Wo, &g->pl);
  if (g->Ho <= 0 || g->Wo <= 0) return set_error(ED2_ERR_BAD_SHAPE, "conv: kernel larger than the 'valid' input");
  return ED2_OK;
}

}  // namespace

int k_conv_out_size(const ConvArgs& a, int* ho, int* wo) {
  ConvGeom g;
  const int st = geom(a, &g);
  if (st != ED2_OK) return st;
  *ho = g.Ho; *wo = g.Wo;
  return ED2_OK;
}

int k_im2col(const ConvArgs& a, const void* x, int dt, void* cols, long long cols_ld, cudaStream_t s, const float* fac) {
  ConvGeom g;
  const int st = geom(a, &g);
  if (st != ED2_OK) return st;
  const long long K = static_cast<long long>(g.kh) * g.kw * g.C;
  if (cols_ld < K) return set_error(ED2_ERR_BAD_ARG, "im2col: cols_ld %lld < kh*kw*C %lld", cols_ld, K);
  const long long rows = static_cast<long long>(g.B) * g.Ho * g.Wo;
  const int vec = dt == kBF16 ? 8 : 4;
  const bool aligned = g.C % vec == 0 && cols_ld % vec == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(cols) & 15) == 0;
  if (aligned && g.kh * g.kw <= kMaxTaps && rows < (1LL << 31) && static_cast<long long>(g.H) * g.W * (g.C / vec) < (1LL << 31)) {
    const int grid = static_cast<int>(std::min<long long>((rows + 7) / 8, 148 * 8));
    const unsigned r = static_cast<unsigned>(rows);
    if (dt == kBF16)
      im2col_rows_kernel<8, uint4><<<grid, kBlock, 0, s>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(cols), cols_ld / 8, g, fac, r);
    else
      im2col_rows_kernel<4, float4><<<grid, kBlock, 0, s>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(cols), cols_ld / 4, g, fac, r);
  } else if (aligned) {
    const long long work = rows * g.kh * g.kw * (g.C / vec);
    if (dt == kBF16)
      im2col_vec_kernel<8, uint4><<<grid_for(work), kBlock, 0, s>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(cols), cols_ld / 8, g, fac);
    else
      im2col_vec_kernel<4, float4><<<grid_for(work), kBlock, 0, s>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(cols), cols_ld / 4, g, fac);
  } else {
    im2col_kernel<<<grid_for(rows * K), kBlock, 0, s>>>(x, dt, cols, cols_ld, g, fac);
  }
  count_launch();
  return check_launch("im2col");
}

int k_col2im(const ConvArgs& a, const void* dcols, int dt, long long cols_ld, void* dx, int dx_dt, cudaStream_t s) {
  ConvGeom g;
  const int st = geom(a, &g);
  if (st != ED2_OK) return st;
  const long long total = static_cast<long long>(g.B) * g.H * g.W * g.C;
  if (dt == kBF16 && g.C % 8 == 0 && cols_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(dcols) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dx) & 15) == 0 && total / 8 < (1LL << 31)) {
    if (g.sh == 1 && g.sw == 1)
      col2im_bf16x8_kernel<true><<<grid_for(total / 8), kBlock, 0, s>>>(reinterpret_cast<const uint4*>(dcols), cols_ld / 8, dx, dx_dt, g);
    else
      col2im_bf16x8_kernel<false><<<grid_for(total / 8), kBlock, 0, s>>>(reinterpret_cast<const uint4*>(dcols), cols_ld / 8, dx, dx_dt, g);
    count_launch();
    return check_launch("col2im");
  }
  col2im_kernel<<<grid_for(total), kBlock, 0, s>>>(dcols, dt, cols_ld, dx, dx_dt, g);
  count_launch();
  return check_launch("col2im");
}

int k_conv_scale_reduce(const ConvArgs& a, const void* gimg, const void* x, int dt, const float* fac, void* dx, int dx_dt,
                        float* dfac, cudaStream_t s) {
  const int HW = a.height * a.width, C = a.channels;
  if (dfac != nullptr) cudaMemsetAsync(dfac, 0, sizeof(float) * static_cast<size_t>(a.batch) * C, s);
  int ppb = 32;
  while (ppb > 4 && static_cast<long long>(a.batch) * ((HW + ppb - 1) / ppb) < 148 * 4) ppb >>= 1;
  dim3 grid((HW + ppb - 1) / ppb, a.batch);
  conv_scale_reduce_kernel<<<grid, kBlock, 0, s>>>(gimg, x, dt, fac, dx, dx_dt, dfac, HW, C, ppb);
  count_launch();
  return check_launch("conv_scale_reduce");
}

int k_expand_table(const float* src, float* out, int B, int C, int rows_per_entry, int reps, cudaStream_t s) {
  if (B <= 0 || C <= 0 || rows_per_entry <= 0 || reps <= 0 || B % rows_per_entry != 0)
    return set_error(ED2_ERR_BAD_SHAPE, "expand_table: B=%d rows_per_entry=%d reps=%d C=%d", B, rows_per_entry, reps, C);
  expand_table_kernel<<<grid_for(static_cast<long long>(B) * reps * C), kBlock, 0, s>>>(src, out, B, C, rows_per_entry, reps);
  count_launch();
  return check_launch("expand_table");
}

int k_expand_table_bwd(const float* dout, float* dsrc, int B, int C, int rows_per_entry, int reps, cudaStream_t s) {
  if (B <= 0 || C <= 0 || rows_per_entry <= 0 || reps <= 0 || B % rows_per_entry != 0)
    return set_error(ED2_ERR_BAD_SHAPE, "expand_table_bwd: B=%d rows_per_entry=%d reps=%d C=%d", B, rows_per_entry, reps, C);
  const int entries = B / rows_per_entry;
  expand_table_bwd_kernel<<<(entries * C + kBlock - 1) / kBlock, kBlock, 0, s>>>(dout, dsrc, entries, C, rows_per_entry, reps);
  count_launch();
  return check_launch("expand_table_bwd");
}

int k_rank1_factor_bwd(const float* df, const float* mu, const float* rho, Noise eps, int E, int n, int C, float lo, float hi,
                       float* dmu, float* drho, cudaStream_t s) {
  if (E <= 0 || n <= 0 || C <= 0) return set_error(ED2_ERR_BAD_SHAPE, "rank1_factor_bwd: empty");
  rank1_factor_bwd_kernel<<<(E * C + kBlock - 1) / kBlock, kBlock, 0, s>>>(df, mu, rho, eps, E, n, C, lo, hi, dmu, drho);
  count_launch();
  return check_launch("rank1_factor_bwd");
}

}  // namespace ed2
