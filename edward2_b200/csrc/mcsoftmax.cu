// Monte-Carlo softmax of the heteroscedastic output layers (edward2/tensorflow/layers/heteroscedastic.py:178-250
// _compute_mc_samples / _compute_predictive_mean / _compute_predictive_variance, :749-785 factor-analysis noise;
// edward2/tensorflow/layers/hetsngp.py:143-170 variance mixing of HeteroscedasticSNGPLayer):
//
//   latent[b,s,k] = loc[b,k] + sum_f V[b,k,f] z[b,s,f] + diag[b,k] e[b,s,k],   z, e ~ N(0, 1)
//   probs_mean[b,k] = mean_s softmax_k(latent / T),   logits = log(clip(probs_mean, eps, 1))
//   diag = softplus(diag_raw) + 1e-3;  eval with an SNGP variance: diag <- sqrt(w_het diag² + w_sngp var[b])
//
// The reference materialises [S, B, F] and [S, B, K] noise tensors and an [B, S, K] latent tensor (S = 1000 samples).
// Here one WARP owns one example and one LANE one sample at a time: the lane draws its sample's F + K normals from
// Philox (sample (b, s) owns the element range [(b S + s) P, +P), P = round8(F) + round8(K): z first, then e, each
// 8-aligned so a Philox block never straddles the two; hardware-approximate Box-Muller, abs error ~1e-6 per draw, which
// averages to ~1e-7 on the class probabilities), evaluates latent -> softmax in registers and accumulates the
// class probabilities; nothing of size B x S ever exists.  The backward regenerates the same noise:
//   u = g / probs_mean / S;  dlatent_s = p_s ∘ (u - <p_s, u>) / T;  dloc = sum_s dlatent_s,
//   ddiag = sum_s dlatent_s ∘ e_s,  dV[k,f] = sum_s dlatent_s[k] z_s[f]  (per-warp shared-memory outer products: the 32
// samples of a round are staged as rows [z | dlatent]; a lane owns one class and four factors).
#include <cuda_runtime.h>

#include <algorithm>

#include "kernels.h"
#include "philox.cuh"
#include "status.h"

namespace ed2 {
namespace {

constexpr int kWarps = 2;   // examples per block
constexpr int kFMax = 16;   // num_factors <= 16
constexpr float kMinScale = 1e-3f;  // MIN_SCALE_MONTE_CARLO, heteroscedastic.py:24

__device__ __forceinline__ float softplus_d(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

struct McArgs {
  const float* loc;       // [B][K]
  const float* fac;       // [B][K*F] factor loadings (row-major [K][F] per example)
  const float* diag_raw;  // [B][K] pre-softplus
  const float* sngp_var;  // [B] or nullptr (eval: variance mixing)
  float w_het, w_sngp;
  Noise noise;            // normal stream; injected: [NB][S][P] fp32
  int B, S, K, F;
  int shared;             // flags — bit 0: share_samples_across_batch; bit 1: one diagonal draw per sample for all
                          // classes; bit 2: sigmoid link instead of softmax
  float inv_t, eps;
  float* probs_mean;      // [B][K]
  float* logits;          // [B][K]
  float* variance;        // [B][K] or nullptr
  // backward
  const float* g;         // [B][K] dL/dlogits
  float* dloc; float* dfac; float* ddiag_raw;
};

// the P = FP + KP normals of sample (nb, s): z[f] and e[k] (8-aligned halves)
// kDiagOne (flag bit 1, compile time): ONE diagonal draw per (example, sample), shared by the classes
template <int KMAX, bool kDiagOne>
__device__ __forceinline__ void draw_sample(const McArgs& a, uint64_t key, long long sample, int FP, int KP, float (&z)[kFMax],
                                            float (&e)[KMAX]) {
  const long long base = sample * (FP + KP);
  constexpr bool diag_one = kDiagOne;
  if (a.noise.injected != nullptr) {
#pragma unroll
    for (int f = 0; f < kFMax; ++f) z[f] = f < a.F ? __ldg(a.noise.injected + base + f) : 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) e[k] = k < a.K ? __ldg(a.noise.injected + base + FP + (diag_one ? 0 : k)) : 0.f;
    return;
  }
#pragma unroll
  for (int fb = 0; fb < kFMax / 8; ++fb) {
    float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (fb * 8 < a.F) philox_normal8_n<true>(key, a.noise.stream, static_cast<uint64_t>((base >> 3) + fb), a.F - fb * 8, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) z[fb * 8 + j] = t[j];
  }
#pragma unroll
  for (int kb = 0; kb < KMAX / 8; ++kb) {
    float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (kb * 8 < a.K && !(diag_one && kb > 0))
      philox_normal8_n<true>(key, a.noise.stream, static_cast<uint64_t>(((base + FP) >> 3) + kb),
                             diag_one ? 1 : a.K - kb * 8, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) e[kb * 8 + j] = t[j];
  }
  if (diag_one) {
#pragma unroll
    for (int k = 1; k < KMAX; ++k) e[k] = e[0];
  }
}

// softmax probabilities of one sample (registers): p[k], k < K
template <int KMAX, bool kSigmoid>
__device__ __forceinline__ void sample_probs(const McArgs& a, const float* sloc, const float* sfac, const float* sdiag,
                                             const float (&z)[kFMax], const float (&e)[KMAX], float (&p)[KMAX]) {
  float m = -3.0e38f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    float lat = -3.0e38f;
    if (k < a.K) {
      lat = fmaf(sdiag[k], e[k], sloc[k]);
#pragma unroll
      for (int f = 0; f < kFMax; ++f)
        if (f < a.F) lat = fmaf(sfac[k * a.F + f], z[f], lat);
      lat *= a.inv_t;
      m = fmaxf(m, lat);
    }
    p[k] = lat;
  }
  if constexpr (kSigmoid) {  // sigmoid link (MCSigmoidDenseFA): independent outputs
#pragma unroll
    for (int k = 0; k < KMAX; ++k) p[k] = k < a.K ? 1.f / (1.f + expf(-p[k])) : 0.f;
    return;
  }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    p[k] = k < a.K ? expf(p[k] - m) : 0.f;
    sum += p[k];
  }
  const float inv = 1.f / sum;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) p[k] *= inv;
}

// per-warp staging of the example's parameters: sloc[K], sdiag[K] (effective scale), sfac[K*F]
__device__ __forceinline__ void stage_params(const McArgs& a, int b, int lane, float* sloc, float* sdiag, float* sfac) {
  for (int k = lane; k < a.K; k += 32) {
    sloc[k] = __ldg(a.loc + static_cast<long long>(b) * a.K + k);
    float d = softplus_d(__ldg(a.diag_raw + static_cast<long long>(b) * a.K + k)) + kMinScale;
    if (a.sngp_var != nullptr) d = sqrtf(a.w_het * d * d + a.w_sngp * __ldg(a.sngp_var + b));
    sdiag[k] = d;
  }
  for (int i = lane; i < a.K * a.F; i += 32) sfac[i] = __ldg(a.fac + static_cast<long long>(b) * a.K * a.F + i);
  __syncwarp();
}

template <int KMAX, bool kSigmoid, bool kDiagOne>
__global__ void __launch_bounds__(kWarps * 32) mc_softmax_fwd_kernel(McArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kWarps + warp;
  if (b >= a.B) return;
  const int per_warp = a.K * a.F + 2 * a.K;
  float* sfac = smem + warp * per_warp;
  float* sloc = sfac + a.K * a.F;
  float* sdiag = sloc + a.K;
  stage_params(a, b, lane, sloc, sdiag, sfac);
  const int FP = (a.F + 7) & ~7, KP = (a.K + 7) & ~7;
  const uint64_t key = noise_key(a.noise);
  const long long nb = (a.shared & 1) ? 0 : (a.noise.injected != nullptr ? b : noise_row(a.noise, b));
  float acc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) acc[k] = 0.f;
  for (int s = lane; s < a.S; s += 32) {
    float z[kFMax], e[KMAX], p[KMAX];
    draw_sample<KMAX, kDiagOne>(a, key, nb * a.S + s, FP, KP, z, e);
    sample_probs<KMAX, kSigmoid>(a, sloc, sfac, sdiag, z, e, p);
#pragma unroll
    for (int k = 0; k < KMAX; ++k) acc[k] += p[k];
  }
  const float inv_s = 1.f / static_cast<float>(a.S);
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < a.K) {
      acc[k] = warp_sum(acc[k]) * inv_s;
      if (lane == 0) {
        a.probs_mean[static_cast<long long>(b) * a.K + k] = acc[k];
        // softmax: log(clip(pm, eps, 1));  sigmoid: inverse sigmoid log(clip(pm, eps)) - log(1 - clip(pm, eps, 1 - eps))
        a.logits[static_cast<long long>(b) * a.K + k] =
            kSigmoid ? logf(fmaxf(acc[k], a.eps)) - logf(1.f - fminf(fmaxf(acc[k], a.eps), 1.f - a.eps))
                     : logf(fminf(fmaxf(acc[k], a.eps), 1.f));
      }
    }
  }
  if (a.variance == nullptr) return;
  // second pass over the same samples: total variance mean_s (p_s - mean)^2 (heteroscedastic.py:246-250)
  float var[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) var[k] = 0.f;
  for (int s = lane; s < a.S; s += 32) {
    float z[kFMax], e[KMAX], p[KMAX];
    draw_sample<KMAX, kDiagOne>(a, key, nb * a.S + s, FP, KP, z, e);
    sample_probs<KMAX, kSigmoid>(a, sloc, sfac, sdiag, z, e, p);
#pragma unroll
    for (int k = 0; k < KMAX; ++k) var[k] += (p[k] - acc[k]) * (p[k] - acc[k]);
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < a.K) {
      const float v = warp_sum(var[k]) * inv_s;
      if (lane == 0) a.variance[static_cast<long long>(b) * a.K + k] = v;
    }
  }
}

// scratch row of one lane's current sample: z[kFMax] then dlatent[KMAX]; the stride is 4 x odd floats, so the eight lanes
// of a 16-byte store phase land in eight different bank groups
template <int KMAX> constexpr int mc_row() { return kFMax + KMAX + 4; }
template <int KMAX> constexpr int mc_bwd_warp_floats(int K, int F) { return 32 * mc_row<KMAX>() + ((K * F + 3 * K + 3) & ~3); }

template <int KMAX, bool kSigmoid, bool kDiagOne>
__global__ void __launch_bounds__(kWarps * 32) mc_softmax_bwd_kernel(McArgs a) {
  extern __shared__ float4 smem4[];
  constexpr int kRow = mc_row<KMAX>();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kWarps + warp;
  if (b >= a.B) return;
  const int KF = a.K * a.F;
  float* scratch = reinterpret_cast<float*>(smem4) + warp * mc_bwd_warp_floats<KMAX>(a.K, a.F);  // [32][kRow], 16-byte aligned
  float* sfac = scratch + 32 * kRow;
  float* sloc = sfac + KF;
  float* sdiag = sloc + a.K;
  float* su = sdiag + a.K;
  stage_params(a, b, lane, sloc, sdiag, sfac);
  const float inv_s = 1.f / static_cast<float>(a.S);
  for (int k = lane; k < a.K; k += 32) {
    const float pm = __ldg(a.probs_mean + static_cast<long long>(b) * a.K + k);
    const float gk = __ldg(a.g + static_cast<long long>(b) * a.K + k);
    if constexpr (kSigmoid) {  // d logit / d pm of the inverse sigmoid; the clips pass no gradient outside their ranges
      const float d = (pm > a.eps ? 1.f / pm : 0.f) + ((pm > a.eps && pm < 1.f - a.eps) ? 1.f / (1.f - pm) : 0.f);
      su[k] = gk * d * inv_s;
    } else {
      su[k] = (pm > a.eps && pm <= 1.f) ? gk / pm * inv_s : 0.f;  // clip: zero gradient outside
    }
  }
  __syncwarp();
  const int FP = (a.F + 7) & ~7, KP = (a.K + 7) & ~7;
  const uint64_t key = noise_key(a.noise);
  const long long nb = (a.shared & 1) ? 0 : (a.noise.injected != nullptr ? b : noise_row(a.noise, b));
  float dl[KMAX], dd[KMAX];
  // outer product dV[k][f] = sum_s dlatent_s[k] z_s[f]: a task is one class k and four consecutive factors; task
  // t = slot * 32 + lane = k * nfg + fg.  Per staged sample a task reads one scalar and one float4 and issues 4 FMAs.
  constexpr int kSlots = KMAX / 8;  // KMAX * (kFMax / 4) tasks over 32 lanes
  const int nfg = (a.F + 3) >> 2, ntasks = a.K * nfg;
  float dv[kSlots][4];
  int tk[kSlots], tf[kSlots];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) { dl[k] = 0.f; dd[k] = 0.f; }
#pragma unroll
  for (int i = 0; i < kSlots; ++i) {
    const int t = i * 32 + lane;
    tk[i] = t < ntasks ? t / nfg : -1;
    tf[i] = t < ntasks ? t - tk[i] * nfg : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) dv[i][j] = 0.f;
  }
  float4* my_row = reinterpret_cast<float4*>(scratch + lane * kRow);
  const int rounds = (a.S + 31) / 32;
  for (int r = 0; r < rounds; ++r) {
    const int s = r * 32 + lane;
    const bool live = s < a.S;
    float z[kFMax], e[KMAX], p[KMAX];
    if (live) {
      draw_sample<KMAX, kDiagOne>(a, key, nb * a.S + s, FP, KP, z, e);
      sample_probs<KMAX, kSigmoid>(a, sloc, sfac, sdiag, z, e, p);
    }
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < a.K && live) dot = fmaf(p[k], su[k], dot);
#pragma unroll
    for (int kq = 0; kq < KMAX / 4; ++kq) {
      if (kq * 4 < a.K) {
        float d4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = kq * 4 + j;
          float d = 0.f;
          if (k < a.K && live) d = kSigmoid ? a.inv_t * su[k] * p[k] * (1.f - p[k]) : a.inv_t * p[k] * (su[k] - dot);
          dl[k] += d;
          dd[k] = fmaf(d, live ? e[k] : 0.f, dd[k]);
          d4[j] = d;
        }
        my_row[kFMax / 4 + kq] = make_float4(d4[0], d4[1], d4[2], d4[3]);
      }
    }
#pragma unroll
    for (int fq = 0; fq < kFMax / 4; ++fq)
      if (fq * 4 < a.F)
        my_row[fq] = live ? make_float4(z[fq * 4], z[fq * 4 + 1], z[fq * 4 + 2], z[fq * 4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kSlots; ++i) {
      if (tk[i] >= 0) {
        const float* pd = scratch + kFMax + tk[i];
        const float4* pz = reinterpret_cast<const float4*>(scratch) + tf[i];
#pragma unroll 8
        for (int l = 0; l < 32; ++l) {
          const float d = pd[l * kRow];
          const float4 zz = pz[l * (kRow / 4)];
          dv[i][0] = fmaf(d, zz.x, dv[i][0]);
          dv[i][1] = fmaf(d, zz.y, dv[i][1]);
          dv[i][2] = fmaf(d, zz.z, dv[i][2]);
          dv[i][3] = fmaf(d, zz.w, dv[i][3]);
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < a.K) {
      const float tl = warp_sum(dl[k]), td = warp_sum(dd[k]);
      if (lane == 0) {
        a.dloc[static_cast<long long>(b) * a.K + k] = tl;
        // diag = softplus(raw) + 1e-3: chain through the softplus
        const float raw = __ldg(a.diag_raw + static_cast<long long>(b) * a.K + k);
        a.ddiag_raw[static_cast<long long>(b) * a.K + k] = td / (1.f + expf(-raw));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kSlots; ++i) {
    if (tk[i] >= 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int f = tf[i] * 4 + j;
        if (f < a.F) a.dfac[static_cast<long long>(b) * KF + tk[i] * a.F + f] = dv[i][j];
      }
    }
  }
}

int check(const McArgs& a) {
  if (a.B <= 0 || a.S <= 0 || a.K < 1 || a.F < 1) return set_error(ED2_ERR_BAD_SHAPE, "mc_softmax: B=%d S=%d K=%d F=%d", a.B, a.S, a.K, a.F);
  if (a.K > 64) return set_error(ED2_ERR_BAD_SHAPE, "mc_softmax: num_classes %d > 64 is not built", a.K);
  if (a.F > kFMax) return set_error(ED2_ERR_BAD_SHAPE, "mc_softmax: num_factors %d > %d is not built", a.F, kFMax);
  return ED2_OK;
}

// the link (bit 2 of `shared`) and the diagonal-noise form (bit 1) are compile-time variants: as run-time branches they
// cost the softmax / per-class form 39 registers in the forward kernel and ~20 % of its time
template <bool kFwd, int KMAX, bool kSigmoid, bool kDiagOne>
void launch_kmax(const McArgs& a, int grid, size_t smem, cudaStream_t s) {
  if constexpr (kFwd) mc_softmax_fwd_kernel<KMAX, kSigmoid, kDiagOne><<<grid, kWarps * 32, smem, s>>>(a);
  else mc_softmax_bwd_kernel<KMAX, kSigmoid, kDiagOne><<<grid, kWarps * 32, smem, s>>>(a);
}
template <bool kFwd, bool kSigmoid, bool kDiagOne>
void launch_link(const McArgs& a, int grid, size_t smem, cudaStream_t s) {
  if (a.K <= 16) launch_kmax<kFwd, 16, kSigmoid, kDiagOne>(a, grid, smem, s);
  else if (a.K <= 32) launch_kmax<kFwd, 32, kSigmoid, kDiagOne>(a, grid, smem, s);
  else launch_kmax<kFwd, 64, kSigmoid, kDiagOne>(a, grid, smem, s);
}
template <bool kFwd>
void launch_variant(const McArgs& a, int grid, size_t smem, cudaStream_t s) {
  const bool sig = (a.shared & 4) != 0, one = (a.shared & 2) != 0;
  if (sig && one) launch_link<kFwd, true, true>(a, grid, smem, s);
  else if (sig) launch_link<kFwd, true, false>(a, grid, smem, s);
  else if (one) launch_link<kFwd, false, true>(a, grid, smem, s);
  else launch_link<kFwd, false, false>(a, grid, smem, s);
}

}  // namespace

int k_mc_softmax_fwd(const float* loc, const float* fac, const float* diag_raw, const float* sngp_var, float w_het,
                     float w_sngp, Noise noise, int B, int S, int K, int F, int shared, float temperature, float eps,
                     float* probs_mean, float* logits, float* variance, cudaStream_t s) {
  McArgs a{};
  a.loc = loc; a.fac = fac; a.diag_raw = diag_raw; a.sngp_var = sngp_var; a.w_het = w_het; a.w_sngp = w_sngp;
  a.noise = noise; a.B = B; a.S = S; a.K = K; a.F = F; a.shared = shared; a.inv_t = 1.f / temperature; a.eps = eps;
  a.probs_mean = probs_mean; a.logits = logits; a.variance = variance;
  const int st = check(a);
  if (st != ED2_OK) return st;
  const size_t smem = sizeof(float) * kWarps * (static_cast<size_t>(K) * F + 2 * K);
  const int grid = (B + kWarps - 1) / kWarps;
  launch_variant<true>(a, grid, smem, s);
  count_launch();
  return check_launch("mc_softmax_fwd");
}

int k_mc_softmax_bwd(const float* g, const float* probs_mean, const float* loc, const float* fac, const float* diag_raw,
                     Noise noise, int B, int S, int K, int F, int shared, float temperature, float eps, float* dloc,
                     float* dfac, float* ddiag_raw, cudaStream_t s) {
  McArgs a{};
  a.loc = loc; a.fac = fac; a.diag_raw = diag_raw; a.noise = noise;
  a.B = B; a.S = S; a.K = K; a.F = F; a.shared = shared; a.inv_t = 1.f / temperature; a.eps = eps;
  a.probs_mean = const_cast<float*>(probs_mean); a.g = g; a.dloc = dloc; a.dfac = dfac; a.ddiag_raw = ddiag_raw;
  const int st = check(a);
  if (st != ED2_OK) return st;
  const size_t smem = sizeof(float) * kWarps *
                      (K <= 16 ? mc_bwd_warp_floats<16>(K, F) : K <= 32 ? mc_bwd_warp_floats<32>(K, F) : mc_bwd_warp_floats<64>(K, F));
  const int grid = (B + kWarps - 1) / kWarps;
  launch_variant<false>(a, grid, smem, s);
  count_launch();
  return check_launch("mc_softmax_bwd");
}

}  // namespace ed2
