// Operand-layout / dtype enums and the fused-epilogue parameter block of the tcgen05 GEMM.
#pragma once

namespace ed2 {

enum : int { kMajorK = 0, kMajorMN = 1 };
enum : int { kActNone = 0, kActRelu = 1, kActCos = 2 };
enum : int { kF32 = 0, kBF16 = 1 };

// Fused epilogue, applied per accumulator element (m, n) in this order:
//   v = alpha * acc
//   raw[m,n] = v                                   (optional second output: the un-scaled product)
//   v *= elem_scale[m,n]                           (Flipout sign_output / per-example rank-1 s)
//   v *= col_scale[(m / group_rows), n]            (BatchEnsemble gamma per member; group_rows=0: one row)
//   v += beta * addend[m,n]                        (Flipout mean term, momentum blend of the precision)
//   v += col_bias[(m / group_rows), n]
//   v = act(v)                                     (none | relu | cos(v) * act_scale)
//   v = mask_src[m,n] > 0 ? v : 0                  (optional: ReLU backward of the layer that produced A's consumer input)
//   col_sum[g(m), n] += sum_m v                    (optional; fp32 atomics: bias gradient of that layer)
//   out[m,n] = v                                   (skipped when out == nullptr)
//   row_sum[m] += sum_n v                          (optional; fp32 atomicAdd, one per row per tile)
struct EpiParams {
  float alpha;
  float beta;
  float act_scale;
  int act;
  int group_rows;
  const void* elem_scale;
  int elem_scale_ld;
  int elem_scale_dt;
  const float* col_scale;
  int col_scale_ld;
  const void* addend;
  int addend_ld;
  int addend_dt;
  const float* col_bias;
  int col_bias_ld;
  void* out;
  int out_ld;
  int out_dt;
  int out_tma;  // 1: `out` goes through swizzled smem + TMA store (needs 16-byte aligned pitch)
  void* raw;
  int raw_ld;
  int raw_dt;
  int raw_neg_sin;  // with act == kActCos: raw receives -sin(v) * act_scale (the derivative factor of the random-feature
                    // map, written next to phi) instead of the pre-activation
  float* row_sum;
  const void* mask_src;
  int mask_ld;
  int mask_dt;
  float* col_sum;
  int col_sum_ld;
  // ---- work decomposition (generic epilogue only)
  int split_k;     // > 1: the K loop of every tile is cut into split_k slices, each a separate work unit whose epilogue
                   //      ADDS alpha * partial into fp32 `out` with atomics (out pre-initialised by the caller; no other
                   //      epilogue term, no TMA store).  For products with few output tiles and a long K (PhiT Phi at
                   //      D = 1024, the weight gradients of narrow layers).
  int lower_only;  // 1: square output, only tiles on or below the diagonal are computed (SYRK: PhiT Phi is symmetric;
                   //      the caller mirrors the strict lower triangle afterwards)
};

// ---- fused rank-1 epilogues (DenseRank1, edward2/tensorflow/layers/dense.py:1096-1144) -------------------------------
// One per-example fast-weight stream f[m,n] = clip(mu[e,n] + sg[e,n] * eps(m,n), lo, hi), e = m / group_rows, with eps
// regenerated from Philox by GLOBAL row (philox.cuh Noise row mapping).  sg == nullptr: deterministic f = mu[e,n].
struct FastSide {
  const float* mu;            // [E][N]
  const float* sg;            // [E][N] softplus(rho) + 1e-7, precomputed (k_fast_prep); nullptr = deterministic
  const float* sgm;           // [E][N] sigmoid(rho)            (backward only)
  unsigned long long seed, stream;
  const unsigned long long* step;
  long long row0;
  int rpg_local, rpg_global;
};

enum : int { kEpiGeneric = 0, kEpiRank1Fwd = 1, kEpiRank1Bwd = 2 };

// kEpiRank1Fwd  (GEMM = (x∘r) W, M = batch, N = out_dim):
//   z = acc -> raw;  y = act(z * s[m,n] + col_bias[e,n]) -> out;  out2 = bf16(y) * r_next[m,n]   (out2 optional: the
//   NEXT layer's x∘r, so that layer needs no scale pass)
// kEpiRank1Bwd  (GEMM = dz Wᵀ, M = batch, N = in_dim of this layer = out_dim of the producer):
//   dxr = acc;  dx = dxr * r[m,n];  dr = dxr * x * 1[r unclipped]  ->  dmu_r[e,n] += dr, drho_r[e,n] += dr*eps_r*sgm_r
//   without a producer:  dx -> out (optional)
//   with a producer (s.mu != nullptr; its y is this layer's x, its z = z_prev):
//     dx -> out2 (optional);  gp = bf16(dx) * act'(x);  dz_prev = gp * s[m,n] -> out;
//     ds = gp * z_prev * 1[s unclipped] -> dmu_s, drho_s;  dbias[e,n] += gp
struct Rank1Epi {
  FastSide s, r;
  float clip_lo, clip_hi;
  void* out2; int out2_ld;            // bf16
  const void* x; int x_ld;            // bf16 [M][N]   (bwd)
  const void* z_prev; int z_prev_ld;  // bf16 [M][N]   (bwd, producer)
  int prev_act;
  float* dmu_r; float* drho_r;        // [E][N] fp32, zero-initialised by the caller (atomics)
  float* dmu_s; float* drho_s; float* dbias;
};

// ---- fused Flipout (DenseFlipout, edward2/tensorflow/layers/dense.py:255-285): gemm_flipout.cuh ------------------------
// A per-example sign stream [rows][N]: element (m, n) = +1 when bit 31 of Philox word (row0 + m) * N + n is set
// (philox.cuh stream contract, identical to the materialised sign tensors of the unfused path).
struct SignSide {
  unsigned long long seed, stream;
  const unsigned long long* step;
  long long row0;
  int on;
};
//   v = acc0 + acc1 * sa[m,n] + bias[n];  v = act(v);  v = mask_src[m,n] > 0 ? v : 0;  out = bf16(v)
//   out2 = out * sb[m,n] (optional);  col_sum[n] += sum_m out (optional, fp32 atomics)
struct FlipoutEpi {
  SignSide sa, sb;
  const float* bias;
  int act;
  const void* mask_src; int mask_ld;  // bf16 [M][N]
  void* out2; int out2_ld;            // bf16 [M][N]
  float* col_sum;                     // [N], zero-initialised by the caller
  // ---- weight-gradient mode (gemm_flipout.cuh kMode 1): sa carries the eps stream of the kernel posterior
  //   dmu[m,n] = acc0 + klg (mu - prior_mean) / prior_std^2
  //   drho[m,n] = (acc1 * eps[m,n] + klg (sigma / prior_std^2 - 1 / sigma)) * sigmoid(rho)
  const float* mu; const float* rho;  // [M][N] fp32 contiguous
  float* dmu; float* drho;            // [M][N] fp32 contiguous
  float klg; const double* kl_upstream;
  float prior_mean, prior_std;
};

}  // namespace ed2
