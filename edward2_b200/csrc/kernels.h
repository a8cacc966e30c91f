// Host launchers of the element-wise / reduction kernels (elementwise.cu, sngp.cu).  All enqueue on `s`.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "philox.cuh"

namespace ed2 {

struct Prior {
  float mean, stddev;
};

int k_philox_fill(int kind /*0 normal 1 uniform 2 sign*/, float* out, long long n, uint64_t seed, uint64_t stream,
                  cudaStream_t s);
int k_philox_raw(uint32_t* out, long long n_blocks, uint64_t seed, uint64_t stream, cudaStream_t s);
int k_cast2d(const void* src, int sdt, long long sld, void* dst, int ddt, long long dld, long long rows,
             long long cols, cudaStream_t s);
int k_split_bf16(const void* src, int sdt, long long sld, void* dst, long long dld, long long rows, int cols, int is_b,
                 int nsplit, cudaStream_t s);
int k_sample_weights(const float* mu, const float* rho, Noise eps, long long rows, long long cols, int mode, void* w,
                     int wdt, long long wld, void* mean_out, long long mean_ld, double* kl_out, float kl_scale,
                     Prior p, cudaStream_t s);
// bf16 layers: KL reduced inside the lean sampling kernel (default) instead of a separate accurate pass
bool fused_kl_enabled();
int k_normal_kl_fwd(const float* mu, const float* rho, long long n, Prior p, float scale, double* out, cudaStream_t s);
int k_normal_kl_bwd(const float* mu, const float* rho, long long n, Prior p, float scale, const float* upstream,
                    float* dmu, float* drho, cudaStream_t s);
// dmu = d_mean + kl*(mu-m)/s^2 ; drho = (d_pert*eps + kl*(sigma/s^2 - 1/sigma)) * sigmoid(rho)
int k_normal_grad_finish(const float* d_mean, const float* d_pert, const float* mu, const float* rho, Noise eps,
                         long long n, float kl_grad_scale, const double* kl_upstream, Prior p, float* dmu, float* drho,
                         bool fast, cudaStream_t s, void* dmu_lp = nullptr, void* drho_lp = nullptr);

// dW given in fp32 or bf16 (an all-reduced bucket); same dW feeds both the mean and the perturbation term
int k_normal_grad_finish_any(const void* dw, int dw_dt, const float* mu, const float* rho, Noise eps, long long n,
                             float kl_grad_scale, const double* kl_upstream, Prior p, float* dmu, float* drho, bool fast,
                             cudaStream_t s);

// ---- forward input preparation:  xo = x ∘ f  (f = alpha[e] | clip(mu+sigma*eps) | sign), dt in/out
// mode 0: per-member vector fw_mu[e][cols] (BatchEnsemble alpha)
// mode 1: per-example sampled fast weight clip(fw_mu[e] + softplus(fw_rho[e])*eps[b], -10, 10) (rank-1)
// mode 2: per-example sign noise (Flipout sign_input)
// s_out (optional, modes 0/1): the factor itself written as a [rows][cols] tensor in dt (rank-1 `s`).
// opts: clip bounds of the sampled factor (mode 1; edward2/tensorflow/layers/dense.py:1108-1110) and the additive
// variant xo = x + f (dense.py:1126-1127).
struct FastOpts {
  float clip_lo = -10.f, clip_hi = 10.f;
  int additive = 0;
};
int k_scale_rows(const void* x, int dt, long long x_ld, long long rows, int cols, int rows_per_member, int mode,
                 const float* fw_mu, const float* fw_rho, Noise nz, void* xo, long long xo_ld, void* f_out,
                 long long f_ld, cudaStream_t s, FastOpts opts = FastOpts());

// sg = softplus(rho) + 1e-7, sgm = sigmoid(rho): the per-(member, column) tables the fused rank-1 GEMM epilogues read
// (rho may be null: tables not needed); z0..z2: optional [n] fp32 accumulators zeroed by the same launch
int k_layernorm_split(const void* x, int xdt, long long x_ld, void* out, long long out_ld, long long rows, int cols,
                      const float* gamma, const float* beta, float eps, cudaStream_t s);
int k_mul_bf16(const void* a, long long a_ld, const void* b, long long b_ld, void* out, long long out_ld, long long rows,
               int cols, cudaStream_t s);
int k_mirror_lower(float* p, int n, cudaStream_t s);          // P[i][j] = P[j][i], i < j (after a lower-triangle SYRK)
int k_fast_prep(const float* rho, long long n, float* sg, float* sgm, cudaStream_t s, float* z0 = nullptr,
                float* z1 = nullptr, float* z2 = nullptr);

// ---- backward preparation on the output side of a layer:
//   gp = gy ∘ act'(y)
//   plain (fast == 0): dz = gp                   dbias[n]    = Σ_m gp
//   BE    (fast == 1): dz = gp ∘ gamma[e]        dgamma[e,n] = Σ gp ∘ z     dbias[e,n] = Σ gp
//   rank1 (fast == 2): dz = gp ∘ s[b,n]          dmu_s[e,n]  = Σ gp∘z∘1[lo<=s_raw<=hi]   drho_s likewise ∘ eps ∘ sigmoid(rho)
//        additive:     dz = gp                   dmu_s[e,n]  = Σ gp∘1[lo<=s_raw<=hi]
//        sampled bias (rho_b != nullptr): d_bias_rho[e,n] = Σ gp ∘ eps_b ∘ sigmoid(rho_b)   (dbias = Σ gp is d mu_b)
//   flipout (fast == 3): dz = gp,  dz2 = gp ∘ T[b,n]   dbias[n] = Σ gp
struct BwdOutArgs {
  int dt;
  long long rows;
  int cols;
  int rows_per_member;  // rows for plain/flipout
  int act;
  int fast;
  const void* gy; long long gy_ld;
  const void* y; long long y_ld;
  const void* z; long long z_ld;
  const void* s_buf; long long s_ld;
  const float* gamma;           // BE
  const float* mu_s; const float* rho_s; Noise eps_s;  // rank-1
  FastOpts opts;                                       // rank-1 clip bounds / additive variant
  int lean_rows = 0;                                   // rows per block of the lean kernels (set by k_bwd_out)
  const float* rho_b; Noise eps_b; float* d_bias_rho;  // rank-1 per-example sampled bias (dense.py:1131-1139)
  Noise sign_out;               // flipout (injected fp32 or philox sign)
  void* dz; long long dz_ld;
  void* dz2; long long dz2_ld;
  float* d_fast_mu;   // dgamma | dmu_s        [E][cols], pre-zeroed by the launcher
  float* d_fast_rho;  // drho_s or nullptr
  float* dbias;       // [E][cols] or nullptr
};
int k_bwd_out(const BwdOutArgs& a, cudaStream_t s);

// ---- backward on the input side: given dxf = d(x∘f) [rows][cols] (dt)
//   BE   : dx = dxf ∘ alpha[e]; dalpha[e,i] = Σ dxf ∘ x
//   rank1: dx = dxf ∘ r[b,i];   dmu_r = Σ dxf∘x∘1[] ; drho_r = Σ dxf∘x∘eps∘sigmoid(rho)∘1[]
struct BwdInArgs {
  int dt;
  long long rows;
  int cols;
  int rows_per_member;
  int fast;  // 1 BE, 2 rank-1
  const void* dxf; long long dxf_ld;
  const void* x; long long x_ld;
  const float* fw_mu; const float* fw_rho; Noise eps;
  FastOpts opts;
  int lean_rows = 0;          // rows per block of the lean kernel (set by k_bwd_in)
  void* dx; long long dx_ld;  // may be nullptr
  float* d_fast_mu; float* d_fast_rho;
};
int k_bwd_in(const BwdInArgs& a, cudaStream_t s);

// flipout: out[b,i] (dt) = a[b,i] (fp32) ∘ S + c ... handled by GEMM epilogue; this materialises a sign tensor
int k_fill_noise(int kind, Noise nz, void* out, int dt, long long ld, long long rows, long long cols, cudaStream_t s);

// ---- SNGP
int k_spectral_norm(const float* w, int in_dim, int out_dim, float* u, float* v, int iterations, float c, int mode,
                    int update_uv, float* w_out, void* w_out_lp, long long w_lp_ld, float* sigma_out, float* ws,
                    cudaStream_t s);
int k_input_normalize(const void* x, int xdt, long long x_ld, void* y, int ydt, long long y_ld, long long rows,
                      int cols, const float* gamma, const float* beta, float eps, int scale_only, float scale,
                      cudaStream_t s);
int k_layernorm_bwd(const void* x, int xdt, long long x_ld, const void* dy, int dydt, long long dy_ld, long long rows,
                    int cols, const float* gamma, float eps, void* dx, int dxdt, long long dx_ld, float* dgamma,
                    float* dbeta, cudaStream_t s);
int k_precision_blend(float* precision, const float* summed, long long n, float momentum, long long batch_total,
                      cudaStream_t s);
int k_mean_field(const float* logits, const float* var, long long batch, int classes, float factor, int likelihood,
                 float var_scale, float* out, cudaStream_t s);
int k_rff_bwd_prep(const void* dphi, int ddt, long long dphi_ld, const float* pre, long long pre_ld, const float* bias, float scale,
                   long long rows, int cols, void* gpre, int gdt, long long gpre_ld, cudaStream_t s);
struct XentExtras {
  int n;
  const double* terms[8];
};
int k_softmax_xent(const void* logits, int dt, long long ld, const long long* labels, long long rows, int cols,
                   float loss_scale, float grad_scale, double* loss, void* dlogits, int ddt, long long dld,
                   const XentExtras& ex, cudaStream_t s);
int k_scale_inplace(float* x, long long n, float a, cudaStream_t s);


// ---- peer-memory gradient exchange of the data-parallel step (peer.cuh explains the layout; peer.cu, elementwise.cu)
constexpr int kMaxPeers = 8;
struct PeerExchange {
  int world, rank;
  long long n;      // elements of the exchanged bf16 gradient (multiple of 8)
  long long chunk;  // elements per owner (multiple of 8); owner o reduces [o*chunk, min(n, (o+1)*chunk))
  long long small_n;  // elements of the fp32 companion vector (0: none)
  const void* bucket[kMaxPeers];
  void* reduced[kMaxPeers];
  const float* small[kMaxPeers];
  unsigned long long* flags[kMaxPeers];
};
// publish this rank's bucket (+ copy `small_src` into the arena): epoch += 1, ready flag on every rank
int k_peer_signal(const PeerExchange& x, const float* small_src, cudaStream_t s);
int k_peer_flag(const PeerExchange& x, int which, cudaStream_t s);
int k_peer_wait(const PeerExchange& x, int which, cudaStream_t s);
// owner pass: wait for every rank's ready flag, pull + sum the owned slice (rank order), publish the reduced flag
// light: small footprint (1 block x 128 threads per SM) for running beside the finishing pass instead of a GEMM
int k_peer_reduce_slice(const PeerExchange& x, int light, cudaStream_t s);
// plain finishing pass: out[i] = float(reduced gradient), slice by slice from its owners
int k_peer_gather_f32(const PeerExchange& x, float* out, cudaStream_t s);
// finishing pass reading the reduced gradient slice by slice from its owners; small_out = sum over ranks of `small`
int k_peer_grad_finish(const PeerExchange& x, const float* mu, const float* rho, Noise eps, float kl_grad_scale,
                       const double* kl_upstream, Prior p, float* dmu, float* drho, float* small_out, bool fast,
                       int blocks_per_sm, cudaStream_t s);
// ---- Conv2D variants on the dense path (conv.cu): NHWC, dilation 1, padding 'valid' (same = 0) or TF 'SAME' (same = 1)
struct ConvArgs {
  int batch, height, width, channels;
  int kh, kw, sh, sw;
  int same;
};
int k_conv_out_size(const ConvArgs& a, int* ho, int* wo);
// cols[(b,ho,wo)][(i,j,c)] = x[b][ho*sh+i-pt][wo*sw+j-pl][c] (zero outside); x contiguous NHWC, cols pitch cols_ld
// fac (optional): per-example channel factors [B][C] fp32 applied while gathering (cols = im2col(x ∘ fac[b]))
int k_im2col(const ConvArgs& a, const void* x, int dt, void* cols, long long cols_ld, cudaStream_t s,
             const float* fac = nullptr);
// image-space tail of a factorised convolution's backward: dx = gimg ∘ fac[b], dfac[b][c] = sum_hw gimg * x
int k_conv_scale_reduce(const ConvArgs& a, const void* gimg, const void* x, int dt, const float* fac, void* dx, int dx_dt,
                        float* dfac, cudaStream_t s);
// adjoint (gather form): dx[b][h][w][c] = sum of the patch entries that read it
int k_col2im(const ConvArgs& a, const void* dcols, int dt, long long cols_ld, void* dx, int dx_dt, cudaStream_t s);
// out[b][t*C + c] = src[b / rows_per_entry][c] (t < reps) and its adjoint; fp32 tables
int k_expand_table(const float* src, float* out, int B, int C, int rows_per_entry, int reps, cudaStream_t s);
int k_expand_table_bwd(const float* dout, float* dsrc, int B, int C, int rows_per_entry, int reps, cudaStream_t s);
// gradient of the rank-1 factor table f[b][c] = clip(mu[e][c] + softplus(rho[e][c]) eps[b][c]) (rho may be null)
int k_rank1_factor_bwd(const float* df, const float* mu, const float* rho, Noise eps, int E, int n, int C, float lo, float hi,
                       float* dmu, float* drho, cudaStream_t s);

// ---- Monte-Carlo softmax of the heteroscedastic heads (mcsoftmax.cu; heteroscedastic.py:178-250,749-785, hetsngp.py:143-170)
int k_mc_softmax_fwd(const float* loc, const float* fac, const float* diag_raw, const float* sngp_var, float w_het,
                     float w_sngp, Noise noise, int B, int S, int K, int F, int shared, float temperature, float eps,
                     float* probs_mean, float* logits, float* variance, cudaStream_t s);
int k_mc_softmax_bwd(const float* g, const float* probs_mean, const float* loc, const float* fac, const float* diag_raw,
                     Noise noise, int B, int S, int K, int F, int shared, float temperature, float eps, float* dloc,
                     float* dfac, float* ddiag_raw, cudaStream_t s);

// ---- pointwise part of the Bayesian LSTM cells (recurrent.cu; recurrent.py:163-184, gate order i | f | c | o)
int k_lstm_pointwise_fwd(const void* z, int dt, long long z_ld, const float* c_prev, int B, int H, float* c, void* h,
                         long long h_ld, cudaStream_t s, const void* z2 = nullptr, long long z2_ld = 0);
int k_lstm_pointwise_bwd(const void* z, int dt, long long z_ld, const float* c_prev, const float* c, const void* dh,
                         long long dh_ld, const float* dc, int B, int H, void* dz, long long dz_ld, float* dc_prev,
                         float* dbias, cudaStream_t s, const void* z2 = nullptr, long long z2_ld = 0);

// ---- SPD inverse without a library call (inverse.cu): in-place block Gauss-Jordan, fp32; ws >= workspace_floats(n)
long long k_spd_inverse_workspace_floats(int n);
int k_spd_inverse(float* a, int n, float* ws, cudaStream_t s);

}  // namespace ed2
