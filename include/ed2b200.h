/* ed2b200.h — C ABI of libed2b200.so: sm_100a (B200) kernels for the Edward2 stochastic-weight dense layers
 * and the SNGP random-feature Gaussian-process head.
 *
 * The reference (google/edward2 0.0.3) is pure Python over TF/JAX and has NO native boundary for this path
 * (SURVEY.md §0.2); this header DEFINES the boundary a TF custom op / jax.ffi custom call / ctypes stub
 * binds to.  Each entry point cites the reference lines whose arithmetic it replaces (paths relative to
 * the reference checkout).
 *
 * Conventions
 *  - every data pointer is a DEVICE pointer owned by the caller; the library never frees caller memory;
 *  - scratch memory comes from a caller-provided workspace sized by the matching *_workspace_bytes();
 *  - every entry point only ENQUEUES work on `stream` (a cudaStream_t passed as void*) and returns;
 *  - return value: 0 = OK, negative = ed2_status; text via ed2_last_error() (thread-local);
 *  - tensors are row-major with an explicit leading dimension in ELEMENTS;
 *  - dtype: ED2_F32 or ED2_BF16 for activations; parameters and gradients of parameters are fp32;
 *  - randomness: every noise tensor is EITHER injected (non-NULL device pointer to fp32 values, used by the
 *    parity tests to inject the reference's draws) OR drawn in-kernel from Philox4x32-10 keyed by
 *    (seed, stream) and indexed by flat element index (ed2_noise.injected == NULL);
 *  - there is no CPU fallback: on a machine without an sm_100 GPU every compute entry point fails.
 */
#ifndef ED2B200_H_
#define ED2B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  ED2_OK = 0,
  ED2_ERR_BAD_SHAPE = -1,
  ED2_ERR_BAD_DTYPE = -2,
  ED2_ERR_BAD_ARG = -3,
  ED2_ERR_UNSUPPORTED_ARCH = -4,
  ED2_ERR_CUDA = -5,
  ED2_ERR_WORKSPACE = -6
} ed2_status;

typedef enum { ED2_F32 = 0, ED2_BF16 = 1 } ed2_dtype;
typedef enum { ED2_ACT_NONE = 0, ED2_ACT_RELU = 1, ED2_ACT_COS = 2 } ed2_activation;
typedef enum { ED2_MAJOR_K = 0, ED2_MAJOR_MN = 1 } ed2_major;

/* A noise tensor: injected values or a Philox stream (see Conventions). */
typedef struct {
  const float* injected; /* NULL -> Philox(seed, stream) */
  uint64_t seed;
  uint64_t stream;
  const uint64_t* step;  /* optional DEVICE counter folded into the key: seed + step[0] * 0x9E3779B97F4A7C15, so a
                            captured CUDA graph draws fresh noise on every replay */
  /* Row mapping of PER-EXAMPLE noise tensors [rows][cols] (rank-1 eps_r / eps_s / eps_b, Flipout sign_in / sign_out)
   * under data parallelism: the Philox element index of local (row m, column c) is global_row(m) * cols + c with
   *   global_row(m) = m + row0                                                     (rows_per_group_local == 0)
   *   global_row(m) = g * rows_per_group_global + row0 + (m - g * rows_per_group_local),  g = m / rows_per_group_local
   * so that a rank holding rows [row0, row0 + n_local) of every ensemble member (member-major layout,
   * edward2/tensorflow/layers/dense.py:572-573, 1101-1102) draws exactly the noise the single-device global batch
   * would draw for those rows.  All zero = identity (single device, and every per-weight noise tensor).  Injected
   * tensors are always indexed by the LOCAL row. */
  long long row0;
  int rows_per_group_local;
  int rows_per_group_global;
} ed2_noise;

const char* ed2_version(void);
const char* ed2_last_error(void);
/* number of kernels this library has launched in this process (bench.py reports the delta) */
long long ed2_launch_count(void);
/* 0 if device `dev` is an sm_100 part this library can run on */
int ed2_check_device(int dev);

/* ------------------------------------------------------------------------------------------------
 * Generic fused GEMM (building block of every layer below; exported for tests and for composition).
 *   out = act( ((alpha * A·B) ∘ elem_scale ∘ col_scale[g(m)]) + beta * addend + col_bias[g(m)] )
 *   raw = alpha * A·B                      (optional)
 *   row_sum[m] += sum_n out[m,n]           (optional, fp32 atomics; `out` may then be NULL)
 *   mask_src / col_sum: ReLU-backward mask and bias-gradient column sums fused into a dX product
 * g(m) = m / group_rows (0 when group_rows == 0).  dtype selects the engine: ED2_BF16 = tcgen05 tensor
 * cores (bf16 in, fp32 accumulate in TMEM); ED2_F32 = fp32 path (exact fp32 products, fp32 accumulate).
 * Layouts: A K-major = [M][lda], MN-major = [K][lda]; B K-major = [N][ldb], MN-major = [K][ldb]. */
typedef struct {
  int dtype; /* operand dtype: ED2_BF16 or ED2_F32 */
  const void* a; int lda; int major_a;
  const void* b; int ldb; int major_b;
  int m, n, k;
  float alpha, beta, act_scale;
  int act;
  int group_rows;
  const void* elem_scale; int elem_scale_ld; int elem_scale_dtype;
  const float* col_scale; int col_scale_ld;
  const void* addend; int addend_ld; int addend_dtype;
  const float* col_bias; int col_bias_ld;
  void* out; int out_ld; int out_dtype;
  void* raw; int raw_ld; int raw_dtype;
  float* row_sum;
  const void* mask_src; int mask_ld; int mask_dtype;  /* out = mask_src[m,n] > 0 ? out : 0 (applied last) */
  float* col_sum; int col_sum_ld;                     /* col_sum[g(m), n] += sum_m out[m,n] (fp32 atomics) */
  /* tuning / test knobs, 0 = automatic */
  int block_n, cta_group, max_ctas, no_tma_store;
  int accumulate;  /* 1 (ED2_BF16, fp32 out): out += alpha * A B into a PRE-INITIALISED out; products with few output tiles
                      and a long K are cut into K slices that accumulate with fp32 atomics so that every SM has work */
  int lower_only;  /* 1: square product, only the tiles on or below the diagonal are computed (symmetric results) */
} ed2_gemm_desc;

int ed2_gemm(const ed2_gemm_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Philox sampler, exposed for the statistical tests.  out[i], i in [0,n), fp32.
 * Replaces tf.random.* / tfp Normal.sample at edward2/tensorflow/random_variable.py:161-174 and
 * edward2/tensorflow/layers/dense.py:261-270. */
int ed2_philox_normal(float* out, long long n, uint64_t seed, uint64_t stream, void* s);
/* the same stream through the hardware-approximate Box-Muller (lg2 / sin / cos / sqrt.approx) that every bf16 path
 * uses — the lean sampling / finishing kernels and the fused rank-1 GEMM epilogues — exposed for the statistical tests */
int ed2_philox_normal_fast(float* out, long long n, uint64_t seed, uint64_t stream, void* s);
int ed2_philox_uniform(float* out, long long n, uint64_t seed, uint64_t stream, void* s);
int ed2_philox_sign(float* out, long long n, uint64_t seed, uint64_t stream, void* s);
int ed2_philox_raw(uint32_t* out, long long n_blocks, uint64_t seed, uint64_t stream, void* s);

/* ------------------------------------------------------------------------------------------------
 * Elementwise building blocks. */
/* dst[r, c] = (dtype_out) src[r, c]; rows x cols with pitches (elements); used to make bf16 working copies */
int ed2_cast(const void* src, int src_dtype, int src_ld, void* dst, int dst_dtype, int dst_ld, long long rows,
             long long cols, void* s);

/* Split an fp32 (or bf16) matrix [rows][cols] into bf16 terms laid out along the contraction axis, so that ONE bf16
 * tensor-core GEMM with K' = 6K (nsplit = 3: 24 mantissa bits) or 3K (nsplit = 2: 16 bits) and fp32 accumulation
 * evaluates an fp32-grade product:  is_b = 0 -> A' = [h|h|h|m|m|l] ([h|h|m]);  is_b = 1 -> B' = [h|m|l|h|m|h] ([h|m|h]).
 * dst: bf16 [rows][dst_ld], dst_ld >= terms * cols.  Used for the random-feature phase x W + b, where bf16 operands
 * would give ~0.1 rad of phase error with the TF defaults (SURVEY.md §7 hard part 6). */
int ed2_split_bf16(const void* src, int src_dtype, int src_ld, void* dst, int dst_ld, long long rows, int cols, int is_b,
                   int nsplit, void* s);

/* TrainableNormal sample (edward2/tensorflow/initializers.py:500-511, constraints.py:62-63):
 *   sigma = softplus(rho) + 1e-7;  w = mu + sigma * eps   (mode 0)   or   delta = sigma * eps   (mode 1)
 * over `rows x cols` fp32 parameters with pitch cols; output `w` in w_dtype with pitch w_ld.  If kl_out is
 * non-NULL also accumulates kl_scale * sum KL(N(mu,sigma) || N(prior_mean, prior_std))
 * (edward2/tensorflow/regularizers.py:175-186) into *kl_out (fp64 device scalar, zeroed by the caller) while mu/rho
 * are in registers.  mean_out (optional) receives (w_dtype) mu — the Flipout mean operand. */
int ed2_sample_normal_weights(const float* mu, const float* rho, ed2_noise eps, long long rows, long long cols,
                              int mode, void* w, int w_dtype, int w_ld, void* mean_out, int mean_ld,
                              double* kl_out, float kl_scale, float prior_mean, float prior_std, void* s);

/* NormalKLDivergence (edward2/tensorflow/regularizers.py:166-186): *out += scale * sum_i KL(q_i || p).
 * sigma = softplus(rho) + 1e-7.  Warp-shuffle + block reduction, fp64 final accumulation. */
int ed2_normal_kl_fwd(const float* mu, const float* rho, long long n, float prior_mean, float prior_std,
                      float scale, double* out, void* s);
/* d_mu += g * scale * (mu - m)/s^2 ;  d_rho += g * scale * (sigma/s^2 - 1/sigma) * sigmoid(rho);
 * g = *upstream (device scalar) or 1 if NULL. */
int ed2_normal_kl_bwd(const float* mu, const float* rho, long long n, float prior_mean, float prior_std,
                      float scale, const float* upstream, float* d_mu, float* d_rho, void* s);

/* Softmax cross-entropy with integer labels (int64), loss and gradient in one pass — the synthetic training loss of
 * the benchmark harness (the reference's scripts use tf.keras.losses.sparse_categorical_crossentropy,
 * experimental/marginalization_mixup/sngp.py:400-409).  *loss (fp64, caller-zeroed) += loss_scale * sum_rows CE;
 * dlogits = grad_scale * (softmax - onehot) (optional).  extra_terms: HOST array of n_extra (<= 8) device pointers to
 * fp64 scalars (the layers' KL regularisers) that are added into *loss by the same kernel. */
int ed2_softmax_xent(const void* logits, int dtype, int ld, const long long* labels, long long rows, int cols,
                     float loss_scale, float grad_scale, double* loss, void* dlogits, int d_dtype, int d_ld,
                     const double* const* extra_terms, int n_extra, void* s);

/* gp = gy ∘ act'(y) (relu' evaluated on the output);  dbias[n] = sum_m gp[m,n] (optional, overwritten).
 * The output-side backward step of a plain dense layer (the layer SpectralNormalization wraps). */
int ed2_act_bwd(int dtype, const void* gy, int gy_ld, const void* y, int y_ld, long long rows, int cols, int act,
                void* gp, int gp_ld, float* dbias, void* s);

/* ------------------------------------------------------------------------------------------------
 * DenseBatchEnsemble (edward2/tensorflow/layers/dense.py:551-584; edward2/jax/nn/dense.py:258-272)
 *   y[e,n,:] = act( ((x[e,n,:] ∘ alpha[e]) W) ∘ gamma[e] + bias[e] ),   batch rows member-major.
 * Saves for backward: xa = x∘alpha (dtype, [B][xa_ld]) and z = (x∘alpha)W (dtype, [B][z_ld]).
 * w_lp: [I][w_ld] working copy of W in `dtype` (== w when dtype is fp32). */
typedef struct {
  int dtype;
  int batch, in_dim, out_dim, ensemble_size;
  int act;
  const void* x; int x_ld;
  const void* w_lp; int w_ld;
  const float* alpha;     /* [E][in_dim]  */
  const float* gamma;     /* [E][out_dim] */
  const float* bias;      /* [E][out_dim] or NULL */
  void* y; int y_ld;
  void* xa; int xa_ld;    /* saved */
  void* z; int z_ld;      /* saved; may be NULL for inference */
} ed2_be_fwd_args;
int ed2_be_dense_fwd(const ed2_be_fwd_args* a, void* stream);

typedef struct {
  int dtype;
  int batch, in_dim, out_dim, ensemble_size;
  int act;
  const void* gy; int gy_ld;   /* dL/dy */
  const void* y; int y_ld;
  const void* z; int z_ld;
  const void* x; int x_ld;
  const void* xa; int xa_ld;
  const void* w_lp; int w_ld;
  const float* alpha; const float* gamma;
  void* dz; int dz_ld;         /* workspace [B][out_dim] in dtype */
  void* dxa; int dxa_ld;       /* workspace [B][in_dim] in dtype */
  void* dx; int dx_ld;         /* NULL to skip (first layer) */
  float* dw;                   /* [I][O] fp32, overwritten */
  float* dalpha; float* dgamma; float* dbias; /* [E][.] fp32, overwritten; dbias may be NULL */
} ed2_be_bwd_args;
int ed2_be_dense_bwd(const ed2_be_bwd_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * DenseRank1 (edward2/tensorflow/layers/dense.py:1096-1144):
 *   r[b,:] = clip(mu_r[e] + sigma_r[e]*eps_r[b,:], clip_lo, clip_hi)   (per example, when sampled; s likewise)
 *   y = act( ((x ∘ r) W) ∘ s + bias )        multiplicative (default)
 *   y = act( ((x + r) W) + s + bias )        additive != 0 (dense.py:1126-1127)
 *   bias[b,:] = bias[e]                                       (rho_b == NULL)
 *   bias[b,:] = bias[e] + sigma_b[e]*eps_b[b,:], unclipped    (rho_b != NULL: per-example sampled, dense.py:1131-1139)
 * If rho_r (rho_s) is NULL the fast weight is deterministic: r[b,:] = mu_r[e].
 * clip_lo == clip_hi == 0 selects the reference defaults [-10, 10] (dense.py:1025-1026). */
typedef struct {
  int dtype;
  int batch, in_dim, out_dim, ensemble_size;
  int act;
  const void* x; int x_ld;
  const void* w_lp; int w_ld;
  const float* mu_r; const float* rho_r; ed2_noise eps_r;  /* [E][in_dim],  noise [B][in_dim]  */
  const float* mu_s; const float* rho_s; ed2_noise eps_s;  /* [E][out_dim], noise [B][out_dim] */
  const float* bias;                                        /* [E][out_dim] or NULL */
  void* y; int y_ld;
  void* xr; int xr_ld;      /* saved x∘r */
  void* s_buf; int s_ld;    /* saved s [B][out_dim] in dtype */
  void* z; int z_ld;        /* saved (x∘r)W; may be NULL for inference */
  int additive; float clip_lo, clip_hi;
  const float* rho_b; ed2_noise eps_b;   /* per-example sampled bias: bias = mean [E][out_dim], rho_b [E][out_dim] */
  void* bias_buf; int bias_ld;           /* workspace [B][out_dim] in dtype, required when rho_b != NULL */
  /* ---- fused path (see ed2_rank1_fusable): s — and the NEXT layer's r — are drawn inside the GEMM epilogue, once per
   * output element, so neither s nor the next layer's x∘r costs a pass over HBM.
   *   fast_ws   caller workspace of ed2_rank1_workspace_floats() floats; NULL selects the unfused path (s_buf required).
   *             On the fused path s_buf may be NULL (the backward regenerates s).
   *   xr_ready  xr already holds x∘r: the producer layer wrote it through its next_xr.
   *   next_*    fast weights of the layer that consumes y ([E][out_dim]; next_rho_r NULL = deterministic) and the
   *             buffer next_xr [B][out_dim] that receives bf16(y) ∘ r_next.  next_mu_r NULL: not produced. */
  float* fast_ws;
  int xr_ready;
  const float* next_mu_r; const float* next_rho_r; ed2_noise next_eps_r;
  void* next_xr; int next_xr_ld;
  void* gemm_ready_event;  /* optional cudaEvent_t recorded on `stream` right BEFORE the layer's GEMM is enqueued: side-stream
                              work that should share the SMs with that GEMM (e.g. the next layer's weight cast) waits on it
                              and becomes runnable at the same moment, so the block scheduler places the higher-priority
                              GEMM CTAs first and the side work takes the SMs the GEMM grid leaves free */
} ed2_rank1_fwd_args;
int ed2_rank1_dense_fwd(const ed2_rank1_fwd_args* a, void* stream);

/* 1 when the fused rank-1 epilogues apply to a layer call of this shape in ED2_BF16 with Philox noise, multiplicative
 * perturbations and a deterministic bias (forward: cols = out_dim; backward dX side: cols = in_dim). */
int ed2_rank1_fusable(int batch, int cols, int ensemble_size);
/* floats of workspace one fused forward or backward call needs */
long long ed2_rank1_workspace_floats(int ensemble_size, int in_dim, int out_dim);

typedef struct {
  int dtype;
  int batch, in_dim, out_dim, ensemble_size;
  int act;
  const void* gy; int gy_ld;
  const void* y; int y_ld;
  const void* z; int z_ld;
  const void* x; int x_ld;
  const void* xr; int xr_ld;
  const void* s_buf; int s_ld;
  const void* w_lp; int w_ld;
  const float* mu_r; const float* rho_r; ed2_noise eps_r;
  const float* mu_s; const float* rho_s; ed2_noise eps_s;
  void* dz; int dz_ld;
  void* dxr; int dxr_ld;
  void* dx; int dx_ld;
  float* dw;                /* fp32 [in_dim][out_dim]; may be NULL when dw_lp is given */
  float* dmu_r; float* drho_r; float* dmu_s; float* drho_s; float* dbias; /* drho_* / dbias may be NULL */
  int additive; float clip_lo, clip_hi;
  const float* rho_b; ed2_noise eps_b; float* drho_b;  /* sampled bias: dbias is d(bias mean), drho_b [E][out_dim] */
  /* ---- fused path (fast_ws != NULL and ed2_rank1_fusable(batch, in_dim, E)): the dX GEMM epilogue computes
   * dx = dxr ∘ r and the dmu_r / drho_r column sums itself (dxr is never written; dxr may be NULL), and — when the
   * PRODUCER layer's tensors are given (prev_mu_s != NULL; its output y is this layer's x) — also the producer's
   * output side:  gp = dx ∘ act'(x),  prev_dz = gp ∘ s_prev,  prev_dmu_s / prev_drho_s / prev_dbias.
   *   dz_ready  this layer's dz, dmu_s, drho_s, dbias were already produced that way by ITS consumer: skip that pass
   *             (gy, y, z, s_buf are then unused). */
  float* fast_ws;
  int dz_ready;
  const void* prev_z; int prev_z_ld;
  const float* prev_mu_s; const float* prev_rho_s; ed2_noise prev_eps_s; int prev_act;
  void* prev_dz; int prev_dz_ld;
  float* prev_dmu_s; float* prev_drho_s; float* prev_dbias;
  /* data parallel: when non-NULL the dW GEMM writes the kernel gradient as ED2_BF16 [in_dim][out_dim] (pitch out_dim)
   * straight into this buffer — the gradient-exchange bucket — instead of fp32 `dw` */
  void* dw_lp;
  void* dw_done_event;  /* optional cudaEvent_t recorded on `stream` right after the dW GEMM is enqueued: the caller's
                           exchange of dW then overlaps the dX GEMM that follows */
} ed2_rank1_bwd_args;
int ed2_rank1_dense_bwd(const ed2_rank1_bwd_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * DenseReparameterization (edward2/tensorflow/layers/dense.py:73-83):  y = act(x (mu + sigma∘eps) + b)
 * w_lp receives the sampled kernel in `dtype` ([I][w_ld]); it stays L2-resident between the sampling pass
 * and the GEMM and is re-used by the backward pass.  kl_out as in ed2_sample_normal_weights. */
typedef struct {
  int dtype;
  int batch, in_dim, out_dim;
  int act;
  const void* x; int x_ld;
  const float* mu; const float* rho; ed2_noise eps;   /* [I][O] fp32 */
  const float* bias;                                   /* [O] fp32 or NULL */
  void* w_lp; int w_ld;
  void* y; int y_ld;
  double* kl_out; float kl_scale; float prior_mean; float prior_std;
} ed2_reparam_fwd_args;
int ed2_reparam_dense_fwd(const ed2_reparam_fwd_args* a, void* stream);

typedef struct {
  int dtype;
  int batch, in_dim, out_dim;
  int act;
  const void* gy; int gy_ld;
  const void* y; int y_ld;
  const void* x; int x_ld;
  const void* w_lp; int w_ld;
  const float* mu; const float* rho; ed2_noise eps;
  void* gp; int gp_ld;        /* workspace [B][O] dtype: gy ∘ act'(y) */
  void* dx; int dx_ld;        /* NULL to skip */
  float* dmu; float* drho;    /* [I][O] fp32, overwritten: likelihood term + kl_grad_scale * dKL */
  float* dbias;               /* [O] or NULL */
  float kl_grad_scale; float prior_mean; float prior_std;   /* kl_grad_scale = 0 -> no KL term */
  const double* kl_upstream;   /* optional fp64 device scalar multiplying kl_grad_scale (dLoss/dKL) */
  /* Cross-layer fusion (optional).  gy_premasked != 0: gy already equals gy ∘ act'(y) and dbias was accumulated by the
   * layer that produced gy (its dX epilogue) -> the output-side pass is skipped and gp is not written.
   * x_is_relu != 0: x is the output of a ReLU layer -> dx := dx ∘ (x > 0) in the dX epilogue and prev_dbias[in_dim]
   * (zeroed here) receives its column sums, i.e. the producing layer's bias gradient. */
  int gy_premasked; int x_is_relu; float* prev_dbias;
  /* phase: 0 = whole backward; 1 = parameter gradients only (output-side pass, dW GEMM, finishing pass);
   * 2 = input gradient only (dX GEMM; phase 1 must have run on the same stream).  The split lets a data-parallel
   * caller launch the gradient all-reduce of this layer between the phases, so it overlaps the dX GEMM.
   * dmu_lp / drho_lp (optional, bf16, [I][O] dense): copies of the gradients for bf16 all-reduce buckets, written by
   * the finishing pass itself.
   * phase 3 = like 1 but WITHOUT the finishing pass: the dW GEMM writes the raw weight gradient as bf16 into dmu_lp
   * (the all-reduce bucket); the caller reduces it and finishes with ed2_normal_grad_finish.  Valid when every rank
   * draws the same eps (same seed), i.e. the data-parallel step reproduces the single-device global-batch step. */
  int phase; void* dmu_lp; void* drho_lp;
} ed2_reparam_bwd_args;
int ed2_reparam_dense_bwd(const ed2_reparam_bwd_args* a, void* stream);

/* Gradient-finishing pass on its own: (dmu, drho) from a weight gradient dW [n] (fp32 or bf16, e.g. a bf16 bucket that
 * was just all-reduced across data-parallel ranks, every rank having drawn the SAME eps):
 *   dmu = dW + klg (mu - m) / s^2 ;  drho = (dW * eps + klg (sigma / s^2 - 1 / sigma)) * sigmoid(rho)
 * with klg = kl_grad_scale * (*kl_upstream).  `fast` selects the hardware-approximate math of the bf16 compute mode. */
int ed2_normal_grad_finish(const void* dw, int dw_dtype, const float* mu, const float* rho, ed2_noise eps, long long n,
                           float kl_grad_scale, const double* kl_upstream, float prior_mean, float prior_std,
                           float* dmu, float* drho, int fast, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Peer-memory gradient exchange of the data-parallel step (one process per GPU on one NVLink / NVSwitch box).
 * The reference leaves the gradient all-reduce to tf.distribute / optimizer.apply_gradients ("Gradient
 * allreduce is implicit in strategy.run"; experimental/rank1_bnns/*: loss pre-divided by num_replicas_in_sync); here
 * the exchange of a DenseReparameterization kernel's RAW bf16 weight gradient is part of the layer's backward:
 *   1. the dW GEMM (ed2_reparam_dense_bwd phase 3) writes this rank's bucket, which lives in an IPC-mapped arena;
 *   2. ed2_peer_signal publishes it (epoch flag written into every peer's memory, release at system scope);
 *      (2-4 run on side streams of the caller: 2 and 3 beside the layer's dX GEMM, 4 as soon as 3 has finished)
 *   3. ed2_peer_reduce_slice — each rank owns n / world elements: it pulls that slice from every rank's bucket over
 *      NVLink, sums in rank order (fp32) and publishes the bf16 result (reduce-scatter by pull);
 *   4. ed2_peer_grad_finish — the finishing pass of every rank reads each slice from its owner (all-gather by pull,
 *      fused into the pass that needed dW anyway) and writes dmu / drho; it also sums the ranks' fp32 companion
 *      vectors (`small`: the layer's bias gradient) one-shot.
 * No collective library call is involved and every kernel can be captured in a CUDA graph (flags carry a step
 * epoch).  A peer that never arrives makes the waiting kernel trap after 20 s instead of hanging the GPU.
 * Arena helpers: ed2_peer_alloc (cudaMalloc + zero + cudaIpcGetMemHandle), ed2_peer_open / _close (map / unmap a
 * peer's arena from its 64-byte handle), ed2_peer_free. */
#define ED2_MAX_PEERS 8
#define ED2_IPC_HANDLE_BYTES 64
typedef struct {
  int world, rank;
  long long n;       /* elements of the exchanged bf16 gradient, multiple of 8 */
  long long chunk;   /* elements per owner, multiple of 8, chunk * world >= n */
  long long small_n; /* elements of the fp32 companion vector (0 = none) */
  const void* bucket[ED2_MAX_PEERS];    /* [r]: rank r's raw bf16 gradient [n]              ([rank] = local memory) */
  void* reduced[ED2_MAX_PEERS];         /* [r]: rank r's reduced buffer [n] (its own slice is valid)              */
  const float* small_vec[ED2_MAX_PEERS];/* [r]: rank r's companion vector, TWO slots [2][small_n] (slot = epoch & 1) */
  void* flags[ED2_MAX_PEERS];           /* [r]: rank r's flag words, (2 * world + 3) x 8 bytes, zero-initialised   */
} ed2_peer_exchange;
int ed2_peer_alloc(size_t bytes, void** ptr_out, void* handle_out /* ED2_IPC_HANDLE_BYTES */);
int ed2_peer_open(const void* handle, void** ptr_out);
int ed2_peer_close(void* ptr);
int ed2_peer_free(void* ptr);
int ed2_peer_signal(const ed2_peer_exchange* e, const float* small_src, void* stream);
/* Copy-engine variant of the exchange (edward2_b200/dist.py DmaExchange): the slices travel by DMA (ed2_dma_copy =
 * cudaMemcpyAsync between IPC-mapped arenas: no SM is involved, so the transfers overlap the persistent GEMMs that own
 * every SM), the sums and the fp32 conversion read LOCAL memory only.  Per gradient: push slice q of the bucket into
 * rank q's receive slot, ed2_peer_signal (ready flags); ed2_peer_reduce_slice on a LOCAL view (bucket[r] = the receive
 * slots) sums the owned slice; push it into every rank's gathered buffer, ed2_peer_flag(which = 1); ed2_peer_gather_f32
 * on a local view converts the gathered bf16 gradient.  ed2_peer_flag publishes the current epoch into flag section
 * `which` (0 ready, 1 reduced) of every rank, after everything enqueued before it on `stream`. */
int ed2_peer_flag(const ed2_peer_exchange* e, int which, void* stream);
/* one-warp kernel that returns when flag section `which` of this rank carries the current epoch from every rank: small
 * enough to sit beside a persistent GEMM CTA, so that the heavy kernels enqueued behind it start only when their data
 * has arrived and never hold SMs while waiting for a peer */
int ed2_peer_wait(const ed2_peer_exchange* e, int which, void* stream);
int ed2_dma_copy(void* dst, const void* src, size_t bytes, void* stream);
int ed2_peer_reduce_slice(const ed2_peer_exchange* e, int light /* 1: small footprint, runs beside the finishing pass */,
                          void* stream);
/* plain finishing pass for gradients that need no further arithmetic (deterministic kernels; any fp32 gradient that was
 * cast into the bucket with ed2_cast): out[n] fp32 = the reduced bf16 gradient, each slice read from its owner */
int ed2_peer_gather_f32(const ed2_peer_exchange* e, float* out, void* stream);
int ed2_peer_grad_finish(const ed2_peer_exchange* e, const float* mu, const float* rho, ed2_noise eps, float kl_grad_scale,
                         const double* kl_upstream, float prior_mean, float prior_std, float* dmu, float* drho,
                         float* small_out, int fast, int blocks_per_sm /* persistent grid: 3 beside a GEMM, 5 alone */,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * DenseFlipout (edward2/tensorflow/layers/dense.py:255-285):
 *   y = act( x·mu + ((x ∘ S)·Delta) ∘ T + b ),  Delta = sigma ∘ eps,  S [B][I], T [B][O].
 * S and T are noise tensors of kind "sign" (Philox draws ±1; injected values may be arbitrary floats — the
 * reference's sign_output is Uniform[-1,3) for float inputs). */
typedef struct {
  int dtype;
  int batch, in_dim, out_dim;
  int act;
  const void* x; int x_ld;
  const float* mu; const float* rho; ed2_noise eps;
  ed2_noise sign_in; ed2_noise sign_out;
  const float* bias;
  void* mu_lp; int mu_ld;        /* [I][O] dtype working copies (outputs, saved) */
  void* delta_lp; int delta_ld;
  void* xs; int xs_ld;           /* saved x∘S [B][I] dtype */
  float* pert; int pert_ld;      /* workspace [B][O] fp32: ((x∘S)Delta)∘T */
  void* y; int y_ld;
  double* kl_out; float kl_scale; float prior_mean; float prior_std;
} ed2_flipout_fwd_args;
int ed2_flipout_dense_fwd(const ed2_flipout_fwd_args* a, void* stream);

typedef struct {
  int dtype;
  int batch, in_dim, out_dim;
  int act;
  const void* gy; int gy_ld;
  const void* y; int y_ld;
  const void* x; int x_ld;
  const void* xs; int xs_ld;
  const void* mu_lp; int mu_ld;
  const void* delta_lp; int delta_ld;
  const float* mu; const float* rho; ed2_noise eps;
  ed2_noise sign_in; ed2_noise sign_out;
  void* gp; int gp_ld;          /* workspace [B][O] dtype */
  void* gt; int gt_ld;          /* workspace [B][O] dtype: gp ∘ T */
  float* tmp; int tmp_ld;       /* workspace [B][I] fp32 */
  void* dx; int dx_ld;
  float* dmu; float* drho; float* dbias;
  float kl_grad_scale; float prior_mean; float prior_std;
  const double* kl_upstream;
} ed2_flipout_bwd_args;
int ed2_flipout_dense_bwd(const ed2_flipout_bwd_args* a, void* stream);

/* Fused DenseFlipout (bf16, Philox signs): ONE tcgen05 launch with two TMEM accumulators computes
 *   y = act( x·mu_lp + ((x∘S)·delta_lp) ∘ T + b )
 * with T regenerated from Philox in the epilogue — the fp32 perturbation [B][O] and the sign tensors never exist in
 * HBM (the call above writes and re-reads them).  mu_lp / delta_lp are drawn beforehand by
 * ed2_sample_normal_weights(mode 1) (any stream; typically a side stream beside the previous layer's GEMM).
 * Stacks: with next_xs != NULL the epilogue also writes y∘S_next — the NEXT Flipout layer's second operand, drawn from
 * that layer's sign_in stream — so only the first layer of a stack runs a sign pass (xs_ready == 0). */
typedef struct {
  int batch, in_dim, out_dim;
  int act;                                  /* ED2_ACT_NONE | ED2_ACT_RELU */
  const void* x; int x_ld;                  /* [B][I] bf16 */
  void* xs; int xs_ld;                      /* [B][I] bf16 x∘S: input when xs_ready, else written from sign_in */
  int xs_ready;
  ed2_noise sign_in; ed2_noise sign_out;    /* Philox streams (injected must be NULL) */
  const void* mu_lp; int mu_ld;             /* [I][O] bf16 */
  const void* delta_lp; int delta_ld;       /* [I][O] bf16 */
  const float* bias;                        /* [O] or NULL; 16-byte aligned */
  void* y; int y_ld;                        /* [B][O] bf16 */
  void* next_xs; int next_xs_ld;            /* optional [B][O] bf16: y ∘ S_next */
  ed2_noise next_sign_in;
} ed2_flipout_fused_fwd_args;
int ed2_flipout_fused_fwd(const ed2_flipout_fused_fwd_args* a, void* stream);

/* Backward of the fused call.  gp = gy∘act'(y) and gt = gp∘T are either computed here from (gy, y) (premade == 0) or were
 * already written by the CONSUMER layer's fused backward (premade == 1: gp / gt / dbias are inputs).
 *   dmu  = xᵀ gp,  dDelta = (x∘S)ᵀ gt   one dual-accumulator launch whose epilogue regenerates eps and adds the KL
 *                                        gradient -> dmu, drho (the raw fp32 products are never written)
 *   dx   = gp·mu_lpᵀ + (gt·delta_lpᵀ) ∘ S                                  one dual-accumulator launch
 * With a producer (prev_gp != NULL; this layer's x is the producer's output, ReLU when prev_relu): the dX epilogue
 * writes prev_gp = dx∘1[x > 0] (or dx), prev_gt = prev_gp∘T_prev (the producer's sign_out stream) and accumulates the
 * producer's bias gradient into prev_dbias (zeroed here), so the producer's backward runs with premade == 1. */
typedef struct {
  int batch, in_dim, out_dim;
  int act;
  int premade;
  const void* gy; int gy_ld;
  const void* y; int y_ld;
  void* gp; int gp_ld;                       /* [B][O] bf16 */
  void* gt; int gt_ld;                       /* [B][O] bf16 */
  const void* x; int x_ld;
  const void* xs; int xs_ld;
  const void* mu_lp; int mu_ld;
  const void* delta_lp; int delta_ld;
  const float* mu; const float* rho; ed2_noise eps;
  ed2_noise sign_in; ed2_noise sign_out;
  float* dmu; float* drho; float* dbias;     /* dbias: output unless premade */
  float kl_grad_scale; float prior_mean; float prior_std;
  const double* kl_upstream;
  void* dx; int dx_ld;                       /* [B][I] bf16, optional (no producer) */
  int prev_relu;
  void* prev_gp; int prev_gp_ld;             /* optional [B][I] bf16 */
  void* prev_gt; int prev_gt_ld;
  ed2_noise prev_sign_out;
  float* prev_dbias;                         /* [I] or NULL */
} ed2_flipout_fused_bwd_args;
int ed2_flipout_fused_bwd(const ed2_flipout_fused_bwd_args* a, void* stream);
/* 1 when the fused calls accept this shape (batch > 128, batch, in_dim and out_dim multiples of 8) */
int ed2_flipout_fusable(int batch, int in_dim, int out_dim);

/* Covariance of the Laplace approximation: a <- a^-1 for the symmetric positive definite a = ridge I + precision
 * (edward2/tensorflow/layers/random_feature.py:447-459 tf.linalg.inv; edward2/jax/nn/random_feature.py:393-404 Cholesky +
 * triangular solves), in place, fp32, no library call: block Gauss-Jordan with 64-wide pivot blocks (each pivot block a
 * Schur complement of an SPD matrix: no pivoting), three fp32 GEMMs per block step, 2 n^3 FLOP.  workspace: at least
 * ed2_spd_inverse_workspace_floats(n) floats; its last word is a status flag (non-zero: a pivot was not positive). */
long long ed2_spd_inverse_workspace_floats(int n);
int ed2_spd_inverse(float* a, int n, float* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Heteroscedastic heads: MCSoftmaxDenseFA (edward2/tensorflow/layers/heteroscedastic.py:508, Monte-Carlo machinery
 * :178-250, factor-analysis noise :749-785) and HeteroscedasticSNGPLayer (edward2/tensorflow/layers/hetsngp.py:25,
 * variance mixing :143-170).  fp32.
 *   latent[b,s,k] = loc[b,k] + sum_f fac[b,k,f] z[b,s,f] + diag[b,k] e[b,s,k],  z, e ~ N(0,1),  s < num_samples
 *   diag = softplus(diag_raw) + 1e-3;  with sngp_var (evaluation): diag <- sqrt(w_het diag² + w_sngp sngp_var[b])
 *   probs_mean = mean_s softmax(latent / temperature);  logits = log(clip(probs_mean, eps, 1));
 *   variance (optional) = mean_s (softmax_s - probs_mean)²
 * Noise: a normal stream; sample (b, s) owns elements [(row(b) S + s) P, +P), P = round8(F) + round8(K) (z, then e);
 * injected: [batch][S][P] fp32.  share_samples bit 0: every example reads row 0 (share_samples_across_batch); bit 1: ONE
 * diagonal draw per (example, sample), element 0 of the e block, shared by all classes — the JAX class
 * (edward2/jax/nn/heteroscedastic_lib.py:186-188: noise of shape [B, S, 1]); the TF class draws one per class.
 * bit 2: SIGMOID link (MCSigmoidDenseFA, heteroscedastic_lib.py:338-594): probs = mean_s sigmoid(latent / T), logits =
 * log(clip(probs, eps)) - log(1 - clip(probs, eps, 1 - eps)).
 * num_classes <= 64, num_factors <= 16.
 * The backward regenerates the samples: dloc, dfac, ddiag_raw from g = dL/dlogits (training form: no sngp_var). */
int ed2_mc_softmax_fa_fwd(const float* loc, const float* fac, const float* diag_raw, const float* sngp_var, float w_het,
                          float w_sngp, ed2_noise noise, int batch, int num_samples, int num_classes, int num_factors,
                          int share_samples, float temperature, float eps, float* probs_mean, float* logits,
                          float* variance, void* stream);
int ed2_mc_softmax_fa_bwd(const float* g, const float* probs_mean, const float* loc, const float* fac,
                          const float* diag_raw, ed2_noise noise, int batch, int num_samples, int num_classes,
                          int num_factors, int share_samples, float temperature, float eps, float* dloc, float* dfac,
                          float* ddiag_raw, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Bayesian LSTM cells (edward2/tensorflow/layers/recurrent.py:30 LSTMCellReparameterization): the weights are sampled once
 * per sequence (call_weights / get_initial_state, :138-161) and reused by every time step, so a step is
 *   z = x·W_lp + h_prev·U_lp + b   (two ed2_gemm calls, the second with addend = z)   [B][4H], gates i | f | c | o
 *   i, f, o = sigmoid(z_.), g = tanh(z_c);  c = f c_prev + i g;  h = o tanh(c)          (:163-184)
 * and its backward dz = [di | df | dg | do] (pre-activation), dc_prev, dbias (column sums, fp32 atomics, zeroed here),
 * followed by the four GEMMs dx = dz·Wᵀ, dh_prev = dz·Uᵀ, dW += xᵀ·dz, dU += h_prevᵀ·dz.  Cell state fp32, z / h / dz in
 * `dtype`. */
int ed2_lstm_pointwise_fwd(const void* z, int dtype, long long z_ld, const float* c_prev, int batch, int units, float* c,
                           void* h, long long h_ld, void* stream);
/* The same gates on z + z2: the input and the recurrent branch kept as two pre-activation tensors — LSTMCellFlipout
 * (recurrent.py:185-358: each branch is a Flipout product with its own sign tensors) and LSTMCellRank1 (:360-830: each
 * branch a rank-1 product with its own alpha / gamma).  dz is the gradient of both. */
int ed2_lstm_gates_fwd(const void* z, const void* z2, int dtype, long long z_ld, long long z2_ld, const float* c_prev,
                       int batch, int units, float* c, void* h, long long h_ld, void* stream);
int ed2_lstm_gates_bwd(const void* z, const void* z2, int dtype, long long z_ld, long long z2_ld, const float* c_prev,
                       const float* c, const void* dh, long long dh_ld, const float* dc, int batch, int units, void* dz,
                       long long dz_ld, float* dc_prev, void* stream);
int ed2_lstm_pointwise_bwd(const void* z, int dtype, long long z_ld, const float* c_prev, const float* c, const void* dh,
                           long long dh_ld, const float* dc, int batch, int units, void* dz, long long dz_ld,
                           float* dc_prev, float* dbias, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Conv2D stochastic-weight layers (edward2/tensorflow/layers/convolutional.py:32 Conv2DReparameterization, :167
 * Conv2DFlipout, :560 Conv2DBatchEnsemble, :2056 Conv2DRank1).  Their per-example factors are channel vectors broadcast
 * over the spatial axes (:235-249, :681-699, :1976-1989), so on the patch matrix
 *   cols[(b,ho,wo)][(i,j,c)] = x[b][ho*sh+i-pt][wo*sw+j-pl][c]
 * every variant is the dense call of the same family with rows-per-member = Ho*Wo and one "member" per example:
 *   ed2_be_dense_fwd(cols, W[kh*kw*Cin][Cout], alpha = tile(f_in[b]), gamma = f_out[b], bias[b], ensemble_size = B)
 *   ed2_reparam_dense_fwd(cols, mu / rho reshaped [kh*kw*Cin][Cout])
 * These entry points are what the dense path lacks.  NHWC, dilation 1, same = 0 'valid' | 1 TF 'SAME'. */
typedef struct {
  int batch, height, width, channels;
  int kh, kw, sh, sw;
  int same;
} ed2_conv_geometry;
int ed2_conv_out_size(const ed2_conv_geometry* g, int* out_h, int* out_w);
/* x contiguous NHWC `dtype`; cols [B*Ho*Wo][cols_ld] `dtype`, cols_ld >= kh*kw*C */
int ed2_im2col(const ed2_conv_geometry* g, const void* x, int dtype, void* cols, long long cols_ld, void* stream);
/* adjoint of ed2_im2col in gather form (no atomics): dx contiguous NHWC `dx_dtype` */
int ed2_col2im(const ed2_conv_geometry* g, const void* dcols, int dtype, long long cols_ld, void* dx, int dx_dtype,
               void* stream);
/* Factorised convolution (Conv2DBatchEnsemble :681-699, Conv2DRank1 :1976-2020, multiplicative):
 *   y[b] = act( conv(x[b] ∘ f_in[b], W) ∘ f_out[b] + bias[b] ),   f_in [B][Cin], f_out / bias [B][Cout] fp32 tables
 * forward : xa = im2col(x ∘ f_in) in ONE gather, then the tcgen05 GEMM xa·W with the per-example scale / bias / activation
 *           in its epilogue (rows-per-member = Ho*Wo); z = the un-scaled product (saved for the backward).
 * backward: dz = gy∘act'(y)∘f_out with df_out / dbias reduced per example, dW = xaᵀ·dz (K slices when the output is a
 *           handful of tiles), dxa = dz·Wᵀ, image-space gradient by col2im, then dx = g∘f_in and df_in = sum_hw g∘x. */
typedef struct {
  ed2_conv_geometry geom;
  int dtype, out_channels, act;
  const void* x;                      /* [B][H][W][Cin] dtype, contiguous */
  const void* w_lp; int w_ld;         /* [kh*kw*Cin][Cout] dtype */
  const float* f_in;                  /* [B][Cin] or NULL */
  const float* f_out;                 /* [B][Cout] */
  const float* bias;                  /* [B][Cout] or NULL */
  void* xa; int xa_ld;                /* [B*Ho*Wo][kh*kw*Cin] dtype: saved */
  void* y; int y_ld; void* z; int z_ld;  /* [B*Ho*Wo][Cout] dtype */
} ed2_conv_factor_fwd_args;
int ed2_conv_factor_fwd(const ed2_conv_factor_fwd_args* a, void* stream);
typedef struct {
  ed2_conv_geometry geom;
  int dtype, out_channels, act;
  const void* gy; int gy_ld; const void* y; int y_ld; const void* z; int z_ld;
  const void* x; const void* xa; int xa_ld; const void* w_lp; int w_ld;
  const float* f_in; const float* f_out;
  void* dz; int dz_ld;                /* workspace [B*Ho*Wo][Cout] dtype */
  void* dxa; int dxa_ld;              /* workspace [B*Ho*Wo][kh*kw*Cin] dtype (NULL: no input-side gradients) */
  void* gimg;                         /* workspace [B][H][W][Cin] dtype */
  void* dx;                           /* [B][H][W][Cin] dtype or NULL */
  float* dw;                          /* [kh*kw*Cin][Cout] fp32 */
  float* df_in; float* df_out; float* dbias;  /* [B][Cin], [B][Cout], [B][Cout] (df_in / dbias may be NULL) */
} ed2_conv_factor_bwd_args;
int ed2_conv_factor_bwd(const ed2_conv_factor_bwd_args* a, void* stream);
/* out[b][t*C + c] = src[b / rows_per_entry][c], t < reps: a per-member ([E][C], rows_per_entry = B / E) or per-example
 * ([B][C], rows_per_entry = 1) channel table laid out for the patch matrix (reps = kh*kw) or the output (reps = 1) */
int ed2_expand_table(const float* src, float* out, int batch, int channels, int rows_per_entry, int reps, void* stream);
int ed2_expand_table_bwd(const float* dout, float* dsrc, int batch, int channels, int rows_per_entry, int reps, void* stream);
/* rank-1 per-example factor table f[b][c] = clip(mu[e][c] + (softplus(rho[e][c]) + 1e-7) eps[b][c], lo, hi), e = b / (B / E)
 * (convolutional.py:1976-1989; rho == NULL: f = mu[e]) and its gradient:
 *   dmu[e][c] = sum_b df 1[unclipped],  drho[e][c] = sum_b df 1[unclipped] eps sigmoid(rho) */
int ed2_rank1_factors(const float* mu, const float* rho, ed2_noise eps, int batch, int ensemble_size, int channels,
                      float clip_lo, float clip_hi, float* f, void* stream);
int ed2_rank1_factors_bwd(const float* df, const float* mu, const float* rho, ed2_noise eps, int batch, int ensemble_size,
                          int channels, float clip_lo, float clip_hi, float* dmu, float* drho, void* stream);
/* a per-example noise TABLE out[rows][cols] (`dtype`, pitch ld) from a noise descriptor — kind 0 normal, 1 sign — with
 * the descriptor's row mapping and device step applied (the sign vectors of Conv2DFlipout, convolutional.py:235-249) */
int ed2_noise_table(int kind, ed2_noise nz, void* out, int dtype, long long ld, long long rows, long long cols, void* stream);
/* the two-input finishing pass of a Flipout-style pair of weight gradients (d_mean = xᵀgp, d_pert = (x∘S)ᵀ(gp∘T)), in place
 * or not:  dmu = d_mean + KL',  drho = (d_pert eps + KL') sigmoid(rho) */
int ed2_normal_grad_finish2(const float* d_mean, const float* d_pert, const float* mu, const float* rho, ed2_noise eps,
                            long long n, float kl_grad_scale, const double* kl_upstream, float prior_mean, float prior_std,
                            float* dmu, float* drho, int fast, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SNGP head.
 * SpectralNormalization (edward2/tensorflow/layers/normalization.py:369-391;
 * edward2/jax/nn/normalization.py:159-178).  w [I][O] fp32, u [O], v [I] (updated in place).
 *   mode 0 (TF):  v <- l2n(u Wᵀ); u <- l2n(v W); sigma = v W uᵀ; w_out = w * min(1, c/sigma)
 *   mode 1 (JAX): v <- l2n(Wᵀ-apply); u_ <- W v; u <- l2n(u_); sigma = <u,u_>; w_out = w / max(sigma/c, 1)
 * (for a Dense kernel both compute the same power iteration; they differ in the l2-normalise epsilon.)
 * w_out may alias w (TF in-place semantics).  sigma_out: device scalar.  w_out_lp (optional): normalised
 * kernel in bf16 with pitch w_lp_ld, written in the same pass.  workspace: >= (I + O + 8) floats. */
int ed2_spectral_norm_update(const float* w, int in_dim, int out_dim, float* u, float* v, int iterations,
                             float norm_multiplier, int mode, int update_uv, float* w_out, void* w_out_lp,
                             int w_lp_ld, float* sigma_out, float* workspace, void* stream);

/* LayerNormalization over the last axis (Keras defaults eps=1e-3, with gamma/beta; Flax eps=1e-6 without):
 * y = (x - mean) / sqrt(var + eps) * gamma + beta   (edward2/tensorflow/layers/random_feature.py:167-171,252;
 * edward2/jax/nn/random_feature.py:98-102).  gamma/beta may be NULL.  Or, with scale_only != 0,
 * y = x * scale (random_feature.py:254-256). */
int ed2_input_normalize(const void* x, int x_dtype, int x_ld, void* y, int y_dtype, int y_ld, long long rows,
                        int cols, const float* gamma, const float* beta, float eps, int scale_only, float scale,
                        void* stream);

/* Backward of the LayerNormalization above (the reference relies on autodiff): dx, and dgamma / dbeta column
 * sums (optional, overwritten). */
int ed2_layernorm_bwd(const void* x, int x_dtype, int x_ld, const void* dy, int dy_dtype, int dy_ld, long long rows,
                      int cols, const float* gamma, float eps, void* dx, int dx_dtype, int dx_ld, float* dgamma,
                      float* dbeta, void* stream);

/* Random Fourier features (edward2/tensorflow/layers/random_feature.py:258-266;
 * edward2/jax/nn/random_feature.py:209-214):  phi = cos(x W + b) * scale.
 * wt: the frozen projection stored TRANSPOSED, [D][in_dim] in `dtype` (K-major B operand).
 * pre (optional, [B][D], for the backward pass): pre_mode 0 — fp32, receives the phase x W (without the bias);
 * pre_mode 1 (ED2_BF16 GEMM) — bf16, receives -scale * sin(x W + b), the derivative factor of the feature map, from the
 * same range reduction that produced phi: half the bytes of the fp32 phase and no sin pass in the backward. */
int ed2_rff_fwd(int dtype, const void* x, int x_ld, const void* wt, int wt_ld, const float* bias, float scale,
                int batch, int in_dim, int num_features, void* phi, int phi_ld, int phi_dtype, void* pre,
                int pre_ld, int pre_mode, void* stream);
/* LayerNorm forward written straight into the bf16 SPLIT A-operand of the fp32-grade feature GEMM (the 2-term split of
 * ed2_split_bf16 with is_b = 0): out [rows][3 * cols] = [h | h | m], h = bf16(y), m = bf16(y - h),
 * y = (x - mean) rsqrt(var + eps) gamma + beta.  cols % 8 == 0. */
int ed2_layernorm_split(const void* x, int x_dtype, int x_ld, void* out, int out_ld, long long rows, int cols,
                        const float* gamma, const float* beta, float eps, void* stream);
/* dx = (-scale * sin(pre + b) ∘ dphi) Wᵀ  (W and b are frozen: only the input gradient exists).
 * w: [in_dim][D] in `dtype`; gpre: workspace [B][D] in `dtype`. */
int ed2_rff_bwd(int dtype, const void* dphi, int dphi_ld, int dphi_dtype, const void* pre, int pre_ld, int pre_mode /* as written by ed2_rff_fwd */,
                const float* bias, const void* w, int w_ld, float scale, int batch, int in_dim, int num_features,
                void* gpre, int gpre_ld, void* dx, int dx_ld, int dx_dtype, void* stream);

/* LaplaceRandomFeatureCovariance precision update (edward2/tensorflow/layers/random_feature.py:394-423;
 * edward2/jax/nn/random_feature.py:350-362).  phi [B][D] in `dtype`.
 *   momentum > 0:  P <- momentum * P + (1 - momentum) * phiᵀphi / batch_total
 *   else        :  P <- P + phiᵀphi
 * partial_out (optional): if non-NULL the un-blended local product phiᵀphi is written there instead and P is
 * untouched — the data-parallel caller all-reduces it over NCCL and finishes with ed2_precision_blend. */
int ed2_laplace_precision_update(int dtype, const void* phi, int phi_ld, int batch, int num_features,
                                 float momentum, long long batch_total, float* precision, float* partial_out,
                                 void* stream);
int ed2_precision_blend(float* precision, const float* summed, int num_features, float momentum,
                        long long batch_total, void* stream);

/* Predictive variance diag(ridge * phi Sigma phiᵀ) (edward2/jax/nn/random_feature.py:393-404 diagonal form;
 * edward2/tensorflow/layers/random_feature.py:482-485 takes its diagonal) fused with mean-field logits
 * (edward2/tensorflow/layers/utils.py:379-424): if logits != NULL,
 *   logits_out = logits / sqrt(1 + var * mean_field_factor)      (likelihood 0: logistic / gaussian)
 *   logits_out = logits / exp(-var * mean_field_factor / 2)      (likelihood 1: poisson)
 * sigma_lp: covariance matrix [D][D] in `dtype`.  var [B] fp32 (zeroed internally). */
int ed2_predictive_variance(int dtype, const void* phi, int phi_ld, const void* sigma_lp, int sigma_ld, int batch,
                            int num_features, float ridge, float* var, const float* logits, int num_classes,
                            float mean_field_factor, int likelihood, float* logits_out, void* stream);
int ed2_mean_field_logits(const float* logits, const float* var, int batch, int num_classes,
                          float mean_field_factor, int likelihood, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ED2B200_H_ */
