// extern "C" entry points of libed2b200.so (declared in include/ed2b200.h).  Each layer-level call composes
// the tcgen05 GEMM with the bandwidth-bound kernels on the caller's stream; nothing here synchronises.
#include "../../include/ed2b200.h"

#include <cuda_runtime.h>

#include "gemm.h"
#include "kernels.h"
#include "status.h"
#include <algorithm>
#include <cstdlib>

using namespace ed2;

namespace {

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline Noise N(const ed2_noise& n) {
  return Noise{n.injected, n.seed, n.stream, n.step, n.row0, n.rows_per_group_local, n.rows_per_group_global};
}
inline FastOpts opts_of(int additive, float lo, float hi) {
  FastOpts o;
  if (lo != 0.f || hi != 0.f) { o.clip_lo = lo; o.clip_hi = hi; }
  o.additive = additive;
  return o;
}

#define ED2_TRY(expr)            \
  do {                           \
    int _st = (expr);            \
    if (_st != ED2_OK) return _st; \
  } while (0)

FastSide fast_side(const float* mu, const float* sg, const float* sgm, const ed2_noise& n) {
  FastSide f{};
  f.mu = mu; f.sg = sg; f.sgm = sgm;
  f.seed = n.seed; f.stream = n.stream; f.step = reinterpret_cast<const unsigned long long*>(n.step);
  f.row0 = n.row0; f.rpg_local = n.rows_per_group_local; f.rpg_global = n.rows_per_group_global;
  return f;
}
inline bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// ED2_FUSED_DW=1 finishes the Reparameterization weight gradient in the epilogue of its own GEMM (gemm_flipout.cuh mode 2).
// OFF by default — measured on B200 (cfg2, profiles/r02_reparam_fused_dw_ab.txt): the step is 0.636 ms with it against
// 0.625 ms without.  The fused launch (256x128 tiles, 222 us per step) is slower than the plain 256x256 dW GEMM (160 us)
// by more than the dX GEMM gains from not sharing the SMs with the finishing pass (186 -> 157 us).  First version, with the
// row-per-thread global accesses of the epilogue (32 lines per warp instruction): 290 us, step 0.699 ms.
inline bool fused_dw_enabled() {
  static const bool v = [] { const char* e = std::getenv("ED2_FUSED_DW"); return e != nullptr && std::atoi(e) != 0; }();
  return v;
}

EpiParams epi_default() {
  EpiParams e{};
  e.alpha = 1.f;
  e.beta = 1.f;
  e.act_scale = 1.f;
  return e;
}

int check_dt(int dt) {
  if (dt != kF32 && dt != kBF16) return set_error(ED2_ERR_BAD_DTYPE, "dtype must be ED2_F32 or ED2_BF16, got %d", dt);
  return ED2_OK;
}

// y[B][O] = epilogue( x[B][I] · w[I][O] )        (A K-major, B MN-major)
int gemm_xw(int dt, const void* x, int x_ld, const void* w, int w_ld, int B, int I, int O, const EpiParams& ep,
            cudaStream_t s) {
  GemmDesc g{};
  g.A = x; g.lda = x_ld; g.major_a = kMajorK;
  g.B = w; g.ldb = w_ld; g.major_b = kMajorMN;
  g.M = B; g.N = O; g.K = I;
  g.ep = ep;
  return gemm(dt, g, s);
}
// dx[B][I] = epilogue( g[B][O] · wᵀ )            (A K-major; B K-major: w[I][O] is [N=I][K=O])
int gemm_gwt(int dt, const void* gz, int g_ld, const void* w, int w_ld, int B, int I, int O, const EpiParams& ep,
             cudaStream_t s) {
  GemmDesc g{};
  g.A = gz; g.lda = g_ld; g.major_a = kMajorK;
  g.B = w; g.ldb = w_ld; g.major_b = kMajorK;
  g.M = B; g.N = I; g.K = O;
  g.ep = ep;
  return gemm(dt, g, s);
}
// dw[I][O] = epilogue( xᵀ[I][B] · g[B][O] )      (A MN-major, B MN-major)
int gemm_xtg(int dt, const void* x, int x_ld, const void* gz, int g_ld, int B, int I, int O, const EpiParams& ep,
             cudaStream_t s) {
  GemmDesc g{};
  g.A = x; g.lda = x_ld; g.major_a = kMajorMN;
  g.B = gz; g.ldb = g_ld; g.major_b = kMajorMN;
  g.M = I; g.N = O; g.K = B;
  g.ep = ep;
  // narrow weight gradient over a long batch axis (conv layers on the patch matrix: [9C x Cout] with K = B*Ho*Wo): a
  // handful of output tiles would leave most SM pairs idle — K slices accumulate with fp32 atomics into a zeroed output
  const long long tiles = ((I + 255) / 256) * static_cast<long long>((O + 255) / 256);
  if (dt == kBF16 && ep.out_dt == kF32 && ep.out != nullptr && tiles * 4 <= 74 && B >= 4096 && ep.addend == nullptr &&
      ep.col_sum == nullptr && ep.raw == nullptr && ep.act == kActNone && ep.out_ld == O) {
    cudaMemsetAsync(ep.out, 0, sizeof(float) * static_cast<size_t>(I) * O, s);
    return gemm_bf16_accumulate(g, s);
  }
  return gemm(dt, g, s);
}

// Fork-join onto a per-thread side stream (capturable: the side stream joins the caller's capture through events).
class ForkJoin {
 public:
  ForkJoin(cudaStream_t main, bool enable) : main_(main), on_(enable) {
    if (!on_) return;
    Ctx& c = ctx();
    if (c.side == nullptr) {
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);  // side work is low priority: GEMM CTAs are placed first
      if (cudaStreamCreateWithPriority(&c.side, cudaStreamNonBlocking, lo) != cudaSuccess ||
          cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&c.joined, cudaEventDisableTiming) != cudaSuccess) {
        c.side = nullptr;
        on_ = false;
        return;
      }
    }
    if (cudaEventRecord(c.fork, main_) != cudaSuccess || cudaStreamWaitEvent(c.side, c.fork, 0) != cudaSuccess) on_ = false;
  }
  cudaStream_t side() { return on_ ? ctx().side : main_; }
  void join() {
    if (!on_) return;
    Ctx& c = ctx();
    cudaEventRecord(c.joined, c.side);
    cudaStreamWaitEvent(main_, c.joined, 0);
    on_ = false;
  }
  ~ForkJoin() { join(); }

 private:
  struct Ctx {
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, joined = nullptr;
  };
  static Ctx& ctx() {
    thread_local Ctx c;
    return c;
  }
  cudaStream_t main_;
  bool on_;
};

}  // namespace

extern "C" {

const char* ed2_version(void) { return "ed2b200 0.1.0 (sm_100a)"; }
const char* ed2_last_error(void) { return last_error(); }
long long ed2_launch_count(void) { return launch_count(); }

int ed2_check_device(int dev) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return set_error(ED2_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (p.major != 10) return set_error(ED2_ERR_UNSUPPORTED_ARCH, "device %d is sm_%d%d; this library is sm_100a only", dev, p.major, p.minor);
  return ED2_OK;
}

int ed2_gemm(const ed2_gemm_desc* d, void* stream) {
  if (d == nullptr) return set_error(ED2_ERR_BAD_ARG, "null desc");
  ED2_TRY(check_dt(d->dtype));
  GemmDesc g{};
  g.A = d->a; g.lda = d->lda; g.major_a = d->major_a;
  g.B = d->b; g.ldb = d->ldb; g.major_b = d->major_b;
  g.M = d->m; g.N = d->n; g.K = d->k;
  EpiParams& e = g.ep;
  e = epi_default();
  e.alpha = d->alpha; e.beta = d->beta; e.act_scale = d->act_scale; e.act = d->act; e.group_rows = d->group_rows;
  e.elem_scale = d->elem_scale; e.elem_scale_ld = d->elem_scale_ld; e.elem_scale_dt = d->elem_scale_dtype;
  e.col_scale = d->col_scale; e.col_scale_ld = d->col_scale_ld;
  e.addend = d->addend; e.addend_ld = d->addend_ld; e.addend_dt = d->addend_dtype;
  e.col_bias = d->col_bias; e.col_bias_ld = d->col_bias_ld;
  e.out = d->out; e.out_ld = d->out_ld; e.out_dt = d->out_dtype;
  e.raw = d->raw; e.raw_ld = d->raw_ld; e.raw_dt = d->raw_dtype;
  e.row_sum = d->row_sum;
  e.mask_src = d->mask_src; e.mask_ld = d->mask_ld; e.mask_dt = d->mask_dtype;
  e.col_sum = d->col_sum; e.col_sum_ld = d->col_sum_ld;
  g.block_n = d->block_n; g.cta_group = d->cta_group; g.max_ctas = d->max_ctas; g.no_tma_store = d->no_tma_store;
  g.ep.lower_only = d->lower_only;
  if (d->accumulate) {
    if (d->dtype != kBF16 || d->out_dtype != kF32) return set_error(ED2_ERR_BAD_ARG, "gemm: accumulate needs ED2_BF16 operands and an fp32 out");
    g.ep.alpha = d->alpha;
    return gemm_bf16_accumulate(g, S(stream));
  }
  return gemm(d->dtype, g, S(stream));
}

int ed2_philox_normal(float* out, long long n, uint64_t seed, uint64_t stream, void* s) { return k_philox_fill(0, out, n, seed, stream, S(s)); }
int ed2_philox_normal_fast(float* out, long long n, uint64_t seed, uint64_t stream, void* s) { return k_philox_fill(3, out, n, seed, stream, S(s)); }
int ed2_philox_uniform(float* out, long long n, uint64_t seed, uint64_t stream, void* s) { return k_philox_fill(1, out, n, seed, stream, S(s)); }
int ed2_philox_sign(float* out, long long n, uint64_t seed, uint64_t stream, void* s) { return k_philox_fill(2, out, n, seed, stream, S(s)); }
int ed2_philox_raw(uint32_t* out, long long nb, uint64_t seed, uint64_t stream, void* s) { return k_philox_raw(out, nb, seed, stream, S(s)); }

int ed2_cast(const void* src, int sdt, int sld, void* dst, int ddt, int dld, long long rows, long long cols, void* s) {
  ED2_TRY(check_dt(sdt));
  ED2_TRY(check_dt(ddt));
  return k_cast2d(src, sdt, sld, dst, ddt, dld, rows, cols, S(s));
}

int ed2_split_bf16(const void* src, int src_dtype, int src_ld, void* dst, int dst_ld, long long rows, int cols, int is_b,
                   int nsplit, void* s) {
  ED2_TRY(check_dt(src_dtype));
  if (nsplit != 2 && nsplit != 3) return set_error(ED2_ERR_BAD_ARG, "split_bf16: nsplit must be 2 or 3");
  return k_split_bf16(src, src_dtype, src_ld, dst, dst_ld, rows, cols, is_b, nsplit, S(s));
}

int ed2_sample_normal_weights(const float* mu, const float* rho, ed2_noise eps, long long rows, long long cols, int mode,
                              void* w, int w_dtype, int w_ld, void* mean_out, int mean_ld, double* kl_out,
                              float kl_scale, float prior_mean, float prior_std, void* s) {
  ED2_TRY(check_dt(w_dtype));
  return k_sample_weights(mu, rho, N(eps), rows, cols, mode, w, w_dtype, w_ld, mean_out, mean_ld, kl_out, kl_scale,
                          Prior{prior_mean, prior_std}, S(s));
}

int ed2_normal_kl_fwd(const float* mu, const float* rho, long long n, float pm, float ps, float scale, double* out, void* s) {
  return k_normal_kl_fwd(mu, rho, n, Prior{pm, ps}, scale, out, S(s));
}
int ed2_normal_kl_bwd(const float* mu, const float* rho, long long n, float pm, float ps, float scale,
                      const float* upstream, float* dmu, float* drho, void* s) {
  return k_normal_kl_bwd(mu, rho, n, Prior{pm, ps}, scale, upstream, dmu, drho, S(s));
}

int ed2_softmax_xent(const void* logits, int dtype, int ld, const long long* labels, long long rows, int cols,
                     float loss_scale, float grad_scale, double* loss, void* dlogits, int d_dtype, int d_ld,
                     const double* const* extra_terms, int n_extra, void* s) {
  ED2_TRY(check_dt(dtype));
  if (n_extra < 0 || n_extra > 8) return set_error(ED2_ERR_BAD_ARG, "softmax_xent: at most 8 extra loss terms");
  XentExtras ex{};
  ex.n = n_extra;
  for (int i = 0; i < n_extra; ++i) ex.terms[i] = extra_terms[i];
  return k_softmax_xent(logits, dtype, ld, labels, rows, cols, loss_scale, grad_scale, loss, dlogits, d_dtype, d_ld, ex, S(s));
}

int ed2_act_bwd(int dtype, const void* gy, int gy_ld, const void* y, int y_ld, long long rows, int cols, int act,
                void* gp, int gp_ld, float* dbias, void* s) {
  ED2_TRY(check_dt(dtype));
  BwdOutArgs o{};
  o.dt = dtype; o.rows = rows; o.cols = cols; o.rows_per_member = static_cast<int>(rows); o.act = act; o.fast = 0;
  o.gy = gy; o.gy_ld = gy_ld; o.y = y; o.y_ld = y_ld;
  o.dz = gp; o.dz_ld = gp_ld;
  o.dbias = dbias;
  return k_bwd_out(o, S(s));
}

// ------------------------------------------------------------------------------------------ BatchEnsemble
int ed2_be_dense_fwd(const ed2_be_fwd_args* a, void* stream) {
  ED2_TRY(check_dt(a->dtype));
  if (a->ensemble_size <= 0 || a->batch % a->ensemble_size != 0)
    return set_error(ED2_ERR_BAD_SHAPE, "batch %d is not divisible by ensemble_size %d", a->batch, a->ensemble_size);
  cudaStream_t s = S(stream);
  const int rpm = a->batch / a->ensemble_size;
  ED2_TRY(k_scale_rows(a->x, a->dtype, a->x_ld, a->batch, a->in_dim, rpm, 0, a->alpha, nullptr, Noise{}, a->xa,
                       a->xa_ld, nullptr, 0, s));
  EpiParams e = epi_default();
  e.group_rows = rpm;
  e.col_scale = a->gamma; e.col_scale_ld = a->out_dim;
  e.col_bias = a->bias; e.col_bias_ld = a->out_dim;
  e.act = a->act;
  e.out = a->y; e.out_ld = a->y_ld; e.out_dt = a->dtype;
  e.raw = a->z; e.raw_ld = a->z_ld; e.raw_dt = a->dtype;
  return gemm_xw(a->dtype, a->xa, a->xa_ld, a->w_lp, a->w_ld, a->batch, a->in_dim, a->out_dim, e, s);
}

int ed2_be_dense_bwd(const ed2_be_bwd_args* a, void* stream) {
  ED2_TRY(check_dt(a->dtype));
  cudaStream_t s = S(stream);
  const int rpm = a->batch / a->ensemble_size;
  BwdOutArgs o{};
  o.dt = a->dtype; o.rows = a->batch; o.cols = a->out_dim; o.rows_per_member = rpm; o.act = a->act; o.fast = 1;
  o.gy = a->gy; o.gy_ld = a->gy_ld; o.y = a->y; o.y_ld = a->y_ld; o.z = a->z; o.z_ld = a->z_ld;
  o.gamma = a->gamma;
  o.dz = a->dz; o.dz_ld = a->dz_ld;
  o.d_fast_mu = a->dgamma; o.dbias = a->dbias;
  ED2_TRY(k_bwd_out(o, s));
  EpiParams e = epi_default();
  e.out = a->dw; e.out_ld = a->out_dim; e.out_dt = kF32;
  ED2_TRY(gemm_xtg(a->dtype, a->xa, a->xa_ld, a->dz, a->dz_ld, a->batch, a->in_dim, a->out_dim, e, s));
  EpiParams e2 = epi_default();
  e2.out = a->dxa; e2.out_ld = a->dxa_ld; e2.out_dt = a->dtype;
  ED2_TRY(gemm_gwt(a->dtype, a->dz, a->dz_ld, a->w_lp, a->w_ld, a->batch, a->in_dim, a->out_dim, e2, s));
  BwdInArgs i{};
  i.dt = a->dtype; i.rows = a->batch; i.cols = a->in_dim; i.rows_per_member = rpm; i.fast = 1;
  i.dxf = a->dxa; i.dxf_ld = a->dxa_ld; i.x = a->x; i.x_ld = a->x_ld;
  i.fw_mu = a->alpha;
  i.dx = a->dx; i.dx_ld = a->dx_ld;
  i.d_fast_mu = a->dalpha;
  return k_bwd_in(i, s);
}

// ------------------------------------------------------------------------------------------ Rank-1
int ed2_rank1_dense_fwd(const ed2_rank1_fwd_args* a, void* stream) {
  ED2_TRY(check_dt(a->dtype));
  if (a->ensemble_size <= 0 || a->batch % a->ensemble_size != 0)
    return set_error(ED2_ERR_BAD_SHAPE, "batch %d is not divisible by ensemble_size %d", a->batch, a->ensemble_size);
  if (a->rho_b != nullptr && (a->bias == nullptr || a->bias_buf == nullptr))
    return set_error(ED2_ERR_BAD_ARG, "rank1_fwd: a sampled bias needs its mean (bias) and the bias_buf workspace");
  cudaStream_t s = S(stream);
  const int rpm = a->batch / a->ensemble_size;
  const FastOpts fo = opts_of(a->additive, a->clip_lo, a->clip_hi);
  if (!a->xr_ready)
    ED2_TRY(k_scale_rows(a->x, a->dtype, a->x_ld, a->batch, a->in_dim, rpm, 1, a->mu_r, a->rho_r, N(a->eps_r), a->xr,
                         a->xr_ld, nullptr, 0, s, fo));
  const bool fused = a->fast_ws != nullptr && a->s_buf == nullptr && a->dtype == kBF16 && !a->additive && a->rho_b == nullptr &&
                     a->eps_s.injected == nullptr && (a->next_mu_r == nullptr || a->next_eps_r.injected == nullptr) &&
                     gemm_rank1_fusable(a->batch, a->out_dim, rpm) && al16p(a->mu_s) &&
                     (a->bias == nullptr || al16p(a->bias)) && (a->next_mu_r == nullptr || al16p(a->next_mu_r));
  if (fused) {
    const long long slot = static_cast<long long>(a->ensemble_size) * std::max(a->in_dim, a->out_dim);
    const long long n = static_cast<long long>(a->ensemble_size) * a->out_dim;
    float* ws = a->fast_ws;
    if (a->rho_s != nullptr) ED2_TRY(k_fast_prep(a->rho_s, n, ws, ws + slot, s));
    if (a->next_mu_r != nullptr && a->next_rho_r != nullptr) ED2_TRY(k_fast_prep(a->next_rho_r, n, ws + 2 * slot, ws + 3 * slot, s));
    Rank1Epi r{};
    r.s = fast_side(a->mu_s, a->rho_s ? ws : nullptr, ws + slot, a->eps_s);
    if (a->next_mu_r != nullptr) {
      r.r = fast_side(a->next_mu_r, a->next_rho_r ? ws + 2 * slot : nullptr, ws + 3 * slot, a->next_eps_r);
      r.out2 = a->next_xr; r.out2_ld = a->next_xr_ld;
    }
    r.clip_lo = fo.clip_lo; r.clip_hi = fo.clip_hi;
    EpiParams e = epi_default();
    e.group_rows = rpm;
    e.col_bias = a->bias; e.col_bias_ld = a->out_dim;
    e.act = a->act;
    e.out = a->y; e.out_ld = a->y_ld; e.out_dt = a->dtype;
    e.raw = a->z; e.raw_ld = a->z_ld; e.raw_dt = a->dtype;
    GemmDesc g{};
    g.A = a->xr; g.lda = a->xr_ld; g.major_a = kMajorK;
    g.B = a->w_lp; g.ldb = a->w_ld; g.major_b = kMajorMN;
    g.M = a->batch; g.N = a->out_dim; g.K = a->in_dim;
    g.ep = e; g.epi_mode = kEpiRank1Fwd; g.r1 = &r;
    if (a->gemm_ready_event != nullptr) {
      cudaError_t ce = cudaEventRecord(reinterpret_cast<cudaEvent_t>(a->gemm_ready_event), s);
      if (ce != cudaSuccess) return set_error(ED2_ERR_CUDA, "cudaEventRecord(gemm_ready_event): %s", cudaGetErrorString(ce));
    }
    return gemm_bf16(g, s);
  }
  if (a->next_mu_r != nullptr)
    return set_error(ED2_ERR_BAD_ARG, "rank1_fwd: next_* needs the fused path (fast_ws, s_buf == NULL, ed2_rank1_fusable)");
  if (a->s_buf == nullptr)
    return set_error(ED2_ERR_BAD_ARG, "rank1_fwd: s_buf is required on the unfused path (this call does not qualify for "
                                      "the fused one: ED2_BF16, Philox noise, multiplicative, deterministic bias, "
                                      "ed2_rank1_fusable(batch, out_dim, E))");
  FastOpts fs = fo;
  fs.additive = 0;  // the factor itself is written; the epilogue multiplies or adds it
  ED2_TRY(k_scale_rows(nullptr, a->dtype, 0, a->batch, a->out_dim, rpm, 1, a->mu_s, a->rho_s, N(a->eps_s), nullptr, 0,
                       a->s_buf, a->s_ld, s, fs));
  EpiParams e = epi_default();
  e.group_rows = rpm;
  if (a->additive) {
    e.addend = a->s_buf; e.addend_ld = a->s_ld; e.addend_dt = a->dtype;
  } else {
    e.elem_scale = a->s_buf; e.elem_scale_ld = a->s_ld; e.elem_scale_dt = a->dtype;
  }
  if (a->rho_b != nullptr) {
    // bias[b,:] = mean[e] + sigma_b[e] * eps_b[b,:], unclipped (dense.py:1131-1139), materialised once per call
    FastOpts fb;
    fb.clip_lo = -3.0e38f; fb.clip_hi = 3.0e38f;
    if (a->additive) {
      // two per-example addends (s and the bias sample): bias_buf = s_buf + bias sample, one addend slot in the epilogue
      fb.additive = 1;
      ED2_TRY(k_scale_rows(a->s_buf, a->dtype, a->s_ld, a->batch, a->out_dim, rpm, 1, a->bias, a->rho_b, N(a->eps_b),
                           a->bias_buf, a->bias_ld, nullptr, 0, s, fb));
    } else {
      ED2_TRY(k_scale_rows(nullptr, a->dtype, 0, a->batch, a->out_dim, rpm, 1, a->bias, a->rho_b, N(a->eps_b), nullptr, 0,
                           a->bias_buf, a->bias_ld, s, fb));
    }
    e.addend = a->bias_buf; e.addend_ld = a->bias_ld; e.addend_dt = a->dtype;
  } else {
    e.col_bias = a->bias; e.col_bias_ld = a->out_dim;
  }
  e.act = a->act;
  e.out = a->y; e.out_ld = a->y_ld; e.out_dt = a->dtype;
  e.raw = a->z; e.raw_ld = a->z_ld; e.raw_dt = a->dtype;
  return gemm_xw(a->dtype, a->xr, a->xr_ld, a->w_lp, a->w_ld, a->batch, a->in_dim, a->out_dim, e, s);
}

int ed2_rank1_dense_bwd(const ed2_rank1_bwd_args* a, void* stream) {
  ED2_TRY(check_dt(a->dtype));
  cudaStream_t s = S(stream);
  const int rpm = a->batch / a->ensemble_size;
  BwdOutArgs o{};
  o.dt = a->dtype; o.rows = a->batch; o.cols = a->out_dim; o.rows_per_member = rpm; o.act = a->act; o.fast = 2;
  o.gy = a->gy; o.gy_ld = a->gy_ld; o.y = a->y; o.y_ld = a->y_ld; o.z = a->z; o.z_ld = a->z_ld;
  o.s_buf = a->s_buf; o.s_ld = a->s_ld;
  o.mu_s = a->mu_s; o.rho_s = a->rho_s; o.eps_s = N(a->eps_s);
  o.opts = opts_of(a->additive, a->clip_lo, a->clip_hi);
  o.rho_b = a->rho_b; o.eps_b = N(a->eps_b); o.d_bias_rho = a->rho_b ? a->drho_b : nullptr;
  o.dz = a->dz; o.dz_ld = a->dz_ld;
  o.d_fast_mu = a->dmu_s; o.d_fast_rho = a->rho_s ? a->drho_s : nullptr; o.dbias = a->dbias;
  if (!a->dz_ready) ED2_TRY(k_bwd_out(o, s));
  EpiParams e = epi_default();
  if (a->dw_lp != nullptr) {
    if (a->dtype != kBF16) return set_error(ED2_ERR_BAD_ARG, "rank1_bwd: dw_lp needs ED2_BF16");
    e.out = a->dw_lp; e.out_ld = a->out_dim; e.out_dt = kBF16;
  } else {
    e.out = a->dw; e.out_ld = a->out_dim; e.out_dt = kF32;
  }
  ED2_TRY(gemm_xtg(a->dtype, a->xr, a->xr_ld, a->dz, a->dz_ld, a->batch, a->in_dim, a->out_dim, e, s));
  if (a->dw_done_event != nullptr) {
    cudaError_t ce = cudaEventRecord(reinterpret_cast<cudaEvent_t>(a->dw_done_event), s);
    if (ce != cudaSuccess) return set_error(ED2_ERR_CUDA, "rank1_bwd: cudaEventRecord: %s", cudaGetErrorString(ce));
  }
  const bool has_prev = a->prev_mu_s != nullptr;
  const bool fused = a->fast_ws != nullptr && a->dtype == kBF16 && !a->additive && a->eps_r.injected == nullptr &&
                     (!has_prev || a->prev_eps_s.injected == nullptr) && gemm_rank1_fusable(a->batch, a->in_dim, rpm) &&
                     al16p(a->mu_r) && al16p(a->x) && a->x_ld % 8 == 0 &&
                     (!has_prev || (al16p(a->prev_mu_s) && al16p(a->prev_z) && a->prev_z_ld % 8 == 0));
  if (fused) {
    const long long slot = static_cast<long long>(a->ensemble_size) * std::max(a->in_dim, a->out_dim);
    const long long n = static_cast<long long>(a->ensemble_size) * a->in_dim;
    const size_t bytes = sizeof(float) * static_cast<size_t>(n);
    float* ws = a->fast_ws;
    // one launch per side: sigma / sigmoid tables + zeroing of the column-sum accumulators the epilogue adds into
    ED2_TRY(k_fast_prep(a->rho_r, n, ws, ws + slot, s, a->dmu_r, a->rho_r != nullptr ? a->drho_r : nullptr));
    (void)bytes;
    Rank1Epi r{};
    r.r = fast_side(a->mu_r, a->rho_r ? ws : nullptr, ws + slot, a->eps_r);
    r.clip_lo = o.opts.clip_lo; r.clip_hi = o.opts.clip_hi;
    r.x = a->x; r.x_ld = a->x_ld;
    r.dmu_r = a->dmu_r; r.drho_r = a->rho_r ? a->drho_r : nullptr;
    EpiParams e2 = epi_default();
    e2.group_rows = rpm;
    if (has_prev) {
      if (a->prev_z == nullptr || a->prev_dz == nullptr || a->prev_dmu_s == nullptr)
        return set_error(ED2_ERR_BAD_ARG, "rank1_bwd: prev_z, prev_dz and prev_dmu_s are required with prev_mu_s");
      ED2_TRY(k_fast_prep(a->prev_rho_s, n, ws + 2 * slot, ws + 3 * slot, s, a->prev_dmu_s,
                          a->prev_rho_s != nullptr ? a->prev_drho_s : nullptr, a->prev_dbias));
      r.s = fast_side(a->prev_mu_s, a->prev_rho_s ? ws + 2 * slot : nullptr, ws + 3 * slot, a->prev_eps_s);
      r.z_prev = a->prev_z; r.z_prev_ld = a->prev_z_ld; r.prev_act = a->prev_act;
      r.dmu_s = a->prev_dmu_s; r.drho_s = a->prev_rho_s ? a->prev_drho_s : nullptr; r.dbias = a->prev_dbias;
      e2.out = a->prev_dz; e2.out_ld = a->prev_dz_ld; e2.out_dt = kBF16;
      r.out2 = a->dx; r.out2_ld = a->dx_ld;
    } else {
      e2.out = a->dx; e2.out_ld = a->dx_ld; e2.out_dt = kBF16;
    }
    GemmDesc g{};
    g.A = a->dz; g.lda = a->dz_ld; g.major_a = kMajorK;
    g.B = a->w_lp; g.ldb = a->w_ld; g.major_b = kMajorK;
    g.M = a->batch; g.N = a->in_dim; g.K = a->out_dim;
    g.ep = e2; g.epi_mode = kEpiRank1Bwd; g.r1 = &r;
    return gemm_bf16(g, s);
  }
  if (has_prev)
    return set_error(ED2_ERR_BAD_ARG, "rank1_bwd: prev_* needs the fused path (fast_ws and ed2_rank1_fusable(batch, in_dim, E))");
  if (a->dxr == nullptr) return set_error(ED2_ERR_BAD_ARG, "rank1_bwd: dxr workspace is required on the unfused path");
  EpiParams e2 = epi_default();
  e2.out = a->dxr; e2.out_ld = a->dxr_ld; e2.out_dt = a->dtype;
  ED2_TRY(gemm_gwt(a->dtype, a->dz, a->dz_ld, a->w_lp, a->w_ld, a->batch, a->in_dim, a->out_dim, e2, s));
  BwdInArgs i{};
  i.dt = a->dtype; i.rows = a->batch; i.cols = a->in_dim; i.rows_per_member = rpm; i.fast = 2;
  i.dxf = a->dxr; i.dxf_ld = a->dxr_ld; i.x = a->x; i.x_ld = a->x_ld;
  i.fw_mu = a->mu_r; i.fw_rho = a->rho_r; i.eps = N(a->eps_r);
  i.opts = opts_of(a->additive, a->clip_lo, a->clip_hi);
  i.dx = a->dx; i.dx_ld = a->dx_ld;
  i.d_fast_mu = a->dmu_r; i.d_fast_rho = a->rho_r ? a->drho_r : nullptr;
  return k_bwd_in(i, s);
}

int ed2_rank1_fusable(int batch, int cols, int ensemble_size) {
  if (ensemble_size <= 0 || batch % ensemble_size != 0) return 0;
  return gemm_rank1_fusable(batch, cols, batch / ensemble_size) ? 1 : 0;
}

long long ed2_rank1_workspace_floats(int ensemble_size, int in_dim, int out_dim) {
  return 4LL * ensemble_size * std::max(in_dim, out_dim);
}

// ------------------------------------------------------------------------------------------ Reparameterization
int ed2_reparam_dense_fwd(const ed2_reparam_fwd_args* a, void* stream) {
  ED2_TRY(check_dt(a->dtype));
  cudaStream_t s = S(stream);
  const Prior prior{a->prior_mean, a->prior_std};
  // bf16: the sampling pass is on the critical path (the GEMM needs W), the KL is not (only the loss needs it):
  // sample with hardware-approximate math and no KL, enqueue the GEMM, then run the accurate KL reduction on the
  // side stream beside the GEMM.  fp32 keeps the single precise pass with the KL fused in.
  const bool split_kl = a->dtype == kBF16 && a->kl_out != nullptr && !fused_kl_enabled();
  ED2_TRY(k_sample_weights(a->mu, a->rho, N(a->eps), a->in_dim, a->out_dim, 0, a->w_lp, a->dtype, a->w_ld, nullptr, 0,
                           split_kl ? nullptr : a->kl_out, a->kl_scale, prior, s));
  ForkJoin fj(s, split_kl);  // fork BEFORE the GEMM so the KL pass depends only on what precedes it
  EpiParams e = epi_default();
  e.col_bias = a->bias; e.col_bias_ld = a->out_dim;
  e.act = a->act;
  e.out = a->y; e.out_ld = a->y_ld; e.out_dt = a->dtype;
  int st = gemm_xw(a->dtype, a->x, a->x_ld, a->w_lp, a->w_ld, a->batch, a->in_dim, a->out_dim, e, s);
  if (split_kl && st == ED2_OK)
    st = k_normal_kl_fwd(a->mu, a->rho, (long long)a->in_dim * a->out_dim, prior, a->kl_scale, a->kl_out, fj.side());
  fj.join();
  return st;
}

int ed2_reparam_dense_bwd(const ed2_reparam_bwd_args* a, void* stream) {
  ED2_TRY(check_dt(a->dtype));
  cudaStream_t s = S(stream);
  const void* gp = a->gp;
  int gp_ld = a->gp_ld;
  if (a->gy_premasked) {
    gp = a->gy;  // the producer's dX epilogue already applied act'(y) and accumulated dbias
    gp_ld = a->gy_ld;
  }
  // The column-sum target of the dX epilogue is cleared up front (phases 0 / 1 / 3; phase 2 continues one of them):
  // a memset between the dW and dX GEMMs would let the forked finishing pass grab every SM before the dX GEMM.
  const bool fuse_prev = a->x_is_relu && a->prev_dbias != nullptr && a->dx != nullptr;
  if (fuse_prev && a->phase != 2) cudaMemsetAsync(a->prev_dbias, 0, sizeof(float) * (size_t)a->in_dim, s);
  if (a->phase == 3) {
    if (a->dmu_lp == nullptr) return set_error(ED2_ERR_BAD_ARG, "reparam_bwd phase 3 needs the bf16 bucket dmu_lp");
    if (!a->gy_premasked) {
      BwdOutArgs o{};
      o.dt = a->dtype; o.rows = a->batch; o.cols = a->out_dim; o.rows_per_member = a->batch; o.act = a->act; o.fast = 0;
      o.gy = a->gy; o.gy_ld = a->gy_ld; o.y = a->y; o.y_ld = a->y_ld;
      o.dz = a->gp; o.dz_ld = a->gp_ld;
      o.dbias = a->dbias;
      ED2_TRY(k_bwd_out(o, s));
    }
    EpiParams e = epi_default();
    e.out = a->dmu_lp; e.out_ld = a->out_dim; e.out_dt = kBF16;
    return gemm_xtg(a->dtype, a->x, a->x_ld, gp, gp_ld, a->batch, a->in_dim, a->out_dim, e, s);
  }
  // Opt-in (ED2_FUSED_DW=1, see fused_dw_enabled): the weight gradient FINISHED in the epilogue of its own GEMM
  // (gemm_flipout.cuh mode 2: dmu = acc + KL', drho = (acc eps + KL') sigmoid(rho), eps regenerated).
  const bool fused_dw = a->phase == 0 && a->dtype == kBF16 && a->eps.injected == nullptr && a->dmu_lp == nullptr &&
                        a->batch > 128 && gemm_flipout_fusable(a->in_dim, a->out_dim, a->batch) && al16p(a->mu) &&
                        al16p(a->rho) && al16p(a->dmu) && al16p(a->drho) && fused_dw_enabled();
  if (a->phase != 2) {
    if (!a->gy_premasked) {
      BwdOutArgs o{};
      o.dt = a->dtype; o.rows = a->batch; o.cols = a->out_dim; o.rows_per_member = a->batch; o.act = a->act; o.fast = 0;
      o.gy = a->gy; o.gy_ld = a->gy_ld; o.y = a->y; o.y_ld = a->y_ld;
      o.dz = a->gp; o.dz_ld = a->gp_ld;
      o.dbias = a->dbias;
      ED2_TRY(k_bwd_out(o, s));
    }
    if (fused_dw) {
      FlipoutGemm w{};
      w.mode = 2; w.major_a = kMajorMN; w.major_b = kMajorMN;
      w.A0 = a->x; w.lda0 = a->x_ld; w.A1 = a->x; w.lda1 = a->x_ld;
      w.B0 = gp; w.ldb0 = gp_ld; w.B1 = gp; w.ldb1 = gp_ld;
      w.M = a->in_dim; w.N = a->out_dim; w.K = a->batch;
      w.epi.sa.seed = a->eps.seed; w.epi.sa.stream = a->eps.stream;
      w.epi.sa.step = reinterpret_cast<const unsigned long long*>(a->eps.step); w.epi.sa.on = 1;
      w.epi.mu = a->mu; w.epi.rho = a->rho; w.epi.dmu = a->dmu; w.epi.drho = a->drho;
      w.epi.klg = a->kl_grad_scale; w.epi.kl_upstream = a->kl_upstream;
      w.epi.prior_mean = a->prior_mean; w.epi.prior_std = a->prior_std;
      ED2_TRY(gemm_flipout(w, s));
    } else {
      // dW = xᵀ gp lands in dmu; the finishing pass turns it into (dmu, drho) in place (+ bf16 bucket copies).
      EpiParams e = epi_default();
      e.out = a->dmu; e.out_ld = a->out_dim; e.out_dt = kF32;
      ED2_TRY(gemm_xtg(a->dtype, a->x, a->x_ld, gp, gp_ld, a->batch, a->in_dim, a->out_dim, e, s));
    }
  }
  // Whole-backward call: the finishing pass (CUDA cores) is forked onto the side stream and the dX GEMM (tensor
  // cores) is enqueued first.  Split call (phase 1): the finishing pass stays on the caller's stream because the
  // caller launches the gradient all-reduce right after it.
  ForkJoin fj(s, a->phase == 0 && a->dx != nullptr && !fused_dw);
  int st = ED2_OK;
  if (a->phase != 1 && a->dx != nullptr) {
    EpiParams e2 = epi_default();
    e2.out = a->dx; e2.out_ld = a->dx_ld; e2.out_dt = a->dtype;
    if (fuse_prev) {
      e2.mask_src = a->x; e2.mask_ld = a->x_ld; e2.mask_dt = a->dtype;
      e2.col_sum = a->prev_dbias; e2.col_sum_ld = a->in_dim;
    }
    st = gemm_gwt(a->dtype, gp, gp_ld, a->w_lp, a->w_ld, a->batch, a->in_dim, a->out_dim, e2, s);
  }
  if (a->phase != 2 && st == ED2_OK && !fused_dw)
    st = k_normal_grad_finish(a->dmu, a->dmu, a->mu, a->rho, N(a->eps), (long long)a->in_dim * a->out_dim,
                              a->kl_grad_scale, a->kl_upstream, Prior{a->prior_mean, a->prior_std}, a->dmu, a->drho,
                              a->dtype == kBF16, fj.side(), a->dmu_lp, a->drho_lp);
  fj.join();
  return st;
}

int ed2_normal_grad_finish(const void* dw, int dw_dtype, const float* mu, const float* rho, ed2_noise eps, long long n,
                           float kl_grad_scale, const double* kl_upstream, float prior_mean, float prior_std,
                           float* dmu, float* drho, int fast, void* stream) {
  ED2_TRY(check_dt(dw_dtype));
  return k_normal_grad_finish_any(dw, dw_dtype, mu, rho, N(eps), n, kl_grad_scale, kl_upstream,
                                  Prior{prior_mean, prior_std}, dmu, drho, fast != 0, S(stream));
}

// ------------------------------------------------------------------------------------------ Flipout
int ed2_flipout_dense_fwd(const ed2_flipout_fwd_args* a, void* stream) {
  ED2_TRY(check_dt(a->dtype));
  cudaStream_t s = S(stream);
  const bool split_kl = a->dtype == kBF16 && a->kl_out != nullptr && !fused_kl_enabled();
  ED2_TRY(k_sample_weights(a->mu, a->rho, N(a->eps), a->in_dim, a->out_dim, 1, a->delta_lp, a->dtype, a->delta_ld,
                           a->mu_lp, a->mu_ld, split_kl ? nullptr : a->kl_out, a->kl_scale,
                           Prior{a->prior_mean, a->prior_std}, s));
  ForkJoin fj(s, split_kl);  // the accurate KL reduction runs beside the GEMMs on the side stream
  ED2_TRY(k_scale_rows(a->x, a->dtype, a->x_ld, a->batch, a->in_dim, a->batch, 2, nullptr, nullptr, N(a->sign_in),
                       a->xs, a->xs_ld, nullptr, 0, s));
  // pert = ((x∘S) Δ) ∘ T : T either injected fp32 (used directly) or materialised from Philox into `y` (dtype,
  // ±1 is exact in bf16) — y is overwritten by the second GEMM afterwards.
  EpiParams e = epi_default();
  if (a->sign_out.injected != nullptr) {
    e.elem_scale = a->sign_out.injected; e.elem_scale_ld = a->out_dim; e.elem_scale_dt = kF32;
  } else {
    ED2_TRY(k_fill_noise(1, N(a->sign_out), a->y, a->dtype, a->y_ld, a->batch, a->out_dim, s));
    e.elem_scale = a->y; e.elem_scale_ld = a->y_ld; e.elem_scale_dt = a->dtype;
  }
  e.out = a->pert; e.out_ld = a->pert_ld; e.out_dt = kF32;
  ED2_TRY(gemm_xw(a->dtype, a->xs, a->xs_ld, a->delta_lp, a->delta_ld, a->batch, a->in_dim, a->out_dim, e, s));
  if (split_kl)
    ED2_TRY(k_normal_kl_fwd(a->mu, a->rho, (long long)a->in_dim * a->out_dim, Prior{a->prior_mean, a->prior_std},
                            a->kl_scale, a->kl_out, fj.side()));
  EpiParams e2 = epi_default();
  e2.addend = a->pert; e2.addend_ld = a->pert_ld; e2.addend_dt = kF32;
  e2.col_bias = a->bias; e2.col_bias_ld = a->out_dim;
  e2.act = a->act;
  e2.out = a->y; e2.out_ld = a->y_ld; e2.out_dt = a->dtype;
  const int st = gemm_xw(a->dtype, a->x, a->x_ld, a->mu_lp, a->mu_ld, a->batch, a->in_dim, a->out_dim, e2, s);
  fj.join();
  return st;
}

int ed2_flipout_dense_bwd(const ed2_flipout_bwd_args* a, void* stream) {
  ED2_TRY(check_dt(a->dtype));
  cudaStream_t s = S(stream);
  BwdOutArgs o{};
  o.dt = a->dtype; o.rows = a->batch; o.cols = a->out_dim; o.rows_per_member = a->batch; o.act = a->act; o.fast = 3;
  o.gy = a->gy; o.gy_ld = a->gy_ld; o.y = a->y; o.y_ld = a->y_ld;
  o.sign_out = N(a->sign_out);
  o.dz = a->gp; o.dz_ld = a->gp_ld; o.dz2 = a->gt; o.dz2_ld = a->gt_ld;
  o.dbias = a->dbias;
  ED2_TRY(k_bwd_out(o, s));
  // dmu_w = xᵀ gp -> dmu ; dΔ = (x∘S)ᵀ (gp∘T) -> drho ; finishing pass combines with eps, sigmoid(rho), KL.
  EpiParams e = epi_default();
  e.out = a->dmu; e.out_ld = a->out_dim; e.out_dt = kF32;
  ED2_TRY(gemm_xtg(a->dtype, a->x, a->x_ld, a->gp, a->gp_ld, a->batch, a->in_dim, a->out_dim, e, s));
  EpiParams e1 = epi_default();
  e1.out = a->drho; e1.out_ld = a->out_dim; e1.out_dt = kF32;
  ED2_TRY(gemm_xtg(a->dtype, a->xs, a->xs_ld, a->gt, a->gt_ld, a->batch, a->in_dim, a->out_dim, e1, s));
  ED2_TRY(k_normal_grad_finish(a->dmu, a->drho, a->mu, a->rho, N(a->eps), (long long)a->in_dim * a->out_dim,
                               a->kl_grad_scale, a->kl_upstream, Prior{a->prior_mean, a->prior_std}, a->dmu, a->drho,
                               a->dtype == kBF16, s));
  if (a->dx != nullptr) {
    // tmp = ((gp∘T) Δᵀ) ∘ S ;  dx = gp μᵀ + tmp
    EpiParams e2 = epi_default();
    const void* sign_in_buf = a->sign_in.injected;
    int sign_dt = kF32, sign_ld = a->in_dim;
    if (sign_in_buf == nullptr) {
      // materialise S into dx (dtype), consumed as elem_scale before dx is overwritten by the last GEMM
      ED2_TRY(k_fill_noise(1, N(a->sign_in), a->dx, a->dtype, a->dx_ld, a->batch, a->in_dim, s));
      sign_in_buf = a->dx; sign_dt = a->dtype; sign_ld = a->dx_ld;
    }
    e2.elem_scale = sign_in_buf; e2.elem_scale_ld = sign_ld; e2.elem_scale_dt = sign_dt;
    e2.out = a->tmp; e2.out_ld = a->tmp_ld; e2.out_dt = kF32;
    ED2_TRY(gemm_gwt(a->dtype, a->gt, a->gt_ld, a->delta_lp, a->delta_ld, a->batch, a->in_dim, a->out_dim, e2, s));
    EpiParams e3 = epi_default();
    e3.addend = a->tmp; e3.addend_ld = a->tmp_ld; e3.addend_dt = kF32;
    e3.out = a->dx; e3.out_ld = a->dx_ld; e3.out_dt = a->dtype;
    ED2_TRY(gemm_gwt(a->dtype, a->gp, a->gp_ld, a->mu_lp, a->mu_ld, a->batch, a->in_dim, a->out_dim, e3, s));
  }
  return ED2_OK;
}

// Fused Flipout: one dual-accumulator launch per product pair (gemm_flipout.cuh)
namespace {
int sign_side(const ed2_noise& n, int batch, SignSide* out, const char* what) {
  if (n.injected != nullptr) return set_error(ED2_ERR_BAD_ARG, "flipout_fused: %s must be a Philox stream (injected signs take the unfused call)", what);
  // Flipout batches are ungrouped: a single group covering the local batch is the same mapping (global row = row0 + m)
  if (n.rows_per_group_local != 0 && n.rows_per_group_local < batch)
    return set_error(ED2_ERR_BAD_ARG, "flipout_fused: %s must be ungrouped (rows_per_group_local %d < batch %d)", what, n.rows_per_group_local, batch);
  out->seed = n.seed; out->stream = n.stream; out->step = reinterpret_cast<const unsigned long long*>(n.step);
  out->row0 = n.row0; out->on = 1;
  return ED2_OK;
}
}  // namespace

int ed2_flipout_fusable(int batch, int in_dim, int out_dim) {
  // the three products: x·W [B x O, K = I], g·Wᵀ [B x I, K = O], xᵀ·g [I x O, K = B]; a 256-row pair tile needs a batch
  // that fills it
  return (batch > 128 && gemm_flipout_fusable(batch, out_dim, in_dim) && gemm_flipout_fusable(batch, in_dim, out_dim) &&
          gemm_flipout_fusable(in_dim, out_dim, batch)) ? 1 : 0;
}

int ed2_flipout_fused_fwd(const ed2_flipout_fused_fwd_args* a, void* stream) {
  cudaStream_t s = S(stream);
  if (!ed2_flipout_fusable(a->batch, a->in_dim, a->out_dim))
    return set_error(ED2_ERR_BAD_SHAPE, "flipout_fused_fwd: shape %d x %d -> %d does not qualify (ed2_flipout_fusable)", a->batch, a->in_dim, a->out_dim);
  FlipoutGemm g{};
  ED2_TRY(sign_side(a->sign_out, a->batch, &g.epi.sa, "sign_out"));
  if (!a->xs_ready)
    ED2_TRY(k_scale_rows(a->x, kBF16, a->x_ld, a->batch, a->in_dim, a->batch, 2, nullptr, nullptr, N(a->sign_in), a->xs,
                         a->xs_ld, nullptr, 0, s));
  g.A0 = a->x; g.lda0 = a->x_ld; g.A1 = a->xs; g.lda1 = a->xs_ld;
  g.B0 = a->mu_lp; g.ldb0 = a->mu_ld; g.B1 = a->delta_lp; g.ldb1 = a->delta_ld; g.major_b = kMajorMN;
  g.M = a->batch; g.N = a->out_dim; g.K = a->in_dim;
  g.out = a->y; g.out_ld = a->y_ld;
  g.epi.bias = a->bias; g.epi.act = a->act;
  if (a->next_xs != nullptr) {
    ED2_TRY(sign_side(a->next_sign_in, a->batch, &g.epi.sb, "next_sign_in"));
    g.epi.out2 = a->next_xs; g.epi.out2_ld = a->next_xs_ld;
  }
  return gemm_flipout(g, s);
}

int ed2_flipout_fused_bwd(const ed2_flipout_fused_bwd_args* a, void* stream) {
  cudaStream_t s = S(stream);
  if (!ed2_flipout_fusable(a->batch, a->in_dim, a->out_dim))
    return set_error(ED2_ERR_BAD_SHAPE, "flipout_fused_bwd: shape %d x %d -> %d does not qualify (ed2_flipout_fusable)", a->batch, a->in_dim, a->out_dim);
  const bool want_dx = a->dx != nullptr || a->prev_gp != nullptr;
  FlipoutGemm g{};
  if (want_dx) {
    ED2_TRY(sign_side(a->sign_in, a->batch, &g.epi.sa, "sign_in"));
    if (a->prev_gt != nullptr) ED2_TRY(sign_side(a->prev_sign_out, a->batch, &g.epi.sb, "prev_sign_out"));
    if (a->prev_dbias != nullptr) cudaMemsetAsync(a->prev_dbias, 0, sizeof(float) * (size_t)a->in_dim, s);
  }
  if (!a->premade) {
    BwdOutArgs o{};
    o.dt = kBF16; o.rows = a->batch; o.cols = a->out_dim; o.rows_per_member = a->batch; o.act = a->act; o.fast = 3;
    o.gy = a->gy; o.gy_ld = a->gy_ld; o.y = a->y; o.y_ld = a->y_ld;
    o.sign_out = N(a->sign_out);
    o.dz = a->gp; o.dz_ld = a->gp_ld; o.dz2 = a->gt; o.dz2_ld = a->gt_ld;
    o.dbias = a->dbias;
    ED2_TRY(k_bwd_out(o, s));
  }
  // both weight-gradient products in ONE launch, finished in its epilogue (eps regenerated, KL gradient added): the raw
  // fp32 gradients are never written
  {
    FlipoutGemm w{};
    w.mode = 1; w.major_a = kMajorMN; w.major_b = kMajorMN;
    w.A0 = a->x; w.lda0 = a->x_ld; w.A1 = a->xs; w.lda1 = a->xs_ld;
    w.B0 = a->gp; w.ldb0 = a->gp_ld; w.B1 = a->gt; w.ldb1 = a->gt_ld;
    w.M = a->in_dim; w.N = a->out_dim; w.K = a->batch;
    if (a->eps.injected != nullptr) return set_error(ED2_ERR_BAD_ARG, "flipout_fused_bwd: eps must be a Philox stream");
    w.epi.sa.seed = a->eps.seed; w.epi.sa.stream = a->eps.stream;
    w.epi.sa.step = reinterpret_cast<const unsigned long long*>(a->eps.step); w.epi.sa.on = 1;
    w.epi.mu = a->mu; w.epi.rho = a->rho; w.epi.dmu = a->dmu; w.epi.drho = a->drho;
    w.epi.klg = a->kl_grad_scale; w.epi.kl_upstream = a->kl_upstream;
    w.epi.prior_mean = a->prior_mean; w.epi.prior_std = a->prior_std;
    ED2_TRY(gemm_flipout(w, s));
  }
  if (want_dx) {
    g.A0 = a->gp; g.lda0 = a->gp_ld; g.A1 = a->gt; g.lda1 = a->gt_ld;
    g.B0 = a->mu_lp; g.ldb0 = a->mu_ld; g.B1 = a->delta_lp; g.ldb1 = a->delta_ld; g.major_b = kMajorK;
    g.M = a->batch; g.N = a->in_dim; g.K = a->out_dim;
    if (a->prev_gp != nullptr) {
      g.out = a->prev_gp; g.out_ld = a->prev_gp_ld;
      if (a->prev_relu) { g.epi.mask_src = a->x; g.epi.mask_ld = a->x_ld; }
      if (a->prev_gt != nullptr) { g.epi.out2 = a->prev_gt; g.epi.out2_ld = a->prev_gt_ld; }
      g.epi.col_sum = a->prev_dbias;
    } else {
      g.out = a->dx; g.out_ld = a->dx_ld;
    }
    ED2_TRY(gemm_flipout(g, s));
  }
  return ED2_OK;
}

long long ed2_spd_inverse_workspace_floats(int n) { return k_spd_inverse_workspace_floats(n); }
int ed2_spd_inverse(float* a, int n, float* workspace, void* stream) {
  if (a == nullptr || workspace == nullptr) return set_error(ED2_ERR_BAD_ARG, "spd_inverse: null pointer");
  return k_spd_inverse(a, n, workspace, S(stream));
}

// ------------------------------------------------------------------------------------------ heteroscedastic heads
int ed2_mc_softmax_fa_fwd(const float* loc, const float* fac, const float* diag_raw, const float* sngp_var, float w_het,
                          float w_sngp, ed2_noise noise, int batch, int num_samples, int num_classes, int num_factors,
                          int share_samples, float temperature, float eps, float* probs_mean, float* logits,
                          float* variance, void* stream) {
  if (temperature <= 0.f) return set_error(ED2_ERR_BAD_ARG, "mc_softmax: temperature must be positive");
  return k_mc_softmax_fwd(loc, fac, diag_raw, sngp_var, w_het, w_sngp, N(noise), batch, num_samples, num_classes,
                          num_factors, share_samples, temperature, eps, probs_mean, logits, variance, S(stream));
}
int ed2_mc_softmax_fa_bwd(const float* g, const float* probs_mean, const float* loc, const float* fac,
                          const float* diag_raw, ed2_noise noise, int batch, int num_samples, int num_classes,
                          int num_factors, int share_samples, float temperature, float eps, float* dloc, float* dfac,
                          float* ddiag_raw, void* stream) {
  if (temperature <= 0.f) return set_error(ED2_ERR_BAD_ARG, "mc_softmax: temperature must be positive");
  return k_mc_softmax_bwd(g, probs_mean, loc, fac, diag_raw, N(noise), batch, num_samples, num_classes, num_factors,
                          share_samples, temperature, eps, dloc, dfac, ddiag_raw, S(stream));
}

// ------------------------------------------------------------------------------------------ LSTM cells
int ed2_lstm_pointwise_fwd(const void* z, int dtype, long long z_ld, const float* c_prev, int batch, int units, float* c,
                           void* h, long long h_ld, void* stream) {
  ED2_TRY(check_dt(dtype));
  return k_lstm_pointwise_fwd(z, dtype, z_ld, c_prev, batch, units, c, h, h_ld, S(stream));
}
int ed2_lstm_pointwise_bwd(const void* z, int dtype, long long z_ld, const float* c_prev, const float* c, const void* dh,
                           long long dh_ld, const float* dc, int batch, int units, void* dz, long long dz_ld,
                           float* dc_prev, float* dbias, void* stream) {
  ED2_TRY(check_dt(dtype));
  return k_lstm_pointwise_bwd(z, dtype, z_ld, c_prev, c, dh, dh_ld, dc, batch, units, dz, dz_ld, dc_prev, dbias, S(stream));
}

int ed2_lstm_gates_fwd(const void* z, const void* z2, int dtype, long long z_ld, long long z2_ld, const float* c_prev,
                       int batch, int units, float* c, void* h, long long h_ld, void* stream) {
  ED2_TRY(check_dt(dtype));
  return k_lstm_pointwise_fwd(z, dtype, z_ld, c_prev, batch, units, c, h, h_ld, S(stream), z2, z2_ld);
}
int ed2_lstm_gates_bwd(const void* z, const void* z2, int dtype, long long z_ld, long long z2_ld, const float* c_prev,
                       const float* c, const void* dh, long long dh_ld, const float* dc, int batch, int units, void* dz,
                       long long dz_ld, float* dc_prev, void* stream) {
  ED2_TRY(check_dt(dtype));
  return k_lstm_pointwise_bwd(z, dtype, z_ld, c_prev, c, dh, dh_ld, dc, batch, units, dz, dz_ld, dc_prev, nullptr, S(stream), z2,
                              z2_ld);
}

// ------------------------------------------------------------------------------------------ Conv2D support
namespace {
ConvArgs conv_args(const ed2_conv_geometry* g) {
  return ConvArgs{g->batch, g->height, g->width, g->channels, g->kh, g->kw, g->sh, g->sw, g->same};
}
}  // namespace

int ed2_conv_out_size(const ed2_conv_geometry* g, int* out_h, int* out_w) {
  if (g == nullptr || out_h == nullptr || out_w == nullptr) return set_error(ED2_ERR_BAD_ARG, "conv_out_size: null argument");
  return k_conv_out_size(conv_args(g), out_h, out_w);
}
int ed2_im2col(const ed2_conv_geometry* g, const void* x, int dtype, void* cols, long long cols_ld, void* stream) {
  ED2_TRY(check_dt(dtype));
  return k_im2col(conv_args(g), x, dtype, cols, cols_ld, S(stream));
}
int ed2_col2im(const ed2_conv_geometry* g, const void* dcols, int dtype, long long cols_ld, void* dx, int dx_dtype,
               void* stream) {
  ED2_TRY(check_dt(dtype));
  ED2_TRY(check_dt(dx_dtype));
  return k_col2im(conv_args(g), dcols, dtype, cols_ld, dx, dx_dtype, S(stream));
}
int ed2_conv_factor_fwd(const ed2_conv_factor_fwd_args* a, void* stream) {
  ED2_TRY(check_dt(a->dtype));
  cudaStream_t s = S(stream);
  const ConvArgs g = conv_args(&a->geom);
  int ho = 0, wo = 0;
  ED2_TRY(k_conv_out_size(g, &ho, &wo));
  const int M = g.batch * ho * wo, K = g.kh * g.kw * g.channels;
  ED2_TRY(k_im2col(g, a->x, a->dtype, a->xa, a->xa_ld, s, a->f_in));
  EpiParams e = epi_default();
  e.group_rows = ho * wo;
  e.col_scale = a->f_out; e.col_scale_ld = a->out_channels;
  e.col_bias = a->bias; e.col_bias_ld = a->out_channels;
  e.act = a->act;
  e.out = a->y; e.out_ld = a->y_ld; e.out_dt = a->dtype;
  e.raw = a->z; e.raw_ld = a->z_ld; e.raw_dt = a->dtype;
  return gemm_xw(a->dtype, a->xa, a->xa_ld, a->w_lp, a->w_ld, M, K, a->out_channels, e, s);
}

int ed2_conv_factor_bwd(const ed2_conv_factor_bwd_args* a, void* stream) {
  ED2_TRY(check_dt(a->dtype));
  cudaStream_t s = S(stream);
  const ConvArgs g = conv_args(&a->geom);
  int ho = 0, wo = 0;
  ED2_TRY(k_conv_out_size(g, &ho, &wo));
  const int M = g.batch * ho * wo, K = g.kh * g.kw * g.channels, O = a->out_channels;
  BwdOutArgs o{};
  o.dt = a->dtype; o.rows = M; o.cols = O; o.rows_per_member = ho * wo; o.act = a->act; o.fast = 1;
  o.gy = a->gy; o.gy_ld = a->gy_ld; o.y = a->y; o.y_ld = a->y_ld; o.z = a->z; o.z_ld = a->z_ld;
  o.gamma = a->f_out;
  o.dz = a->dz; o.dz_ld = a->dz_ld;
  o.d_fast_mu = a->df_out; o.dbias = a->dbias;
  ED2_TRY(k_bwd_out(o, s));
  EpiParams e = epi_default();
  e.out = a->dw; e.out_ld = O; e.out_dt = kF32;
  ED2_TRY(gemm_xtg(a->dtype, a->xa, a->xa_ld, a->dz, a->dz_ld, M, K, O, e, s));
  if (a->dxa == nullptr) return ED2_OK;
  EpiParams e2 = epi_default();
  e2.out = a->dxa; e2.out_ld = a->dxa_ld; e2.out_dt = a->dtype;
  ED2_TRY(gemm_gwt(a->dtype, a->dz, a->dz_ld, a->w_lp, a->w_ld, M, K, O, e2, s));
  if (a->f_in == nullptr && a->dx != nullptr) return k_col2im(g, a->dxa, a->dtype, a->dxa_ld, a->dx, a->dtype, s);
  ED2_TRY(k_col2im(g, a->dxa, a->dtype, a->dxa_ld, a->gimg, a->dtype, s));
  return k_conv_scale_reduce(g, a->gimg, a->x, a->dtype, a->f_in, a->dx, a->dtype, a->df_in, s);
}

int ed2_expand_table(const float* src, float* out, int batch, int channels, int rows_per_entry, int reps, void* stream) {
  return k_expand_table(src, out, batch, channels, rows_per_entry, reps, S(stream));
}
int ed2_expand_table_bwd(const float* dout, float* dsrc, int batch, int channels, int rows_per_entry, int reps, void* stream) {
  return k_expand_table_bwd(dout, dsrc, batch, channels, rows_per_entry, reps, S(stream));
}
int ed2_rank1_factors(const float* mu, const float* rho, ed2_noise eps, int batch, int ensemble_size, int channels,
                      float clip_lo, float clip_hi, float* f, void* stream) {
  if (ensemble_size <= 0 || batch % ensemble_size != 0)
    return set_error(ED2_ERR_BAD_SHAPE, "batch %d is not divisible by ensemble_size %d", batch, ensemble_size);
  FastOpts fo;
  fo.clip_lo = clip_lo; fo.clip_hi = clip_hi;
  return k_scale_rows(nullptr, kF32, 0, batch, channels, batch / ensemble_size, rho != nullptr ? 1 : 0, mu, rho, N(eps), nullptr,
                      0, f, channels, S(stream), fo);
}
int ed2_rank1_factors_bwd(const float* df, const float* mu, const float* rho, ed2_noise eps, int batch, int ensemble_size,
                          int channels, float clip_lo, float clip_hi, float* dmu, float* drho, void* stream) {
  if (ensemble_size <= 0 || batch % ensemble_size != 0)
    return set_error(ED2_ERR_BAD_SHAPE, "batch %d is not divisible by ensemble_size %d", batch, ensemble_size);
  return k_rank1_factor_bwd(df, mu, rho, N(eps), ensemble_size, batch / ensemble_size, channels, clip_lo, clip_hi, dmu, drho,
                            S(stream));
}
int ed2_noise_table(int kind, ed2_noise nz, void* out, int dtype, long long ld, long long rows, long long cols, void* stream) {
  ED2_TRY(check_dt(dtype));
  if (kind != 0 && kind != 1) return set_error(ED2_ERR_BAD_ARG, "noise_table: kind must be 0 (normal) or 1 (sign)");
  if (nz.injected != nullptr) return set_error(ED2_ERR_BAD_ARG, "noise_table: the descriptor holds injected values already");
  return k_fill_noise(kind, N(nz), out, dtype, ld, rows, cols, S(stream));
}
int ed2_normal_grad_finish2(const float* d_mean, const float* d_pert, const float* mu, const float* rho, ed2_noise eps,
                            long long n, float kl_grad_scale, const double* kl_upstream, float prior_mean, float prior_std,
                            float* dmu, float* drho, int fast, void* stream) {
  return k_normal_grad_finish(d_mean, d_pert, mu, rho, N(eps), n, kl_grad_scale, kl_upstream, Prior{prior_mean, prior_std},
                              dmu, drho, fast != 0, S(stream));
}

// ------------------------------------------------------------------------------------------ SNGP head
int ed2_spectral_norm_update(const float* w, int in_dim, int out_dim, float* u, float* v, int iterations,
                             float norm_multiplier, int mode, int update_uv, float* w_out, void* w_out_lp, int w_lp_ld,
                             float* sigma_out, float* workspace, void* stream) {
  if (in_dim <= 0 || out_dim <= 0) return set_error(ED2_ERR_BAD_SHAPE, "spectral_norm: empty kernel");
  return k_spectral_norm(w, in_dim, out_dim, u, v, iterations, norm_multiplier, mode, update_uv, w_out, w_out_lp,
                         w_lp_ld, sigma_out, workspace, S(stream));
}

int ed2_input_normalize(const void* x, int x_dtype, int x_ld, void* y, int y_dtype, int y_ld, long long rows, int cols,
                        const float* gamma, const float* beta, float eps, int scale_only, float scale, void* stream) {
  ED2_TRY(check_dt(x_dtype));
  ED2_TRY(check_dt(y_dtype));
  return k_input_normalize(x, x_dtype, x_ld, y, y_dtype, y_ld, rows, cols, gamma, beta, eps, scale_only, scale, S(stream));
}

int ed2_layernorm_bwd(const void* x, int x_dtype, int x_ld, const void* dy, int dy_dtype, int dy_ld, long long rows,
                      int cols, const float* gamma, float eps, void* dx, int dx_dtype, int dx_ld, float* dgamma,
                      float* dbeta, void* stream) {
  ED2_TRY(check_dt(x_dtype));
  ED2_TRY(check_dt(dy_dtype));
  return k_layernorm_bwd(x, x_dtype, x_ld, dy, dy_dtype, dy_ld, rows, cols, gamma, eps, dx, dx_dtype, dx_ld, dgamma,
                         dbeta, S(stream));
}

int ed2_layernorm_split(const void* x, int x_dtype, int x_ld, void* out, int out_ld, long long rows, int cols,
                        const float* gamma, const float* beta, float eps, void* stream) {
  ED2_TRY(check_dt(x_dtype));
  return k_layernorm_split(x, x_dtype, x_ld, out, out_ld, rows, cols, gamma, beta, eps, S(stream));
}

int ed2_rff_fwd(int dtype, const void* x, int x_ld, const void* wt, int wt_ld, const float* bias, float scale, int batch,
                int in_dim, int num_features, void* phi, int phi_ld, int phi_dtype, void* pre, int pre_ld, int pre_mode,
                void* stream) {
  ED2_TRY(check_dt(dtype));
  GemmDesc g{};
  g.A = x; g.lda = x_ld; g.major_a = kMajorK;
  g.B = wt; g.ldb = wt_ld; g.major_b = kMajorK;
  g.M = batch; g.N = num_features; g.K = in_dim;
  g.ep = epi_default();
  g.ep.col_bias = bias; g.ep.col_bias_ld = num_features;
  g.ep.act = kActCos; g.ep.act_scale = scale;
  g.ep.out = phi; g.ep.out_ld = phi_ld; g.ep.out_dt = phi_dtype;
  // pre_mode 0: the fp32 phase x W (bias not included);  1: -scale * sin(x W + b) in bf16, the derivative factor
  g.ep.raw = pre; g.ep.raw_ld = pre_ld; g.ep.raw_dt = pre_mode == 1 ? kBF16 : kF32;
  g.ep.raw_neg_sin = pre_mode == 1 ? 1 : 0;
  if (pre_mode == 1 && dtype != kBF16) return set_error(ED2_ERR_BAD_ARG, "rff_fwd: pre_mode 1 needs the ED2_BF16 GEMM");
  return gemm(dtype, g, S(stream));
}

int ed2_laplace_precision_update(int dtype, const void* phi, int phi_ld, int batch, int num_features, float momentum,
                                 long long batch_total, float* precision, float* partial_out, void* stream) {
  ED2_TRY(check_dt(dtype));
  GemmDesc g{};
  g.A = phi; g.lda = phi_ld; g.major_a = kMajorMN;
  g.B = phi; g.ldb = phi_ld; g.major_b = kMajorMN;
  g.M = num_features; g.N = num_features; g.K = batch;
  g.ep = epi_default();
  cudaStream_t s = S(stream);
  // PhiT Phi is symmetric: only the tiles on or below the diagonal are computed (SYRK), the strict upper triangle is
  // mirrored afterwards; with few output tiles and a long K (D = 1024: 10 tiles for 74 SM pairs) the K loop is split and
  // the slices accumulate with fp32 atomics
  const bool bf16_path = dtype == kBF16;
  g.ep.lower_only = bf16_path ? 1 : 0;
  g.ep.split_k = bf16_path ? -1 : 0;
  float* target = partial_out != nullptr ? partial_out : precision;
  const long long nn = static_cast<long long>(num_features) * num_features;
  const float a = (partial_out == nullptr && momentum > 0.f) ? (1.f - momentum) / static_cast<float>(batch_total) : 1.f;
  if (bf16_path) {
    // accumulate form for every case: target <- base, then += a * PhiT Phi by the GEMM
    if (partial_out != nullptr) cudaMemsetAsync(partial_out, 0, sizeof(float) * nn, s);
    else if (momentum > 0.f) ED2_TRY(k_scale_inplace(precision, nn, momentum, s));
    g.ep.alpha = a;
    g.ep.out = target; g.ep.out_ld = num_features; g.ep.out_dt = kF32;
    // split_k == 1 after the automatic choice: plain store epilogue needs the addend form instead of atomics
    g.ep.split_k = -1;
    ED2_TRY(gemm_bf16_accumulate(g, s));
    return k_mirror_lower(target, num_features, s);
  }
  if (partial_out != nullptr) {
    g.ep.out = partial_out; g.ep.out_ld = num_features; g.ep.out_dt = kF32;
  } else {
    // in-place blend: every element of P is read (addend) and written (out) by the same thread
    g.ep.addend = precision; g.ep.addend_ld = num_features; g.ep.addend_dt = kF32;
    if (momentum > 0.f) {
      g.ep.alpha = (1.f - momentum) / static_cast<float>(batch_total);
      g.ep.beta = momentum;
    }
    g.ep.out = precision; g.ep.out_ld = num_features; g.ep.out_dt = kF32;
  }
  return gemm(dtype, g, s);
}

int ed2_precision_blend(float* precision, const float* summed, int num_features, float momentum, long long batch_total,
                        void* stream) {
  return k_precision_blend(precision, summed, (long long)num_features * num_features, momentum, batch_total, S(stream));
}

int ed2_predictive_variance(int dtype, const void* phi, int phi_ld, const void* sigma_lp, int sigma_ld, int batch,
                            int num_features, float ridge, float* var, const float* logits, int num_classes,
                            float mean_field_factor, int likelihood, float* logits_out, void* stream) {
  ED2_TRY(check_dt(dtype));
  cudaStream_t s = S(stream);
  cudaMemsetAsync(var, 0, sizeof(float) * (size_t)batch, s);
  GemmDesc g{};
  g.A = phi; g.lda = phi_ld; g.major_a = kMajorK;
  g.B = sigma_lp; g.ldb = sigma_ld; g.major_b = kMajorK;  // Sigma is symmetric
  g.M = batch; g.N = num_features; g.K = num_features;
  g.ep = epi_default();
  g.ep.alpha = ridge;
  g.ep.elem_scale = phi; g.ep.elem_scale_ld = phi_ld; g.ep.elem_scale_dt = dtype;
  g.ep.row_sum = var;
  ED2_TRY(gemm(dtype, g, s));
  if (logits != nullptr && logits_out != nullptr)
    ED2_TRY(k_mean_field(logits, var, batch, num_classes, mean_field_factor, likelihood, 1.f, logits_out, s));
  return ED2_OK;
}

int ed2_mean_field_logits(const float* logits, const float* var, int batch, int num_classes, float mean_field_factor,
                          int likelihood, float* out, void* stream) {
  return k_mean_field(logits, var, batch, num_classes, mean_field_factor, likelihood, 1.f, out, S(stream));
}

int ed2_rff_bwd(int dtype, const void* dphi, int dphi_ld, int dphi_dtype, const void* pre, int pre_ld, int pre_mode,
                const float* bias, const void* w, int w_ld, float scale, int batch, int in_dim, int num_features,
                void* gpre, int gpre_ld, void* dx, int dx_ld, int dx_dtype, void* stream) {
  ED2_TRY(check_dt(dtype));
  cudaStream_t s = S(stream);
  if (pre_mode == 1) {
    if (dtype != kBF16 || dphi_dtype != kBF16) return set_error(ED2_ERR_BAD_ARG, "rff_bwd: pre_mode 1 is the ED2_BF16 path");
    ED2_TRY(k_mul_bf16(dphi, dphi_ld, pre, pre_ld, gpre, gpre_ld, batch, num_features, s));
  } else {
    ED2_TRY(k_rff_bwd_prep(dphi, dphi_dtype, dphi_ld, reinterpret_cast<const float*>(pre), pre_ld, bias, scale, batch,
                           num_features, gpre, dtype, gpre_ld, s));
  }
  EpiParams e = epi_default();
  e.out = dx; e.out_ld = dx_ld; e.out_dt = dx_dtype;
  return gemm_gwt(dtype, gpre, gpre_ld, w, w_ld, batch, in_dim, num_features, e, s);
}

}  // extern "C"
