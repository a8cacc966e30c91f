"""Op-level Python wrappers over the C ABI (thin: fill a struct, pass pointers, check the status)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ACT_COS, ACT_NONE, ACT_RELU, BF16, F32, MAJOR_K, MAJOR_MN  # noqa: F401

ACTIVATIONS = {None: ACT_NONE, "linear": ACT_NONE, "relu": ACT_RELU, "cos": ACT_COS}


def act_id(activation) -> int:
    if callable(activation):
        name = getattr(activation, "__name__", None)
        if name in ACTIVATIONS:
            return ACTIVATIONS[name]
        raise ValueError(f"activation {activation!r} is not fused by libed2b200 (supported: None, 'relu', 'cos')")
    if activation not in ACTIVATIONS:
        raise ValueError(f"activation {activation!r} is not fused by libed2b200 (supported: None, 'relu', 'cos')")
    return ACTIVATIONS[activation]


def gemm(a: torch.Tensor, b: torch.Tensor, m: int, n: int, k: int, *, major_a=MAJOR_K, major_b=MAJOR_MN,
         out: torch.Tensor | None = None, out_dtype: torch.dtype | None = None, alpha=1.0, beta=1.0, act=ACT_NONE,
         act_scale=1.0, group_rows=0, elem_scale=None, col_scale=None, addend=None, col_bias=None, raw=None,
         row_sum=None, want_out=True, block_n=0, cta_group=0, max_ctas=0, no_tma_store=0, accumulate=False,
         lower_only=False) -> torch.Tensor | None:
    """out = act(((alpha*A·B) ∘ elem_scale ∘ col_scale[g]) + beta*addend + col_bias[g]); see ed2b200.h."""
    lib = _lib.load()
    d = _lib.GemmDesc()
    d.dtype = _lib.dt_of(a)
    assert a.dtype == b.dtype
    d.a, d.lda, d.major_a = a.data_ptr(), _lib.ld(a), major_a
    d.b, d.ldb, d.major_b = b.data_ptr(), _lib.ld(b), major_b
    d.m, d.n, d.k = m, n, k
    d.alpha, d.beta, d.act_scale, d.act, d.group_rows = alpha, beta, act_scale, act, group_rows
    if elem_scale is not None:
        d.elem_scale, d.elem_scale_ld, d.elem_scale_dtype = elem_scale.data_ptr(), _lib.ld(elem_scale), _lib.dt_of(elem_scale)
    if col_scale is not None:
        assert col_scale.dtype == torch.float32
        d.col_scale, d.col_scale_ld = col_scale.data_ptr(), _lib.ld(col_scale)
    if addend is not None:
        d.addend, d.addend_ld, d.addend_dtype = addend.data_ptr(), _lib.ld(addend), _lib.dt_of(addend)
    if col_bias is not None:
        assert col_bias.dtype == torch.float32
        d.col_bias, d.col_bias_ld = col_bias.data_ptr(), _lib.ld(col_bias)
    if out is None and want_out:
        out = _lib.empty_padded(m, n, out_dtype or a.dtype, a.device)
    if out is not None:
        d.out, d.out_ld, d.out_dtype = out.data_ptr(), _lib.ld(out), _lib.dt_of(out)
    if raw is not None:
        d.raw, d.raw_ld, d.raw_dtype = raw.data_ptr(), _lib.ld(raw), _lib.dt_of(raw)
    if row_sum is not None:
        assert row_sum.dtype == torch.float32
        d.row_sum = row_sum.data_ptr()
    d.block_n, d.cta_group, d.max_ctas, d.no_tma_store = block_n, cta_group, max_ctas, no_tma_store
    d.accumulate, d.lower_only = int(accumulate), int(lower_only)
    _lib.check(lib.ed2_gemm(C.byref(d), _lib.stream_ptr()), "ed2_gemm")
    return out


def philox(kind: str, n: int, seed: int, stream: int, device="cuda") -> torch.Tensor:
    lib = _lib.load()
    if kind == "raw":
        out = torch.empty(n * 4, dtype=torch.int32, device=device)
        _lib.check(lib.ed2_philox_raw(out.data_ptr(), n, seed, stream, _lib.stream_ptr()), "philox_raw")
        return out
    out = torch.empty(n, dtype=torch.float32, device=device)
    fn = {"normal": lib.ed2_philox_normal, "normal_fast": lib.ed2_philox_normal_fast, "uniform": lib.ed2_philox_uniform,
          "sign": lib.ed2_philox_sign}[kind]
    _lib.check(fn(out.data_ptr(), n, seed, stream, _lib.stream_ptr()), f"philox_{kind}")
    return out


def cast(src: torch.Tensor, dtype: torch.dtype, out: torch.Tensor | None = None) -> torch.Tensor:
    """2-D cast into a (possibly pitch-padded) buffer of `dtype`."""
    lib = _lib.load()
    rows, cols = src.shape
    if out is None:
        out = _lib.empty_padded(rows, cols, dtype, src.device)
    _lib.check(lib.ed2_cast(src.data_ptr(), _lib.dt_of(src), _lib.ld(src), out.data_ptr(), _lib.dt_of(out),
                            _lib.ld(out), rows, cols, _lib.stream_ptr()), "ed2_cast")
    return out


def normal_kl(mu: torch.Tensor, rho: torch.Tensor, prior_mean=0.0, prior_std=1.0, scale=1.0) -> torch.Tensor:
    lib = _lib.load()
    out = torch.zeros((), dtype=torch.float64, device=mu.device)
    _lib.check(lib.ed2_normal_kl_fwd(mu.data_ptr(), rho.data_ptr(), mu.numel(), prior_mean, prior_std, scale,
                                     out.data_ptr(), _lib.stream_ptr()), "normal_kl_fwd")
    return out


def normal_kl_bwd(mu, rho, dmu, drho, prior_mean=0.0, prior_std=1.0, scale=1.0, upstream=None) -> None:
    lib = _lib.load()
    _lib.check(lib.ed2_normal_kl_bwd(mu.data_ptr(), rho.data_ptr(), mu.numel(), prior_mean, prior_std, scale,
                                     _lib.ptr(upstream), dmu.data_ptr(), drho.data_ptr(), _lib.stream_ptr()),
               "normal_kl_bwd")


def act_bwd(gy: torch.Tensor, y: torch.Tensor, gp: torch.Tensor, dbias: torch.Tensor | None, act: int) -> None:
    """gp = gy ∘ act'(y); dbias = column sums of gp."""
    lib = _lib.load()
    rows, cols = y.shape
    _lib.check(lib.ed2_act_bwd(_lib.dt_of(y), gy.data_ptr(), _lib.ld(gy), y.data_ptr(), _lib.ld(y), rows, cols, act,
                               gp.data_ptr(), _lib.ld(gp), _lib.ptr(dbias), _lib.stream_ptr()), "ed2_act_bwd")


# ------------------------------------------------------------------------------------------------- SNGP ops
def spectral_norm_update(w: torch.Tensor, u: torch.Tensor, v: torch.Tensor, iterations: int, norm_multiplier: float,
                         mode: int, update_uv: bool, w_out: torch.Tensor | None, w_out_lp: torch.Tensor | None = None):
    """Power iteration + normalisation (ed2_spectral_norm_update).  Returns (sigma, factor) device scalars
    (factor = multiplicative scale applied to w)."""
    lib = _lib.load()
    in_dim, out_dim = w.shape
    assert w.is_contiguous() and w.dtype == torch.float32
    ws = torch.empty(in_dim + out_dim + 8, dtype=torch.float32, device=w.device)
    _lib.check(lib.ed2_spectral_norm_update(w.data_ptr(), in_dim, out_dim, u.data_ptr(), v.data_ptr(), iterations,
                                            norm_multiplier, mode, int(update_uv), _lib.ptr(w_out), _lib.ptr(w_out_lp),
                                            0 if w_out_lp is None else _lib.ld(w_out_lp), None, ws.data_ptr(),
                                            _lib.stream_ptr()), "ed2_spectral_norm_update")
    scal = ws[in_dim + out_dim:]
    return scal[0], scal[1]


def input_normalize(x: torch.Tensor, out_dtype: torch.dtype, gamma=None, beta=None, eps=1e-3, scale_only=False,
                    scale=1.0) -> torch.Tensor:
    lib = _lib.load()
    rows, cols = x.shape
    y = _lib.empty_padded(rows, cols, out_dtype, x.device)
    _lib.check(lib.ed2_input_normalize(x.data_ptr(), _lib.dt_of(x), _lib.ld(x), y.data_ptr(), _lib.dt_of(y), _lib.ld(y),
                                       rows, cols, _lib.ptr(gamma), _lib.ptr(beta), eps, int(scale_only), scale,
                                       _lib.stream_ptr()), "ed2_input_normalize")
    return y


def layernorm_bwd(x, dy, gamma, eps, want_dx=True, want_params=True):
    lib = _lib.load()
    rows, cols = x.shape
    dx = _lib.empty_padded(rows, cols, x.dtype, x.device) if want_dx else None
    dgamma = torch.empty(cols, dtype=torch.float32, device=x.device) if want_params else None
    dbeta = torch.empty(cols, dtype=torch.float32, device=x.device) if want_params else None
    _lib.check(lib.ed2_layernorm_bwd(x.data_ptr(), _lib.dt_of(x), _lib.ld(x), dy.data_ptr(), _lib.dt_of(dy), _lib.ld(dy),
                                     rows, cols, _lib.ptr(gamma), eps, _lib.ptr(dx), _lib.dt_of(x),
                                     0 if dx is None else _lib.ld(dx), _lib.ptr(dgamma), _lib.ptr(dbeta),
                                     _lib.stream_ptr()), "ed2_layernorm_bwd")
    return dx, dgamma, dbeta


def rff_fwd(x: torch.Tensor, wt: torch.Tensor, bias: torch.Tensor, scale: float, phi_dtype: torch.dtype,
            want_pre: bool):
    """phi = scale * cos(x W + b); wt = Wᵀ [D, d] in x.dtype.  Returns (phi, pre or None)."""
    lib = _lib.load()
    B, d = x.shape
    D = wt.shape[0]
    phi = _lib.empty_padded(B, D, phi_dtype, x.device)
    pre = torch.empty(B, D, dtype=torch.float32, device=x.device) if want_pre else None
    _lib.check(lib.ed2_rff_fwd(_lib.dt_of(x), x.data_ptr(), _lib.ld(x), wt.data_ptr(), _lib.ld(wt), bias.data_ptr(),
                               scale, B, d, D, phi.data_ptr(), _lib.ld(phi), _lib.dt_of(phi), _lib.ptr(pre),
                               0 if pre is None else D, 0, _lib.stream_ptr()), "ed2_rff_fwd")
    return phi, pre


def rff_bwd(dphi: torch.Tensor, pre: torch.Tensor, bias: torch.Tensor, w: torch.Tensor, scale: float,
            dx_dtype: torch.dtype) -> torch.Tensor:
    """dx = (-scale * sin(pre + b) ∘ dphi) Wᵀ;  w [d, D] in the compute dtype."""
    lib = _lib.load()
    B, D = pre.shape
    d = w.shape[0]
    gpre = _lib.empty_padded(B, D, w.dtype, w.device)
    dx = _lib.empty_padded(B, d, dx_dtype, w.device)
    mode = 1 if pre.dtype == torch.bfloat16 else 0  # bf16: -scale*sin(xW+b) written by the forward (pre_mode 1)
    if mode == 1 and dphi.dtype != torch.bfloat16:
        dphi = cast(dphi, torch.bfloat16)
    _lib.check(lib.ed2_rff_bwd(_lib.dt_of(w), dphi.data_ptr(), _lib.ld(dphi), _lib.dt_of(dphi), pre.data_ptr(), _lib.ld(pre), mode,
                               bias.data_ptr(), w.data_ptr(), _lib.ld(w), scale, B, d, D, gpre.data_ptr(),
                               _lib.ld(gpre), dx.data_ptr(), _lib.ld(dx), _lib.dt_of(dx), _lib.stream_ptr()),
               "ed2_rff_bwd")
    return dx


def laplace_precision_update(phi: torch.Tensor, precision: torch.Tensor, momentum: float, batch_total: int | None = None,
                             partial_out: torch.Tensor | None = None) -> None:
    """P <- m P + (1-m) phiᵀphi / batch_total  (m > 0)  |  P + phiᵀphi; or the raw local product into partial_out."""
    lib = _lib.load()
    B, D = phi.shape
    assert precision.is_contiguous() and precision.dtype == torch.float32
    _lib.check(lib.ed2_laplace_precision_update(_lib.dt_of(phi), phi.data_ptr(), _lib.ld(phi), B, D, momentum,
                                                batch_total if batch_total is not None else B, precision.data_ptr(),
                                                _lib.ptr(partial_out), _lib.stream_ptr()), "ed2_laplace_precision_update")


def precision_blend(precision: torch.Tensor, summed: torch.Tensor, momentum: float, batch_total: int) -> None:
    lib = _lib.load()
    D = precision.shape[0]
    _lib.check(lib.ed2_precision_blend(precision.data_ptr(), summed.data_ptr(), D, momentum, batch_total,
                                       _lib.stream_ptr()), "ed2_precision_blend")


def predictive_variance(phi: torch.Tensor, sigma_lp: torch.Tensor, ridge: float, logits: torch.Tensor | None = None,
                        mean_field_factor: float = -1.0, likelihood: int = 0):
    """var[b] = ridge * phi[b] Sigma phi[b]ᵀ (+ optional fused mean-field logits). Returns (var, logits_out)."""
    lib = _lib.load()
    B, D = phi.shape
    var = torch.empty(B, dtype=torch.float32, device=phi.device)
    out = None
    K = 0
    if logits is not None:
        logits = logits.to(torch.float32).contiguous()
        K = logits.shape[1]
        out = torch.empty_like(logits)
    _lib.check(lib.ed2_predictive_variance(_lib.dt_of(phi), phi.data_ptr(), _lib.ld(phi), sigma_lp.data_ptr(),
                                           _lib.ld(sigma_lp), B, D, ridge, var.data_ptr(), _lib.ptr(logits), K,
                                           mean_field_factor, likelihood, _lib.ptr(out), _lib.stream_ptr()),
               "ed2_predictive_variance")
    return var, out


def mean_field_logits(logits: torch.Tensor, var: torch.Tensor, mean_field_factor: float, likelihood: int) -> torch.Tensor:
    lib = _lib.load()
    lg = logits.to(torch.float32).contiguous()
    B, K = lg.shape
    out = torch.empty_like(lg)
    _lib.check(lib.ed2_mean_field_logits(lg.data_ptr(), var.to(torch.float32).contiguous().data_ptr(), B, K,
                                         mean_field_factor, likelihood, out.data_ptr(), _lib.stream_ptr()),
               "ed2_mean_field_logits")
    return out


def split_bf16(src: torch.Tensor, is_b: bool, nsplit: int = 3) -> torch.Tensor:
    """[rows, cols] fp32 -> [rows, terms*cols] bf16 split operand (see ed2_split_bf16)."""
    lib = _lib.load()
    rows, cols = src.shape
    terms = 6 if nsplit == 3 else 3
    assert (terms * cols) % 8 == 0, "split operand pitch must be a multiple of 8 bf16"
    out = torch.empty(rows, terms * cols, dtype=torch.bfloat16, device=src.device)
    _lib.check(lib.ed2_split_bf16(src.data_ptr(), _lib.dt_of(src), _lib.ld(src), out.data_ptr(), terms * cols, rows, cols,
                                  int(is_b), nsplit, _lib.stream_ptr()), "ed2_split_bf16")
    return out


def layernorm_split(x: torch.Tensor, gamma=None, beta=None, eps=1e-3) -> torch.Tensor:
    """LayerNorm(x) written directly as the 2-term bf16 split A-operand [rows, 3 * cols] (ed2_layernorm_split)."""
    lib = _lib.load()
    rows, cols = x.shape
    out = torch.empty(rows, 3 * cols, dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.ed2_layernorm_split(x.data_ptr(), _lib.dt_of(x), _lib.ld(x), out.data_ptr(), 3 * cols, rows, cols,
                                       _lib.ptr(gamma), _lib.ptr(beta), eps, _lib.stream_ptr()), "ed2_layernorm_split")
    return out


def rff_fwd_split(x: torch.Tensor, wt_split: torch.Tensor, bias: torch.Tensor, scale: float, num_features: int,
                  phi_dtype: torch.dtype, want_pre: bool, nsplit: int = 3, presplit: bool = False):
    """phi = scale * cos(x W + b) with an fp32-grade phase on the bf16 tensor cores: x (fp32 or bf16, [B, d]) is split
    on the fly, wt_split = split_bf16(W^T, is_b=True) was made once (W is frozen)."""
    lib = _lib.load()
    xs = x if presplit else split_bf16(x, False, nsplit)
    B, Kp = xs.shape
    D = num_features
    phi = _lib.empty_padded(B, D, phi_dtype, xs.device)
    # the backward needs only -scale*sin(phase): bf16, written by the same epilogue (ed2_rff_fwd pre_mode 1)
    pre = _lib.empty_padded(B, D, torch.bfloat16, xs.device) if want_pre else None
    _lib.check(lib.ed2_rff_fwd(BF16, xs.data_ptr(), _lib.ld(xs), wt_split.data_ptr(), _lib.ld(wt_split), bias.data_ptr(),
                               scale, B, Kp, D, phi.data_ptr(), _lib.ld(phi), _lib.dt_of(phi), _lib.ptr(pre),
                               0 if pre is None else _lib.ld(pre), 1, _lib.stream_ptr()), "ed2_rff_fwd[split]")
    return phi, pre


def spd_inverse(a: torch.Tensor):
    """Inverse of a symmetric positive definite fp32 matrix on the GPU without a library call (ed2_spd_inverse: block
    Gauss-Jordan).  Returns (inverse, status): a new [n, n] tensor and a one-element int32 device tensor that is
    non-zero when a pivot was not positive (the input was not positive definite).  Reading the status synchronises, so
    it is left to the caller."""
    lib = _lib.load()
    n = a.shape[0]
    assert a.dim() == 2 and a.shape[1] == n and a.dtype == torch.float32 and a.is_cuda
    out = a.contiguous().clone()
    ws = torch.empty(int(lib.ed2_spd_inverse_workspace_floats(n)), dtype=torch.float32, device=a.device)
    _lib.check(lib.ed2_spd_inverse(out.data_ptr(), n, ws.data_ptr(), _lib.stream_ptr()), "ed2_spd_inverse")
    return out, ws[-4:].view(torch.int32)[0:1]
