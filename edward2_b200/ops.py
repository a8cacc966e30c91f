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
         row_sum=None, want_out=True, block_n=0, cta_group=0, max_ctas=0, no_tma_store=0) -> torch.Tensor | None:
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
    _lib.check(lib.ed2_gemm(C.byref(d), _lib.stream_ptr()), "ed2_gemm")
    return out


def philox(kind: str, n: int, seed: int, stream: int, device="cuda") -> torch.Tensor:
    lib = _lib.load()
    if kind == "raw":
        out = torch.empty(n * 4, dtype=torch.int32, device=device)
        _lib.check(lib.ed2_philox_raw(out.data_ptr(), n, seed, stream, _lib.stream_ptr()), "philox_raw")
        return out
    out = torch.empty(n, dtype=torch.float32, device=device)
    fn = {"normal": lib.ed2_philox_normal, "uniform": lib.ed2_philox_uniform, "sign": lib.ed2_philox_sign}[kind]
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
