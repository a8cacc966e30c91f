"""`edward2.jax.nn.SpectralNormalization` on libed2b200 (edward2/jax/nn/normalization.py:54-185).

Functional semantics: the wrapped layer's parameters live under this module (`params[<LayerName>]`), `u` / `v` in the
`spectral_stats` collection, updated only when training and that collection is mutable; the normalised kernel is
w / stop_gradient(max(sigma / c, 1)) (:175-178).  Power iteration = the bandwidth-bound matvec kernels of sngp.cu.
"""
from __future__ import annotations

from typing import Any, Callable, Optional

import torch

from .. import functional as F, ops
from ..layers.base import resolve_dtype
from . import module as m


class Dense(m.Module):
    """flax.linen.Dense (the layer SpectralNormalization wraps in the reference's tests, normalization_test.py)."""
    features: int
    use_bias: bool = True
    dtype: Optional[Any] = None
    kernel_init: Callable = m.lecun_normal()
    bias_init: Callable = m.zeros
    activation: Optional[Callable] = None

    def __call__(self, inputs, _w_override=None, _w_scale=None):
        dtype = resolve_dtype(self.dtype) if self.dtype is not None else (
            inputs.dtype if inputs.dtype in (torch.float32, torch.bfloat16) else torch.float32)
        kernel = self.param("kernel", self.kernel_init, (inputs.shape[-1], self.features), dtype)
        bias = self.param("bias", self.bias_init, (self.features,), dtype) if self.use_bias else None
        x2 = inputs.reshape(-1, inputs.shape[-1])
        w_eff = kernel if _w_override is None else _SubstituteValue.apply(kernel, _w_override)
        y = F.Dense.apply(x2, w_eff, bias, ops.act_id(self.activation), dtype, _w_scale)
        return y.reshape(*inputs.shape[:-1], self.features)


class _SubstituteValue(torch.autograd.Function):
    """Forward value = `value` (the normalised kernel computed by the CUDA kernel), gradient passes to `param`
    unchanged; the stop-gradient scale is applied by F.Dense's w_scale."""

    @staticmethod
    def forward(ctx, param, value):
        return value.view_as(param)

    @staticmethod
    def backward(ctx, g):
        return g, None


class SpectralNormalization(m.Module):
    layer: Any
    iteration: int = 1
    norm_multiplier: float = 0.95
    u_init: Callable = m.normal(stddev=0.05)
    v_init: Callable = m.normal(stddev=0.05)
    kernel_apply_kwargs: Optional[Any] = None
    kernel_name: str = "kernel"
    layer_name: Optional[str] = None

    def __call__(self, inputs, training: bool = True):
        if self.kernel_apply_kwargs is not None:
            raise NotImplementedError("custom kernel_apply_kwargs (conv interpretation) is outside the hot path")
        layer_name = type(self.layer).__name__ if self.layer_name is None else self.layer_name
        child = self.sub(self.layer, layer_name)
        initializing = self.is_initializing()
        if initializing:
            child(inputs)  # creates params[layer_name]
        w = self._scope.collection("params")[layer_name][self.kernel_name]
        in_dim, out_dim = w.shape[0], w.shape[-1]
        u = self.variable("spectral_stats", "u", lambda: self.u_init(self.make_rng("u"), (out_dim,)))
        v = self.variable("spectral_stats", "v", lambda: self.v_init(self.make_rng("v"), (in_dim,)))
        update = bool(training) and not initializing
        with torch.no_grad():
            u_new, v_new = u.value.clone(), v.value.clone()
            w_hat = torch.empty_like(w)
            _, factor = ops.spectral_norm_update(w.detach().contiguous(), u_new, v_new, self.iteration,
                                                 self.norm_multiplier, 1, update, w_hat)
        if update and self.is_mutable_collection("spectral_stats"):
            u.value, v.value = u_new, v_new
        self.sow("intermediates", "w", w_hat)
        return child(inputs, _w_override=w_hat, _w_scale=factor.detach())
