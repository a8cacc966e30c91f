"""`edward2.jax.nn.DenseBatchEnsemble` on libed2b200 (edward2/jax/nn/dense.py:215-272):
einsum('E...C,EC,CD,ED->E...D') + bias + activation as ONE fused tcgen05 GEMM."""
from __future__ import annotations

from typing import Any, Callable, Optional

import torch

from .. import functional as F, ops
from ..layers.base import resolve_dtype
from . import module as m


class DenseBatchEnsemble(m.Module):
    features: int
    ens_size: int
    activation: Optional[Callable] = None
    use_bias: bool = True
    dtype: Optional[Any] = None
    alpha_init: Callable = m.ones
    gamma_init: Callable = m.ones
    kernel_init: Callable = m.lecun_normal()
    bias_init: Callable = m.zeros

    def __call__(self, inputs):
        """inputs [ens_size * batch_size, ..., input_dim] -> [ens_size * batch_size, ..., features]."""
        dtype = resolve_dtype(self.dtype) if self.dtype is not None else (
            inputs.dtype if inputs.dtype in (torch.float32, torch.bfloat16) else torch.float32)
        input_dim = inputs.shape[-1]
        kernel = self.param("kernel", self.kernel_init, (input_dim, self.features), dtype)
        alpha = self.param("fast_weight_alpha", self.alpha_init, (self.ens_size, input_dim), dtype)
        gamma = self.param("fast_weight_gamma", self.gamma_init, (self.ens_size, self.features), dtype)
        bias = self.param("bias", self.bias_init, (self.ens_size, self.features), dtype) if self.use_bias else None
        if inputs.shape[0] % self.ens_size != 0:
            raise ValueError(f"leading dimension {inputs.shape[0]} is not divisible by ens_size {self.ens_size}")
        x2 = inputs.reshape(-1, input_dim)  # member-major rows are preserved by the flattening
        y = F.BatchEnsembleDense.apply(x2, kernel, alpha, gamma, bias, ops.act_id(self.activation), self.ens_size, dtype)
        return y.reshape(*inputs.shape[:-1], self.features)
