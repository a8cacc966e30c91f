"""`edward2.jax.nn.utils` on libed2b200: make_sign_initializer (utils.py:28-51), mean_field_logits (:54-101)."""
from __future__ import annotations

import torch

from .. import ops
from .module import PRNG


def make_sign_initializer(random_sign_init: float):
    """> 0: random sign vector with P(+1) = random_sign_init; <= 0: Normal(1, -random_sign_init)."""
    if random_sign_init > 0:
        def initializer(key, shape, dtype=None):
            p = torch.full(tuple(shape), float(random_sign_init))
            return 2 * torch.bernoulli(p, generator=PRNG(key).generator()) - 1.0
        return initializer

    def initializer(key, shape, dtype=None):
        return torch.randn(tuple(shape), generator=PRNG(key).generator()) * (-random_sign_init) + 1.0
    return initializer


def mean_field_logits(logits, covmat=None, mean_field_factor=1.0, likelihood="logistic"):
    """logits [batch, classes]; covmat = the [batch] predictive VARIANCE vector (utils.py:54-101), or None."""
    if likelihood not in ("logistic", "binary_logistic", "poisson"):
        raise ValueError(
            f'Likelihood" must be one of (\'logistic\', \'binary_logistic\', \'poisson\'), got {likelihood}.')
    if mean_field_factor < 0:
        return logits
    squeeze = logits.dim() == 1
    lg = logits.reshape(-1, 1) if squeeze else logits
    variances = (torch.ones(lg.shape[0], dtype=torch.float32, device=lg.device) if covmat is None
                 else (torch.diagonal(covmat) if covmat.dim() == 2 else covmat))
    out = ops.mean_field_logits(lg, variances, float(mean_field_factor), 1 if likelihood == "poisson" else 0)
    return out.reshape(-1) if squeeze else out
