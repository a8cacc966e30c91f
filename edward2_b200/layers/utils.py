"""`ed.layers.utils` members on the hot path: mean_field_logits (edward2/tensorflow/layers/utils.py:379-424)."""
from __future__ import annotations

import torch

from .. import ops


def mean_field_logits(logits, covmat=None, mean_field_factor=1.0, likelihood="logistic"):
    """Adjust the model logits so its softmax approximates the posterior mean.

    logits [batch, classes]; covmat [batch, batch] (its diagonal is used) or None (identity).
    """
    if likelihood not in ("logistic", "binary_logistic", "poisson"):
        raise ValueError(
            f'Likelihood" must be one of (\'logistic\', \'binary_logistic\', \'poisson\'), got {likelihood}.')
    if mean_field_factor < 0:
        return logits
    squeeze = logits.dim() == 1
    lg = logits.reshape(-1, 1) if squeeze else logits
    if covmat is None:
        variances = torch.ones(lg.shape[0], dtype=torch.float32, device=lg.device)
    else:
        variances = torch.diagonal(covmat) if covmat.dim() == 2 else covmat
    out = ops.mean_field_logits(lg, variances, float(mean_field_factor), 1 if likelihood == "poisson" else 0)
    return out.reshape(-1) if squeeze else out
