"""Regularizers of the hot path: `NormalKLDivergence` (edward2/tensorflow/regularizers.py:166-186) — the analytic KL
regularizer, evaluated by the warp-shuffle reduction kernel (or fused into the weight-sampling pass by the layers)."""
from __future__ import annotations

import torch

from . import functional as F


class NormalKLDivergence:
    """KL divergence regularizer from an input to the normal distribution: scale_factor * KL(q || N(mean, stddev))."""

    def __init__(self, mean=0.0, stddev=1.0, scale_factor=1.0):
        self.mean, self.stddev, self.scale_factor = float(mean), float(stddev), float(scale_factor)

    def __call__(self, x) -> torch.Tensor:
        """x: a trainable-normal initializer (owning mean / stddev(=rho) parameters) or a (mean, rho) pair.
        A deterministic x contributes KL of a point mass only through its location, which the reference
        evaluates as a Normal with scale 0 is undefined — there it raises; here too."""
        if hasattr(x, "params"):
            mean, rho = x.params()
        else:
            mean, rho = x
        if rho is None:
            raise ValueError("NormalKLDivergence needs a Normal posterior (trainable_normal initializer)")
        return F.NormalKL.apply(mean, rho, self.mean, self.stddev, self.scale_factor)

    def get_config(self):
        return {"mean": self.mean, "stddev": self.stddev, "scale_factor": self.scale_factor}


class L2:
    """tf.keras.regularizers.l2 used by the GP output layer (random_feature.py:191)."""

    def __init__(self, l2=0.01):
        self.l2 = float(l2)

    def __call__(self, w):
        return self.l2 * (w * w).sum()

    def get_config(self):
        return {"l2": self.l2}


def get(identifier):
    """edward2/tensorflow/regularizers.py:403-422."""
    if identifier is None:
        return None
    if isinstance(identifier, str):
        key = identifier.lower()
        if key in ("normal_kl_divergence", "normalkldivergence"):
            return NormalKLDivergence()
        raise ValueError(f"Unknown regularizer: {identifier} (the hot path implements 'normal_kl_divergence')")
    if callable(identifier):
        return identifier
    raise ValueError(f"Could not interpret regularizer identifier: {identifier}")


def serialize(r):
    return None if r is None else {"class_name": type(r).__name__, "config": r.get_config() if hasattr(r, "get_config") else {}}
