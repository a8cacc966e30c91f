"""Constraints of the hot path: only `Softplus` (edward2/tensorflow/constraints.py:56-66) — the formula is fused in
the sampling / KL kernels (sigma = softplus(rho) + epsilon), this module keeps the lookup contract."""
from __future__ import annotations

import torch

EPSILON = 1e-7  # tf.keras.backend.epsilon()


class Softplus:
    """Softplus constraint: w -> softplus(w) + epsilon."""

    def __init__(self, epsilon=EPSILON):
        self.epsilon = epsilon

    def __call__(self, w: torch.Tensor) -> torch.Tensor:
        return torch.nn.functional.softplus(w) + self.epsilon

    def get_config(self):
        return {"epsilon": self.epsilon}


def get(identifier):
    """edward2/tensorflow/constraints.py:93-112."""
    if identifier is None:
        return None
    if isinstance(identifier, str):
        if identifier.lower() == "softplus":
            return Softplus()
        raise ValueError(f"Unknown constraint: {identifier} (the hot path fuses 'softplus' only)")
    if callable(identifier):
        return identifier
    raise ValueError(f"Could not interpret constraint identifier: {identifier}")


def serialize(c):
    return None if c is None else {"class_name": type(c).__name__, "config": c.get_config()}
