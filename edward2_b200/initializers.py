"""Initializers of the hot path (host-side, one-off): the parameterisation contract of the reference's
`TrainableNormal`, `TrainableDeterministic`, `RandomSign`, `OrthogonalRandomFeatures` plus the Keras string aliases
the layers accept (edward2/tensorflow/initializers.py:176-225, 454-528, 667-694, 759-830, 866-885).

A "trainable initializer" owns the variational parameters of a weight: calling it returns the distribution
parameters `(mean, rho)` with sigma = softplus(rho) + 1e-7 (constraints.py:56-63); the layers hand those to the
fused sampling kernels instead of materialising a RandomVariable.
"""
from __future__ import annotations

import math

import torch
from torch import nn


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed) if seed is not None else torch.seed() % (2**31))
    return g


def _fans(shape):
    if len(shape) < 1:
        return 1, 1
    if len(shape) == 1:
        return shape[0], shape[0]
    receptive = 1
    for s in shape[:-2]:
        receptive *= s
    return shape[-2] * receptive, shape[-1] * receptive


def _truncated_normal(shape, mean, stddev, gen):
    t = torch.empty(tuple(shape), dtype=torch.float32)
    nn.init.trunc_normal_(t, mean=mean, std=stddev, a=mean - 2 * stddev, b=mean + 2 * stddev, generator=gen)
    return t


class Initializer:
    """Plain (non-trainable) initializer: __call__(shape) -> fp32 CPU tensor."""

    def __call__(self, shape, dtype=None):
        raise NotImplementedError

    def get_config(self):
        return {}


class Zeros(Initializer):
    def __call__(self, shape, dtype=None):
        return torch.zeros(tuple(shape))


class Ones(Initializer):
    def __call__(self, shape, dtype=None):
        return torch.ones(tuple(shape))


class Constant(Initializer):
    def __init__(self, value=0.0):
        self.value = value

    def __call__(self, shape, dtype=None):
        return torch.full(tuple(shape), float(self.value))

    def get_config(self):
        return {"value": self.value}


class RandomNormal(Initializer):
    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev, self.seed = mean, stddev, seed

    def __call__(self, shape, dtype=None):
        return torch.randn(tuple(shape), generator=_gen(self.seed)) * self.stddev + self.mean

    def get_config(self):
        return {"mean": self.mean, "stddev": self.stddev, "seed": self.seed}


class RandomUniform(Initializer):
    def __init__(self, minval=-0.05, maxval=0.05, seed=None):
        self.minval, self.maxval, self.seed = minval, maxval, seed

    def __call__(self, shape, dtype=None):
        return torch.rand(tuple(shape), generator=_gen(self.seed)) * (self.maxval - self.minval) + self.minval

    def get_config(self):
        return {"minval": self.minval, "maxval": self.maxval, "seed": self.seed}


class TruncatedNormal(Initializer):
    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev, self.seed = mean, stddev, seed

    def __call__(self, shape, dtype=None):
        return _truncated_normal(shape, self.mean, self.stddev, _gen(self.seed))

    def get_config(self):
        return {"mean": self.mean, "stddev": self.stddev, "seed": self.seed}


class GlorotUniform(Initializer):
    def __init__(self, seed=None):
        self.seed = seed

    def __call__(self, shape, dtype=None):
        fi, fo = _fans(shape)
        lim = math.sqrt(6.0 / (fi + fo))
        return (torch.rand(tuple(shape), generator=_gen(self.seed)) * 2 - 1) * lim


class HeNormal(Initializer):
    def __init__(self, seed=None):
        self.seed = seed

    def __call__(self, shape, dtype=None):
        fi, _ = _fans(shape)
        std = math.sqrt(2.0 / fi) / 0.87962566103423978  # Keras truncated-normal variance correction
        return _truncated_normal(shape, 0.0, std, _gen(self.seed))


class GlorotNormal(Initializer):
    """tf.keras.initializers.GlorotNormal: truncated normal with stddev sqrt(2 / (fan_in + fan_out)) (variance-
    corrected for the +-2 sigma truncation, as Keras does)."""

    def __init__(self, seed=None):
        self.seed = seed

    def __call__(self, shape, dtype=None):
        fi, fo = _fans(shape)
        std = math.sqrt(2.0 / (fi + fo)) / 0.87962566103423978
        return _truncated_normal(shape, 0.0, std, _gen(self.seed))


class RandomSign(Initializer):
    """+-1 with P(+1) = probs (edward2/tensorflow/initializers.py:667-694; jax/nn/utils.py:28-51)."""

    def __init__(self, probs=1.0, seed=None, dtype=None):
        self.probs, self.seed = probs, seed

    def __call__(self, shape, dtype=None):
        b = torch.bernoulli(torch.full(tuple(shape), float(self.probs)), generator=_gen(self.seed))
        return 2.0 * b - 1.0

    def get_config(self):
        return {"probs": self.probs, "seed": self.seed}


class Orthogonal(Initializer):
    def __init__(self, gain=1.0, seed=None):
        self.gain, self.seed = gain, seed

    def __call__(self, shape, dtype=None):
        rows, cols = shape
        a = torch.randn(max(rows, cols), min(rows, cols), generator=_gen(self.seed), dtype=torch.float64)
        q, r = torch.linalg.qr(a)
        q = q * torch.sign(torch.diagonal(r))
        if rows < cols:
            q = q.t()
        return (self.gain * q[:rows, :cols]).to(torch.float32)


class OrthogonalRandomFeatures(Initializer):
    """W = stddev * Q @ S (edward2/tensorflow/initializers.py:759-830): stacked square orthogonal blocks when
    rows < cols (:788-803), chi-like column norms from squared normals (:806-821) or sqrt(rows) if random_norm=False."""

    def __init__(self, stddev=1.0, random_norm=True, seed=None):
        self.stddev, self.random_norm, self.seed = stddev, random_norm, seed

    def __call__(self, shape, dtype=None):
        rows, cols = shape
        gen = _gen(self.seed)

        def ortho(r, c):
            a = torch.randn(max(r, c), min(r, c), generator=gen, dtype=torch.float64)
            q, rr = torch.linalg.qr(a)
            q = q * torch.sign(torch.diagonal(rr))
            if r < c:
                q = q.t()
            return self.stddev * q[:r, :c]

        if rows < cols:
            blocks, n = [], 0
            while n < cols:
                blocks.append(ortho(rows, rows))
                n += rows
            mat = torch.cat(blocks, dim=-1)[:, :cols]
        else:
            mat = ortho(rows, cols)
        if self.random_norm:
            norms_sq = torch.randn(mat.shape, generator=gen, dtype=torch.float64) ** 2
        else:
            norms_sq = torch.ones(mat.shape, dtype=torch.float64)
        norms = torch.sqrt(norms_sq.sum(0))
        return (mat * norms).to(torch.float32)

    def get_config(self):
        return {"stddev": self.stddev, "random_norm": self.random_norm, "seed": self.seed}


# --------------------------------------------------------------------------------------------- trainable initializers
class TrainableInitializer(nn.Module):
    """Base of initializers that own parameters (the reference makes them Keras Layers)."""

    built = False
    is_stochastic = False

    def build(self, shape, device=None):
        raise NotImplementedError


class TrainableDeterministic(TrainableInitializer):
    """Point-mass 'distribution' with a trainable location (initializers.py:176-225)."""

    def __init__(self, loc_initializer=None, seed=None, **kwargs):
        super().__init__()
        self.loc_initializer = get(loc_initializer if loc_initializer is not None else TruncatedNormal(stddev=1e-5))
        self.seed = seed

    def build(self, shape, device=None):
        self.mean = nn.Parameter(self.loc_initializer(shape).to(device))
        self.built = True

    def params(self):
        return self.mean, None


class TrainableNormal(TrainableInitializer):
    """q = Normal(mean, softplus(stddev) + 1e-7) with trainable mean / stddev(=rho)
    (initializers.py:454-528): mean ~ TruncN(0, 1e-5), rho ~ TruncN(-3, 0.1) by default."""

    is_stochastic = True

    def __init__(self, mean_initializer=None, stddev_initializer=None, mean_regularizer=None, stddev_regularizer=None,
                 mean_constraint=None, stddev_constraint="softplus", seed=None, **kwargs):
        super().__init__()
        self.mean_initializer = get(mean_initializer if mean_initializer is not None else TruncatedNormal(stddev=1e-5))
        self.stddev_initializer = get(stddev_initializer if stddev_initializer is not None
                                      else TruncatedNormal(mean=-3.0, stddev=0.1))
        if stddev_constraint not in ("softplus", None) and type(stddev_constraint).__name__ != "Softplus":
            raise ValueError("libed2b200 fuses the Softplus stddev constraint only")
        self.seed = seed

    def build(self, shape, device=None):
        self.mean = nn.Parameter(self.mean_initializer(shape).to(device))
        self.stddev = nn.Parameter(self.stddev_initializer(shape).to(device))
        self.built = True

    def params(self):
        return self.mean, self.stddev


class TrainableHeNormal(TrainableNormal):
    """initializers.py:531-560: mean ~ he_normal(seed); stddev keeps the TrainableNormal default TruncN(-3, 0.1)."""

    def __init__(self, seed=None, **kwargs):
        kwargs.pop("mean_initializer", None)
        super().__init__(mean_initializer=HeNormal(seed), seed=seed, **kwargs)

    def get_config(self):
        return {"seed": self.seed}


class TrainableGlorotNormal(TrainableNormal):
    """initializers.py:563-590: mean ~ GlorotNormal(seed); stddev keeps the TrainableNormal default."""

    def __init__(self, seed=None, **kwargs):
        kwargs.pop("mean_initializer", None)
        super().__init__(mean_initializer=GlorotNormal(seed), seed=seed, **kwargs)

    def get_config(self):
        return {"seed": self.seed}


def _softplus_inverse(y):
    return math.log(math.expm1(y))


_ALIASES = {
    "zeros": Zeros, "zero": Zeros, "ones": Ones, "one": Ones,
    "glorot_uniform": GlorotUniform, "glorot_normal": GlorotNormal, "he_normal": HeNormal,
    "random_normal": RandomNormal, "normal": RandomNormal, "random_uniform": RandomUniform,
    "truncated_normal": TruncatedNormal, "orthogonal": Orthogonal,
    "random_sign": RandomSign, "orthogonal_random_features": OrthogonalRandomFeatures,
    "trainable_normal": TrainableNormal, "trainable_deterministic": TrainableDeterministic,
    "trainable_he_normal": TrainableHeNormal, "trainable_glorot_normal": TrainableGlorotNormal,
}


def get(identifier):
    """String / instance / class lookup (edward2/tensorflow/initializers.py:866-885)."""
    if identifier is None:
        return None
    if isinstance(identifier, (Initializer, TrainableInitializer)):
        return identifier
    if isinstance(identifier, str):
        key = identifier.lower()
        if key not in _ALIASES:
            raise ValueError(f"Unknown initializer: {identifier}")
        return _ALIASES[key]()
    if isinstance(identifier, type):
        return identifier()
    if callable(identifier):
        return identifier
    raise ValueError(f"Could not interpret initializer identifier: {identifier}")


def serialize(initializer):
    if initializer is None:
        return None
    cfg = initializer.get_config() if hasattr(initializer, "get_config") else {}
    return {"class_name": type(initializer).__name__, "config": cfg}
