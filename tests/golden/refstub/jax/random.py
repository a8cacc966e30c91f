"""`jax.random` on NumPy generators.  The streams are NOT threefry: values drawn here are stored in the fixture and
injected wherever they are needed again."""
import numpy as _np


def PRNGKey(seed):  # noqa: N802
    return _np.array([0, int(seed) & 0xFFFFFFFF], dtype=_np.uint32)


def _gen(key):
    return _np.random.default_rng([int(k) for k in _np.asarray(key).reshape(-1)])


def split(key, num=2):
    g = _gen(key)
    return [g.integers(0, 2 ** 32, size=2, dtype=_np.uint32) for _ in range(num)]


def fold_in(key, data):
    return _np.array([int(_np.asarray(key).reshape(-1)[-1]) ^ 0x9E3779B9, int(data) & 0xFFFFFFFF], dtype=_np.uint32)


def normal(key, shape=(), dtype=None):
    return _gen(key).standard_normal(tuple(shape))


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    return _gen(key).uniform(minval, maxval, tuple(shape))


def bernoulli(key, p=0.5, shape=None):
    return _gen(key).uniform(size=tuple(shape)) < p


def truncated_normal(key, lower, upper, shape=(), dtype=None):
    g = _gen(key)
    z = g.standard_normal(tuple(shape))
    while (bad := (z < lower) | (z > upper)).any():
        z[bad] = g.standard_normal(int(bad.sum()))
    return z
