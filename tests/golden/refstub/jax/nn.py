"""`jax.nn` activations used by the reference's nn files."""
import numpy as _np


def sigmoid(x):
    return 1.0 / (1.0 + _np.exp(-_np.asarray(x, dtype=_np.float64)))


def softplus(x):
    return _np.logaddexp(0.0, _np.asarray(x, dtype=_np.float64))


def relu(x):
    return _np.maximum(x, 0.0)


def softmax(x, axis=-1):
    x = _np.asarray(x, dtype=_np.float64)
    x = x - x.max(axis=axis, keepdims=True)
    e = _np.exp(x)
    return e / e.sum(axis=axis, keepdims=True)
