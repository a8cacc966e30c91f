"""`jax.numpy` as float64 NumPy: every floating dtype the reference names maps to float64 so that the reference's own
statements are evaluated in double precision (the fixture is compared with the float64 oracle)."""
import numpy as _np
from numpy import *  # noqa: F401,F403

ndarray = _np.ndarray
float32 = _np.float64
float64 = _np.float64
bfloat16 = _np.float64
float16 = _np.float64
pi = _np.pi


def asarray(a, dtype=None):
    a = _np.asarray(a)
    if dtype is not None or a.dtype.kind == "f":
        return a.astype(_np.float64) if (dtype is None or _np.dtype(dtype).kind == "f") else a.astype(dtype)
    return a


array = asarray


def vdot(a, b):
    return _np.vdot(_np.asarray(a), _np.asarray(b))
