"""NumPy-backed stand-in for the slice of `jax` the reference's edward2/jax/nn files use (see ../README.md)."""
import numpy as _np

from . import lax, nn, numpy, random  # noqa: F401


class Array:  # scipy's array-API shim probes sys.modules['jax'].Array
    pass


class ShapeDtypeStruct:
    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(int(s) for s in shape), dtype


def eval_shape(fn, *specs):
    out = fn(*[_np.zeros(s.shape, dtype=_np.float64) for s in specs])
    return ShapeDtypeStruct(out.shape, out.dtype)


def vjp(fn, primal):
    """(fn(primal), vjp_fn) for a LINEAR `fn` (the only use in the reference: the kernel operator of spectral
    normalisation, normalization.py:160).  The Jacobian is assembled column by column from basis vectors."""
    primal = _np.asarray(primal, dtype=_np.float64)
    out = _np.asarray(fn(primal))
    jac = _np.empty((out.size, primal.size), dtype=_np.float64)
    for i in range(primal.size):
        e = _np.zeros(primal.size)
        e[i] = 1.0
        jac[:, i] = _np.asarray(fn(e.reshape(primal.shape))).reshape(-1)
    # linearity check: fn(primal) must equal J @ primal
    assert _np.allclose(jac @ primal.reshape(-1), out.reshape(-1), rtol=1e-10, atol=1e-12), "vjp stub: fn is not linear"

    def vjp_fn(cotangent):
        return ((jac.T @ _np.asarray(cotangent, dtype=_np.float64).reshape(-1)).reshape(primal.shape),)

    return out, vjp_fn
