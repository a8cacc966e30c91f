"""`jax.lax` primitives used by the reference's nn files."""
import enum
import types

import numpy as _np


class Precision(enum.Enum):
    DEFAULT = 0
    HIGH = 1
    HIGHEST = 2


def dot_general(lhs, rhs, dimension_numbers, precision=None, preferred_element_type=None):
    (lc, rc), (lb, rb) = dimension_numbers
    assert tuple(lb) == () and tuple(rb) == (), "stub: batch dimensions not needed by the reference call sites"
    return _np.tensordot(_np.asarray(lhs, dtype=_np.float64), _np.asarray(rhs, dtype=_np.float64), axes=(tuple(lc), tuple(rc)))


def rsqrt(x):
    return 1.0 / _np.sqrt(x)


def square(x):
    return _np.square(x)


def real(x):
    return _np.real(x)


def imag(x):
    return _np.imag(x)


def stop_gradient(x):
    return x


def pmean(x, axis_name=None, axis_index_groups=None):
    return x


def scan(f, init, xs, length=None):
    carry, ys = init, []
    n = length if xs is None else len(xs)
    for i in range(n):
        carry, y = f(carry, None if xs is None else xs[i])
        ys.append(y)
    return carry, (None if all(y is None for y in ys) else _np.stack(ys))


def _cholesky(a, symmetrize_input=True):
    a = _np.asarray(a, dtype=_np.float64)
    if symmetrize_input:
        a = (a + a.T) / 2
    return _np.linalg.cholesky(a)


def _triangular_solve(a, b, left_side=False, lower=False, transpose_a=False, conjugate_a=False, unit_diagonal=False):
    import scipy.linalg

    a, b = _np.asarray(a, dtype=_np.float64), _np.asarray(b, dtype=_np.float64)
    assert not unit_diagonal
    if left_side:   # solve op(a) x = b
        return scipy.linalg.solve_triangular(a, b, lower=lower, trans=1 if transpose_a else 0)
    # solve x op(a) = b  <=>  op(a)^T x^T = b^T
    return scipy.linalg.solve_triangular(a, b.T, lower=lower, trans=0 if transpose_a else 1).T


linalg = types.SimpleNamespace(cholesky=_cholesky, triangular_solve=_triangular_solve)
