"""On-GPU covariance inverse (include/ed2b200.h: ed2_spd_inverse — block Gauss-Jordan, no library call) against a float64
NumPy inverse: Sigma = (ridge I + precision)^-1 of edward2/tensorflow/layers/random_feature.py:447-459 /
edward2/jax/nn/random_feature.py:393-404."""
import numpy as np
import pytest
import torch

from tests.util import rel_err, to_np

pytestmark = pytest.mark.gpu


def _spd(n, ridge, rng, rows=None):
    phi = rng.standard_normal((rows or 2 * n, n)) * np.sqrt(2.0 / n)
    return ridge * np.eye(n) + phi.T @ phi


@pytest.mark.parametrize("n,ridge", [(48, 1.0), (64, 1.0), (200, 1e-1), (1024, 1.0), (1000, 1e-2)])
def test_spd_inverse_matches_float64(cuda, n, ridge):
    from edward2_b200 import ops

    rng = np.random.default_rng(n)
    a = _spd(n, ridge, rng).astype(np.float32)
    x, status = ops.spd_inverse(torch.tensor(a, device=cuda))
    ref = np.linalg.inv(a.astype(np.float64))
    cond = np.linalg.cond(a.astype(np.float64))
    assert int(status.item()) == 0
    err = rel_err(to_np(x), ref)
    assert err < max(2e-5, 2e-6 * cond), (err, cond)            # fp32: ~cond * eps
    xa = to_np(x) @ a.astype(np.float64)
    assert np.abs(xa - np.eye(n)).max() < max(1e-4, 1e-5 * cond)
    assert rel_err(to_np(x), to_np(x).T) < 1e-4                # the inverse of a symmetric matrix stays symmetric


def test_spd_inverse_flags_indefinite_input(cuda):
    from edward2_b200 import ops

    a = torch.eye(96, device=cuda)
    a[70, 70] = -1.0
    _, status = ops.spd_inverse(a)
    assert int(status.item()) != 0


def test_laplace_covariance_uses_the_gpu_inverse(cuda):
    """LaplaceRandomFeatureCovariance: the covariance matrix at evaluation is (ridge I + P)^-1 from ed2_spd_inverse."""
    import edward2_b200 as ed

    D, B = 256, 512
    rng = np.random.default_rng(5)
    phi = torch.tensor(rng.standard_normal((B, D)) * 0.2, dtype=torch.float32, device=cuda)
    cov = ed.layers.LaplaceRandomFeatureCovariance(momentum=-1.0, ridge_penalty=0.5, dtype="float32")
    cov(phi, None, training=True)
    cov(phi, None, training=False)
    p = to_np(cov.precision_matrix)
    ref = np.linalg.inv(0.5 * np.eye(D) + p)
    assert rel_err(to_np(cov.covariance_matrix), ref) < 5e-5
