"""GPU tests of the Flax-signature mirrors (`edward2_b200.nn`): the assertions of edward2/jax/nn/dense_test.py,
random_feature_test.py, normalization_test.py and utils_test.py restated."""
import numpy as np
import pytest
import torch

from oracle import dense as od, sngp as og
from tests.util import rel_err, to_np

pytestmark = pytest.mark.gpu


def _ed():
    import edward2_b200 as ed

    return ed


@pytest.mark.parametrize("inputs_shape", [(8, 5), (8, 3, 5), (8, 2, 3, 5)])
@pytest.mark.parametrize("random_sign_init", [0.5, -0.5])
def test_dense_batch_ensemble_matches_loop(cuda, inputs_shape, random_sign_init):
    """jax dense_test.py:34-71: vectorised == per-member loop, rtol = atol = 1e-6, 2-D / 3-D / 4-D inputs, fast weights
    initialised to random signs or N(1, 0.5)."""
    ed = _ed()
    ens, feats = 4, 6
    init = ed.nn.make_sign_initializer(random_sign_init)
    layer = ed.nn.DenseBatchEnsemble(features=feats, ens_size=ens, alpha_init=init, gamma_init=init,
                                     bias_init=ed.nn.module.normal(0.1))
    rng = np.random.default_rng(0)
    x = torch.tensor(rng.standard_normal(inputs_shape), dtype=torch.float32, device=cuda)
    variables = layer.init(0, x)
    assert set(variables["params"]) == {"kernel", "fast_weight_alpha", "fast_weight_gamma", "bias"}
    out = layer.apply(variables, x)
    assert out.shape == inputs_shape[:-1] + (feats,)
    p = {k: to_np(v) for k, v in variables["params"].items()}
    xe = to_np(x).reshape((ens, -1) + inputs_shape[1:])
    loop = np.stack([((xe[e] * p["fast_weight_alpha"][e]) @ p["kernel"]) * p["fast_weight_gamma"][e] + p["bias"][e]
                     for e in range(ens)]).reshape(inputs_shape[:-1] + (feats,))
    np.testing.assert_allclose(to_np(out), loop, rtol=1e-5, atol=1e-5)
    # gradients flow to every parameter
    out.sum().backward()
    assert all(v.grad is not None for v in variables["params"].values())


def test_rfgp_collections_and_posterior(cuda):
    """jax random_feature_test.py:208-222 (collections), :417-452 (exact precision update), :346 (None while updating),
    :502-528 (diag == diag(full)) and the posterior vs the float64 oracle."""
    ed = _ed()
    rng = np.random.default_rng(1)
    B, d, D, K, ridge = 48, 12, 64, 3, 0.7
    x = torch.tensor(rng.standard_normal((B, d)), dtype=torch.float32, device=cuda)
    gp = ed.nn.RandomFeatureGaussianProcess(features=K, hidden_features=D, normalize_input=True,
                                            hidden_kwargs=dict(feature_scale=None),
                                            covmat_kwargs=dict(ridge_penalty=ridge))
    variables = gp.init(0, x)
    assert set(variables) == {"params", "random_features", "laplace_covariance"}
    assert set(variables["params"]) == {"output_layer"}
    assert set(variables["random_features"]["hidden_layer"]) == {"kernel", "bias"}
    p0 = to_np(variables["laplace_covariance"]["covmat_layer"]["precision_matrix"])
    np.testing.assert_allclose(p0, ridge * np.eye(D))
    # training: laplace_covariance mutable -> covmat is None and the precision is updated exactly
    (logits, covmat), mutated = gp.apply(variables, x, mutable=["laplace_covariance"])
    assert covmat is None
    w = to_np(variables["random_features"]["hidden_layer"]["kernel"])
    b = to_np(variables["random_features"]["hidden_layer"]["bias"])
    phi, _ = og.rff(og.layer_norm(to_np(x), eps=1e-6), w, b, np.sqrt(2.0 / D))
    p1 = to_np(mutated["laplace_covariance"]["covmat_layer"]["precision_matrix"])
    assert rel_err(p1, ridge * np.eye(D) + phi.T @ phi) < 1e-5
    beta = to_np(variables["params"]["output_layer"]["kernel"])
    bias = to_np(variables["params"]["output_layer"]["bias"])
    assert rel_err(to_np(logits), og.gp_output(phi, beta, bias)) < 1e-5
    # inference: diagonal (default) and full covariance
    new_vars = dict(variables, laplace_covariance=mutated["laplace_covariance"])
    _, var = gp.apply(new_vars, x)
    _, full = gp.apply(new_vars, x, return_full_covmat=True)
    ref_full = og.predictive_covariance_jax(phi, ridge * np.eye(D) + phi.T @ phi, ridge, diagonal_only=False)
    assert var.shape == (B,) and full.shape == (B, B)
    assert rel_err(to_np(full), ref_full) < 1e-4
    assert rel_err(to_np(var), np.diag(ref_full)) < 1e-4
    adj = ed.nn.mean_field_logits(logits, var, mean_field_factor=2.0)
    assert rel_err(to_np(adj), og.mean_field_logits(to_np(logits), to_np(var), 2.0)) < 1e-5


def test_laplace_momentum_and_errors(cuda):
    """jax random_feature_test.py:464-497 (momentum update), :357-379 (errors)."""
    ed = _ed()
    rng = np.random.default_rng(2)
    B, D, ridge, m = 40, 16, 1.0, 0.9
    phi = torch.tensor(rng.standard_normal((B, D)), dtype=torch.float32, device=cuda)
    layer = ed.nn.LaplaceRandomFeatureCovariance(hidden_features=D, ridge_penalty=ridge, momentum=m)
    v = layer.init(0, phi)
    _, mut = layer.apply(v, phi, mutable=["laplace_covariance"])
    ref = og.precision_update_jax(ridge * np.eye(D), to_np(phi), m)
    assert rel_err(to_np(mut["laplace_covariance"]["precision_matrix"]), ref) < 1e-5
    with pytest.raises(ValueError):
        ed.nn.LaplaceRandomFeatureCovariance(hidden_features=D, momentum=1.5)
    with pytest.raises(ValueError):
        ed.nn.LaplaceRandomFeatureCovariance(hidden_features=D, likelihood="bogus")
    lp = ed.nn.LaplaceRandomFeatureCovariance(hidden_features=D, likelihood="poisson")
    vp = lp.init(0, phi, torch.zeros(B, 1, device=cuda))
    with pytest.raises(ValueError):
        lp.apply(vp, phi, None, mutable=["laplace_covariance"])
    with pytest.raises(ValueError):
        lp.apply(vp, phi, torch.zeros(B, 3, device=cuda), mutable=["laplace_covariance"])


def test_spectral_normalization_functional(cuda):
    """jax normalization_test.py:52-104: u / v live in `spectral_stats`, are updated only when that collection is
    mutable and training; a kernel rescaled to sigma = 10 is normalised back to norm_multiplier (rtol 1e-2); the
    gradient reaches the wrapped layer's kernel scaled by the stop-gradient factor."""
    ed = _ed()
    rng = np.random.default_rng(3)
    x = torch.tensor(rng.standard_normal((32, 20)), dtype=torch.float32, device=cuda)
    sn = ed.nn.SpectralNormalization(ed.nn.Dense(features=16, use_bias=False), iteration=100, norm_multiplier=0.95)
    variables = sn.init(0, x)
    assert set(variables) == {"params", "spectral_stats"} and set(variables["spectral_stats"]) == {"u", "v"}
    k = variables["params"]["Dense"]["kernel"]
    with torch.no_grad():
        k.mul_(10.0 / np.linalg.svd(to_np(k), compute_uv=False)[0])
    u0 = to_np(variables["spectral_stats"]["u"]).copy()
    out_eval = sn.apply(variables, x, training=False)
    np.testing.assert_allclose(to_np(variables["spectral_stats"]["u"]), u0)
    state = variables
    for _ in range(5):
        out, mut = sn.apply(state, x, training=True, mutable=["spectral_stats", "intermediates"])
        state = dict(state, spectral_stats=mut["spectral_stats"])
    w_hat = to_np(mut["intermediates"]["w"])
    np.testing.assert_allclose(np.linalg.svd(w_hat, compute_uv=False)[0], 0.95, rtol=1e-2)
    assert rel_err(to_np(out), to_np(x) @ w_hat) < 1e-5
    out.sum().backward()
    g = to_np(k.grad)
    g_ref = (to_np(x).T @ np.ones((32, 16))) * (0.95 / 10.0)
    assert rel_err(g, g_ref) < 2e-2
    assert out_eval.shape == (32, 16)
