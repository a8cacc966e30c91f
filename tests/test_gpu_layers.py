"""GPU tests of the drop-in layer classes: the reference's own test assertions, restated against `edward2_b200`
(edward2/tensorflow/layers/dense_test.py, random_feature_test.py, normalization_test.py; edward2/jax/nn/*_test.py)."""
import numpy as np
import pytest
import torch

from oracle import dense as od, sngp as og
from tests.util import rel_err, to_np

pytestmark = pytest.mark.gpu


def _ed():
    import edward2_b200 as ed

    return ed


@pytest.mark.parametrize("layer_name", ["DenseReparameterization", "DenseFlipout"])
@pytest.mark.parametrize("kernel_init,bias_init,all_close", [
    ("zeros", "zeros", True),
    ("trainable_normal", "zeros", False),
    ("zeros", "trainable_normal", False),
])
def test_dense_kernel(cuda, layer_name, kernel_init, bias_init, all_close):
    """dense_test.py:85-105 (testDenseKernel): two calls agree iff the initializers are deterministic; relu >= 0;
    rank-3 inputs [5, 3, 12] -> [5, 3, 4]."""
    ed = _ed()
    inputs = torch.tensor(np.random.default_rng(0).random((5, 3, 12)), dtype=torch.float32, device=cuda)
    layer = getattr(ed.layers, layer_name)(4, kernel_initializer=kernel_init, bias_initializer=bias_init, activation="relu")
    # widen the posterior so that two samples differ visibly (default rho=-3 -> sigma 0.049)
    o1 = layer(inputs)
    o2 = layer(inputs)
    torch.cuda.synchronize()
    assert o1.shape == (5, 3, 4)
    assert (to_np(o1) >= 0).all()
    if all_close:
        np.testing.assert_allclose(to_np(o1), to_np(o2), atol=1e-6)
    else:
        assert not np.allclose(to_np(o1), to_np(o2))


@pytest.mark.parametrize("layer_name", ["DenseReparameterization", "DenseFlipout"])
def test_dense_mean(cuda, layer_name):
    """dense_test.py:113-129 (testDenseMean): with the kernel at its posterior mean (~TruncN(0,1e-5)) the output is ~0."""
    ed = _ed()
    inputs = torch.tensor(np.random.default_rng(0).random((5, 3, 7)), dtype=torch.float32, device=cuda)
    layer = getattr(ed.layers, layer_name)(4, activation="relu")
    layer.inject(eps=torch.zeros(7, 4, device=cuda))
    out = layer(inputs)
    np.testing.assert_allclose(to_np(out), np.zeros((5, 3, 4)), atol=1e-4)


@pytest.mark.parametrize("layer_name,n_losses", [("DenseReparameterization", 1), ("DenseFlipout", 1), ("DenseRank1", 2)])
def test_dense_loss_and_gradients(cuda, layer_name, n_losses):
    """dense_test.py:138-213: gradients exist w.r.t. the posterior mean and stddev for NLL and KL, across two epochs;
    one loss per Bayesian layer (two for DenseRank1)."""
    ed = _ed()
    rng = np.random.default_rng(1)
    x = torch.tensor(rng.random((8, 12)), dtype=torch.float32, device=cuda)
    t = torch.tensor(rng.random((8, 2)), dtype=torch.float32, device=cuda)
    kw = dict(ensemble_size=2) if layer_name == "DenseRank1" else {}
    layer = getattr(ed.layers, layer_name)(2, **kw)
    for _ in range(2):
        out = layer(x)
        nll = ((out.float() - t) ** 2).sum()
        assert len(layer.losses) == n_losses
        kl = sum(layer.losses)
        params = [p for p in layer.parameters()]
        g_nll = torch.autograd.grad(nll, params, retain_graph=True, allow_unused=True)
        g_kl = torch.autograd.grad(kl, params, allow_unused=True)
        names = [n for n, _ in layer.named_parameters()]
        for n, g in zip(names, g_nll):
            assert g is not None, f"no NLL gradient for {n}"
        for n, g in zip(names, g_kl):
            if n.endswith(("mean", "stddev")):
                assert g is not None, f"no KL gradient for {n}"
        assert float(kl) >= 0


def test_kl_scale_factor_linearity(cuda):
    """regularizers_test.py:65-80: KL >= 0 and linear in scale_factor; trainable_normal scale ~ softplus(-3)."""
    ed = _ed()
    init = ed.initializers.TrainableNormal()
    init.build((40, 30), device=cuda)
    k1 = ed.regularizers.NormalKLDivergence()(init)
    k2 = ed.regularizers.NormalKLDivergence(scale_factor=0.25)(init)
    assert float(k1) >= 0
    assert abs(float(k2) - 0.25 * float(k1)) <= 1e-5 * float(k1)
    mean, rho = init.params()
    assert abs(float(k1) - od.normal_kl(to_np(mean), to_np(rho))) <= 1e-5 * float(k1)
    np.testing.assert_allclose(to_np(mean), 0, atol=1e-4)
    np.testing.assert_allclose(od.stddev(to_np(rho)), 0.048587, atol=5e-2)


@pytest.mark.parametrize("dtype,tol", [("float32", 1e-5), ("bfloat16", 2e-2)])
def test_batch_ensemble_vectorised_equals_loop(cuda, dtype, tol):
    """dense_test.py:289-316 (testDenseBatchEnsemble): the vectorised layer equals a per-member loop of
    Dense(x * alpha_e) * gamma_e + bias_e."""
    ed = _ed()
    E, n, I, O = 3, 4, 5, 5
    rng = np.random.default_rng(2)
    x = torch.tensor(rng.standard_normal((E * n, I)), dtype=torch.float32, device=cuda)
    layer = ed.layers.DenseBatchEnsemble(O, ensemble_size=E, alpha_initializer="he_normal", gamma_initializer="he_normal",
                                         activation=None, dtype=dtype)
    out = layer(x)
    torch.cuda.synchronize()
    xin = to_np(x.to(layer.compute_dtype))
    w = to_np(layer.kernel.to(layer.compute_dtype))
    alpha, gamma, bias = to_np(layer.alpha), to_np(layer.gamma), to_np(layer.ensemble_bias)
    loop = np.concatenate([((xin[e * n:(e + 1) * n] * alpha[e]) @ w) * gamma[e] + bias[e] for e in range(E)], 0)
    assert rel_err(to_np(out), loop) < tol
    with pytest.raises(ValueError):
        layer(x[:-1])


@pytest.mark.parametrize("dtype,tol", [("float32", 1e-5), ("bfloat16", 2e-2)])
def test_batch_ensemble_rank_r(cuda, dtype, tol):
    """DenseBatchEnsemble(rank=3) (dense.py:518-523, 560-570): forward against the oracle's statement-by-statement
    restatement, gradients against float64 autograd of the same formula."""
    from oracle import dense as od

    ed = _ed()
    R, E, n, I, O = 3, 2, 8, 24, 16
    rng = np.random.default_rng(11)
    layer = ed.layers.DenseBatchEnsemble(O, rank=R, ensemble_size=E, alpha_initializer="he_normal",
                                         gamma_initializer="he_normal", bias_initializer="he_normal", activation="relu",
                                         dtype=dtype)
    x = torch.tensor(rng.standard_normal((E * n, I)), dtype=torch.float32, device=cuda).to(layer.compute_dtype).requires_grad_(True)
    y = layer(x)
    assert tuple(layer.alpha.shape) == (R, E, I) and tuple(layer.gamma.shape) == (R, E, O)
    gy = torch.tensor(rng.standard_normal((E * n, O)), dtype=torch.float32, device=cuda)
    (y.float() * gy).sum().backward()
    torch.cuda.synchronize()
    xin, w = to_np(x.detach()), to_np(layer.kernel.detach().to(layer.compute_dtype))
    alpha, gamma, bias = to_np(layer.alpha.detach()), to_np(layer.gamma.detach()), to_np(layer.ensemble_bias.detach())
    assert rel_err(to_np(y), od.batch_ensemble_rank_fwd(xin, w, alpha, gamma, bias, "relu")) < tol
    t = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in
         dict(x=xin, w=w, a=alpha, g=gamma, b=bias).items()}
    a_rows = t["a"].repeat_interleave(n, dim=1)          # [R, B, I]: member e owns rows e*n .. e*n + n
    g_rows = t["g"].repeat_interleave(n, dim=1)
    pre = (((t["x"][None] * a_rows) @ t["w"]) * g_rows).sum(0) + t["b"].repeat_interleave(n, dim=0)
    mask = torch.tensor(to_np(y) > 0)                    # bf16: relu' conditioned on the kernel's own pattern
    ((pre * mask) * torch.tensor(to_np(gy), dtype=torch.float64)).sum().backward()
    assert rel_err(to_np(x.grad), t["x"].grad.numpy()) < tol
    assert rel_err(to_np(layer.kernel.grad), t["w"].grad.numpy()) < tol
    assert rel_err(to_np(layer.alpha.grad), t["a"].grad.numpy()) < tol
    assert rel_err(to_np(layer.gamma.grad), t["g"].grad.numpy()) < tol
    assert rel_err(to_np(layer.ensemble_bias.grad), t["b"].grad.numpy()) < tol


def test_rank1_deterministic_equals_batch_ensemble(cuda):
    """dense_test.py:430-460 (testDenseRank1BatchEnsemble): with deterministic fast weights DenseRank1 == BatchEnsemble."""
    ed = _ed()
    E, n, I, O = 3, 4, 5, 6
    rng = np.random.default_rng(3)
    x = torch.tensor(rng.standard_normal((E * n, I)), dtype=torch.float32, device=cuda)
    r1 = ed.layers.DenseRank1(O, alpha_initializer="he_normal", gamma_initializer="he_normal", alpha_regularizer=None,
                              gamma_regularizer=None, ensemble_size=E)
    out = r1(x)
    y_ref, _ = od.batch_ensemble_fwd(to_np(x), to_np(r1.kernel), to_np(r1.alpha), to_np(r1.gamma), to_np(r1.ensemble_bias))
    assert rel_err(to_np(out), y_ref) < 1e-5
    assert r1.losses == []


def test_rank1_sampling_toggles(cuda):
    """dense_test.py:529-550 (testDenseRank1AlphaGamma), rows that are consistent with dense.py:1105-1129: sampled
    fast weights make two calls differ; deterministic ones do not."""
    ed = _ed()
    x = torch.tensor(np.random.default_rng(4).standard_normal((8, 6)), dtype=torch.float32, device=cuda)
    det = ed.layers.DenseRank1(4, alpha_initializer="ones", gamma_initializer="ones", ensemble_size=2)
    assert np.allclose(to_np(det(x)), to_np(det(x)))
    sto = ed.layers.DenseRank1(4, alpha_initializer="trainable_normal", gamma_initializer="ones", ensemble_size=2)
    # default posterior mean ~1e-5: move it to 1 so the perturbation is visible
    sto(x)
    with torch.no_grad():
        sto.alpha_initializer.mean.fill_(1.0)
    assert not np.allclose(to_np(sto(x)), to_np(sto(x)))
    assert len(sto.losses) == 1


def test_orthogonal_random_features(cuda):
    """initializers_test.py:105-133: squared column norms == rows * stddev^2 (random_norm=False), blocks orthogonal."""
    ed = _ed()
    rows, cols, std = 8, 20, 2.0
    w = ed.initializers.OrthogonalRandomFeatures(stddev=std, random_norm=False, seed=0)((rows, cols)).double().numpy()
    np.testing.assert_allclose((w ** 2).sum(0), rows * std ** 2 * np.ones(cols), atol=1e-5, rtol=1e-5)
    blk = w[:, :rows]
    gram = blk.T @ blk
    np.testing.assert_allclose(gram - np.diag(np.diag(gram)), 0, atol=1e-3)


def test_laplace_covariance_state_machine(cuda):
    """random_feature_test.py:175-244: training returns eye and raises the update flag; the first inference call inverts
    and clears it; precision follows the momentum rule; covariance == ridge * phi inv(ridge I + P) phiᵀ."""
    ed = _ed()
    rng = np.random.default_rng(5)
    B, D, ridge, m = 24, 16, 1e-2, 0.9
    phi = torch.tensor(rng.standard_normal((B, D)), dtype=torch.float32, device=cuda)
    layer = ed.layers.LaplaceRandomFeatureCovariance(momentum=m, ridge_penalty=ridge)
    out = layer(phi, training=True)
    np.testing.assert_allclose(to_np(out), np.eye(B), atol=0)
    assert bool(layer.if_update_covariance)
    p_ref = og.precision_update_tf(np.zeros((D, D)), to_np(phi), m)
    assert rel_err(to_np(layer.precision_matrix), p_ref) < 1e-5
    cov = layer(phi, training=False)
    assert not bool(layer.if_update_covariance)
    sigma = og.covariance_tf(p_ref, ridge)
    assert rel_err(to_np(layer.covariance_matrix), sigma) < 1e-4
    assert rel_err(to_np(cov), og.predictive_covariance_tf(to_np(phi), sigma, ridge)) < 1e-4
    var, _ = layer.compute_predictive_variance(phi)
    assert rel_err(to_np(var), np.diag(og.predictive_covariance_tf(to_np(phi), sigma, ridge))) < 1e-4
    layer.reset_precision_matrix()
    assert float(layer.precision_matrix.abs().max()) == 0.0
    with pytest.raises(ValueError):
        ed.layers.LaplaceRandomFeatureCovariance(likelihood="bogus")
    lp = ed.layers.LaplaceRandomFeatureCovariance(likelihood="poisson")
    with pytest.raises(ValueError):
        lp(phi, logits=None, training=True)
    with pytest.raises(ValueError):
        lp(phi, logits=torch.zeros(B, 3, device=cuda), training=True)


def test_laplace_covariance_ema_converges(cuda):
    """random_feature_test.py:84-104: minibatch EMA converges to the population XᵀX / N (atol 1e-3, rtol 5e-2 there,
    with 1000 epochs; fewer steps and a faster momentum here)."""
    ed = _ed()
    rng = np.random.default_rng(6)
    N, D, bs = 400, 8, 50
    x = rng.standard_normal((N, D))
    pop = x.T @ x / N
    layer = ed.layers.LaplaceRandomFeatureCovariance(momentum=0.98, ridge_penalty=0.0)
    xt = torch.tensor(x, dtype=torch.float32, device=cuda)
    for _ in range(60):
        for i in range(0, N, bs):
            layer(xt[i:i + bs], training=True)
    np.testing.assert_allclose(to_np(layer.precision_matrix), pop, atol=5e-2, rtol=2e-1)


def test_rfgp_end_to_end(cuda):
    """RFGP forward vs the oracle composition (random_feature.py:247-281) with the layer's own frozen W, b, beta;
    list return [logits, covmat]; training covmat is eye; gradients reach the input and the output weights."""
    ed = _ed()
    rng = np.random.default_rng(7)
    B, d, D, K = 64, 24, 128, 10
    x = torch.tensor(rng.standard_normal((B, d)), dtype=torch.float32, device=cuda, requires_grad=True)
    gp = ed.layers.RandomFeatureGaussianProcess(K, num_inducing=D, gp_cov_momentum=-1, gp_cov_ridge_penalty=1.0,
                                                return_random_features=True)
    logits, cov, feats = gp(x, training=True)
    assert logits.shape == (B, K) and cov.shape == (B, B) and feats.shape == (B, D)
    np.testing.assert_allclose(to_np(cov), np.eye(B))
    ln = gp._input_norm_layer
    xn = og.layer_norm(to_np(x), to_np(ln.gamma), to_np(ln.beta), 1e-3)
    phi_ref, _ = og.rff(xn, to_np(gp._random_feature.kernel), to_np(gp._random_feature.bias), np.sqrt(2.0 / D))
    assert rel_err(to_np(feats), phi_ref) < 1e-5
    logits_ref = og.gp_output(phi_ref, to_np(gp._gp_output_layer.kernel), to_np(gp._gp_output_bias))
    assert rel_err(to_np(logits), logits_ref) < 1e-5
    logits.sum().backward()
    assert x.grad is not None and gp._gp_output_layer.kernel.grad is not None
    assert rel_err(to_np(gp._gp_cov_layer.precision_matrix), phi_ref.T @ phi_ref) < 1e-5
    logits2, cov2, _ = gp(x.detach(), training=False)
    sigma = og.covariance_tf(phi_ref.T @ phi_ref, 1.0)
    assert rel_err(to_np(cov2), og.predictive_covariance_tf(phi_ref, sigma, 1.0)) < 1e-4
    adj = ed.layers.utils.mean_field_logits(logits2, cov2, mean_field_factor=1.0)
    assert rel_err(to_np(adj), og.mean_field_logits(to_np(logits2), to_np(cov2), 1.0)) < 1e-5


def test_rfgp_linear_kernel(cuda):
    """gp_kernel_type='linear' (random_feature.py:221-223): the features are the normalised inputs times
    sqrt(2 / num_inducing); output layer, precision update and predictive covariance act on the input width."""
    ed = _ed()
    rng = np.random.default_rng(17)
    B, d, K = 48, 24, 5
    x = torch.tensor(rng.standard_normal((B, d)), dtype=torch.float32, device=cuda, requires_grad=True)
    gp = ed.layers.RandomFeatureGaussianProcess(K, num_inducing=64, gp_kernel_type="linear", gp_cov_momentum=-1,
                                                gp_cov_ridge_penalty=1.0, return_random_features=True)
    logits, cov, feats = gp(x, training=True)
    assert feats.shape == (B, d) and gp._random_feature is None
    ln = gp._input_norm_layer
    phi_ref = og.layer_norm(to_np(x), to_np(ln.gamma), to_np(ln.beta), 1e-3) * np.sqrt(2.0 / 64)
    assert rel_err(to_np(feats), phi_ref) < 1e-5
    logits_ref = og.gp_output(phi_ref, to_np(gp._gp_output_layer.kernel), to_np(gp._gp_output_bias))
    assert rel_err(to_np(logits), logits_ref) < 1e-5
    logits.sum().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()
    assert rel_err(to_np(gp._gp_cov_layer.precision_matrix), phi_ref.T @ phi_ref) < 1e-5
    _, cov2, _ = gp(x.detach(), training=False)
    sigma = og.covariance_tf(phi_ref.T @ phi_ref, 1.0)
    assert rel_err(to_np(cov2), og.predictive_covariance_tf(phi_ref, sigma, 1.0)) < 1e-4
    # without input normalisation the inputs are rescaled by 1 / sqrt(gp_kernel_scale) (:253-257)
    gp2 = ed.layers.RandomFeatureGaussianProcess(K, num_inducing=64, gp_kernel_type="linear", normalize_input=False,
                                                 gp_kernel_scale=4.0, scale_random_features=False,
                                                 return_gp_cov=False, return_random_features=True)
    _, feats2 = gp2(x.detach())
    assert rel_err(to_np(feats2), to_np(x) * 0.5) < 1e-6


def test_spectral_normalization_layer(cuda):
    """normalization_test.py:68-102: after training iterations the wrapped kernel's spectral norm is <= ~norm_multiplier
    and the layer is Lipschitz with that constant; kernels already below the bound are left unchanged."""
    ed = _ed()
    rng = np.random.default_rng(8)
    x = torch.tensor(rng.standard_normal((32, 20)), dtype=torch.float32, device=cuda)
    dense = ed.layers.Dense(16, use_bias=False, kernel_initializer=ed.initializers.RandomNormal(stddev=1.0, seed=1))
    sn = ed.layers.SpectralNormalization(dense, iteration=50, norm_multiplier=0.95)
    for _ in range(20):
        y = sn(x, training=True)
    s = np.linalg.svd(to_np(dense.kernel), compute_uv=False)[0]
    assert abs(s - 0.95) < 1e-2
    x2 = x + 0.1
    ratio = (sn(x2, training=False) - sn(x, training=False)).float().norm() / (x2 - x).norm()
    assert float(ratio) <= 0.95 + 1e-2
    small = ed.layers.Dense(16, use_bias=False, kernel_initializer=ed.initializers.RandomNormal(stddev=1e-3, seed=2))
    sn2 = ed.layers.SpectralNormalization(small, iteration=5, norm_multiplier=0.95)
    sn2(x)
    k0 = to_np(small.kernel).copy()
    sn2(x)
    np.testing.assert_allclose(to_np(small.kernel), k0)
    with pytest.raises(ValueError):
        ed.layers.SpectralNormalization(torch.nn.Linear(3, 3))


@pytest.mark.parametrize("dtype,tol", [("float32", 1e-5), ("bfloat16", 2e-2)])
def test_two_layer_mlp_cross_layer_fusion(cuda, dtype, tol):
    """Two stacked DenseReparameterization layers: the second layer's dX epilogue applies the first layer's ReLU mask
    and accumulates its bias gradient (functional.ReluLink).  Gradients must equal the oracle's layer-by-layer chain,
    with and without the presample() scheduling hint."""
    ed = _ed()
    from edward2_b200 import functional as F

    rng = np.random.default_rng(11)
    B, I, H, O = 136, 72, 264, 40
    tdt = torch.float32 if dtype == "float32" else torch.bfloat16
    x = torch.tensor(rng.standard_normal((B, I)), dtype=torch.float32, device=cuda).to(tdt)
    eps1 = torch.tensor(rng.standard_normal((I, H)), dtype=torch.float32, device=cuda)
    eps2 = torch.tensor(rng.standard_normal((H, O)), dtype=torch.float32, device=cuda)
    gy = torch.tensor(rng.standard_normal((B, O)), dtype=torch.float32, device=cuda).to(tdt)
    bias1 = torch.tensor(rng.standard_normal(H) * 0.1, dtype=torch.float32, device=cuda)
    results = []
    for hint in (False, True):
        l1 = ed.layers.DenseReparameterization(H, activation="relu", dtype=dtype, seed=5,
                                               kernel_initializer=ed.initializers.TrainableNormal(
                                                   mean_initializer=ed.initializers.RandomNormal(stddev=0.1, seed=1),
                                                   stddev_initializer=ed.initializers.TruncatedNormal(-3.0, 0.1, seed=3)))
        l2 = ed.layers.DenseReparameterization(O, activation=None, dtype=dtype, seed=6,
                                               kernel_initializer=ed.initializers.TrainableNormal(
                                                   mean_initializer=ed.initializers.RandomNormal(stddev=0.1, seed=2),
                                                   stddev_initializer=ed.initializers.TruncatedNormal(-3.0, 0.1, seed=4)))
        l1.inject(eps=eps1)
        l2.inject(eps=eps2)
        with torch.no_grad():
            l2(l1(x))  # build
            l1.bias.copy_(bias1)
        if hint:
            l2.presample()
        h = l1(x)
        link = F.take_link(h)
        assert link is not None
        out = l2(h)
        out.backward(gy)
        torch.cuda.synchronize()
        assert link.fused, "the fused dX epilogue path was not taken"
        results.append((l1, l2, out))
    (l1, l2, out) = results[0]
    m1, r1 = l1.kernel_initializer.params()
    m2, r2 = l2.kernel_initializer.params()
    xin = to_np(x)
    h_ref, _ = od.reparam_fwd(xin, to_np(m1), to_np(r1), to_np(eps1), to_np(l1.bias), "relu")
    if dtype == "bfloat16":  # the kernel feeds layer 2 with its own bf16-rounded activations
        h_ref_in = to_np(l1(x).detach())
    else:
        h_ref_in = h_ref
    y_ref, _ = od.reparam_fwd(h_ref_in, to_np(m2), to_np(r2), to_np(eps2), to_np(l2.bias), None)
    assert rel_err(to_np(out), y_ref) < tol
    g2 = od.reparam_bwd(h_ref_in, to_np(m2), to_np(r2), to_np(eps2), to_np(l2.bias), None, to_np(gy))
    g1 = od.reparam_bwd(xin, to_np(m1), to_np(r1), to_np(eps1), to_np(l1.bias), "relu", g2["dx"], y_mask=h_ref_in)
    assert rel_err(to_np(m2.grad), g2["dmu"]) < tol
    assert rel_err(to_np(m1.grad), g1["dmu"]) < tol
    assert rel_err(to_np(r1.grad), g1["drho"]) < tol
    assert rel_err(to_np(l1.bias.grad), g1["dbias"]) < tol
    # the presample() hint changes scheduling only
    (p1, p2, pout) = results[1]
    assert rel_err(to_np(pout), to_np(out)) < (1e-6 if dtype == "float32" else tol)
    assert rel_err(to_np(p1.kernel_initializer.mean.grad), to_np(m1.grad)) < (1e-5 if dtype == "float32" else tol)


def test_relu_link_falls_back_under_fan_out(cuda):
    """A ReLU layer whose output feeds the next DenseReparameterization AND a skip connection: autograd sums the two
    gradients, the producer must not treat the sum as pre-masked (functional.ReluLink.matches).  Gradients equal the
    oracle's chain with the skip term included."""
    ed = _ed()
    from edward2_b200 import functional as F

    rng = np.random.default_rng(12)
    B, I, H = 72, 40, 136
    x = torch.tensor(rng.standard_normal((B, I)), dtype=torch.float32, device=cuda)
    eps1 = torch.tensor(rng.standard_normal((I, H)), dtype=torch.float32, device=cuda)
    eps2 = torch.tensor(rng.standard_normal((H, H)), dtype=torch.float32, device=cuda)
    gy = torch.tensor(rng.standard_normal((B, H)), dtype=torch.float32, device=cuda)
    mk = lambda seed: ed.layers.DenseReparameterization(  # noqa: E731
        H, activation="relu" if seed == 1 else None, seed=seed, kernel_regularizer=None,
        kernel_initializer=ed.initializers.TrainableNormal(
            mean_initializer=ed.initializers.RandomNormal(stddev=0.1, seed=seed),
            stddev_initializer=ed.initializers.TruncatedNormal(-3.0, 0.1, seed=seed + 10)))
    for order in (0, 1):  # both arrival orders of the two gradient contributions
        l1, l2 = mk(1), mk(2)
        l1.inject(eps=eps1)
        l2.inject(eps=eps2)
        h = l1(x)
        link = F.take_link(h)
        out = (l2(h) + 0.5 * h) if order == 0 else (0.5 * h + l2(h))
        out.backward(gy)
        torch.cuda.synchronize()
        assert link is not None and not link.fused
        m1, r1 = l1.kernel_initializer.params()
        m2, r2 = l2.kernel_initializer.params()
        h_ref, _ = od.reparam_fwd(to_np(x), to_np(m1), to_np(r1), to_np(eps1), to_np(l1.bias), "relu")
        g2 = od.reparam_bwd(h_ref, to_np(m2), to_np(r2), to_np(eps2), to_np(l2.bias), None, to_np(gy))
        g1 = od.reparam_bwd(to_np(x), to_np(m1), to_np(r1), to_np(eps1), to_np(l1.bias), "relu",
                            g2["dx"] + 0.5 * to_np(gy))
        assert rel_err(to_np(m1.grad), g1["dmu"]) < 1e-5
        assert rel_err(to_np(r1.grad), g1["drho"]) < 1e-5
        assert rel_err(to_np(l1.bias.grad), g1["dbias"]) < 1e-5


def test_sngp_head_restores_from_state_dict(cuda):
    """Loading a checkpoint with a DIFFERENT random-feature kernel and a mid-training covariance state must change the
    forward: the operand caches of the RFF layer and the host mirror of `if_update_covariance` follow the restored
    buffers (random_feature.py:346-378 variable set)."""
    ed = _ed()
    torch.manual_seed(0)
    x = torch.randn(64, 24, device=cuda)

    def make(seed):
        return ed.layers.RandomFeatureGaussianProcess(
            3, num_inducing=128, gp_cov_momentum=-1, gp_cov_ridge_penalty=1.0,
            custom_random_features_initializer=ed.initializers.OrthogonalRandomFeatures(stddev=1.0, seed=seed))

    src, dst = make(1), make(2)
    src.train()
    for _ in range(2):
        src(x)  # precision accumulated, if_update_covariance == True, covariance not yet inverted
    dst.eval()
    logits_before, cov_before = dst(x)
    dst.load_state_dict(src.state_dict())
    logits_after, cov_after = dst(x)
    src.eval()
    logits_src, cov_src = src(x)
    torch.cuda.synchronize()
    assert rel_err(to_np(logits_after), to_np(logits_src)) < 1e-6
    assert rel_err(to_np(cov_after), to_np(cov_src)) < 1e-5
    assert rel_err(to_np(cov_before), to_np(cov_src)) > 1e-2  # the restore really changed the result
    # training identity: cached, still the identity
    src.train()
    eye = src(x)[1]
    assert eye.shape == (64, 64) and torch.equal(eye, torch.eye(64, device=cuda))
