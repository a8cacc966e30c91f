"""CPU tests of the oracle: known-answer vectors, the reference's algebraic identities, closed-form gradients vs
float64 autograd, and the committed golden vectors (tests/golden/oracle_vectors.npz)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import dense as od, philox as op, sngp as og

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_philox_known_answers():
    kat = json.load(open(os.path.join(GOLD, "philox_kat.json")))
    for v in kat["vectors"]:
        c = np.array([int(x, 16) for x in v["counter"]], dtype=np.uint32)
        k = np.array([int(x, 16) for x in v["key"]], dtype=np.uint32)
        out = op.philox4x32_10(c, k)
        assert [f"{int(x):08x}" for x in out] == v["output"]


def test_philox_stream_moments_and_ks():
    """moment and KS tests of the stream recipe (host restatement; the device sampler is compared with
    it value by value in the GPU tests)."""
    from scipy import stats

    n = 1_000_000
    z = op.normal(7, 1, n)
    assert abs(z.mean()) < 4e-3 and abs(z.var() - 1) < 6e-3
    assert abs(stats.skew(z)) < 1e-2 and abs(stats.kurtosis(z)) < 2e-2
    assert stats.kstest(z[:200_000], "norm").pvalue > 1e-3
    u = op.uniform(7, 2, n)
    assert 0 < u.min() and u.max() < 1
    assert stats.kstest(u[:200_000], "uniform").pvalue > 1e-3
    s = op.sign(7, 3, n)
    assert set(np.unique(s)) == {-1.0, 1.0}
    assert stats.binomtest(int((s > 0).sum()), n, 0.5).pvalue > 1e-3
    # element i depends only on (seed, stream, i): any tiling regenerates the same values
    assert np.array_equal(op.normal(7, 1, 1000)[400:800], op.normal(7, 1, 800)[400:])
    assert not np.array_equal(op.normal(7, 1, 100), op.normal(8, 1, 100))
    assert not np.array_equal(op.normal(7, 1, 100), op.normal(7, 2, 100))


def test_mean_field_logits_known_answers():
    """edward2/tensorflow/layers/utils_test.py:201-241; edward2/jax/nn/utils_test.py:28-68."""
    rng = np.random.RandomState(0)
    logits = rng.randn(10, 12)
    covmat = np.diag([1.5] * 10)
    np.testing.assert_allclose(og.mean_field_logits(logits, covmat, 2.0), logits / 2.0, atol=1e-4)
    np.testing.assert_allclose(og.mean_field_logits(logits, covmat, 2.0, "poisson"), logits * np.exp(1.5), atol=1e-4)
    np.testing.assert_allclose(og.mean_field_logits(logits, None, -1), logits, atol=1e-4)
    np.testing.assert_allclose(og.mean_field_logits(logits, None, 3.0), logits / 2.0, atol=1e-4)
    np.testing.assert_allclose(og.mean_field_logits(logits, np.full(10, 1.5), 2.0), logits / 2.0, atol=1e-4)
    with pytest.raises(ValueError):
        og.mean_field_logits(logits, covmat, likelihood="gaussian")


def test_batch_ensemble_equals_member_loop():
    """dense_test.py:289-316 / jax dense_test.py:34-71."""
    rng = np.random.default_rng(0)
    E, n, I, O = 3, 4, 5, 5
    x, w = rng.standard_normal((E * n, I)), rng.standard_normal((I, O))
    a, g, b = rng.standard_normal((E, I)), rng.standard_normal((E, O)), rng.standard_normal((E, O))
    y, _ = od.batch_ensemble_fwd(x, w, a, g, b)
    loop = np.concatenate([((x[e * n:(e + 1) * n] * a[e]) @ w) * g[e] + b[e] for e in range(E)])
    np.testing.assert_allclose(y, loop, rtol=1e-12, atol=1e-12)


def test_batch_ensemble_rank_r_equals_sum_of_rank_one_members():
    """dense.py:560-570: the rank > 1 branch is the sum over r of rank-1 BatchEnsemble products, one bias per member."""
    rng = np.random.default_rng(7)
    R, E, n, I, O = 3, 2, 4, 5, 6
    x, w = rng.standard_normal((E * n, I)), rng.standard_normal((I, O))
    a, g, b = rng.standard_normal((R, E, I)), rng.standard_normal((R, E, O)), rng.standard_normal((E, O))
    y = od.batch_ensemble_rank_fwd(x, w, a, g, b, "relu")
    loop = np.concatenate([sum(((x[e * n:(e + 1) * n] * a[r, e]) @ w) * g[r, e] for r in range(R)) + b[e] for e in range(E)])
    np.testing.assert_allclose(y, np.maximum(loop, 0), rtol=1e-12, atol=1e-12)
    y1 = od.batch_ensemble_rank_fwd(x, w, a[:1], g[:1], b, None)
    np.testing.assert_allclose(y1, od.batch_ensemble_fwd(x, w, a[0], g[0], b)[0], rtol=1e-12, atol=1e-12)


def test_kl_properties():
    """regularizers_test.py:65-80: KL >= 0, linear in scale_factor, 0 for q == p; TFP's expm1 form agrees."""
    rng = np.random.default_rng(1)
    mu, rho = rng.standard_normal((20, 10)), rng.standard_normal((20, 10))
    k = od.normal_kl(mu, rho)
    assert k >= 0
    assert abs(od.normal_kl(mu, rho, scale_factor=0.3) - 0.3 * k) < 1e-12 * k
    rho0 = np.log(np.expm1(1.0 - 1e-7))
    assert abs(od.normal_kl(np.zeros(5), np.full(5, rho0))) < 1e-12
    sigma = od.stddev(rho)
    ratio = sigma / 1.0
    tfp = (0.5 * (mu / 1.0) ** 2 + 0.5 * np.expm1(2 * np.log(ratio)) - np.log(ratio)).sum()
    assert abs(tfp - k) < 1e-10 * k
    np.testing.assert_allclose(od.stddev(-3.0), 0.048587, atol=1e-5)  # initializers_test.py:73-86


def _t(a, grad=True):
    return torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=grad)


def test_closed_form_gradients_match_autograd():
    rng = np.random.default_rng(2)
    E, n, I, O = 2, 5, 7, 6
    B = E * n
    x, w = rng.standard_normal((B, I)), rng.standard_normal((I, O))
    gy = rng.standard_normal((B, O))
    sp = lambda r: torch.nn.functional.softplus(r) + 1e-7  # noqa: E731

    # BatchEnsemble
    a, g, b = rng.standard_normal((E, I)), rng.standard_normal((E, O)), rng.standard_normal((E, O))
    xt, wt, at, gt, bt = _t(x), _t(w), _t(a), _t(g), _t(b)
    y = torch.relu(((xt.reshape(E, n, I) * at[:, None]) @ wt) * gt[:, None] + bt[:, None]).reshape(B, O)
    y.backward(_t(gy, False))
    got = od.batch_ensemble_bwd(x, w, a, g, b, "relu", gy)
    for k_, t_ in (("dx", xt), ("dw", wt), ("dalpha", at), ("dgamma", gt), ("dbias", bt)):
        np.testing.assert_allclose(got[k_], t_.grad.numpy(), rtol=1e-10, atol=1e-10)

    # Reparameterisation and Flipout
    mu, rho, eps, b1 = rng.standard_normal((I, O)), rng.standard_normal((I, O)), rng.standard_normal((I, O)), rng.standard_normal(O)
    s_in, s_out = rng.choice([-1.0, 1.0], (B, I)), rng.uniform(-1, 3, (B, O))
    xt, mt, rt, bt = _t(x), _t(mu), _t(rho), _t(b1)
    torch.relu(xt @ (mt + sp(rt) * _t(eps, False)) + bt).backward(_t(gy, False))
    got = od.reparam_bwd(x, mu, rho, eps, b1, "relu", gy)
    for k_, t_ in (("dx", xt), ("dmu", mt), ("drho", rt), ("dbias", bt)):
        np.testing.assert_allclose(got[k_], t_.grad.numpy(), rtol=1e-10, atol=1e-10)
    xt, mt, rt, bt = _t(x), _t(mu), _t(rho), _t(b1)
    torch.relu(xt @ mt + ((xt * _t(s_in, False)) @ (sp(rt) * _t(eps, False))) * _t(s_out, False) + bt).backward(_t(gy, False))
    got = od.flipout_bwd(x, mu, rho, eps, s_in, s_out, b1, "relu", gy)
    for k_, t_ in (("dx", xt), ("dmu", mt), ("drho", rt), ("dbias", bt)):
        np.testing.assert_allclose(got[k_], t_.grad.numpy(), rtol=1e-10, atol=1e-10)

    # Rank-1 with clipping active on some draws
    mu_r, mu_s = rng.standard_normal((E, I)), rng.standard_normal((E, O))
    rho_r, rho_s = rng.standard_normal((E, I)), rng.standard_normal((E, O))
    eps_r, eps_s = rng.standard_normal((B, I)), rng.standard_normal((B, O))
    eps_r[0, :2] = 60.0
    eps_s[3, :2] = -60.0
    xt, wt, mrt, rrt, mst, rst, bt = _t(x), _t(w), _t(mu_r), _t(rho_r), _t(mu_s), _t(rho_s), _t(b)
    rep = lambda t_: t_.repeat_interleave(n, dim=0)  # noqa: E731
    r = torch.clamp(rep(mrt) + rep(sp(rrt)) * _t(eps_r, False), -10, 10)
    s = torch.clamp(rep(mst) + rep(sp(rst)) * _t(eps_s, False), -10, 10)
    torch.relu(((xt * r) @ wt) * s + rep(bt)).backward(_t(gy, False))
    got = od.rank1_bwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, b, "relu", gy)
    for k_, t_ in (("dx", xt), ("dw", wt), ("dmu_r", mrt), ("drho_r", rrt), ("dmu_s", mst), ("drho_s", rst), ("dbias", bt)):
        np.testing.assert_allclose(got[k_], t_.grad.numpy(), rtol=1e-9, atol=1e-9)

    # KL gradient
    mt, rt = _t(mu), _t(rho)
    sg = sp(rt)
    (0.7 * (torch.log(1.3 / sg) + (sg ** 2 + (mt - 0.2) ** 2) / (2 * 1.3 ** 2) - 0.5).sum()).backward()
    dmu, drho = od.normal_kl_grad(mu, rho, 0.2, 1.3, 0.7)
    np.testing.assert_allclose(dmu, mt.grad.numpy(), rtol=1e-10)
    np.testing.assert_allclose(drho, rt.grad.numpy(), rtol=1e-10)

    # RFF input gradient
    W, bb = rng.standard_normal((I, 9)), rng.uniform(0, 6.28, 9)
    xt = _t(x)
    (0.5 * torch.cos(xt @ _t(W, False) + _t(bb, False))).backward(_t(rng.standard_normal((B, 9)), False) * 0 + 1)
    np.testing.assert_allclose(og.rff_bwd(x, W, bb, 0.5, np.ones((B, 9))), xt.grad.numpy(), rtol=1e-10, atol=1e-12)


def test_laplace_precision_identities():
    """jax random_feature_test.py:381-528: exact update P = ridge I + ΦᵀΦ; momentum update; orthogonal data ->
    (1 + ridge) I; diag == diag(full); TF (momentum <= 0) and JAX (momentum None) predictive covariances agree."""
    rng = np.random.default_rng(3)
    B, D, ridge = 40, 12, 0.7
    phi = rng.standard_normal((B, D))
    p_jax = og.precision_update_jax(ridge * np.eye(D), phi)
    np.testing.assert_allclose(p_jax, ridge * np.eye(D) + phi.T @ phi, rtol=1e-12)
    m = 0.9
    np.testing.assert_allclose(og.precision_update_jax(ridge * np.eye(D), phi, m),
                               m * ridge * np.eye(D) + (1 - m) * phi.T @ phi / B, rtol=1e-12)
    q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    np.testing.assert_allclose(og.precision_update_jax(ridge * np.eye(D), q), (1 + ridge) * np.eye(D), atol=1e-10)
    full = og.predictive_covariance_jax(phi, p_jax, ridge, diagonal_only=False)
    np.testing.assert_allclose(og.predictive_covariance_jax(phi, p_jax, ridge), np.diag(full), rtol=1e-10)
    p_tf = og.precision_update_tf(np.zeros((D, D)), phi, momentum=-1)
    cov_tf = og.predictive_covariance_tf(phi, og.covariance_tf(p_tf, ridge), ridge)
    np.testing.assert_allclose(cov_tf, full, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(og.predictive_variance(phi, og.covariance_tf(p_tf, ridge), ridge), np.diag(cov_tf), rtol=1e-10)
    with pytest.raises(ValueError):
        og.precision_update_tf(p_tf, phi, logits=np.zeros((B, 3)), likelihood="poisson")


def test_primal_equals_dual_posterior():
    """jax random_feature_test.py:155-180: the random-feature (primal) posterior variance equals the kernel (dual) form
    K** - K*ᵀ (K + ridge I)^-1 K* with K = ΦΦᵀ (max diff < 1e-6)."""
    rng = np.random.default_rng(4)
    D, ridge = 32, 1.0
    phi_tr, phi_ts = rng.standard_normal((50, D)) / np.sqrt(D), rng.standard_normal((20, D)) / np.sqrt(D)
    prec = ridge * np.eye(D) + phi_tr.T @ phi_tr
    primal = og.predictive_covariance_jax(phi_ts, prec, ridge, diagonal_only=False)
    k_tr, k_x, k_ts = phi_tr @ phi_tr.T, phi_tr @ phi_ts.T, phi_ts @ phi_ts.T
    dual = k_ts - k_x.T @ np.linalg.solve(k_tr + ridge * np.eye(50), k_x)
    assert np.abs(primal - dual).max() < 1e-6


def test_rff_approximates_rbf_kernel():
    """jax random_feature_test.py:284-302: Φ Φᵀ approximates exp(-|x-y|^2 / 2) (atol 5e-2, D = 10240)."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal((10, 6)) * 0.7
    D = 10240
    phi, _ = og.rff(x, rng.standard_normal((6, D)), rng.uniform(0, 2 * np.pi, D), np.sqrt(2.0 / D))
    d2 = ((x[:, None] - x[None]) ** 2).sum(-1)
    np.testing.assert_allclose(phi @ phi.T, np.exp(-d2 / 2), atol=5e-2)


def test_spectral_norm_variants_agree_and_converge():
    """normalization_test.py:68-102; jax normalization_test.py:52-104: sigma -> largest singular value, normalised
    kernel has spectral norm min(sigma, c); the TF and JAX formulations compute the same iteration."""
    rng = np.random.default_rng(6)
    w, u, v = rng.standard_normal((30, 20)), rng.standard_normal(20), rng.standard_normal(30)
    wt, ut, vt, st = og.spectral_norm_tf(w, u, v, 200, 0.95)
    wj, uj, vj, sj = og.spectral_norm_jax(w, u, v, 200, 0.95)
    s_true = np.linalg.svd(w, compute_uv=False)[0]
    assert abs(st - s_true) < 1e-6 * s_true and abs(sj - s_true) < 1e-6 * s_true
    np.testing.assert_allclose(wt, wj, rtol=1e-9)
    assert abs(np.linalg.svd(wt, compute_uv=False)[0] - 0.95) < 1e-6
    small = w * (0.5 / s_true)
    ws, _, _, _ = og.spectral_norm_tf(small, u, v, 200, 0.95)
    np.testing.assert_allclose(ws, small)
    w1, _, _, s1 = og.spectral_norm_tf(w, u, v, 1, 0.95, training=False)
    assert abs(s1 - float((v.reshape(1, -1) @ w @ u.reshape(-1, 1)).item())) < 1e-12


def test_golden_vectors_pin_the_oracle():
    g = np.load(os.path.join(GOLD, "oracle_vectors.npz"))
    y, _ = od.batch_ensemble_fwd(g["be_x"], g["be_w"], g["be_alpha"], g["be_gamma"], g["be_bias"], "relu")
    np.testing.assert_allclose(y, g["be_y"], rtol=1e-12)
    y, _ = od.reparam_fwd(g["be_x"], g["bn_mu"], g["bn_rho"], g["bn_eps"], g["bn_bias"], "relu")
    np.testing.assert_allclose(y, g["rp_y"], rtol=1e-12)
    y, _ = od.flipout_fwd(g["be_x"], g["bn_mu"], g["bn_rho"], g["bn_eps"], g["bn_sign_in"], g["bn_sign_out"], g["bn_bias"], "relu")
    np.testing.assert_allclose(y, g["fl_y"], rtol=1e-12)
    y, _, _, _ = od.rank1_fwd(g["be_x"], g["be_w"], g["r1_mu_r"], g["r1_rho_r"], g["r1_eps_r"], g["r1_mu_s"], g["r1_rho_s"],
                              g["r1_eps_s"], g["be_bias"], "relu")
    np.testing.assert_allclose(y, g["r1_y"], rtol=1e-12)
    assert abs(od.normal_kl(g["bn_mu"], g["bn_rho"], 0.5, 0.7, 0.25) - float(g["kl"])) < 1e-12
    np.testing.assert_allclose(op.normal(2026, 3, 64), g["philox_normal"], rtol=1e-14)
    wn, _, _, sg = og.spectral_norm_tf(g["sn_w_in"], g["sn_u_in"], g["sn_v_in"], 2, 0.95, True)
    np.testing.assert_allclose(wn, g["sn_tf_w"], rtol=1e-12)


def test_rank1_equals_member_loop_and_clip_gradient():
    """dense_test.py:430-460: the vectorised rank-1 layer equals a per-member loop with per-example fast weights;
    values beyond the [-10, 10] clip (dense.py:1109-1111) pass no gradient to mu / rho."""
    rng = np.random.default_rng(3)
    E, n, I, O = 2, 3, 4, 5
    x, w = rng.standard_normal((E * n, I)), rng.standard_normal((I, O))
    mu_r, rho_r = rng.standard_normal((E, I)), rng.standard_normal((E, I)) - 1
    mu_s, rho_s = rng.standard_normal((E, O)), rng.standard_normal((E, O)) - 1
    eps_r, eps_s, bias = rng.standard_normal((E * n, I)), rng.standard_normal((E * n, O)), rng.standard_normal((E, O))
    y, _, _, _ = od.rank1_fwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, "relu")
    rows = []
    for e in range(E):
        for i in range(n):
            b = e * n + i
            r = np.clip(mu_r[e] + od.stddev(rho_r[e]) * eps_r[b], -10, 10)
            s = np.clip(mu_s[e] + od.stddev(rho_s[e]) * eps_s[b], -10, 10)
            rows.append(np.maximum(((x[b] * r) @ w) * s + bias[e], 0.0))
    np.testing.assert_allclose(y, np.stack(rows), rtol=1e-12, atol=1e-12)
    # saturate one fast weight: its mean / rho gradient contribution from that example must vanish
    mu_big = mu_r.copy()
    mu_big[0, 0] = 50.0
    g = od.rank1_bwd(x, w, mu_big, rho_r, eps_r, mu_s, rho_s, eps_s, bias, None, np.ones((E * n, O)))
    assert g["dmu_r"][0, 0] == 0.0 and g["drho_r"][0, 0] == 0.0 and np.abs(g["dmu_r"][1]).sum() > 0


def test_posterior_mean_forward_is_near_deterministic():
    """dense_test.py:113-129: with the trainable-normal defaults (stddev = softplus(-3) ~ 0.049 only at init; the test
    pins rho very negative) the sampled forward equals the mean forward to 1e-4; Flipout with zero perturbation too."""
    rng = np.random.default_rng(4)
    x, mu = rng.standard_normal((6, 5)), rng.standard_normal((5, 4))
    rho = np.full((5, 4), -20.0)
    eps = rng.standard_normal((5, 4))
    y, _ = od.reparam_fwd(x, mu, rho, eps, None, None)
    np.testing.assert_allclose(y, x @ mu, atol=1e-4)
    s_in, s_out = np.sign(rng.standard_normal((6, 5))), np.sign(rng.standard_normal((6, 4)))
    yf = od.flipout_fwd(x, mu, rho, eps, s_in, s_out, None, None)[0]
    np.testing.assert_allclose(yf, x @ mu, atol=1e-4)


def test_precision_ema_converges_and_exact_sum():
    """random_feature_test.py (TF) :84-104: with momentum the precision matrix converges to E[phi^T phi / B]; with
    momentum <= 0 it is the exact running sum; the JAX form (momentum None) starts from ridge * I."""
    rng = np.random.default_rng(5)
    D, B = 6, 4000
    target = None
    p = np.zeros((D, D))
    for _ in range(400):
        phi = rng.standard_normal((B, D)) * np.arange(1, D + 1)
        p = og.precision_update_tf(p, phi, momentum=0.95)
    target = np.diag(np.arange(1, D + 1) ** 2.0)
    np.testing.assert_allclose(p, target, rtol=0.05, atol=0.3)
    phis = [rng.standard_normal((8, D)) for _ in range(3)]
    acc = np.zeros((D, D))
    for ph in phis:
        acc = og.precision_update_tf(acc, ph, momentum=-1.0)
    np.testing.assert_allclose(acc, sum(ph.T @ ph for ph in phis), rtol=1e-12)
    pj = 1e-3 * np.eye(D)
    for ph in phis:
        pj = og.precision_update_jax(pj, ph, momentum=None)
    np.testing.assert_allclose(pj, 1e-3 * np.eye(D) + acc, rtol=1e-12)
    # TF inverse of (ridge I + P) and the JAX Cholesky form of the same precision give the same variance
    cov = og.covariance_tf(acc, 1e-3)
    v_tf = og.predictive_variance(phis[0], cov, 1e-3)
    v_jx = og.predictive_covariance_jax(phis[0], pj, 1e-3, diagonal_only=True)
    np.testing.assert_allclose(v_tf, v_jx, rtol=1e-8)


def test_conv_oracle_pinned_against_torch_and_reference_identities():
    """oracle/conv.py (the next row): im2col restatement == torch conv2d in float64 for 'valid' / 'same'
    and strides; BatchEnsemble conv == per-member loop (convolutional_test.py's loop identity); Flipout with zero
    perturbation == mean conv; closed-form gradients == float64 autograd."""
    import torch.nn.functional as tf_

    from oracle import conv as oc

    rng = np.random.default_rng(7)
    B, H, W, C, K, kh, kw = 4, 7, 6, 3, 5, 3, 2
    x, k = rng.standard_normal((B, H, W, C)), rng.standard_normal((kh, kw, C, K))
    xt = torch.tensor(x).permute(0, 3, 1, 2)
    kt = torch.tensor(k).permute(3, 2, 0, 1)
    for strides, padding in (((1, 1), "valid"), ((2, 1), "valid"), ((1, 1), "same"), ((2, 2), "same")):
        y = oc.conv2d(x, k, strides, padding)
        if padding == "same":  # TF SAME: asymmetric padding, extra at the bottom / right
            ho, pt, pb = oc._out_size(H, kh, strides[0], "same")
            wo, pl, pr = oc._out_size(W, kw, strides[1], "same")
            ref = tf_.conv2d(tf_.pad(xt, (pl, pr, pt, pb)), kt, stride=strides)
        else:
            ref = tf_.conv2d(xt, kt, stride=strides)
        np.testing.assert_allclose(y, ref.permute(0, 2, 3, 1).numpy(), rtol=1e-12, atol=1e-12)

    # BatchEnsemble conv == loop over members
    E, n = 2, 2
    alpha, gamma, bias = rng.standard_normal((E, C)), rng.standard_normal((E, K)), rng.standard_normal((E, K))
    y = oc.batch_ensemble_conv2d(x, k, alpha, gamma, bias, "relu", (1, 1), "same")
    loop = np.concatenate([np.maximum(oc.conv2d(x[e * n:(e + 1) * n] * alpha[e], k, (1, 1), "same") * gamma[e] + bias[e], 0)
                           for e in range(E)])
    np.testing.assert_allclose(y, loop, rtol=1e-12, atol=1e-12)

    # Flipout: zero perturbation -> mean conv; Reparameterization: eps = 0 -> mean conv
    mu, rho, eps = k, np.full(k.shape, -30.0), rng.standard_normal(k.shape)
    s_in, s_out = np.sign(rng.standard_normal((B, C))), np.sign(rng.standard_normal((B, K)))
    np.testing.assert_allclose(oc.flipout_conv2d(x, mu, rho, eps, s_in, s_out)[0], oc.conv2d(x, mu), atol=1e-4)  # sigma >= 1e-7
    np.testing.assert_allclose(oc.reparam_conv2d(x, mu, rho - 0, np.zeros_like(eps))[0], oc.conv2d(x, mu), atol=1e-12)
    # Flipout == per-example loop with W_b = mu + (s_in[b] outer s_out[b]) * delta
    rho2 = rng.standard_normal(k.shape) - 2
    yf, delta = oc.flipout_conv2d(x, mu, rho2, eps, s_in, s_out)
    for b in range(B):
        wb = mu + delta * s_in[b][None, None, :, None] * s_out[b][None, None, None, :]
        np.testing.assert_allclose(yf[b:b + 1], oc.conv2d(x[b:b + 1], wb), rtol=1e-10, atol=1e-10)

    # rank-1 conv == per-example loop
    mu_r, rho_r = rng.standard_normal((E, C)), rng.standard_normal((E, C)) - 1
    mu_s, rho_s = rng.standard_normal((E, K)), rng.standard_normal((E, K)) - 1
    e_r, e_s = rng.standard_normal((B, C)), rng.standard_normal((B, K))
    y1, r, s = oc.rank1_conv2d(x, k, mu_r, rho_r, e_r, mu_s, rho_s, e_s, bias, None, (1, 1), "valid")
    for b in range(B):
        np.testing.assert_allclose(y1[b:b + 1], oc.conv2d(x[b:b + 1] * r[b], k) * s[b] + bias[b // n], rtol=1e-12, atol=1e-12)

    # closed-form gradients vs float64 autograd (through torch conv2d)
    gy = rng.standard_normal(oc.conv2d(x, k, (2, 1), "valid").shape)
    xt2 = torch.tensor(x, requires_grad=True)
    mu_t, rho_t = torch.tensor(mu, requires_grad=True), torch.tensor(rho2, requires_grad=True)
    bt = torch.tensor(bias[0], requires_grad=True)
    w_t = mu_t + (torch.nn.functional.softplus(rho_t) + 1e-7) * torch.tensor(eps)
    yt = torch.relu(tf_.conv2d(xt2.permute(0, 3, 1, 2), w_t.permute(3, 2, 0, 1), stride=(2, 1)).permute(0, 2, 3, 1) + bt)
    (yt * torch.tensor(gy)).sum().backward()
    g = oc.reparam_conv2d_bwd(x, mu, rho2, eps, bias[0], "relu", gy, (2, 1), "valid")
    for a, b_ in ((g["dx"], xt2.grad), (g["dmu"], mu_t.grad), (g["drho"], rho_t.grad), (g["dbias"], bt.grad)):
        np.testing.assert_allclose(a, b_.numpy(), rtol=1e-9, atol=1e-9)

    xt3 = torch.tensor(x, requires_grad=True)
    kt3, at3, gt3, bt3 = (torch.tensor(t, requires_grad=True) for t in (k, alpha, gamma, bias))
    rep = lambda t: t.repeat_interleave(n, 0)[:, None, None, :]  # noqa: E731
    z = tf_.conv2d((xt3 * rep(at3)).permute(0, 3, 1, 2), kt3.permute(3, 2, 0, 1)).permute(0, 2, 3, 1)
    yb = torch.relu(z * rep(gt3) + rep(bt3))
    gy2 = rng.standard_normal(tuple(yb.shape))
    (yb * torch.tensor(gy2)).sum().backward()
    g = oc.batch_ensemble_conv2d_bwd(x, k, alpha, gamma, bias, "relu", gy2)
    for a, b_ in ((g["dx"], xt3.grad), (g["dkernel"], kt3.grad), (g["dalpha"], at3.grad), (g["dgamma"], gt3.grad),
                  (g["dbias"], bt3.grad)):
        np.testing.assert_allclose(a, b_.numpy(), rtol=1e-9, atol=1e-9)

    # Flipout conv and rank-1 conv: closed-form gradients vs float64 autograd
    sp = lambda t: torch.nn.functional.softplus(t) + 1e-7  # noqa: E731
    conv = lambda a, w: tf_.conv2d(a.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1)).permute(0, 2, 3, 1)  # noqa: E731
    xt4 = torch.tensor(x, requires_grad=True)
    mu4, rho4, b4 = torch.tensor(mu, requires_grad=True), torch.tensor(rho2, requires_grad=True), torch.tensor(bias[0], requires_grad=True)
    si, so = torch.tensor(s_in)[:, None, None, :], torch.tensor(s_out)[:, None, None, :]
    yf4 = torch.relu(conv(xt4, mu4) + conv(xt4 * si, sp(rho4) * torch.tensor(eps)) * so + b4)
    gy4 = rng.standard_normal(tuple(yf4.shape))
    (yf4 * torch.tensor(gy4)).sum().backward()
    g = oc.flipout_conv2d_bwd(x, mu, rho2, eps, s_in, s_out, bias[0], "relu", gy4)
    for a, b_ in ((g["dx"], xt4.grad), (g["dmu"], mu4.grad), (g["drho"], rho4.grad), (g["dbias"], b4.grad)):
        np.testing.assert_allclose(a, b_.numpy(), rtol=1e-9, atol=1e-9)

    xt5, kt5 = torch.tensor(x, requires_grad=True), torch.tensor(k, requires_grad=True)
    mr, rr, ms, rs, bb = (torch.tensor(t, requires_grad=True) for t in (mu_r, rho_r, mu_s, rho_s, bias))
    rep1 = lambda t: t.repeat_interleave(n, 0)  # noqa: E731
    r_t = torch.clamp(rep1(mr) + rep1(sp(rr)) * torch.tensor(e_r), -10, 10)[:, None, None, :]
    s_t = torch.clamp(rep1(ms) + rep1(sp(rs)) * torch.tensor(e_s), -10, 10)[:, None, None, :]
    y5 = torch.relu(conv(xt5 * r_t, kt5) * s_t + rep1(bb)[:, None, None, :])
    gy5 = rng.standard_normal(tuple(y5.shape))
    (y5 * torch.tensor(gy5)).sum().backward()
    g = oc.rank1_conv2d_bwd(x, k, mu_r, rho_r, e_r, mu_s, rho_s, e_s, bias, "relu", gy5)
    for a, b_ in ((g["dx"], xt5.grad), (g["dkernel"], kt5.grad), (g["dmu_r"], mr.grad), (g["drho_r"], rr.grad),
                  (g["dmu_s"], ms.grad), (g["drho_s"], rs.grad), (g["dbias"], bb.grad)):
        np.testing.assert_allclose(a, b_.numpy(), rtol=1e-9, atol=1e-9)


# ------------------------------------------------------------------------------------------------------------------
# Reference-generated fixture: tests/golden/reference_vectors.npz holds outputs of the UNMODIFIED reference files
# edward2/jax/nn/{dense,random_feature,normalization,utils}.py executed under NumPy-backed jax / flax stand-ins
# (tests/golden/make_reference_golden.py).  These tests pin the oracle to the reference itself.
@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLD, "reference_vectors.npz"))


REF_TOL = 1e-12  # both sides are float64 evaluations of the same formulas


def _close(a, b, tol=REF_TOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), np.abs(a - b).max()


@pytest.mark.parametrize("tag", ["be2d", "be3d"])
def test_reference_batch_ensemble(ref, tag):
    """jax/nn/dense.py:238-272 run as shipped; N-D inputs are [E*n, ..., I] with the member axis leading."""
    x = ref[f"{tag}_x"]
    E = ref[f"{tag}_alpha"].shape[0]
    # the oracle takes member-major 2-D rows: [E, n, ..., I] -> [E * n * ..., I] keeps each member's rows contiguous
    xe = x.reshape((E, -1) + x.shape[1:])
    x2 = xe.reshape(E, -1, x.shape[-1]).reshape(-1, x.shape[-1])
    y, _ = od.batch_ensemble_fwd(x2, ref[f"{tag}_kernel"], ref[f"{tag}_alpha"], ref[f"{tag}_gamma"], ref[f"{tag}_bias"], "relu")
    _close(y.reshape(ref[f"{tag}_y"].shape), ref[f"{tag}_y"])


@pytest.mark.parametrize("tag,scale", [("rff_unit", 1.0), ("rff_auto", None)])
def test_reference_random_fourier_features(ref, tag, scale):
    """jax/nn/random_feature.py:183-214."""
    D = ref[f"{tag}_kernel"].shape[1]
    phi, _ = og.rff(ref[f"{tag}_x"], ref[f"{tag}_kernel"], ref[f"{tag}_bias"], np.sqrt(2.0 / D) if scale is None else scale)
    _close(phi, ref[f"{tag}_phi"])


@pytest.mark.parametrize("tag,momentum,likelihood", [("cov_exact", None, "gaussian"), ("cov_mom", 0.9, "gaussian"),
                                                     ("cov_logistic", None, "binary_logistic"),
                                                     ("cov_poisson", 0.99, "poisson")])
def test_reference_laplace_covariance(ref, tag, momentum, likelihood):
    """jax/nn/random_feature.py:307-309 (initial precision), :340-362 (update), :393-404 (predictive covariance)."""
    ridge = float(ref[f"{tag}_ridge"])
    D = ref["cov_phi1"].shape[1]
    _close(ridge * np.eye(D), ref[f"{tag}_p0"])
    p1 = og.precision_update_jax(ref[f"{tag}_p0"], ref["cov_phi1"], momentum, ref["cov_logits1"], likelihood)
    _close(p1, ref[f"{tag}_p1"])
    p2 = og.precision_update_jax(p1, ref["cov_phi2"], momentum, ref["cov_logits2"], likelihood)
    _close(p2, ref[f"{tag}_p2"])
    _close(og.predictive_covariance_jax(ref["cov_phit"], p2, ridge, True), ref[f"{tag}_var"], 1e-10)
    _close(og.predictive_covariance_jax(ref["cov_phit"], p2, ridge, False), ref[f"{tag}_cov"], 1e-10)
    # the TF form of the same quantity (tensorflow/layers/random_feature.py:447-459,482-485) for momentum-free updates:
    # TF keeps P without the ridge and adds it at inversion time
    if momentum is None:
        sigma = og.covariance_tf(p2 - ridge * np.eye(D), ridge)
        _close(og.predictive_covariance_tf(ref["cov_phit"], sigma, ridge), ref[f"{tag}_cov"], 1e-9)
        _close(og.predictive_variance(ref["cov_phit"], sigma, ridge), ref[f"{tag}_var"], 1e-9)


def test_reference_rfgp_end_to_end(ref):
    """jax/nn/random_feature.py:112-140: LayerNorm -> RFF -> Dense, two precision updates, predictive variance."""
    D = ref["gp_rff_kernel"].shape[1]
    scale = np.sqrt(2.0 / D)

    def feats(x):
        return og.rff(og.layer_norm(x, eps=1e-6), ref["gp_rff_kernel"], ref["gp_rff_bias"], scale)[0]

    pa, pb, pt = feats(ref["gp_xa"]), feats(ref["gp_xb"]), feats(ref["gp_xt"])
    _close(pt, ref["gp_features"], 1e-11)
    _close(og.gp_output(pa, ref["gp_out_kernel"], ref["gp_out_bias"]), ref["gp_logits_a"], 1e-11)
    p = og.precision_update_jax(og.precision_update_jax(np.eye(D), pa), pb)
    _close(p, ref["gp_precision"], 1e-11)
    _close(og.gp_output(pt, ref["gp_out_kernel"], ref["gp_out_bias"]), ref["gp_logits"], 1e-11)
    _close(og.predictive_covariance_jax(pt, p, 1.0, True), ref["gp_var"], 1e-10)
    _close(og.predictive_covariance_jax(pt, p, 1.0, False), ref["gp_cov"], 1e-10)


@pytest.mark.parametrize("tag", ["sn1", "sn3", "sn_small"])
def test_reference_spectral_normalization(ref, tag):
    """jax/nn/normalization.py:123-185 over flax Dense: power iterations, sigma, w / max(sigma / c, 1)."""
    w, u0, v0 = ref[f"{tag}_w"], ref[f"{tag}_u0"], ref[f"{tag}_v0"]
    it, c = int(ref[f"{tag}_iters"]), float(ref[f"{tag}_c"])
    w_hat, u1, v1, _ = og.spectral_norm_jax(w, u0, v0, it, c, training=True)
    _close(u1, ref[f"{tag}_u1"], 1e-11)
    _close(v1, ref[f"{tag}_v1"], 1e-11)
    _close(w_hat, ref[f"{tag}_what"], 1e-11)
    _close(ref[f"{tag}_x"] @ w_hat + ref[f"{tag}_b"], ref[f"{tag}_y"], 1e-11)
    w_eval, u2, v2, _ = og.spectral_norm_jax(w, u1, v1, it, c, training=False)
    assert np.array_equal(u2, np.asarray(u1).reshape(-1)) and np.array_equal(v2, np.asarray(v1).reshape(-1))
    _close(w_eval, ref[f"{tag}_what_eval"], 1e-11)
    if tag == "sn_small":  # spectral norm already below the bound: the kernel is returned unchanged
        _close(w_hat, w)


def test_reference_mean_field_logits(ref):
    """jax/nn/utils.py:54-101."""
    lg, var = ref["mf_logits"], ref["mf_var"]
    for lik in ("logistic", "binary_logistic", "poisson"):
        _close(og.mean_field_logits(lg, var, 0.7, lik), ref[f"mf_{lik}"])
    _close(og.mean_field_logits(lg, None, 3.0), ref["mf_nocov"])
    _close(og.mean_field_logits(lg, var, -1.0), ref["mf_negative"])


@pytest.mark.parametrize("additive", [False, True])
@pytest.mark.parametrize("sampled_bias", [False, True])
def test_rank1_variants_match_autograd_and_member_loop(additive, sampled_bias):
    """dense.py:1096-1144 with use_additive_perturbation (:1126-1127), custom clip bounds (:1108-1110) and the
    per-example sampled bias (:1131-1139): the oracle's forward equals a literal per-example loop, and its closed-form
    gradients equal float64 autograd of an op-for-op torch restatement."""
    rng = np.random.default_rng(5)
    E, n, I, O = 2, 4, 5, 6
    B = E * n
    clip = (-0.8, 1.3)
    x, w = rng.standard_normal((B, I)), rng.standard_normal((I, O))
    mu_r, rho_r = rng.standard_normal((E, I)), rng.standard_normal((E, I)) - 1
    mu_s, rho_s = rng.standard_normal((E, O)), rng.standard_normal((E, O)) - 1
    eps_r, eps_s, eps_b = rng.standard_normal((B, I)), rng.standard_normal((B, O)), rng.standard_normal((B, O))
    bias, rho_b = rng.standard_normal((E, O)), (rng.standard_normal((E, O)) - 1 if sampled_bias else None)
    gy = rng.standard_normal((B, O))
    kw = dict(additive=additive, clip=clip, rho_b=rho_b, eps_b=eps_b if sampled_bias else None)
    y, _, _, _ = od.rank1_fwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, "relu", **kw)
    rows = []
    for e in range(E):
        for i in range(n):
            b = e * n + i
            r = np.clip(mu_r[e] + od.stddev(rho_r[e]) * eps_r[b], *clip)
            s = np.clip(mu_s[e] + od.stddev(rho_s[e]) * eps_s[b], *clip)
            bb = bias[e] + (od.stddev(rho_b[e]) * eps_b[b] if sampled_bias else 0.0)
            pre = ((x[b] + r) @ w + s) if additive else ((x[b] * r) @ w) * s
            rows.append(np.maximum(pre + bb, 0.0))
    np.testing.assert_allclose(y, np.stack(rows), rtol=1e-12, atol=1e-12)

    T = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)  # noqa: E731
    tx, tw, tmr, trr, tms, trs, tb = (T(a) for a in (x, w, mu_r, rho_r, mu_s, rho_s, bias))
    trb = T(rho_b) if sampled_bias else None
    sp = lambda r: torch.nn.functional.softplus(r) + 1e-7  # noqa: E731
    rep = lambda a: a.repeat_interleave(n, 0)  # noqa: E731
    r = torch.clamp(rep(tmr) + rep(sp(trr)) * torch.tensor(eps_r), *clip)
    s = torch.clamp(rep(tms) + rep(sp(trs)) * torch.tensor(eps_s), *clip)
    pre = ((tx + r) @ tw + s) if additive else ((tx * r) @ tw) * s
    pre = pre + rep(tb) + (rep(sp(trb)) * torch.tensor(eps_b) if sampled_bias else 0.0)
    torch.relu(pre).backward(torch.tensor(gy))
    g = od.rank1_bwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, "relu", gy, **kw)
    for name, t in (("dx", tx), ("dw", tw), ("dmu_r", tmr), ("drho_r", trr), ("dmu_s", tms), ("drho_s", trs),
                    ("dbias", tb)):
        np.testing.assert_allclose(g[name], t.grad.numpy(), rtol=1e-10, atol=1e-12, err_msg=name)
    if sampled_bias:
        np.testing.assert_allclose(g["drho_b"], trb.grad.numpy(), rtol=1e-10, atol=1e-12)


def test_per_example_noise_is_a_function_of_the_global_row():
    """ed2_noise row mapping (include/ed2b200.h): the rows a rank draws under the member-major shard
    are exactly the matching rows of the single-device global tensor, for every world size."""
    E, n, cols = 4, 12, 7
    full = op.per_example("normal", 9, 4, E * n, cols)
    np.testing.assert_array_equal(full, op.normal(9, 4, E * n * cols).reshape(E * n, cols))
    for world in (1, 2, 3, 4):
        per = n // world
        for rank in range(world):
            local = op.per_example("normal", 9, 4, E * per, cols, rank * per, per, n)
            want = full.reshape(E, n, cols)[:, rank * per:(rank + 1) * per].reshape(E * per, cols)
            np.testing.assert_array_equal(local, want)
    # ungrouped (Flipout signs): a rank's rows are a contiguous block of the global batch
    s_full = op.per_example("sign", 3, 2, 24, 5)
    np.testing.assert_array_equal(op.per_example("sign", 3, 2, 8, 5, 16, 0, 0), s_full[16:24])


def test_heteroscedastic_oracle_gradients_and_identities():
    """oracle/heteroscedastic.py: closed-form gradients == float64 autograd of the reference recipe
    (heteroscedastic.py:178-226, 749-785); zero scale -> softmax(loc); probabilities sum to one; the HetSNGP variance
    mixing (hetsngp.py:160-168) reduces to the heteroscedastic scale when sngp_var_weight = 0."""
    from oracle import heteroscedastic as oh

    rng = np.random.default_rng(11)
    B, S, K, Fn, T = 5, 40, 6, 3, 0.7
    loc, fac, draw = rng.standard_normal((B, K)), rng.standard_normal((B, K * Fn)) * 0.5, rng.standard_normal((B, K))
    z, e = rng.standard_normal((B, S, Fn)), rng.standard_normal((B, S, K))
    logits, pm, var = oh.mc_softmax_fa(loc, fac, draw, z, e, T)
    np.testing.assert_allclose(pm.sum(-1), 1.0, rtol=1e-12)
    assert (var >= 0).all()
    lt, ft, dt_ = (torch.tensor(t, requires_grad=True) for t in (loc, fac, draw))
    diag = torch.nn.functional.softplus(dt_) + 1e-3
    lat = lt[:, None, :] + torch.einsum("ijk,iak->iaj", ft.reshape(B, K, Fn), torch.tensor(z)) + torch.tensor(e) * diag[:, None, :]
    pt = torch.softmax(lat / T, -1).mean(1)
    lg = torch.log(torch.clamp(pt, 1e-7, 1.0))
    g = rng.standard_normal((B, K))
    (lg * torch.tensor(g)).sum().backward()
    np.testing.assert_allclose(logits, lg.detach().numpy(), rtol=1e-12)
    got = oh.mc_softmax_fa_bwd(loc, fac, draw, z, e, g, T)
    for a, b_ in ((got["dloc"], lt.grad), (got["dfac"], ft.grad), (got["ddiag_raw"], dt_.grad)):
        np.testing.assert_allclose(a, b_.numpy(), rtol=1e-9, atol=1e-12)
    # shared samples: batch dimension 1 broadcasts
    l2, _, _ = oh.mc_softmax_fa(loc, fac, draw, z[:1], e[:1], T)
    l3, _, _ = oh.mc_softmax_fa(loc, fac, draw, np.repeat(z[:1], B, 0), np.repeat(e[:1], B, 0), T)
    np.testing.assert_allclose(l2, l3, rtol=1e-12)
    # no noise -> plain softmax; variance mixing with zero SNGP weight == heteroscedastic scale
    l0, p0, _ = oh.mc_softmax_fa(loc, np.zeros_like(fac), np.full_like(draw, -800.0), z * 0, e * 0, 1.0)
    ref = np.exp(loc - loc.max(-1, keepdims=True))
    np.testing.assert_allclose(p0, ref / ref.sum(-1, keepdims=True), rtol=1e-12)
    np.testing.assert_allclose(oh.effective_diag(draw, rng.random(B), 1.0, 0.0), oh.effective_diag(draw), rtol=1e-12)
    # native noise layout: z and e are disjoint 8-aligned slices of one stream
    zz, ee = oh.native_noise(5, 7, 3, 4, Fn, K)
    assert zz.shape == (3, 4, Fn) and ee.shape == (3, 4, K)
    flat = __import__("oracle.philox", fromlist=["normal"]).normal(5, 7, 3 * 4 * 16).reshape(3, 4, 16)
    np.testing.assert_array_equal(zz, flat[:, :, :Fn])
    np.testing.assert_array_equal(ee, flat[:, :, 8:8 + K])


def test_recurrent_oracle_bptt_matches_autograd():
    """oracle/recurrent.py: the closed-form backpropagation through time of the LSTM cell
    (recurrent.py:163-183, gate order i | f | c | o) == float64 autograd."""
    from oracle import recurrent as orc

    rng = np.random.default_rng(21)
    T, B, I, H = 4, 3, 5, 6
    xs, h0, c0 = rng.standard_normal((T, B, I)), rng.standard_normal((B, H)), rng.standard_normal((B, H))
    W, U, b = rng.standard_normal((I, 4 * H)) * 0.4, rng.standard_normal((H, 4 * H)) * 0.4, rng.standard_normal(4 * H) * 0.1
    hs, cT, caches = orc.lstm_sequence(xs, h0, c0, W, U, b)
    dhs, dcl = rng.standard_normal((T, B, H)), rng.standard_normal((B, H))
    g = orc.lstm_sequence_bwd(caches, W, U, dhs, dcl)
    t = {k: torch.tensor(v, requires_grad=True) for k, v in dict(xs=xs, h0=h0, c0=c0, W=W, U=U, b=b).items()}
    h, c, loss = t["h0"], t["c0"], 0.0
    for s in range(T):
        z = t["xs"][s] @ t["W"] + h @ t["U"] + t["b"]
        i, f, gg, o = torch.sigmoid(z[:, :H]), torch.sigmoid(z[:, H:2 * H]), torch.tanh(z[:, 2 * H:3 * H]), torch.sigmoid(z[:, 3 * H:])
        c = f * c + i * gg
        h = o * torch.tanh(c)
        loss = loss + (h * torch.tensor(dhs[s])).sum()
    (loss + (c * torch.tensor(dcl)).sum()).backward()
    np.testing.assert_allclose(hs[-1], h.detach().numpy(), rtol=1e-12)
    for key, ref in (("dxs", t["xs"].grad), ("dh0", t["h0"].grad), ("dc0", t["c0"].grad), ("dW", t["W"].grad),
                     ("dU", t["U"].grad), ("db", t["b"].grad)):
        np.testing.assert_allclose(g[key], ref.numpy(), rtol=1e-9, atol=1e-12)
    # the two-branch driver (LSTMCellFlipout / LSTMCellRank1 oracles) with plain linear branches reproduces the same BPTT
    bx, bh = (lambda x: x @ W + b), (lambda h: h @ U)
    gx_ = lambda x, dz: dict(dx=dz @ W.T, dW=x.T @ dz, db=dz.sum(0))  # noqa: E731
    gh_ = lambda h, dz: dict(dx=dz @ U.T, dU=h.T @ dz)  # noqa: E731
    hs2, dxs2, gx2, gh2 = orc.two_branch_sequence(xs, h0, c0, bx, bh, gx_, gh_, dhs)
    g0 = orc.lstm_sequence_bwd(caches, W, U, dhs)
    np.testing.assert_allclose(hs2, hs, rtol=1e-12)
    for a, b_ in ((dxs2, g0["dxs"]), (gx2["dW"], g0["dW"]), (gh2["dU"], g0["dU"]), (gx2["db"], g0["db"])):
        np.testing.assert_allclose(a, b_, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("tag", ["het", "het_shared"])
def test_reference_mc_softmax_dense_fa(ref, tag):
    """jax/nn/heteroscedastic_lib.py:43-336 (MCSoftmaxDenseFA) run as shipped, with the draws of its jax.random.normal calls
    recorded: loc / scale / diag Dense layers, latent = loc + V z + diag * n, softmax / T, mean over samples, clip, log.
    (The JAX class draws ONE diagonal-noise scalar per (example, sample), shape [B, S, 1] (:186-188); the TF class draws one
    per class (heteroscedastic.py:676-707) — the oracle takes the noise as an explicit [B, S, K] argument, so both are the
    same formula.)"""
    from oracle import heteroscedastic as oh

    x = ref[f"{tag}_x"]
    lin = lambda n: x @ ref[f"{tag}_{n}_kernel"] + ref[f"{tag}_{n}_bias"]  # noqa: E731
    loc, fac, draw = lin("loc_layer"), lin("scale_layer"), lin("diag_layer")
    K = loc.shape[1]
    z = ref[f"{tag}_factor_noise"]
    e = np.broadcast_to(ref[f"{tag}_diag_noise"], ref[f"{tag}_diag_noise"].shape[:2] + (K,))
    logits, probs, _ = oh.mc_softmax_fa(loc, fac, draw, z, e, float(ref[f"{tag}_temperature"]), 1e-7)
    _close(probs, ref[f"{tag}_probs"])
    _close(logits, ref[f"{tag}_logits"])
    _close(logits, ref[f"{tag}_log_probs"])


def test_reference_mc_sigmoid_dense_fa(ref):
    """jax/nn/heteroscedastic_lib.py:338-594 (MCSigmoidDenseFA) run as shipped with its draws recorded: sigmoid link, clip
    from below, inverse-sigmoid logits with the second clip; plus the closed-form gradient against float64 autograd."""
    from oracle import heteroscedastic as oh

    x = ref["hsig_x"]
    lin = lambda n: x @ ref[f"hsig_{n}_kernel"] + ref[f"hsig_{n}_bias"]  # noqa: E731
    loc, fac, draw = lin("loc_layer"), lin("scale_layer"), lin("diag_layer")
    K = loc.shape[1]
    z = ref["hsig_factor_noise"]
    e = np.broadcast_to(ref["hsig_diag_noise"], ref["hsig_diag_noise"].shape[:2] + (K,))
    T = float(ref["hsig_temperature"])
    logits, log_probs, probs = oh.mc_sigmoid_fa(loc, fac, draw, z, e, T, 1e-7)
    _close(probs, ref["hsig_probs"])
    _close(log_probs, ref["hsig_log_probs"])
    _close(logits, ref["hsig_logits"])
    lt, ft, dt_ = (torch.tensor(t, requires_grad=True) for t in (loc, fac, draw))
    B, Fn = loc.shape[0], z.shape[-1]
    diag = torch.nn.functional.softplus(dt_) + 1e-3
    lat = lt[:, None, :] + torch.einsum("ijk,iak->iaj", ft.reshape(B, K, Fn), torch.tensor(z)) + torch.tensor(e.copy()) * diag[:, None, :]
    pm = torch.sigmoid(lat / T).mean(1)
    lg = torch.log(torch.clamp(pm, min=1e-7)) - torch.log(1 - torch.clamp(pm, 1e-7, 1 - 1e-7))
    g = np.random.default_rng(0).standard_normal(loc.shape)
    (lg * torch.tensor(g)).sum().backward()
    got = oh.mc_sigmoid_fa_bwd(loc, fac, draw, z, e, g, T)
    for a, b_ in ((got["dloc"], lt.grad), (got["dfac"], ft.grad), (got["ddiag_raw"], dt_.grad)):
        np.testing.assert_allclose(a, b_.numpy(), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("tag", ["hsngp", "hsngpbe"])
def test_reference_mc_sigmoid_dense_fa_sngp(ref, tag):
    """jax/nn/random_feature.py:407-735 (MCSigmoidDenseFASNGP / ...BE) run as shipped with its draws recorded: a training
    call (GP location, precision update, Monte-Carlo sigmoid head) and an evaluation call whose diagonal scale mixes in
    the GP variance — restated from the oracle's pieces."""
    from oracle import dense as od, heteroscedastic as oh, sngp as osn

    D = ref[f"{tag}_rff_kernel"].shape[1]

    def lin(name, x):
        k, b = ref[f"{tag}_{name}_kernel"], ref[f"{tag}_{name}_bias"]
        if tag == "hsngpbe":
            return od.batch_ensemble_fwd(x, k, ref[f"{tag}_{name}_fast_weight_alpha"],
                                         ref[f"{tag}_{name}_fast_weight_gamma"], b, None)[0]
        return x @ k + b

    def head(x):
        phi, _ = osn.rff(osn.layer_norm(x, eps=1e-6), ref[f"{tag}_rff_kernel"], ref[f"{tag}_rff_bias"], 1.0)
        return phi, osn.gp_output(phi, ref[f"{tag}_out_kernel"], ref[f"{tag}_out_bias"])

    def noise(prefix, K):
        e = ref[f"{tag}_{prefix}diag_noise"]
        return ref[f"{tag}_{prefix}factor_noise"], np.broadcast_to(e, e.shape[:2] + (K,))

    x, xt = ref[f"{tag}_x"], ref[f"{tag}_xt"]
    phi, loc = head(x)
    K = loc.shape[1]
    z, e = noise("", K)
    logits, log_probs, probs = oh.mc_sigmoid_fa(loc, lin("scale_layer", x), lin("diag_layer", x), z, e, 0.8, 1e-7)
    _close(probs, ref[f"{tag}_probs"])
    _close(log_probs, ref[f"{tag}_log_probs"])
    _close(logits, ref[f"{tag}_logits"])
    precision = osn.precision_update_jax(np.eye(D), phi, None)
    _close(precision, ref[f"{tag}_precision"])
    phit, loct = head(xt)
    var = osn.predictive_covariance_jax(phit, precision, 1.0, diagonal_only=True)
    _close(var, ref[f"{tag}_eval_var"])
    z, e = noise("eval_", K)
    logits, log_probs, probs = oh.mc_sigmoid_fa(loct, lin("scale_layer", xt), lin("diag_layer", xt), z, e, 0.8, 1e-7,
                                                sngp_var=var, w_het=0.7, w_sngp=1.3)
    _close(probs, ref[f"{tag}_eval_probs"])
    _close(log_probs, ref[f"{tag}_eval_log_probs"])
    _close(logits, ref[f"{tag}_eval_logits"])


# -------------------------------------------------------------------------------------- the reference's TF classes, executed
# edward2/tensorflow/layers/dense.py (+ layers/utils.py add_weight, initializers.py TrainableNormal, regularizers.py
# NormalKLDivergence, constraints.py Softplus) run UNMODIFIED under NumPy-backed tensorflow / tfp stand-ins with every random
# draw recorded (tests/golden/make_reference_golden_tf.py).  These tests pin oracle/dense.py to those classes.
@pytest.fixture(scope="module")
def ref_tf():
    return np.load(os.path.join(GOLD, "reference_vectors_tf.npz"))


def test_reference_tf_dense_reparameterization(ref_tf):
    """dense.py:33-83: kernel = mean + (softplus(raw stddev) + 1e-7) * eps re-sampled by call_weights, Keras Dense call (2-D
    matmul and the tensordot branch for 3-D inputs); the layer's loss is NormalKLDivergence of the kernel variable."""
    r = ref_tf
    _close(od.stddev(r["rho_rp"]), r["rp_stddev"])
    y, w = od.reparam_fwd(r["rp_x"], r["rp_mu"], r["rho_rp"], r["rp_eps"], r["rp_bias"], "relu")
    _close(w, r["rp_kernel"])
    _close(y, r["rp_y"])
    _close(od.normal_kl(r["rp_mu"], r["rho_rp"]), r["rp_kl"])
    m, s, f = r["rp_kl_prior_args"]
    _close(od.normal_kl(r["rp_mu"], r["rho_rp"], prior_mean=m, prior_std=s, scale_factor=f), r["rp_kl_prior"])
    x3 = r["rp_x3"]
    y3, _ = od.reparam_fwd(x3.reshape(-1, x3.shape[-1]), r["rp_mu"], r["rho_rp"], r["rp_eps3"], r["rp_bias"], "relu")
    _close(y3.reshape(r["rp_y3"].shape), r["rp_y3"])


def test_reference_tf_dense_flipout(ref_tf):
    """dense.py:255-285 as executed: int32 sign_input in {-1, +1}, FLOAT sign_output = 2 U[0, 2) - 1 in [-1, 3) (the
    reference's dtype quirk), perturbation = kernel - mean, 2-D matmul and 3-D tensordot branches."""
    r = ref_tf
    assert set(np.unique(r["fo_sign_in"])) <= {-1.0, 1.0}
    so = r["fo_sign_out"]
    assert so.min() >= -1.0 and so.max() < 3.0 and len(np.unique(so)) > 2  # real-valued, not a sign tensor
    y, _ = od.flipout_fwd(r["fo_x"], r["fo_mu"], r["fo_rho"], r["fo_eps"], r["fo_sign_in"], so, r["fo_bias"], "relu")
    _close(y, r["fo_y"])
    x3 = r["fo_x3"]
    flat = lambda a: a.reshape(-1, a.shape[-1])  # noqa: E731
    y3, _ = od.flipout_fwd(flat(x3), r["fo_mu"], r["fo_rho"], r["fo_eps3"], flat(r["fo_sign_in3"]), flat(r["fo_sign_out3"]),
                           r["fo_bias"], "relu")
    _close(y3.reshape(r["fo_y3"].shape), r["fo_y3"])


def test_reference_tf_dense_batch_ensemble(ref_tf):
    """dense.py:551-584, TF class: rank 1 (reshape to [E, n, I]) and rank 3 (tile / reshape of [rank, E, dim] fast weights)."""
    r = ref_tf
    y, _ = od.batch_ensemble_fwd(r["be_x"], r["be_kernel"], r["be_alpha"], r["be_gamma"], r["be_bias"], "relu")
    _close(y, r["be_y"])
    y3 = od.batch_ensemble_rank_fwd(r["be_r3_x"], r["be_r3_kernel"], r["be_r3_alpha"], r["be_r3_gamma"], r["be_r3_bias"], "relu")
    _close(y3, r["be_r3_y"])


@pytest.mark.parametrize("tag", ["r1", "r1_add", "r1_bias"])
def test_reference_tf_dense_rank1(ref_tf, tag):
    """dense.py:1096-1144 as executed: distribution.sample(examples_per_model) -> [n, E, dim], clip, transpose to
    [E, n, dim] (recorded draws re-laid to member-major rows), multiplicative / additive perturbations, deterministic /
    per-example sampled ensemble bias; the layer's losses are the KLs of alpha and gamma (and the sampled bias)."""
    r = ref_tf
    lo, hi = r[f"{tag}_clip"]
    sampled_bias = f"{tag}_rho_b" in r.files
    y, _, fr, fs = od.rank1_fwd(r[f"{tag}_x"], r[f"{tag}_kernel"], r[f"{tag}_mu_r"], r[f"{tag}_rho_r"], r[f"{tag}_eps_r"],
                                r[f"{tag}_mu_s"], r[f"{tag}_rho_s"], r[f"{tag}_eps_s"], r[f"{tag}_bias"], "relu",
                                additive=(tag == "r1_add"), clip=(lo, hi),
                                rho_b=r[f"{tag}_rho_b"] if sampled_bias else None,
                                eps_b=r[f"{tag}_eps_b"] if sampled_bias else None)
    _close(y, r[f"{tag}_y"])
    if tag == "r1":  # the narrow clip bounds of this case are active
        assert (fr == lo).any() or (fr == hi).any() or (fs == lo).any() or (fs == hi).any()
    kls = [od.normal_kl(r[f"{tag}_mu_r"], r[f"{tag}_rho_r"]), od.normal_kl(r[f"{tag}_mu_s"], r[f"{tag}_rho_s"])]
    got = sorted(float(v) for v in r[f"{tag}_kl"])
    for k in kls:
        assert min(abs(k - g) for g in got) <= 1e-12 * max(1.0, abs(k))


@pytest.mark.parametrize("tag", ["sn1", "sn3", "sn_small"])
def test_reference_tf_spectral_normalization(ref_tf, tag):
    """normalization.py:284-395, TF class as executed: power iteration from the stored (u, v), sigma, the kernel normalised
    IN PLACE (self.w aliases layer.kernel), the wrapped Dense output; then an evaluation call (no power iteration) on the
    already normalised kernel."""
    r = ref_tf
    it, c = int(r[f"{tag}_iters"]), float(r[f"{tag}_c"])
    w1, u1, v1, _ = og.spectral_norm_tf(r[f"{tag}_w0"], r[f"{tag}_u0"], r[f"{tag}_v0"], it, c, True)
    _close(u1, r[f"{tag}_u1"])
    _close(v1, r[f"{tag}_v1"])
    _close(w1, r[f"{tag}_w1"])
    _close(r[f"{tag}_x"] @ w1 + r[f"{tag}_b"], r[f"{tag}_y"])
    w2, _, _, _ = og.spectral_norm_tf(w1, u1, v1, it, c, False)
    _close(w2, r[f"{tag}_w2"])
    _close(r[f"{tag}_x"] @ w2 + r[f"{tag}_b"], r[f"{tag}_y_eval"])
    if tag == "sn_small":  # a kernel already below the bound is left unchanged (normalization.py:384-386)
        np.testing.assert_array_equal(r[f"{tag}_w1"], r[f"{tag}_w0"])


@pytest.mark.parametrize("tag", ["cov_exact", "cov_mom", "cov_logistic", "cov_poisson"])
def test_reference_tf_laplace_covariance(ref_tf, tag):
    """random_feature.py:284-549, TF class as executed: two training calls (exact-sum / momentum updates, Gaussian /
    logistic / Poisson likelihoods), then the evaluation call that inverts ridge I + P and returns ridge Phi Sigma Phi^T."""
    r = ref_tf
    momentum, ridge = r[f"{tag}_args"]
    lik = {"cov_exact": "gaussian", "cov_mom": "gaussian", "cov_logistic": "binary_logistic", "cov_poisson": "poisson"}[tag]
    D = r["cov_phi1"].shape[1]
    p1 = og.precision_update_tf(np.zeros((D, D)), r["cov_phi1"], momentum, r["cov_logits1"], lik)
    _close(p1, r[f"{tag}_p1"])
    p2 = og.precision_update_tf(p1, r["cov_phi2"], momentum, r["cov_logits2"], lik)
    _close(p2, r[f"{tag}_p2"])
    sigma = og.covariance_tf(p2, ridge)
    _close(sigma, r[f"{tag}_sigma"], 1e-10)
    _close(og.predictive_covariance_tf(r["cov_phit"], sigma, ridge), r[f"{tag}_cov"], 1e-10)


def test_reference_tf_rfgp_end_to_end(ref_tf):
    """random_feature.py:33-281, TF class as executed: Keras LayerNormalization (eps 1e-3), frozen random-feature Dense with
    cos activation, sqrt(2 / D) feature scale, output weights + bias variable, exact-sum precision, evaluation covariance;
    the reference's OrthogonalRandomFeatures draws mutually orthogonal directions within a square block."""
    r = ref_tf
    D = r["gp_rff_kernel"].shape[1]

    def head(x):
        xn = og.layer_norm(x, r["gp_ln_gamma"], r["gp_ln_beta"], 1e-3)
        phi, _ = og.rff(xn, r["gp_rff_kernel"], r["gp_rff_bias"], np.sqrt(2.0 / D))
        return phi, og.gp_output(phi, r["gp_out_kernel"], r["gp_out_bias"])

    phi_a, logits_a = head(r["gp_xa"])
    _close(phi_a, r["gp_features_a"])
    _close(logits_a, r["gp_logits_a"])
    precision = og.precision_update_tf(og.precision_update_tf(np.zeros((D, D)), phi_a, -1.0), phi_a, -1.0)
    _close(precision, r["gp_precision"])
    phi_t, logits_t = head(r["gp_xt"])
    _close(phi_t, r["gp_features"])
    _close(logits_t, r["gp_logits"])
    _close(og.predictive_covariance_tf(phi_t, og.covariance_tf(precision, 1.0), 1.0), r["gp_cov"], 1e-10)
    gram = r["orf_gram"]
    assert np.abs(gram - np.diag(np.diag(gram))).max() < 1e-10 * np.abs(gram).max()
    assert r["gp_rff_bias"].min() >= 0.0 and r["gp_rff_bias"].max() <= 2.0 * np.pi


def test_reference_tf_mean_field_logits(ref_tf):
    """layers/utils.py:379-424 as executed: the [B, B] covariance enters through its diagonal."""
    r = ref_tf
    for lik in ("logistic", "binary_logistic", "poisson"):
        _close(og.mean_field_logits(r["mf_logits"], r["mf_cov"], 0.7, lik), r[f"mf_{lik}"])
    _close(og.mean_field_logits(r["mf_logits"], None, 3.0), r["mf_nocov"])
    _close(og.mean_field_logits(r["mf_logits"], r["mf_cov"], -1.0), r["mf_negative"])


def test_reference_tf_conv2d_variants(ref_tf):
    """layers/convolutional.py as executed (NHWC): Conv2DReparameterization ('same', stride 1; :88-98), Conv2DFlipout
    ('valid', stride 2, int32 signs of shape [B, 1, 1, C]; :225-249), Conv2DBatchEnsemble (tile / reshape of the fast
    weights; :681-699) and Conv2DRank1 (per-example fast weights, strides (2, 1); :1968-2020) against oracle/conv.py."""
    from oracle import conv as oc

    r = ref_tf
    y, _ = oc.reparam_conv2d(r["crp_x"], r["crp_mu"], r["crp_rho"], r["crp_eps"], r["crp_bias"], "relu", (1, 1), "same")
    _close(y, r["crp_y"])
    _close(od.normal_kl(r["crp_mu"], r["crp_rho"]), r["crp_kl"])
    assert set(np.unique(r["cfo_sign_in"])) <= {-1.0, 1.0} and set(np.unique(r["cfo_sign_out"])) <= {-1.0, 1.0}
    y, _ = oc.flipout_conv2d(r["cfo_x"], r["cfo_mu"], r["cfo_rho"], r["cfo_eps"], r["cfo_sign_in"], r["cfo_sign_out"],
                             r["cfo_bias"], "relu", (2, 2), "valid")
    _close(y, r["cfo_y"])
    y = oc.batch_ensemble_conv2d(r["cbe_x"], r["cbe_kernel"], r["cbe_alpha"], r["cbe_gamma"], r["cbe_bias"], "relu",
                                 (1, 1), "same")
    _close(y, r["cbe_y"])
    y, _, _ = oc.rank1_conv2d(r["cr1_x"], r["cr1_kernel"], r["cr1_mu_r"], r["cr1_rho_r"], r["cr1_eps_r"], r["cr1_mu_s"],
                              r["cr1_rho_s"], r["cr1_eps_s"], r["cr1_bias"], "relu", (2, 1), "same")
    _close(y, r["cr1_y"])


@pytest.mark.parametrize("tag", ["lrp", "lrp1", "lfo", "lfo1"])
def test_reference_tf_lstm_cell_reparameterization_and_flipout(ref_tf, tag):
    """layers/recurrent.py:30-183 (LSTMCellReparameterization) and :185-358 (LSTMCellFlipout) as executed over three steps,
    both Keras implementations (1: per-gate slices through _compute_carry_and_output; 2: fused): ONE weight sample / sign
    draw per sequence (get_initial_state), reused by every step."""
    from oracle import recurrent as orc

    r = ref_tf
    xs = r[f"{tag}_xs"]
    T, B, _ = xs.shape
    H = r[f"{tag}_mu_u"].shape[0]
    h, c = np.zeros((B, H)), np.zeros((B, H))
    if tag.startswith("lrp"):
        W = orc.reparam_kernel(r[f"{tag}_mu_k"], r[f"{tag}_rho_k"], r[f"{tag}_eps_k"])
        U = orc.reparam_kernel(r[f"{tag}_mu_u"], r[f"{tag}_rho_u"], r[f"{tag}_eps_u"])
        hs, c_last, _ = orc.lstm_sequence(xs, h, c, W, U, r[f"{tag}_bias"])
    else:
        hs = []
        for t in range(T):
            z = od.flipout_fwd(xs[t], r[f"{tag}_mu_k"], r[f"{tag}_rho_k"], r[f"{tag}_eps_k"], r[f"{tag}_sign_in"],
                               r[f"{tag}_sign_out"])[0]
            z = z + od.flipout_fwd(h, r[f"{tag}_mu_u"], r[f"{tag}_rho_u"], r[f"{tag}_eps_u"], r[f"{tag}_rsign_in"],
                                   r[f"{tag}_rsign_out"])[0] + r[f"{tag}_bias"]
            h, c, _ = orc.gates(z, c)
            hs.append(h)
        hs, c_last = np.stack(hs), c
    _close(hs, r[f"{tag}_hs"])
    _close(c_last, r[f"{tag}_c"])
    kls = sorted([od.normal_kl(r[f"{tag}_mu_k"], r[f"{tag}_rho_k"]), od.normal_kl(r[f"{tag}_mu_u"], r[f"{tag}_rho_u"])])
    np.testing.assert_allclose(sorted(r[f"{tag}_kl"]), kls, rtol=1e-12)


@pytest.mark.parametrize("tag", ["lr1", "lr1_add"])
def test_reference_tf_lstm_cell_rank1(ref_tf, tag):
    """layers/recurrent.py:360-830 (LSTMCellRank1) as executed: per-example fast weights of the input and the recurrent
    branch drawn once per sequence ([n, E, dim] -> member-major rows, no clipping), multiplicative and additive
    perturbations, one bias row per member."""
    from oracle import recurrent as orc

    r = ref_tf
    xs = r[f"{tag}_xs"]
    T, B, _ = xs.shape
    E = r[f"{tag}_bias"].shape[0]
    n = B // E
    H = r[f"{tag}_rkernel"].shape[0]
    f = {name: od.rank1_factors(r[f"{tag}_{name}_mu"], r[f"{tag}_{name}_rho"], r[f"{tag}_{name}_eps"], E, None)[0]
         for name in ("alpha", "gamma", "recurrent_alpha", "recurrent_gamma")}
    bias = np.repeat(r[f"{tag}_bias"], n, axis=0)
    h, c, hs = np.zeros((B, H)), np.zeros((B, H)), []
    for t in range(T):
        if tag == "lr1_add":
            z = (xs[t] + f["alpha"]) @ r[f"{tag}_kernel"] + f["gamma"]
            z = z + (h + f["recurrent_alpha"]) @ r[f"{tag}_rkernel"] + f["recurrent_gamma"]
        else:
            z = ((xs[t] * f["alpha"]) @ r[f"{tag}_kernel"]) * f["gamma"]
            z = z + ((h * f["recurrent_alpha"]) @ r[f"{tag}_rkernel"]) * f["recurrent_gamma"]
        h, c, _ = orc.gates(z + bias, c)
        hs.append(h)
    _close(np.stack(hs), r[f"{tag}_hs"])
    _close(c, r[f"{tag}_c"])


@pytest.mark.parametrize("tag", ["hsm", "hsm_shared", "hsg"])
def test_reference_tf_mc_heads(ref_tf, tag):
    """layers/heteroscedastic.py as executed: MCSoftmaxDenseFA (:508; per-example and shared noise) and MCSigmoidDenseFA
    (:1382) — Dense parameter layers, softplus + 1e-3 diagonal, ONE diagonal draw per class, einsum factor noise,
    temperature, clips; the predictive variance is evaluated on the samples of the predictive mean."""
    from oracle import heteroscedastic as oh

    r = ref_tf
    x = r[f"{tag}_x"]
    lin = lambda n: x @ r[f"{tag}_{n}_kernel"] + r[f"{tag}_{n}_bias"]  # noqa: E731
    loc, fac, draw = lin("loc_layer"), lin("scale_layer"), lin("diag_layer")
    z, e = r[f"{tag}_factor_noise"], r[f"{tag}_diag_noise"]
    assert e.shape[0] == (1 if tag == "hsm_shared" else x.shape[0]) and e.shape[2] == loc.shape[1]
    if tag == "hsg":
        logits, log_probs, probs = oh.mc_sigmoid_fa(loc, fac, draw, z, e, 0.7, 1e-7)
        _close(np.clip(probs, 1e-7, 1 - 1e-7), r[f"{tag}_probs"])
    else:
        logits, probs, var = oh.mc_softmax_fa(loc, fac, draw, z, e, 0.7, 1e-7)
        log_probs = logits
        _close(probs, r[f"{tag}_probs"])
        _close(var, r[f"{tag}_var"])
    _close(log_probs, r[f"{tag}_log_probs"])
    _close(logits, r[f"{tag}_logits"])


def test_reference_tf_heteroscedastic_sngp_layer(ref_tf):
    """layers/hetsngp.py:25-245 as executed: location = RandomFeatureGaussianProcess logits, a training call (exact-sum
    precision update, plain diagonal scale) and an evaluation call whose diagonal scale is
    sqrt(het_var_weight diag^2 + sngp_var_weight diag(cov))."""
    from oracle import heteroscedastic as oh

    r = ref_tf
    w_het, w_sngp, T = r["hsn_weights"]
    D = r["hsn_rff_kernel"].shape[1]

    def gp(x):
        xn = og.layer_norm(x, r["hsn_ln_gamma"], r["hsn_ln_beta"], 1e-3)
        phi, _ = og.rff(xn, r["hsn_rff_kernel"], r["hsn_rff_bias"], np.sqrt(2.0 / D))
        return phi, og.gp_output(phi, r["hsn_out_kernel"], r["hsn_out_bias"])

    def scale(x):
        return (x @ r["hsn_scale_layer_kernel"] + r["hsn_scale_layer_bias"],
                x @ r["hsn_diag_layer_kernel"] + r["hsn_diag_layer_bias"])

    x, xt = r["hsn_x"], r["hsn_xt"]
    phi, loc = gp(x)
    fac, draw = scale(x)
    logits, probs, _ = oh.mc_softmax_fa(loc, fac, draw, r["hsn_factor_noise"], r["hsn_diag_noise"], T, 1e-7)
    _close(probs, r["hsn_probs"])
    _close(logits, r["hsn_logits"])
    precision = og.precision_update_tf(np.zeros((D, D)), phi, -1.0)
    _close(precision, r["hsn_precision"])
    sigma = og.covariance_tf(precision, 1.0)
    _close(sigma, r["hsn_eval_sigma"], 1e-10)
    phit, loct = gp(xt)
    var = np.diag(og.predictive_covariance_tf(phit, sigma, 1.0))
    fac, draw = scale(xt)
    logits, probs, _ = oh.mc_softmax_fa(loct, fac, draw, r["hsn_eval_factor_noise"], r["hsn_eval_diag_noise"], T, 1e-7,
                                        sngp_var=var, w_het=w_het, w_sngp=w_sngp)
    _close(probs, r["hsn_eval_probs"], 1e-10)
    _close(logits, r["hsn_eval_logits"], 1e-10)
