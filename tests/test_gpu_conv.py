"""Conv2D stochastic-weight layers on the CUDA path (edward2_b200/layers/convolutional.py; include/ed2b200.h: ed2_im2col,
ed2_col2im, ed2_expand_table, ed2_rank1_factors, ed2_noise_table) against the float64 oracle (oracle/conv.py, following
edward2/tensorflow/layers/convolutional.py:88-98, 225-249, 681-699, 1968-2020): forward and every gradient, fp32 at 1e-5
and bf16 at 2e-2, 'valid' / 'same' padding, strides, injected noise and the native Philox streams."""
import numpy as np
import pytest
import torch

from oracle import conv as oc
from oracle import dense as od
from oracle import philox as op
from tests.util import TOL, rel_err, round_to, to_np

pytestmark = pytest.mark.gpu

GEOMS = [((3, 3), (1, 1), "same"), ((3, 2), (2, 1), "valid"), ((1, 1), (1, 1), "valid"), ((3, 3), (2, 2), "same")]


def _dt(dtype):
    return "bfloat16" if dtype == torch.bfloat16 else "float32"


def _t(a, dtype, cuda, grad=False):
    t = torch.tensor(np.asarray(a), dtype=torch.float32, device=cuda).to(dtype)
    return t.requires_grad_(True) if grad else t


def _normal_init(ed, std, seed):
    return ed.initializers.TrainableNormal(mean_initializer=ed.initializers.RandomNormal(stddev=std, seed=seed),
                                           stddev_initializer=ed.initializers.TruncatedNormal(-2.5, 0.2, seed=seed + 1))


def test_im2col_col2im_match_oracle(cuda):
    """The patch gather and its adjoint, both dtypes, vector and scalar channel counts."""
    from edward2_b200 import functional as F

    rng = np.random.default_rng(0)
    for C in (8, 3):
        for ks, st, pad in GEOMS:
            x = rng.standard_normal((3, 9, 7, C))
            for dtype in (torch.float32, torch.bfloat16):
                xr = round_to(x, dtype)
                xt = _t(xr, dtype, cuda, grad=True)
                cols = F.Im2Col.apply(xt, ks, st, pad, dtype)
                ref = oc.im2col(xr, ks[0], ks[1], st, pad)
                assert tuple(cols.shape) == (ref.shape[0] * ref.shape[1] * ref.shape[2], ref.shape[3])
                np.testing.assert_array_equal(to_np(cols), ref.reshape(cols.shape))
                g = round_to(rng.standard_normal(tuple(cols.shape)), dtype)
                cols.backward(_t(g, dtype, cuda))
                dref = oc.col2im(g, x.shape, ks[0], ks[1], st, pad)
                assert rel_err(to_np(xt.grad), dref) < (1e-6 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("geom", GEOMS[:2])
def test_conv2d_reparameterization(cuda, dtype, geom):
    import edward2_b200 as ed
    from edward2_b200.layers import dense as ld

    ks, st, pad = geom
    B, H, W, C, Fo = 4, 10, 9, 8, 16
    rng = np.random.default_rng(1)
    x = round_to(rng.standard_normal((B, H, W, C)), dtype)
    layer = ed.layers.Conv2DReparameterization(Fo, ks, strides=st, padding=pad, activation="relu", dtype=_dt(dtype), seed=5,
                                               kernel_initializer=_normal_init(ed, 0.2, 3), kernel_regularizer=None,
                                               bias_initializer=ed.initializers.RandomNormal(stddev=0.1, seed=9))
    xt = _t(x, dtype, cuda, grad=True)
    y = layer(xt)
    gy = round_to(rng.standard_normal(tuple(y.shape)), dtype)
    y.backward(_t(gy, dtype, cuda))
    mu, rho = (to_np(t) for t in layer.kernel_initializer.params())
    eps = op.normal(layer.seed, ld.S_KERNEL, mu.size).reshape(mu.shape)
    bias = to_np(layer.bias)
    tol = TOL[dtype]
    y_ref, _ = oc.reparam_conv2d(x, mu, rho, eps, bias, "relu", st, pad)
    assert rel_err(to_np(y), y_ref) < tol
    # bf16: the oracle's gradient with the kernel's own activation pattern
    w = mu + od.stddev(rho) * eps
    gp = gy * (to_np(y) > 0)
    dx, dw = oc.conv2d_bwd(x, w, gp, st, pad)
    gmu, grho = (p.grad for p in layer.kernel_initializer.params())
    assert rel_err(to_np(xt.grad), dx) < tol
    assert rel_err(to_np(gmu), dw) < tol
    assert rel_err(to_np(grho), dw * eps * od.sigmoid(rho)) < tol
    assert rel_err(to_np(layer.bias.grad), gp.sum((0, 1, 2))) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("inject", [True, False])
def test_conv2d_flipout(cuda, dtype, inject):
    import edward2_b200 as ed
    from edward2_b200.layers import dense as ld

    ks, st, pad = (3, 3), (1, 1), "same"
    B, H, W, C, Fo = 6, 8, 8, 8, 24
    rng = np.random.default_rng(2)
    x = round_to(rng.standard_normal((B, H, W, C)), dtype)
    layer = ed.layers.Conv2DFlipout(Fo, ks, strides=st, padding=pad, activation="relu", dtype=_dt(dtype), seed=6,
                                    kernel_initializer=_normal_init(ed, 0.2, 4), kernel_regularizer=None)
    xt = _t(x, dtype, cuda, grad=True)
    with torch.no_grad():
        layer(xt)  # build
    layer._calls = 0
    mu, rho = (to_np(t) for t in layer.kernel_initializer.params())
    if inject:
        eps = rng.standard_normal(mu.shape)
        s_in, s_out = np.sign(rng.standard_normal((B, C))), np.sign(rng.standard_normal((B, Fo)))
        layer.inject(eps=_t(eps, torch.float32, cuda), sign_input=_t(s_in, torch.float32, cuda),
                     sign_output=_t(s_out, torch.float32, cuda))
    else:
        eps = op.normal(layer.seed, ld.S_KERNEL, mu.size).reshape(mu.shape)
        s_in = op.sign(layer.seed, ld.S_SIGN_IN, B * C).reshape(B, C)
        s_out = op.sign(layer.seed, ld.S_SIGN_OUT, B * Fo).reshape(B, Fo)
    y = layer(xt)
    gy = round_to(rng.standard_normal(tuple(y.shape)), dtype)
    y.backward(_t(gy, dtype, cuda))
    bias = to_np(layer.bias)
    tol = TOL[dtype]
    y_ref, _ = oc.flipout_conv2d(x, mu, rho, eps, s_in, s_out, bias, "relu", st, pad)
    assert rel_err(to_np(y), y_ref) < tol
    g = oc.flipout_conv2d_bwd(x, mu, rho, eps, s_in, s_out, bias, "relu", gy, st, pad,
                              y_mask=None if dtype == torch.float32 else to_np(y))
    gmu, grho = (p.grad for p in layer.kernel_initializer.params())
    assert rel_err(to_np(xt.grad), g["dx"]) < tol
    assert rel_err(to_np(gmu), g["dmu"]) < tol
    assert rel_err(to_np(grho), g["drho"]) < tol
    assert rel_err(to_np(layer.bias.grad), g["dbias"]) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("geom", [GEOMS[0], GEOMS[3]])
def test_conv2d_batch_ensemble(cuda, dtype, geom):
    import edward2_b200 as ed

    ks, st, pad = geom
    E, n, H, W, C, Fo = 2, 3, 9, 8, 8, 16
    B = E * n
    rng = np.random.default_rng(3)
    x = round_to(rng.standard_normal((B, H, W, C)), dtype)
    layer = ed.layers.Conv2DBatchEnsemble(Fo, ks, ensemble_size=E, strides=st, padding=pad, activation="relu",
                                          dtype=_dt(dtype), alpha_initializer=ed.initializers.RandomNormal(1.0, 0.3, seed=1),
                                          gamma_initializer=ed.initializers.RandomNormal(1.0, 0.3, seed=2),
                                          bias_initializer=ed.initializers.RandomNormal(stddev=0.2, seed=3),
                                          kernel_initializer=ed.initializers.RandomNormal(stddev=0.2, seed=4))
    xt = _t(x, dtype, cuda, grad=True)
    y = layer(xt)
    gy = round_to(rng.standard_normal(tuple(y.shape)), dtype)
    y.backward(_t(gy, dtype, cuda))
    k, alpha, gamma, bias = (to_np(t) for t in (layer.kernel, layer.alpha, layer.gamma, layer.ensemble_bias))
    kq = round_to(k, dtype)  # the kernel is cast to the compute dtype
    tol = TOL[dtype]
    assert rel_err(to_np(y), oc.batch_ensemble_conv2d(x, kq, alpha, gamma, bias, "relu", st, pad)) < tol
    gy_eff = gy if dtype == torch.float32 else gy * (to_np(y) > 0)  # bf16: the kernel's own activation pattern
    g = oc.batch_ensemble_conv2d_bwd(x, kq, alpha, gamma, bias, "relu" if dtype == torch.float32 else None, gy_eff, st, pad)
    assert rel_err(to_np(xt.grad), g["dx"]) < tol
    assert rel_err(to_np(layer.kernel.grad), g["dkernel"]) < tol
    assert rel_err(to_np(layer.alpha.grad), g["dalpha"]) < tol
    assert rel_err(to_np(layer.gamma.grad), g["dgamma"]) < tol
    assert rel_err(to_np(layer.ensemble_bias.grad), g["dbias"]) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("inject", [True, False])
def test_conv2d_rank1(cuda, dtype, inject):
    import edward2_b200 as ed
    from edward2_b200.layers import dense as ld

    ks, st, pad = (3, 3), (1, 1), "same"
    E, n, H, W, C, Fo = 2, 3, 8, 8, 8, 16
    B = E * n
    rng = np.random.default_rng(4)
    x = round_to(rng.standard_normal((B, H, W, C)), dtype)
    fast = lambda seed: ed.initializers.TrainableNormal(  # noqa: E731
        mean_initializer=ed.initializers.RandomNormal(1.0, 0.3, seed=seed),
        stddev_initializer=ed.initializers.TruncatedNormal(-1.5, 0.2, seed=seed + 1))
    layer = ed.layers.Conv2DRank1(Fo, ks, ensemble_size=E, strides=st, padding=pad, activation="relu", dtype=_dt(dtype),
                                  seed=8, alpha_initializer=fast(11), gamma_initializer=fast(13), alpha_regularizer=None,
                                  gamma_regularizer=None, kernel_initializer=ed.initializers.RandomNormal(stddev=0.2, seed=4),
                                  bias_initializer=ed.initializers.RandomNormal(stddev=0.2, seed=5))
    xt = _t(x, dtype, cuda, grad=True)
    with torch.no_grad():
        layer(xt)
    layer._calls = 0
    if inject:
        e_r, e_s = rng.standard_normal((B, C)), rng.standard_normal((B, Fo))
        e_r[0, 0], e_s[1, 2] = 60.0, -60.0  # force the [-10, 10] clip on both sides
        layer.inject(eps_alpha=_t(e_r, torch.float32, cuda), eps_gamma=_t(e_s, torch.float32, cuda))
    else:
        e_r = op.per_example("normal", layer.seed, ld.S_R, B, C)
        e_s = op.per_example("normal", layer.seed, ld.S_S, B, Fo)
    y = layer(xt)
    gy = round_to(rng.standard_normal(tuple(y.shape)), dtype)
    y.backward(_t(gy, dtype, cuda))
    mu_r, rho_r = (to_np(t) for t in layer.alpha_initializer.params())
    mu_s, rho_s = (to_np(t) for t in layer.gamma_initializer.params())
    k, bias = round_to(to_np(layer.kernel), dtype), to_np(layer.ensemble_bias)
    tol = TOL[dtype]
    y_ref, _, _ = oc.rank1_conv2d(x, k, mu_r, rho_r, e_r, mu_s, rho_s, e_s, bias, "relu", st, pad)
    assert rel_err(to_np(y), y_ref) < tol
    g = oc.rank1_conv2d_bwd(x, k, mu_r, rho_r, e_r, mu_s, rho_s, e_s, bias, "relu", gy, st, pad,
                            y_mask=None if dtype == torch.float32 else to_np(y))
    pr, ps = layer.alpha_initializer.params(), layer.gamma_initializer.params()
    assert rel_err(to_np(xt.grad), g["dx"]) < tol
    assert rel_err(to_np(layer.kernel.grad), g["dkernel"]) < tol
    assert rel_err(to_np(pr[0].grad), g["dmu_r"]) < tol
    assert rel_err(to_np(pr[1].grad), g["drho_r"]) < tol
    assert rel_err(to_np(ps[0].grad), g["dmu_s"]) < tol
    assert rel_err(to_np(ps[1].grad), g["drho_s"]) < tol
    assert rel_err(to_np(layer.ensemble_bias.grad), g["dbias"]) < tol


def test_conv2d_value_errors_and_config(cuda):
    import edward2_b200 as ed

    with pytest.raises(ValueError):
        ed.layers.Conv2DReparameterization(4, 3, padding="causal")
    layer = ed.layers.Conv2DBatchEnsemble(4, 3, ensemble_size=4)
    with pytest.raises(ValueError):
        layer(torch.zeros(6, 5, 5, 2, device=cuda))  # batch not divisible by ensemble_size
    with pytest.raises(ValueError):
        ed.layers.Conv2DRank1(4, 3)(torch.zeros(5, 5, 2, device=cuda))  # not 4-D
    cfg = ed.layers.Conv2DFlipout(7, (3, 2), strides=2, padding="SAME").get_config()
    assert cfg["filters"] == 7 and cfg["kernel_size"] == (3, 2) and cfg["strides"] == (2, 2) and cfg["padding"] == "same"
    # losses: the kernel KL of the Bayesian conv layers, alpha / gamma KLs of rank-1
    rp = ed.layers.Conv2DReparameterization(4, 3)
    rp(torch.zeros(2, 5, 5, 2, device=cuda))
    assert len(rp.losses) == 1 and float(rp.losses[0]) >= 0
    r1 = ed.layers.Conv2DRank1(4, 3, ensemble_size=2)
    r1(torch.zeros(2, 5, 5, 2, device=cuda))
    assert len(r1.losses) == 2
