"""GPU parity of the DenseRank1 branches the default configuration does not take (edward2/tensorflow/layers/dense.py):
additive perturbations (:1126-1127), non-default clip bounds (:1025-1026, :1108-1110, :1118-1120) and the per-example
sampled ensemble bias (:1131-1139), against the float64 oracle on injected noise and on the native Philox streams."""
import numpy as np
import pytest
import torch

from oracle import dense as od
from oracle import philox as op
from tests.util import TOL, dev, rel_err, round_to, to_np

pytestmark = pytest.mark.gpu

RELU = 1


def _check(name, got, want, tol):
    e = rel_err(to_np(got), want)
    assert e <= tol, f"{name}: rel err {e:.3e} > {tol:.1e}"


def _f32(a):
    return np.float32(a).astype(np.float64)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("additive", [False, True])
@pytest.mark.parametrize("clip", [(-10.0, 10.0), (-0.5, 1.25)])
@pytest.mark.parametrize("sampled_bias", [False, True])
def test_rank1_variants_injected(cuda, dtype, additive, clip, sampled_bias):
    from edward2_b200 import functional as F

    E, n, I, O = 3, 16, 72, 40
    B = E * n
    rng = np.random.default_rng(11)
    x = round_to(rng.standard_normal((B, I)), dtype)
    w = round_to(rng.standard_normal((I, O)) / np.sqrt(I), dtype)
    mu_r, mu_s = _f32(rng.choice([-1.0, 1.0], (E, I))), _f32(rng.choice([-1.0, 1.0], (E, O)))
    rho_r, rho_s = _f32(rng.standard_normal((E, I)) * 0.3 - 1.0), _f32(rng.standard_normal((E, O)) * 0.3 - 1.0)
    eps_r, eps_s, eps_b = (_f32(rng.standard_normal(s)) for s in ((B, I), (B, O), (B, O)))
    bias = _f32(rng.standard_normal((E, O)) * 0.1)
    rho_b = _f32(rng.standard_normal((E, O)) * 0.3 - 2.0) if sampled_bias else None
    gy = round_to(rng.standard_normal((B, O)), dtype)
    tol = TOL[dtype]

    xt, wt = dev(x, dtype, cuda, True), dev(w, torch.float32, cuda, True)
    mr, ms = dev(mu_r, torch.float32, cuda, True), dev(mu_s, torch.float32, cuda, True)
    rr, rs = dev(rho_r, torch.float32, cuda, True), dev(rho_s, torch.float32, cuda, True)
    bt = dev(bias, torch.float32, cuda, True)
    rb = dev(rho_b, torch.float32, cuda, True) if sampled_bias else None
    R = lambda a: F.Randomness(dev(a, torch.float32, cuda))  # noqa: E731
    y = F.Rank1Dense.apply(xt, wt, mr, rr, ms, rs, bt, RELU, E, dtype, R(eps_r), R(eps_s), additive, clip, rb,
                           R(eps_b) if sampled_bias else None)
    y.backward(dev(gy, dtype, cuda))
    torch.cuda.synchronize()

    kw = dict(additive=additive, clip=clip, rho_b=rho_b, eps_b=eps_b if sampled_bias else None)
    y_ref, _, _, _ = od.rank1_fwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, "relu", **kw)
    _check("y", y, y_ref, tol)
    g = od.rank1_bwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, "relu", gy,
                     y_mask=None if dtype == torch.float32 else to_np(y), **kw)
    for name, got in (("dx", xt.grad), ("dw", wt.grad), ("dmu_r", mr.grad), ("drho_r", rr.grad), ("dmu_s", ms.grad),
                      ("drho_s", rs.grad), ("dbias", bt.grad)):
        _check(name, got, g[name], tol)
    if sampled_bias:
        _check("drho_b", rb.grad, g["drho_b"], tol)


def test_rank1_layer_variants_native_sampler(cuda):
    """The drop-in class with use_additive_perturbation, custom clip bounds and a trainable-normal ensemble bias, on the
    in-kernel Philox streams; the oracle is fed the host restatement of the same streams (fp32: precise sampler)."""
    import edward2_b200 as ed
    from edward2_b200.layers import dense as ld

    E, n, I, O = 2, 24, 40, 56
    B = E * n
    torch.manual_seed(0)
    layer = ed.layers.DenseRank1(O, activation="relu", ensemble_size=E, use_additive_perturbation=True,
                                 min_perturbation_value=-0.9, max_perturbation_value=1.1,
                                 bias_initializer="trainable_normal", seed=77)
    x = torch.randn(B, I, device=cuda)
    y = layer(x)
    gy = torch.randn(B, O, device=cuda)
    y.backward(gy)
    torch.cuda.synchronize()
    key = layer.seed  # first call: key = seed + 0 * GOLDEN
    eps_r = op.normal(key, ld.S_R, B * I).reshape(B, I)
    eps_s = op.normal(key, ld.S_S, B * O).reshape(B, O)
    eps_b = op.normal(key, ld.S_B, B * O).reshape(B, O)
    mu_r, rho_r = (to_np(t) for t in layer.alpha_initializer.params())
    mu_s, rho_s = (to_np(t) for t in layer.gamma_initializer.params())
    mu_b, rho_b = (to_np(t) for t in layer.ensemble_bias_initializer.params())
    kw = dict(additive=True, clip=(-0.9, 1.1), rho_b=rho_b, eps_b=eps_b)
    y_ref, _, _, _ = od.rank1_fwd(to_np(x), to_np(layer.kernel), mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, mu_b, "relu", **kw)
    _check("y", y, y_ref, 1e-5)
    g = od.rank1_bwd(to_np(x), to_np(layer.kernel), mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, mu_b, "relu", to_np(gy), **kw)
    _check("dw", layer.kernel.grad, g["dw"], 1e-5)
    _check("drho_r", layer.alpha_initializer.params()[1].grad, g["drho_r"], 1e-5)
    _check("drho_s", layer.gamma_initializer.params()[1].grad, g["drho_s"], 1e-5)
    _check("dmu_b", layer.ensemble_bias_initializer.params()[0].grad, g["dbias"], 1e-5)
    _check("drho_b", layer.ensemble_bias_initializer.params()[1].grad, g["drho_b"], 1e-5)
    assert len(layer.losses) == 2  # alpha + gamma KL (dense.py:1017-1018 defaults); bias_regularizer=None
    cfg = layer.get_config()
    assert cfg["use_additive_perturbation"] is True and cfg["min_perturbation_value"] == -0.9
