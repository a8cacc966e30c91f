"""Heteroscedastic heads on the CUDA path (edward2_b200/layers/heteroscedastic.py; include/ed2b200.h:
ed2_mc_softmax_fa_fwd / _bwd) against the float64 oracle (oracle/heteroscedastic.py, following
edward2/tensorflow/layers/heteroscedastic.py:178-253, 749-785 and edward2/tensorflow/layers/hetsngp.py:143-240):
op level with injected and native Philox noise, shared samples, predictive variance, the layer classes end to end."""
import numpy as np
import pytest
import torch

from oracle import heteroscedastic as oh
from tests.util import rel_err, to_np

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _f32(a, cuda, grad=False):
    t = torch.tensor(np.asarray(a), dtype=torch.float32, device=cuda)
    return t.requires_grad_(True) if grad else t


@pytest.mark.parametrize("shape", [(37, 100, 10, 10), (8, 64, 3, 1), (5, 33, 40, 16), (3, 50, 64, 7)])
@pytest.mark.parametrize("inject", [True, False])
def test_mc_softmax_fa_op(cuda, shape, inject):
    from edward2_b200 import functional as F

    B, S, K, Fn = shape
    rng = np.random.default_rng(B)
    T, eps = 0.8, 1e-7
    loc = rng.standard_normal((B, K)).astype(np.float32).astype(np.float64)
    fac = (rng.standard_normal((B, K * Fn)) * 0.5).astype(np.float32).astype(np.float64)
    draw = rng.standard_normal((B, K)).astype(np.float32).astype(np.float64)
    FP, KP = (Fn + 7) // 8 * 8, (K + 7) // 8 * 8
    if inject:
        buf = rng.standard_normal((B, S, FP + KP)).astype(np.float32)
        z, e = buf[:, :, :Fn].astype(np.float64), buf[:, :, FP:FP + K].astype(np.float64)
        noise = F.Randomness(torch.tensor(buf, device=cuda))
    else:
        z, e = oh.native_noise(99, 7, B, S, Fn, K)
        noise = F.Randomness(seed=99, stream=7)
    lt, ft, dt_ = _f32(loc, cuda, True), _f32(fac, cuda, True), _f32(draw, cuda, True)
    logits, pm, var = F.MCSoftmaxFA.apply(lt, ft, dt_, None, 1.0, 1.0, noise, S, Fn, False, T, eps, True)
    l_ref, pm_ref, var_ref = oh.mc_softmax_fa(loc, fac, draw, z, e, T, eps)
    assert rel_err(to_np(pm), pm_ref) < TOL
    assert rel_err(to_np(logits), l_ref) < TOL
    assert rel_err(to_np(var), var_ref) < 1e-4
    g = rng.standard_normal((B, K))
    logits.backward(_f32(g, cuda))
    ref = oh.mc_softmax_fa_bwd(loc, fac, draw, z, e, g, T, eps)
    assert rel_err(to_np(lt.grad), ref["dloc"]) < 2e-5
    assert rel_err(to_np(ft.grad), ref["dfac"]) < 2e-5
    assert rel_err(to_np(dt_.grad), ref["ddiag_raw"]) < 2e-5


def test_mc_softmax_shared_samples_and_sngp_mixing(cuda):
    from edward2_b200 import functional as F

    B, S, K, Fn = 9, 128, 10, 4
    rng = np.random.default_rng(3)
    loc, fac, draw = rng.standard_normal((B, K)), rng.standard_normal((B, K * Fn)) * 0.3, rng.standard_normal((B, K))
    sv = rng.random(B) * 2
    z, e = oh.native_noise(5, 7, B, S, Fn, K, share=True)
    logits, pm, _ = F.MCSoftmaxFA.apply(_f32(loc, cuda), _f32(fac, cuda), _f32(draw, cuda), _f32(sv, cuda), 0.7, 1.3,
                                        F.Randomness(seed=5, stream=7), S, Fn, True, 1.0, 1e-7, False)
    f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)  # noqa: E731
    l_ref, pm_ref, _ = oh.mc_softmax_fa(f32(loc), f32(fac), f32(draw), z, e, 1.0, 1e-7, f32(sv), 0.7, 1.3)
    assert rel_err(to_np(pm), pm_ref) < TOL and rel_err(to_np(logits), l_ref) < TOL


def test_mc_softmax_dense_fa_layer(cuda):
    """Layer end to end: three Dense parameter layers + the MC kernel; outputs and every parameter gradient against the
    oracle evaluated on the layer's own weights; tuple outputs; errors of the reference's unsupported branches."""
    import edward2_b200 as ed
    from edward2_b200.layers import heteroscedastic as lh

    B, D, K, Fn, S = 24, 16, 7, 3, 200
    rng = np.random.default_rng(8)
    x = rng.standard_normal((B, D)).astype(np.float32)
    layer = ed.layers.MCSoftmaxDenseFA(K, Fn, temperature=1.5, train_mc_samples=S, test_mc_samples=S // 2,
                                       compute_pred_variance=True, logits_only=False, seed=77)
    xt = _f32(x, cuda, True)
    logits, log_probs, probs, var = layer(xt, training=True)
    assert logits is log_probs and tuple(probs.shape) == (B, K) and tuple(var.shape) == (B, K)
    w = {n: (to_np(l.kernel), to_np(l.bias)) for n, l in (("loc", layer._loc_layer), ("scale", layer._scale_layer),
                                                          ("diag", layer._diag_layer))}
    x64 = x.astype(np.float64)
    loc, fac, draw = (x64 @ w[n][0] + w[n][1] for n in ("loc", "scale", "diag"))
    z, e = oh.native_noise(layer.seed, lh.S_MC, B, S, Fn, K)
    l_ref, pm_ref, var_ref = oh.mc_softmax_fa(loc, fac, draw, z, e, 1.5)
    assert rel_err(to_np(logits), l_ref) < 1e-4 and rel_err(to_np(probs), pm_ref) < 1e-4 and rel_err(to_np(var), var_ref) < 1e-3
    g = rng.standard_normal((B, K))
    logits.backward(_f32(g, cuda))
    gr = oh.mc_softmax_fa_bwd(loc, fac, draw, z, e, g, 1.5)
    for name, l, d in (("loc", layer._loc_layer, gr["dloc"]), ("scale", layer._scale_layer, gr["dfac"]),
                       ("diag", layer._diag_layer, gr["ddiag_raw"])):
        assert rel_err(to_np(l.kernel.grad), x64.T @ d) < 1e-4, name
        assert rel_err(to_np(l.bias.grad), d.sum(0)) < 1e-4, name
    dx = gr["dloc"] @ w["loc"][0].T + gr["dfac"] @ w["scale"][0].T + gr["ddiag_raw"] @ w["diag"][0].T
    assert rel_err(to_np(xt.grad), dx) < 1e-4
    # evaluation uses test_mc_samples; an explicit seed reproduces
    a = layer(xt.detach(), training=False, seed=123)[0]
    b = layer(xt.detach(), training=False, seed=123)[0]
    c = layer(xt.detach(), training=False, seed=124)[0]
    assert torch.equal(a, b) and not torch.equal(a, c)
    with pytest.raises(NotImplementedError):
        ed.layers.MCSoftmaxDenseFA(2, 3)
    with pytest.raises(NotImplementedError):
        ed.layers.MCSoftmaxDenseFA(5, 3, parameter_efficient=True)


def test_heteroscedastic_sngp_layer(cuda):
    """HeteroscedasticSNGPLayer: training = MC softmax around the GP logits with the heteroscedastic scale only; evaluation
    mixes in the SNGP marginal variance diag(covmat) (hetsngp.py:160-168); covariance reset is forwarded."""
    import edward2_b200 as ed
    from edward2_b200.layers import heteroscedastic as lh

    B, D, K, Fn, S = 64, 32, 5, 4, 256
    rng = np.random.default_rng(9)
    x = _f32(rng.standard_normal((B, D)), cuda)
    layer = ed.layers.HeteroscedasticSNGPLayer(K, num_factors=Fn, train_mc_samples=S, test_mc_samples=S, num_inducing=128,
                                               gp_cov_momentum=-1.0, gp_cov_ridge_penalty=1.0, sngp_var_weight=0.6,
                                               het_var_weight=1.4, logits_only=True, seed=31, dtype="float32")
    layer.train()
    logits_tr = layer(x, training=True)
    assert tuple(logits_tr.shape) == (B, K)
    logits_tr.sum().backward()
    assert layer.sngp_layer._gp_output_layer.kernel.grad is not None and layer._scale_layer.kernel.grad is not None
    # oracle for both modes from the layer's own intermediate tensors
    layer.eval()
    with torch.no_grad():
        gp_logits, covmat = layer.sngp_layer(x, training=False)
        fac, draw = layer._scale_params(x)
        calls = layer._calls
        logits_ev = layer(x, training=False)
    z, e = oh.native_noise((layer.seed + calls * lh.GOLDEN) & lh.MASK64, lh.S_MC, B, S, Fn, K)
    sv = np.diagonal(to_np(covmat))
    l_ref, _, _ = oh.mc_softmax_fa(to_np(gp_logits), to_np(fac), to_np(draw), z, e, 1.0, 1e-7, sv, 1.4, 0.6)
    assert rel_err(to_np(logits_ev), l_ref) < 1e-4
    assert (sv > 0).all()
    layer.reset_covariance_matrix()
    assert float(layer.sngp_layer._gp_cov_layer.precision_matrix.abs().sum()) >= 0


@pytest.mark.parametrize("inject", [True, False])
def test_mc_sigmoid_fa_op_and_layer(cuda, inject):
    """Sigmoid link (MCSigmoidDenseFA): forward and the gradients of the inverse-sigmoid logits against the oracle; the TF
    layer class end to end (one diagonal draw per class)."""
    import edward2_b200 as ed
    from edward2_b200 import functional as F
    from edward2_b200.layers import heteroscedastic as lh

    B, S, K, Fn, T = 21, 96, 6, 3, 1.3
    rng = np.random.default_rng(17)
    f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)  # noqa: E731
    loc, fac, draw = f32(rng.standard_normal((B, K))), f32(rng.standard_normal((B, K * Fn)) * 0.5), f32(rng.standard_normal((B, K)))
    FP, KP = 8, 8
    if inject:
        buf = rng.standard_normal((B, S, FP + KP)).astype(np.float32)
        z, e = buf[:, :, :Fn].astype(np.float64), buf[:, :, FP:FP + K].astype(np.float64)
        noise = F.Randomness(torch.tensor(buf, device=cuda))
    else:
        z, e = oh.native_noise(4, 7, B, S, Fn, K)
        noise = F.Randomness(seed=4, stream=7)
    lt, ft, dt_ = _f32(loc, cuda, True), _f32(fac, cuda, True), _f32(draw, cuda, True)
    logits, pm, _ = F.MCSoftmaxFA.apply(lt, ft, dt_, None, 1.0, 1.0, noise, S, Fn, 4, T, 1e-7, False)
    l_ref, _, pm_ref = oh.mc_sigmoid_fa(loc, fac, draw, z, e, T)
    assert rel_err(to_np(pm), pm_ref) < TOL and rel_err(to_np(logits), l_ref) < TOL
    g = rng.standard_normal((B, K))
    logits.backward(_f32(g, cuda))
    gr = oh.mc_sigmoid_fa_bwd(loc, fac, draw, z, e, g, T)
    assert rel_err(to_np(lt.grad), gr["dloc"]) < 2e-5
    assert rel_err(to_np(ft.grad), gr["dfac"]) < 2e-5
    assert rel_err(to_np(dt_.grad), gr["ddiag_raw"]) < 2e-5
    if not inject:
        layer = ed.layers.MCSigmoidDenseFA(2, 3, train_mc_samples=64, seed=5)  # two outputs: allowed for the sigmoid head
        x = torch.randn(10, 12, device=cuda)
        lg, lp, pr, var = layer(x, training=True)
        assert tuple(lg.shape) == (10, 2) and var is None and float(pr.min()) > 0 and float(pr.max()) < 1
        x64 = to_np(x)
        lin = lambda l: x64 @ to_np(l.kernel) + to_np(l.bias)  # noqa: E731
        zz, ee = oh.native_noise(layer.seed, lh.S_MC, 10, 64, 3, 2)
        l2, lp2, p2 = oh.mc_sigmoid_fa(lin(layer._loc_layer), lin(layer._scale_layer), lin(layer._diag_layer), zz, ee, 1.0)
        assert rel_err(to_np(lg), l2) < 1e-4 and rel_err(to_np(lp), lp2) < 1e-4 and rel_err(to_np(pr), p2) < 1e-4
