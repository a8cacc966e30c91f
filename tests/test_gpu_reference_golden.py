"""GPU parity against the REFERENCE-generated fixture tests/golden/reference_vectors.npz (outputs of the unmodified
edward2/jax/nn/{dense,random_feature,normalization,utils}.py executed under NumPy-backed jax / flax stand-ins by
tests/golden/make_reference_golden.py).  The CUDA path is driven through the Flax-signature mirrors `edward2_b200.nn`
with the reference's own variables loaded; tolerance is the fp32 bar (1e-5 tensor-scale relative error) unless a test
states otherwise."""
import os

import numpy as np
import pytest
import torch

from tests.util import rel_err, to_np

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
FP32 = 1e-5


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLD, "reference_vectors.npz"))


def _t(a, cuda, grad=False):
    t = torch.tensor(np.asarray(a), dtype=torch.float32, device=cuda)
    return t.requires_grad_(True) if grad else t


def _ed():
    import edward2_b200 as ed

    return ed


@pytest.mark.parametrize("tag", ["be2d", "be3d"])
def test_batch_ensemble_vs_reference(cuda, ref, tag):
    """jax/nn/dense.py:238-272 (einsum('E...C,EC,CD,ED->E...D') + bias + relu) on N-D inputs."""
    ed = _ed()
    E, O = ref[f"{tag}_alpha"].shape[0], ref[f"{tag}_kernel"].shape[1]
    layer = ed.nn.DenseBatchEnsemble(features=O, ens_size=E, activation=torch.relu)
    variables = {"params": {"kernel": _t(ref[f"{tag}_kernel"], cuda), "fast_weight_alpha": _t(ref[f"{tag}_alpha"], cuda),
                            "fast_weight_gamma": _t(ref[f"{tag}_gamma"], cuda), "bias": _t(ref[f"{tag}_bias"], cuda)}}
    y = layer.apply(variables, _t(ref[f"{tag}_x"], cuda))
    assert tuple(y.shape) == ref[f"{tag}_y"].shape
    assert rel_err(to_np(y), ref[f"{tag}_y"]) < FP32


@pytest.mark.parametrize("tag,scale", [("rff_unit", 1.0), ("rff_auto", None)])
def test_random_fourier_features_vs_reference(cuda, ref, tag, scale):
    """jax/nn/random_feature.py:183-214."""
    ed = _ed()
    D = ref[f"{tag}_kernel"].shape[1]
    layer = ed.nn.RandomFourierFeatures(features=D, feature_scale=scale)
    variables = {"random_features": {"kernel": _t(ref[f"{tag}_kernel"], cuda), "bias": _t(ref[f"{tag}_bias"], cuda)}}
    phi = layer.apply(variables, _t(ref[f"{tag}_x"], cuda))
    assert rel_err(to_np(phi), ref[f"{tag}_phi"]) < FP32


@pytest.mark.parametrize("tag,momentum,likelihood", [("cov_exact", None, "gaussian"), ("cov_mom", 0.9, "gaussian"),
                                                     ("cov_logistic", None, "binary_logistic"),
                                                     ("cov_poisson", 0.99, "poisson")])
def test_laplace_covariance_vs_reference(cuda, ref, tag, momentum, likelihood):
    """jax/nn/random_feature.py:307-309, :340-362, :393-404: two precision updates, then diagonal and full covariance."""
    ed = _ed()
    D, ridge = ref["cov_phi1"].shape[1], float(ref[f"{tag}_ridge"])
    layer = ed.nn.LaplaceRandomFeatureCovariance(hidden_features=D, ridge_penalty=ridge, momentum=momentum,
                                                 likelihood=likelihood)
    v = layer.init(0, _t(ref["cov_phi1"], cuda), _t(ref["cov_logits1"], cuda))
    assert rel_err(to_np(v["laplace_covariance"]["precision_matrix"]), ref[f"{tag}_p0"]) < 1e-7
    r, v1 = layer.apply(v, _t(ref["cov_phi1"], cuda), _t(ref["cov_logits1"], cuda), mutable=["laplace_covariance"])
    assert r is None
    assert rel_err(to_np(v1["laplace_covariance"]["precision_matrix"]), ref[f"{tag}_p1"]) < FP32
    _, v2 = layer.apply(v1, _t(ref["cov_phi2"], cuda), _t(ref["cov_logits2"], cuda), mutable=["laplace_covariance"])
    assert rel_err(to_np(v2["laplace_covariance"]["precision_matrix"]), ref[f"{tag}_p2"]) < FP32
    # predictive covariance with the reference's precision injected (the inverse is a library call; 1e-4 as in
    # tests/test_gpu_nn.py)
    vr = {"laplace_covariance": {"precision_matrix": _t(ref[f"{tag}_p2"], cuda)}}
    var = layer.apply(vr, _t(ref["cov_phit"], cuda), None, diagonal_only=True)
    cov = layer.apply(vr, _t(ref["cov_phit"], cuda), None, diagonal_only=False)
    assert rel_err(to_np(var), ref[f"{tag}_var"]) < 1e-4
    assert rel_err(to_np(cov), ref[f"{tag}_cov"]) < 1e-4


def test_rfgp_end_to_end_vs_reference(cuda, ref):
    """jax/nn/random_feature.py:112-140 with the reference's random features and output layer loaded."""
    ed = _ed()
    D, K = ref["gp_rff_kernel"].shape[1], ref["gp_out_kernel"].shape[1]
    gp = ed.nn.RandomFeatureGaussianProcess(features=K, hidden_features=D, hidden_kwargs=dict(feature_scale=None),
                                            covmat_kwargs=dict(ridge_penalty=1.0))
    variables = {
        "params": {"output_layer": {"kernel": _t(ref["gp_out_kernel"], cuda), "bias": _t(ref["gp_out_bias"], cuda)}},
        "random_features": {"hidden_layer": {"kernel": _t(ref["gp_rff_kernel"], cuda), "bias": _t(ref["gp_rff_bias"], cuda)}},
        "laplace_covariance": {"covmat_layer": {"precision_matrix": _t(np.eye(D), cuda)}},
    }
    (lg_a, cov_a), st1 = gp.apply(variables, _t(ref["gp_xa"], cuda), mutable=["laplace_covariance"])
    assert cov_a is None
    assert rel_err(to_np(lg_a), ref["gp_logits_a"]) < FP32
    (_, _), st2 = gp.apply(dict(variables, **st1), _t(ref["gp_xb"], cuda), mutable=["laplace_covariance"])
    assert rel_err(to_np(st2["laplace_covariance"]["covmat_layer"]["precision_matrix"]), ref["gp_precision"]) < FP32
    full = dict(variables, **st2)
    logits, var, feats = gp.apply(full, _t(ref["gp_xt"], cuda), return_random_features=True)
    assert rel_err(to_np(feats), ref["gp_features"]) < FP32
    assert rel_err(to_np(logits), ref["gp_logits"]) < FP32
    assert rel_err(to_np(var), ref["gp_var"]) < 1e-4
    _, cov = gp.apply(full, _t(ref["gp_xt"], cuda), return_full_covmat=True)
    assert rel_err(to_np(cov), ref["gp_cov"]) < 1e-4


@pytest.mark.parametrize("tag", ["sn1", "sn3", "sn_small"])
def test_spectral_normalization_vs_reference(cuda, ref, tag):
    """jax/nn/normalization.py:123-185 over a Dense layer: updated u, v, the normalised kernel and the layer output."""
    ed = _ed()
    it, c = int(ref[f"{tag}_iters"]), float(ref[f"{tag}_c"])
    O = ref[f"{tag}_w"].shape[1]
    layer = ed.nn.SpectralNormalization(ed.nn.Dense(features=O), iteration=it, norm_multiplier=c)
    variables = {"params": {"Dense": {"kernel": _t(ref[f"{tag}_w"], cuda), "bias": _t(ref[f"{tag}_b"], cuda)}},
                 "spectral_stats": {"u": _t(ref[f"{tag}_u0"], cuda), "v": _t(ref[f"{tag}_v0"], cuda)}}
    y, st = layer.apply(variables, _t(ref[f"{tag}_x"], cuda), training=True, mutable=["spectral_stats", "intermediates"])
    assert rel_err(to_np(st["spectral_stats"]["u"]), ref[f"{tag}_u1"]) < FP32
    assert rel_err(to_np(st["spectral_stats"]["v"]), ref[f"{tag}_v1"]) < FP32
    assert rel_err(to_np(st["intermediates"]["w"]), ref[f"{tag}_what"]) < FP32
    assert rel_err(to_np(y), ref[f"{tag}_y"]) < FP32
    v2 = dict(variables, spectral_stats=st["spectral_stats"])
    y_eval = layer.apply(v2, _t(ref[f"{tag}_x"], cuda), training=False)
    assert rel_err(to_np(y_eval), ref[f"{tag}_y_eval"]) < FP32


def test_mean_field_logits_vs_reference(cuda, ref):
    """jax/nn/utils.py:54-101."""
    ed = _ed()
    lg, var = _t(ref["mf_logits"], cuda), _t(ref["mf_var"], cuda)
    for lik in ("logistic", "binary_logistic", "poisson"):
        out = ed.nn.mean_field_logits(lg, var, mean_field_factor=0.7, likelihood=lik)
        assert rel_err(to_np(out), ref[f"mf_{lik}"]) < FP32
    assert rel_err(to_np(ed.nn.mean_field_logits(lg, None, mean_field_factor=3.0)), ref["mf_nocov"]) < FP32
    assert rel_err(to_np(ed.nn.mean_field_logits(lg, var, mean_field_factor=-1.0)), ref["mf_negative"]) < 1e-7


@pytest.mark.parametrize("tag", ["het", "het_shared"])
def test_mc_softmax_dense_fa_vs_reference(cuda, ref, tag):
    """`edward2_b200.nn.MCSoftmaxDenseFA` with the reference's variables loaded and its recorded noise injected, against
    the outputs of the reference's own MCSoftmaxDenseFA (edward2/jax/nn/heteroscedastic_lib.py:43-336): Dense parameter
    layers + Monte-Carlo softmax kernel in its JAX form (one diagonal draw per sample)."""
    ed = _ed()
    x = ref[f"{tag}_x"]
    z, n = ref[f"{tag}_factor_noise"], ref[f"{tag}_diag_noise"]
    rows, S, Fn = z.shape
    K = ref[f"{tag}_loc_layer_bias"].shape[0]
    FP, KP = (Fn + 7) // 8 * 8, (K + 7) // 8 * 8
    buf = np.zeros((rows, S, FP + KP), dtype=np.float32)
    buf[:, :, :Fn] = z
    buf[:, :, FP] = n[:, :, 0]
    layer = ed.nn.MCSoftmaxDenseFA(num_classes=K, num_factors=Fn, temperature=float(ref[f"{tag}_temperature"]),
                                   train_mc_samples=S, test_mc_samples=S, share_samples_across_batch=(tag == "het_shared"))
    variables = {"params": {name: {"kernel": _t(ref[f"{tag}_{name}_kernel"], cuda, True), "bias": _t(ref[f"{tag}_{name}_bias"], cuda, True)}
                            for name in ("loc_layer", "scale_layer", "diag_layer")}}
    logits, log_probs, probs = layer.apply(variables, _t(x, cuda), training=True, mc_noise=torch.tensor(buf, device=cuda))
    assert rel_err(to_np(probs), ref[f"{tag}_probs"]) < FP32
    assert rel_err(to_np(logits), ref[f"{tag}_logits"]) < FP32
    assert rel_err(to_np(log_probs), ref[f"{tag}_log_probs"]) < FP32
    # init / apply with named rng streams: fresh noise per seed, reproducible per seed; gradients reach every parameter
    v = layer.init({"params": 0, "diag_noise_samples": 1, "standard_norm_noise_samples": 2}, _t(x, cuda))
    assert set(v["params"]) == {"loc_layer", "scale_layer", "diag_layer"}
    a = layer.apply(v, _t(x, cuda), rngs={"diag_noise_samples": 5, "standard_norm_noise_samples": 6})[0]
    b = layer.apply(v, _t(x, cuda), rngs={"diag_noise_samples": 5, "standard_norm_noise_samples": 6})[0]
    c = layer.apply(v, _t(x, cuda), rngs={"diag_noise_samples": 7, "standard_norm_noise_samples": 6})[0]
    assert torch.equal(a, b) and not torch.equal(a, c)
    a.sum().backward()
    assert all(v["params"][nm]["kernel"].grad is not None for nm in v["params"])


def test_mc_sigmoid_dense_fa_vs_reference(cuda, ref):
    """`edward2_b200.nn.MCSigmoidDenseFA` against the reference's own MCSigmoidDenseFA outputs
    (edward2/jax/nn/heteroscedastic_lib.py:338-594) with its recorded noise injected."""
    ed = _ed()
    x, z, n = ref["hsig_x"], ref["hsig_factor_noise"], ref["hsig_diag_noise"]
    rows, S, Fn = z.shape
    K = ref["hsig_loc_layer_bias"].shape[0]
    FP, KP = (Fn + 7) // 8 * 8, (K + 7) // 8 * 8
    buf = np.zeros((rows, S, FP + KP), dtype=np.float32)
    buf[:, :, :Fn] = z
    buf[:, :, FP] = n[:, :, 0]
    layer = ed.nn.MCSigmoidDenseFA(num_outputs=K, num_factors=Fn, temperature=float(ref["hsig_temperature"]),
                                   train_mc_samples=S, test_mc_samples=S)
    variables = {"params": {name: {"kernel": _t(ref[f"hsig_{name}_kernel"], cuda, True), "bias": _t(ref[f"hsig_{name}_bias"], cuda, True)}
                            for name in ("loc_layer", "scale_layer", "diag_layer")}}
    logits, log_probs, probs = layer.apply(variables, _t(x, cuda), training=True, mc_noise=torch.tensor(buf, device=cuda))
    assert rel_err(to_np(probs), ref["hsig_probs"]) < FP32
    assert rel_err(to_np(log_probs), ref["hsig_log_probs"]) < FP32
    assert rel_err(to_np(logits), ref["hsig_logits"]) < FP32


@pytest.mark.parametrize("tag", ["hsngp", "hsngpbe"])
def test_mc_sigmoid_dense_fa_sngp_vs_reference(cuda, ref, tag):
    """`edward2_b200.nn.MCSigmoidDenseFASNGP` / `...BE` against the reference's own outputs
    (edward2/jax/nn/random_feature.py:407-735) with its recorded noise injected: a training call (precision update), then
    an evaluation call whose diagonal scale mixes in the GP variance."""
    ed = _ed()
    K, D = ref[f"{tag}_out_bias"].shape[0], ref[f"{tag}_rff_kernel"].shape[1]
    Fn, S = ref[f"{tag}_factor_noise"].shape[2], ref[f"{tag}_factor_noise"].shape[1]
    FP, KP = (Fn + 7) // 8 * 8, (K + 7) // 8 * 8

    def noise(prefix):
        z, n = ref[f"{tag}_{prefix}factor_noise"], ref[f"{tag}_{prefix}diag_noise"]
        buf = np.zeros((z.shape[0], S, FP + KP), dtype=np.float32)
        buf[:, :, :Fn] = z
        buf[:, :, FP] = n[:, :, 0]
        return torch.tensor(buf, device=cuda)

    kw = dict(num_outputs=K, num_factors=Fn, temperature=0.8, train_mc_samples=S, test_mc_samples=S, hidden_features=D,
              het_var_weight=0.7, sngp_var_weight=1.3, covmat_kwargs=dict(ridge_penalty=1.0))
    layer = ed.nn.MCSigmoidDenseFASNGPBE(ens_size=2, **kw) if tag == "hsngpbe" else ed.nn.MCSigmoidDenseFASNGP(**kw)
    leaves = ("kernel", "bias") + (("fast_weight_alpha", "fast_weight_gamma") if tag == "hsngpbe" else ())
    params = {name: {leaf: _t(ref[f"{tag}_{name}_{leaf}"], cuda, True) for leaf in leaves}
              for name in ("scale_layer", "diag_layer")}
    params["loc_layer"] = {"output_layer": {"kernel": _t(ref[f"{tag}_out_kernel"], cuda, True),
                                            "bias": _t(ref[f"{tag}_out_bias"], cuda, True)}}
    variables = {
        "params": params,
        "random_features": {"loc_layer": {"hidden_layer": {"kernel": _t(ref[f"{tag}_rff_kernel"], cuda),
                                                           "bias": _t(ref[f"{tag}_rff_bias"], cuda)}}},
        "laplace_covariance": {"loc_layer": {"covmat_layer": {"precision_matrix": _t(np.eye(D), cuda)}}},
    }
    (logits, cov, log_probs, probs), st = layer.apply(variables, _t(ref[f"{tag}_x"], cuda), training=True,
                                                      mc_noise=noise(""), mutable=["laplace_covariance"])
    assert cov is None
    assert rel_err(to_np(probs), ref[f"{tag}_probs"]) < FP32
    assert rel_err(to_np(log_probs), ref[f"{tag}_log_probs"]) < FP32
    assert rel_err(to_np(logits), ref[f"{tag}_logits"]) < FP32
    assert rel_err(to_np(st["laplace_covariance"]["loc_layer"]["covmat_layer"]["precision_matrix"]),
                   ref[f"{tag}_precision"]) < FP32
    logits.sum().backward()   # the training form differentiates through the Monte-Carlo head into all three layers
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in
               (params["scale_layer"]["kernel"], params["diag_layer"]["kernel"], params["loc_layer"]["output_layer"]["kernel"]))
    logits, var, log_probs, probs = layer.apply(dict(variables, **st), _t(ref[f"{tag}_xt"], cuda), training=False,
                                                mc_noise=noise("eval_"))
    assert rel_err(to_np(var), ref[f"{tag}_eval_var"]) < 1e-4
    assert rel_err(to_np(probs), ref[f"{tag}_eval_probs"]) < 1e-4
    assert rel_err(to_np(log_probs), ref[f"{tag}_eval_log_probs"]) < 1e-4
    assert rel_err(to_np(logits), ref[f"{tag}_eval_logits"]) < 1e-4
