"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — float64 NumPy restatement of the Monte-Carlo softmax heads:
MCSoftmaxDenseFA (edward2/tensorflow/layers/heteroscedastic.py:508; _compute_mc_samples :178-202,
_compute_predictive_mean :204-226, _compute_predictive_variance :228-253, factor-analysis noise :749-785) and the
variance mixing of HeteroscedasticSNGPLayer (edward2/tensorflow/layers/hetsngp.py:143-170, call :186-240).

Every random tensor is an explicit argument: z [B or 1, S, F] standard-normal factor samples, e [B or 1, S, K] diagonal
noise samples (batch dimension 1 = share_samples_across_batch).  PINNED to the reference: the unmodified
edward2/jax/nn/heteroscedastic_lib.py (MCSoftmaxDenseFA, :43-336) is executed under the NumPy-backed jax / flax stand-ins
with its jax.random.normal draws recorded (tests/golden/make_reference_golden.py); tests/test_oracle.py::
test_reference_mc_softmax_dense_fa checks this restatement against those outputs at 1e-12 and
tests/test_gpu_reference_golden.py checks the CUDA kernel against them.  The TF classes (one diagonal draw per class) are pinned the same way: layers/heteroscedastic.py
MCSoftmaxDenseFA / MCSigmoidDenseFA and layers/hetsngp.py HeteroscedasticSNGPLayer run unmodified under the tensorflow / tfp
stand-ins (tests/golden/make_reference_golden_tf.py; tests/test_oracle.py::test_reference_tf_mc_heads,
test_reference_tf_heteroscedastic_sngp_layer).  The closed-form gradient is pinned against float64 autograd in tests/test_oracle.py."""
from __future__ import annotations

import numpy as np

from . import dense as od
from . import philox as op

MIN_SCALE_MONTE_CARLO = 1e-3  # heteroscedastic.py:24


def softplus(x):
    x = np.asarray(x, dtype=np.float64)
    return np.where(x > 30, x, np.log1p(np.exp(np.minimum(x, 30))))


def effective_diag(diag_raw, sngp_var=None, w_het=1.0, w_sngp=1.0):
    """_diag_layer has a softplus activation and MIN_SCALE_MONTE_CARLO is added (:660-674); HetSNGP at evaluation mixes
    in the SNGP marginal variance (hetsngp.py:160-168)."""
    d = softplus(diag_raw) + MIN_SCALE_MONTE_CARLO
    if sngp_var is not None:
        d = np.sqrt(w_het * d * d + w_sngp * np.asarray(sngp_var, dtype=np.float64)[:, None])
    return d


def sample_probs(loc, fac, diag, z, e, temperature=1.0):
    """[B, S, K] softmax samples: latents = loc + einsum('ijk,iak->iaj', V, z) + e * diag (:775-785, :196-202)."""
    loc = np.asarray(loc, dtype=np.float64)
    B, K = loc.shape
    V = np.asarray(fac, dtype=np.float64).reshape(B, K, -1)
    z = np.broadcast_to(np.asarray(z, dtype=np.float64), (B,) + tuple(np.shape(z)[1:]))
    e = np.broadcast_to(np.asarray(e, dtype=np.float64), (B,) + tuple(np.shape(e)[1:]))
    lat = loc[:, None, :] + np.einsum("ijk,iak->iaj", V, z) + e * diag[:, None, :]
    lat = lat / temperature
    lat = lat - lat.max(-1, keepdims=True)
    p = np.exp(lat)
    return p / p.sum(-1, keepdims=True), z, e


def mc_softmax_fa(loc, fac, diag_raw, z, e, temperature=1.0, eps=1e-7, sngp_var=None, w_het=1.0, w_sngp=1.0):
    """Returns (logits, probs_mean, variance) for num_classes > 2 (:186-240 of hetsngp.py, same tail in the base class)."""
    diag = effective_diag(diag_raw, sngp_var, w_het, w_sngp)
    p, _, _ = sample_probs(loc, fac, diag, z, e, temperature)
    pm = p.mean(1)
    var = ((p - pm[:, None, :]) ** 2).mean(1)
    return np.log(np.clip(pm, eps, 1.0)), pm, var


def mc_softmax_fa_bwd(loc, fac, diag_raw, z, e, g, temperature=1.0, eps=1e-7):
    """Closed-form gradients of logits = log(clip(mean_s softmax(latent_s / T))) w.r.t. loc, fac, diag_raw."""
    diag = effective_diag(diag_raw)
    p, zb, eb = sample_probs(loc, fac, diag, z, e, temperature)
    S = p.shape[1]
    pm = p.mean(1)
    u = np.where((pm > eps) & (pm <= 1.0), np.asarray(g, dtype=np.float64) / pm, 0.0) / S
    dot = (p * u[:, None, :]).sum(-1, keepdims=True)
    dlat = p * (u[:, None, :] - dot) / temperature                 # [B, S, K]
    dloc = dlat.sum(1)
    ddiag = (dlat * eb).sum(1)
    dV = np.einsum("iaj,iak->ijk", dlat, zb)                        # [B, K, F]
    return dict(dloc=dloc, dfac=dV.reshape(dV.shape[0], -1), ddiag_raw=ddiag * od.sigmoid(diag_raw))


def native_noise(seed, stream, batch, num_samples, num_factors, num_classes, share=False, row0=0):
    """(z, e) as the kernel draws them: sample (b, s) owns elements [((row0 + b) S + s) P, +P) of the normal stream,
    P = round8(F) + round8(K), z first then e (edward2_b200/csrc/mcsoftmax.cu)."""
    FP, KP = (num_factors + 7) // 8 * 8, (num_classes + 7) // 8 * 8
    P = FP + KP
    rows = 1 if share else batch
    first = 0 if share else row0
    flat = op.normal(seed, stream, (first + rows) * num_samples * P).reshape(-1, num_samples, P)[first:first + rows]
    return flat[:, :, :num_factors], flat[:, :, FP:FP + num_classes]


# ---- sigmoid link: MCSigmoidDenseFA (edward2/jax/nn/heteroscedastic_lib.py:338-594; TF heteroscedastic.py, same recipe)
def sample_sigmoids(loc, fac, diag, z, e, temperature=1.0):
    loc = np.asarray(loc, dtype=np.float64)
    B, K = loc.shape
    V = np.asarray(fac, dtype=np.float64).reshape(B, K, -1)
    z = np.broadcast_to(np.asarray(z, dtype=np.float64), (B,) + tuple(np.shape(z)[1:]))
    e = np.broadcast_to(np.asarray(e, dtype=np.float64), (B,) + tuple(np.shape(e)[1:]))
    lat = (loc[:, None, :] + np.einsum("ijk,iak->iaj", V, z) + e * diag[:, None, :]) / temperature
    return 1.0 / (1.0 + np.exp(-lat)), z, e


def mc_sigmoid_fa(loc, fac, diag_raw, z, e, temperature=1.0, eps=1e-7, sngp_var=None, w_het=1.0, w_sngp=1.0):
    """Returns (logits, log_probs, probs_mean): probs = mean_s sigmoid(latent / T) clipped from below at eps (:575-576);
    logits = inverse sigmoid with the second clip to [eps, 1 - eps] (:578-580).  sngp_var [B]: the evaluation form of
    MCSigmoidDenseFASNGP (edward2/jax/nn/random_feature.py:531-536): diag <- sqrt(w_het diag² + w_sngp var)."""
    p, _, _ = sample_sigmoids(loc, fac, effective_diag(diag_raw, sngp_var, w_het, w_sngp), z, e, temperature)
    pm = np.clip(p.mean(1), eps, None)
    log_probs = np.log(pm)
    return log_probs - np.log(1.0 - np.clip(pm, eps, 1.0 - eps)), log_probs, pm


def mc_sigmoid_fa_bwd(loc, fac, diag_raw, z, e, g, temperature=1.0, eps=1e-7):
    """Closed-form gradients of the inverse-sigmoid logits w.r.t. loc, fac, diag_raw (g = dL/dlogits)."""
    diag = effective_diag(diag_raw)
    p, zb, eb = sample_sigmoids(loc, fac, diag, z, e, temperature)
    S = p.shape[1]
    pm = p.mean(1)
    d = np.where(pm > eps, 1.0 / pm, 0.0) + np.where((pm > eps) & (pm < 1.0 - eps), 1.0 / (1.0 - pm), 0.0)
    u = np.asarray(g, dtype=np.float64) * d / S
    dlat = u[:, None, :] * p * (1.0 - p) / temperature
    dV = np.einsum("iaj,iak->ijk", dlat, zb)
    return dict(dloc=dlat.sum(1), dfac=dV.reshape(dV.shape[0], -1), ddiag_raw=(dlat * eb).sum(1) * od.sigmoid(diag_raw))
