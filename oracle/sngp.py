"""float64 NumPy restatement of the reference's SNGP head: SpectralNormalization, RandomFeatureGaussianProcess /
RandomFourierFeatures, LaplaceRandomFeatureCovariance, mean_field_logits.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Citations are to /root/reference/edward2/...
"""
import numpy as np


# ------------------------------------------------------------------------------------------- SpectralNormalization
def l2_normalize(x, eps=1e-12):
    """tf.nn.l2_normalize with axis=None (3p) == edward2/jax/nn/normalization.py:50-51."""
    x = np.asarray(x, dtype=np.float64)
    return x / np.sqrt(max(np.sum(x * x), eps))


def spectral_norm_tf(w, u, v, iteration=1, norm_multiplier=0.95, training=True):
    """edward2/tensorflow/layers/normalization.py:369-391.  w [I, O]; u [1, O]; v [1, I].
    Returns (w_norm, u_hat, v_hat, sigma)."""
    w = np.asarray(w, dtype=np.float64)
    u_hat = np.asarray(u, dtype=np.float64).reshape(1, -1)
    v_hat = np.asarray(v, dtype=np.float64).reshape(1, -1)
    if training:
        for _ in range(iteration):
            v_hat = l2_normalize(u_hat @ w.T)
            u_hat = l2_normalize(v_hat @ w)
    sigma = float((v_hat @ w @ u_hat.T).reshape(()))
    w_norm = (norm_multiplier / sigma) * w if (norm_multiplier / sigma) < 1 else w
    return w_norm, u_hat, v_hat, sigma


def spectral_norm_jax(w, u, v, iteration=1, norm_multiplier=0.95, training=True):
    """edward2/jax/nn/normalization.py:159-178 (Dense kernel: kernel_apply(x) = x @ w).  u [O], v [I]."""
    w = np.asarray(w, dtype=np.float64)
    u_hat = np.asarray(u, dtype=np.float64).reshape(-1)
    v_hat = np.asarray(v, dtype=np.float64).reshape(-1)
    u_ = v_hat @ w
    if training:
        for _ in range(iteration):
            v_ = w @ u_hat          # kernel_transpose(u_hat)
            v_hat = l2_normalize(v_)
            u_ = v_hat @ w
            u_hat = l2_normalize(u_)
    sigma = float(np.vdot(u_hat, u_))
    div = max(sigma / norm_multiplier, 1.0)
    return w / div, u_hat, v_hat, sigma


# ------------------------------------------------------------------------------------------- input normalisation
def layer_norm(x, gamma=None, beta=None, eps=1e-3):
    """Keras LayerNormalization over the last axis (3p; eps=1e-3, trainable gamma/beta) used at
    edward2/tensorflow/layers/random_feature.py:167-171,252; Flax nn.LayerNorm(use_bias=False, use_scale=False),
    eps=1e-6 (3p) at edward2/jax/nn/random_feature.py:98-102."""
    x = np.asarray(x, dtype=np.float64)
    mean = x.mean(-1, keepdims=True)
    var = ((x - mean) ** 2).mean(-1, keepdims=True)
    y = (x - mean) / np.sqrt(var + eps)
    if gamma is not None:
        y = y * gamma
    if beta is not None:
        y = y + beta
    return y


# ------------------------------------------------------------------------------------------- random features
def rff(x, w, b, scale):
    """edward2/tensorflow/layers/random_feature.py:258-266; edward2/jax/nn/random_feature.py:209-214:
    phi = scale * cos(x w + b).  x [B, d]; w [d, D]; b [D]."""
    pre = np.asarray(x, dtype=np.float64) @ np.asarray(w, dtype=np.float64) + np.asarray(b, dtype=np.float64)
    return scale * np.cos(pre), pre


def rff_bwd(x, w, b, scale, dphi):
    _, pre = rff(x, w, b, scale)
    return (-scale * np.sin(pre) * np.asarray(dphi, dtype=np.float64)) @ np.asarray(w, dtype=np.float64).T


def gp_output(phi, beta, bias):
    """random_feature.py:269: Dense(units, use_bias=False)(phi) + bias variable."""
    return np.asarray(phi, dtype=np.float64) @ np.asarray(beta, dtype=np.float64) + np.asarray(bias, dtype=np.float64)


# ------------------------------------------------------------------------------------------- Laplace covariance
def prob_multiplier(logits, likelihood):
    """random_feature.py:398-406; jax random_feature.py:341-348."""
    if likelihood == "gaussian":
        return 1.0
    logits = np.asarray(logits, dtype=np.float64)
    if logits.ndim > 1 and logits.shape[-1] != 1:
        raise ValueError(f"likelihood={likelihood} only support univariate logits. Got logits dimension: {logits.shape[-1]}")
    if likelihood == "binary_logistic":
        p = 1.0 / (1.0 + np.exp(-logits))
        return p * (1.0 - p)
    if likelihood == "poisson":
        return np.exp(logits)
    raise ValueError(likelihood)


def precision_update_tf(precision, phi, momentum=0.999, logits=None, likelihood="gaussian"):
    """edward2/tensorflow/layers/random_feature.py:394-423."""
    phi = np.sqrt(prob_multiplier(logits, likelihood)) * np.asarray(phi, dtype=np.float64)
    mb = phi.T @ phi
    if momentum > 0:
        return momentum * precision + (1.0 - momentum) * mb / phi.shape[0]
    return precision + mb


def precision_update_jax(precision, phi, momentum=None, logits=None, likelihood="gaussian"):
    """edward2/jax/nn/random_feature.py:350-362 (initial precision = ridge * I, :307-309)."""
    phi = np.sqrt(prob_multiplier(logits, likelihood)) * np.asarray(phi, dtype=np.float64)
    mb = phi.T @ phi
    if momentum is None:
        return precision + mb
    return momentum * precision + (1 - momentum) * mb / phi.shape[0]


def covariance_tf(precision, ridge):
    """random_feature.py:447-459: inv(ridge * I + P)."""
    d = precision.shape[0]
    return np.linalg.inv(ridge * np.eye(d) + np.asarray(precision, dtype=np.float64))


def predictive_covariance_tf(phi, covariance, ridge):
    """random_feature.py:482-485: phi (Sigma phiᵀ * ridge), full [B, B]."""
    phi = np.asarray(phi, dtype=np.float64)
    return phi @ (covariance @ phi.T * ridge)


def predictive_covariance_jax(phi, precision, ridge, diagonal_only=True):
    """edward2/jax/nn/random_feature.py:393-404: Cholesky of the precision, triangular solve."""
    phi = np.asarray(phi, dtype=np.float64)
    chol = np.linalg.cholesky(np.asarray(precision, dtype=np.float64))
    a = np.linalg.solve(chol, phi.T)  # lower-triangular solve
    if diagonal_only:
        return ridge * np.square(a).sum(0)
    return ridge * a.T @ a


def predictive_variance(phi, covariance, ridge):
    """diag of predictive_covariance_tf without forming [B, B]: ridge * rowsum((phi Sigma) * phi)."""
    phi = np.asarray(phi, dtype=np.float64)
    return ridge * ((phi @ covariance) * phi).sum(1)


# ------------------------------------------------------------------------------------------- mean-field logits
def mean_field_logits(logits, covmat=None, mean_field_factor=1.0, likelihood="logistic"):
    """edward2/tensorflow/layers/utils.py:379-424 (covmat [B, B] -> its diagonal);
    edward2/jax/nn/utils.py:54-101 (covmat is the [B] variance vector)."""
    if likelihood not in ("logistic", "binary_logistic", "poisson"):
        raise ValueError(f'Likelihood" must be one of (\'logistic\', \'binary_logistic\', \'poisson\'), got {likelihood}.')
    logits = np.asarray(logits, dtype=np.float64)
    if mean_field_factor < 0:
        return logits
    if covmat is None:
        variances = 1.0
    else:
        covmat = np.asarray(covmat, dtype=np.float64)
        variances = np.diag(covmat) if covmat.ndim == 2 else covmat
    if likelihood == "poisson":
        scale = np.exp(-variances * mean_field_factor / 2.0)
    else:
        scale = np.sqrt(1.0 + variances * mean_field_factor)
    if np.ndim(scale) > 0 and logits.ndim > 1:
        scale = np.expand_dims(scale, -1)
    return logits / scale
