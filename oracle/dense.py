"""float64 NumPy restatement of the reference's stochastic-weight dense layers, with closed-form gradients.

TEST INFRASTRUCTURE (see oracle/__init__.py).  All randomness (eps, signs, per-example noise) is an argument.
Citations are to /root/reference/edward2/... ; "(3p)" marks behaviour of un-vendored TF/TFP/Keras code.

Batch layout for ensemble layers: rows are member-major, row b = e * n + i  (dense.py:572-573, 1101-1102).
Per-example noise for DenseRank1 is passed as eps[b, :] with the same row order; the reference draws it as
`distribution.sample(n)` -> [n, E, I] and transposes to [E, n, I] (dense.py:1105-1112), i.e.
eps_ref[i, e, :] == eps[e * n + i, :].
"""
import numpy as np

SOFTPLUS_EPS = 1e-7  # tf.keras.backend.epsilon() (3p), used at edward2/tensorflow/constraints.py:56-63
CLIP = 10.0          # dense.py:1025-1026 defaults min/max_perturbation_value


def softplus(x):
    x = np.asarray(x, dtype=np.float64)
    return np.logaddexp(0.0, x)


def sigmoid(x):
    x = np.asarray(x, dtype=np.float64)
    return 1.0 / (1.0 + np.exp(-x))


def stddev(rho):
    """edward2/tensorflow/constraints.py:62-63 Softplus constraint applied at initializers.py:506-509."""
    return softplus(rho) + SOFTPLUS_EPS


def activation(z, act):
    if act in (None, "linear"):
        return z
    if act == "relu":
        return np.maximum(z, 0.0)
    raise ValueError(act)


def activation_grad(y, gy, act):
    """dL/dz from dL/dy; relu' is evaluated on the output (y > 0), as the kernels do."""
    if act in (None, "linear"):
        return gy
    if act == "relu":
        return gy * (y > 0)
    raise ValueError(act)


# ------------------------------------------------------------------------------------------- NormalKLDivergence
def normal_kl(mu, rho, prior_mean=0.0, prior_std=1.0, scale_factor=1.0):
    """edward2/tensorflow/regularizers.py:175-186: scale_factor * sum KL(N(mu, sigma) || N(m, s)).
    TFP Normal.kl_divergence (3p) == log(s/sigma) + (sigma^2 + (mu-m)^2) / (2 s^2) - 1/2."""
    sigma = stddev(rho)
    mu = np.asarray(mu, dtype=np.float64)
    kl = np.log(prior_std / sigma) + (sigma ** 2 + (mu - prior_mean) ** 2) / (2.0 * prior_std ** 2) - 0.5
    return scale_factor * kl.sum()


def normal_kl_grad(mu, rho, prior_mean=0.0, prior_std=1.0, scale_factor=1.0):
    sigma = stddev(rho)
    mu = np.asarray(mu, dtype=np.float64)
    dmu = scale_factor * (mu - prior_mean) / prior_std ** 2
    drho = scale_factor * (sigma / prior_std ** 2 - 1.0 / sigma) * sigmoid(rho)
    return dmu, drho


# ------------------------------------------------------------------------------------------- DenseBatchEnsemble
def batch_ensemble_fwd(x, w, alpha, gamma, bias=None, act=None):
    """edward2/tensorflow/layers/dense.py:551-584 (rank-1 branch); edward2/jax/nn/dense.py:258-272.
    x [E*n, I]; w [I, O]; alpha [E, I]; gamma [E, O]; bias [E, O] or None."""
    x = np.asarray(x, dtype=np.float64)
    E = alpha.shape[0]
    n = x.shape[0] // E
    xe = x.reshape(E, n, -1)
    z = (xe * alpha[:, None, :]) @ np.asarray(w, dtype=np.float64)
    pre = z * gamma[:, None, :]
    if bias is not None:
        pre = pre + bias[:, None, :]
    y = activation(pre, act)
    return y.reshape(E * n, -1), z.reshape(E * n, -1)


def batch_ensemble_rank_fwd(x, w, alpha, gamma, bias=None, act=None):
    """edward2/tensorflow/layers/dense.py:560-570 (rank > 1 branch), statement by statement: alpha [R, E, I] and gamma
    [R, E, O] are tiled along their last axis and reshaped to [R, B, dim], the perturbed inputs go through the shared
    kernel, the products are weighted by gamma and summed over the rank axis."""
    x = np.asarray(x, dtype=np.float64)
    R, E, I = alpha.shape
    B = x.shape[0]
    n = B // E
    a = np.tile(alpha, (1, 1, n)).reshape(R, B, I)
    g = np.tile(gamma, (1, 1, n)).reshape(R, B, -1)
    out = (((x[None] * a) @ np.asarray(w, dtype=np.float64)) * g).sum(0)
    out = out.reshape(E, n, -1)
    if bias is not None:
        out = out + np.asarray(bias, dtype=np.float64)[:, None, :]
    return activation(out, act).reshape(B, -1)


def batch_ensemble_bwd(x, w, alpha, gamma, bias, act, gy, y_mask=None):
    """Closed-form gradients of batch_ensemble_fwd.  y_mask: optional forward output whose sign
    pattern defines relu' (bf16 parity is conditioned on the kernel's own activation pattern)."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    E = alpha.shape[0]
    n = x.shape[0] // E
    y, z = batch_ensemble_fwd(x, w, alpha, gamma, bias, act)
    gp = activation_grad(y if y_mask is None else y_mask, np.asarray(gy, dtype=np.float64), act).reshape(E, n, -1)
    ze = z.reshape(E, n, -1)
    xe = x.reshape(E, n, -1)
    dz = gp * gamma[:, None, :]
    dgamma = (gp * ze).sum(1)
    dbias = gp.sum(1)
    xa = xe * alpha[:, None, :]
    dw = np.einsum("eni,eno->io", xa, dz)
    dxa = dz @ w.T
    dx = dxa * alpha[:, None, :]
    dalpha = (dxa * xe).sum(1)
    return dict(dx=dx.reshape(E * n, -1), dw=dw, dalpha=dalpha, dgamma=dgamma, dbias=dbias)


# ------------------------------------------------------------------------------------------- DenseRank1
def rank1_factors(mu, rho, eps, E, clip=(-CLIP, CLIP)):
    """Per-example fast weights (dense.py:1105-1112): clip(mu[e] + sigma[e] * eps[b], min, max); rho None ->
    deterministic broadcast (dense.py:1114).  Returns (factor [B, D], pass-through mask, eps).  clip=None: unclipped
    (the per-example ensemble bias, dense.py:1131-1137)."""
    B = eps.shape[0] if eps is not None else None
    if rho is None:
        n = B // E
        f = np.repeat(np.asarray(mu, dtype=np.float64), n, axis=0)
        return f, np.ones_like(f), np.zeros_like(f)
    n = B // E
    m = np.repeat(np.asarray(mu, dtype=np.float64), n, axis=0)
    s = np.repeat(stddev(rho), n, axis=0)
    raw = m + s * np.asarray(eps, dtype=np.float64)
    if clip is None:
        return raw, np.ones_like(raw), np.asarray(eps, dtype=np.float64)
    lo, hi = clip
    mask = ((raw >= lo) & (raw <= hi)).astype(np.float64)  # tf.clip_by_value gradient (3p)
    return np.clip(raw, lo, hi), mask, np.asarray(eps, dtype=np.float64)


def rank1_fwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias=None, act=None, additive=False, clip=(-CLIP, CLIP),
              rho_b=None, eps_b=None):
    """edward2/tensorflow/layers/dense.py:1096-1144: multiplicative (:1128-1129) or additive (:1126-1127)
    perturbations, clip bounds (:1108-1110, :1118-1120), deterministic or per-example sampled bias (:1131-1139).
    eps_r [B, I] / eps_s [B, O] / eps_b [B, O] (ignored when the matching rho is None; pass zeros for the shape)."""
    x = np.asarray(x, dtype=np.float64)
    E = mu_r.shape[0]
    n = x.shape[0] // E
    r, _, _ = rank1_factors(mu_r, rho_r, eps_r, E, clip)
    s, _, _ = rank1_factors(mu_s, rho_s, eps_s, E, clip)
    w = np.asarray(w, dtype=np.float64)
    if additive:
        z = (x + r) @ w
        pre = z + s
    else:
        z = (x * r) @ w
        pre = z * s
    if bias is not None:
        if rho_b is not None:
            b, _, _ = rank1_factors(bias, rho_b, eps_b, E, None)
            pre = pre + b
        else:
            pre = pre + np.repeat(np.asarray(bias, dtype=np.float64), n, axis=0)
    return activation(pre, act), z, r, s


def rank1_bwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, act, gy, y_mask=None, additive=False,
              clip=(-CLIP, CLIP), rho_b=None, eps_b=None):
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    E = mu_r.shape[0]
    n = x.shape[0] // E
    y, z, r, s = rank1_fwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, act, additive, clip, rho_b, eps_b)
    _, mask_r, er = rank1_factors(mu_r, rho_r, eps_r, E, clip)
    _, mask_s, es = rank1_factors(mu_s, rho_s, eps_s, E, clip)
    gp = activation_grad(y if y_mask is None else y_mask, np.asarray(gy, dtype=np.float64), act)
    if additive:
        dz = gp
        ds = gp * mask_s
    else:
        dz = gp * s
        ds = gp * z * mask_s
    dmu_s = ds.reshape(E, n, -1).sum(1)
    drho_s = None if rho_s is None else (ds * es).reshape(E, n, -1).sum(1) * sigmoid(rho_s)
    dbias = gp.reshape(E, n, -1).sum(1)
    drho_b = None
    if rho_b is not None:
        drho_b = (gp * np.asarray(eps_b, dtype=np.float64)).reshape(E, n, -1).sum(1) * sigmoid(rho_b)
    dxr = dz @ w.T
    if additive:
        dw = (x + r).T @ dz
        dx = dxr
        dr = dxr * mask_r
    else:
        dw = (x * r).T @ dz
        dx = dxr * r
        dr = dxr * x * mask_r
    dmu_r = dr.reshape(E, n, -1).sum(1)
    drho_r = None if rho_r is None else (dr * er).reshape(E, n, -1).sum(1) * sigmoid(rho_r)
    return dict(dx=dx, dw=dw, dmu_r=dmu_r, drho_r=drho_r, dmu_s=dmu_s, drho_s=drho_s, dbias=dbias, drho_b=drho_b)


# ------------------------------------------------------------------------------------------- DenseReparameterization
def reparam_fwd(x, mu, rho, eps, bias=None, act=None):
    """edward2/tensorflow/layers/dense.py:73-83 + initializers.py:500-511: y = act(x (mu + sigma*eps) + b);
    one kernel sample per call, shared by the batch."""
    w = np.asarray(mu, dtype=np.float64) + stddev(rho) * np.asarray(eps, dtype=np.float64)
    pre = np.asarray(x, dtype=np.float64) @ w
    if bias is not None:
        pre = pre + np.asarray(bias, dtype=np.float64)
    return activation(pre, act), w


def reparam_bwd(x, mu, rho, eps, bias, act, gy, y_mask=None):
    x = np.asarray(x, dtype=np.float64)
    y, w = reparam_fwd(x, mu, rho, eps, bias, act)
    gp = activation_grad(y if y_mask is None else y_mask, np.asarray(gy, dtype=np.float64), act)
    dw = x.T @ gp
    return dict(dx=gp @ w.T, dmu=dw, drho=dw * np.asarray(eps, dtype=np.float64) * sigmoid(rho), dbias=gp.sum(0))


# ------------------------------------------------------------------------------------------- DenseFlipout
def flipout_fwd(x, mu, rho, eps, sign_in, sign_out, bias=None, act=None):
    """edward2/tensorflow/layers/dense.py:255-285: y = act(x mu + ((x*S) Delta) * T + b), Delta = kernel - mean.
    sign_out may be any float tensor (the reference draws Uniform[-1,3) for float inputs, dense.py:266-270)."""
    x = np.asarray(x, dtype=np.float64)
    mu = np.asarray(mu, dtype=np.float64)
    delta = stddev(rho) * np.asarray(eps, dtype=np.float64)
    pre = x @ mu + ((x * sign_in) @ delta) * sign_out
    if bias is not None:
        pre = pre + np.asarray(bias, dtype=np.float64)
    return activation(pre, act), delta


def flipout_bwd(x, mu, rho, eps, sign_in, sign_out, bias, act, gy, y_mask=None):
    x = np.asarray(x, dtype=np.float64)
    mu = np.asarray(mu, dtype=np.float64)
    y, delta = flipout_fwd(x, mu, rho, eps, sign_in, sign_out, bias, act)
    gp = activation_grad(y if y_mask is None else y_mask, np.asarray(gy, dtype=np.float64), act)
    gt = gp * sign_out
    dmu = x.T @ gp
    ddelta = (x * sign_in).T @ gt
    drho = ddelta * np.asarray(eps, dtype=np.float64) * sigmoid(rho)
    dx = gp @ mu.T + (gt @ delta.T) * sign_in
    return dict(dx=dx, dmu=dmu, drho=drho, dbias=gp.sum(0))


# ------------------------------------------------------------------------------------------- loss used by benches
def softmax_xent(logits, labels):
    """Mean softmax cross-entropy and its gradient w.r.t. logits (SURVEY.md §8d synthetic loss)."""
    z = np.asarray(logits, dtype=np.float64)
    z = z - z.max(1, keepdims=True)
    p = np.exp(z)
    p /= p.sum(1, keepdims=True)
    B = z.shape[0]
    loss = -np.log(p[np.arange(B), labels]).mean()
    g = p.copy()
    g[np.arange(B), labels] -= 1.0
    return loss, g / B
