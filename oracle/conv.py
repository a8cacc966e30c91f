"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — float64 NumPy restatement of the Conv2D variants that SURVEY.md
§8f-1 names as the next row: Conv2DReparameterization, Conv2DFlipout, Conv2DBatchEnsemble (rank 1) and Conv2DRank1
(multiplicative perturbations), NHWC ("channels_last"), strides, 'valid' / 'same' padding, dilation 1.

The convolution is restated as im2col + one matrix product on purpose: that is the shape in which the CUDA path will
reuse the tcgen05 GEMM (patches [B*Ho*Wo, kh*kw*Cin] · kernel [kh*kw*Cin, Cout]), and every stochastic-weight variant
then reduces to the dense algebra of oracle/dense.py applied to the patch matrix:
  * Reparameterization (edward2/tensorflow/layers/convolutional.py:88-98 + initializers.py:500-511):
      y = conv(x, mu + softplus(rho) * eps) + b
  * Flipout (:225-249; the conv version draws BOTH sign tensors as int32 +-1 of shape [B, 1, 1, C]):
      y = conv(x, mu) + conv(x * S_in, sigma * eps) * S_out + b
  * BatchEnsemble rank 1 (:651-699): y = act(conv(x * alpha[e]) * gamma[e] + bias[e]), member-major batch
  * Rank1 (:1968-2020): alpha / gamma sampled PER EXAMPLE from N(mean[e], softplus(rho[e])), clipped to [-10, 10],
      broadcast over the spatial axes.
Parity is unpinned against the reference itself (TF is not installable here); the restatement is pinned against
torch.nn.functional.conv2d in float64 and the reference tests' identities (tests/test_oracle.py)."""
from __future__ import annotations

import numpy as np

from . import dense as od


def _out_size(n, k, stride, padding):
    if padding == "valid":
        return (n - k) // stride + 1, 0, 0
    if padding == "same":  # TF 'SAME': out = ceil(n / stride), extra padding goes to the bottom / right
        out = -(-n // stride)
        total = max((out - 1) * stride + k - n, 0)
        return out, total // 2, total - total // 2
    raise ValueError(padding)


def im2col(x, kh, kw, strides=(1, 1), padding="valid"):
    """x [B, H, W, C] -> patches [B, Ho, Wo, kh*kw*C] (kernel-row major, then kernel-col, then channel: the flattening
    of a Keras kernel [kh, kw, Cin, Cout])."""
    x = np.asarray(x, dtype=np.float64)
    B, H, W, C = x.shape
    sh, sw = strides
    ho, pt, pb = _out_size(H, kh, sh, padding)
    wo, pl, pr = _out_size(W, kw, sw, padding)
    xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    cols = np.empty((B, ho, wo, kh, kw, C))
    for i in range(kh):
        for j in range(kw):
            cols[:, :, :, i, j, :] = xp[:, i:i + sh * ho:sh, j:j + sw * wo:sw, :]
    return cols.reshape(B, ho, wo, kh * kw * C)


def col2im(dcols, x_shape, kh, kw, strides=(1, 1), padding="valid"):
    """Adjoint of im2col: scatter-add patch gradients back to the input gradient."""
    B, H, W, C = x_shape
    sh, sw = strides
    ho, pt, pb = _out_size(H, kh, sh, padding)
    wo, pl, pr = _out_size(W, kw, sw, padding)
    dxp = np.zeros((B, H + pt + pb, W + pl + pr, C))
    d = np.asarray(dcols, dtype=np.float64).reshape(B, ho, wo, kh, kw, C)
    for i in range(kh):
        for j in range(kw):
            dxp[:, i:i + sh * ho:sh, j:j + sw * wo:sw, :] += d[:, :, :, i, j, :]
    return dxp[:, pt:pt + H, pl:pl + W, :]


def conv2d(x, kernel, strides=(1, 1), padding="valid"):
    """Keras Conv2D.convolution_op, NHWC, kernel [kh, kw, Cin, Cout]."""
    kh, kw, cin, cout = kernel.shape
    cols = im2col(x, kh, kw, strides, padding)
    return cols @ np.asarray(kernel, dtype=np.float64).reshape(kh * kw * cin, cout)


def conv2d_bwd(x, kernel, gy, strides=(1, 1), padding="valid"):
    kh, kw, cin, cout = kernel.shape
    cols = im2col(x, kh, kw, strides, padding)
    g = np.asarray(gy, dtype=np.float64)
    dk = (cols.reshape(-1, kh * kw * cin).T @ g.reshape(-1, cout)).reshape(kernel.shape)
    dcols = g @ np.asarray(kernel, dtype=np.float64).reshape(kh * kw * cin, cout).T
    return col2im(dcols, np.shape(x), kh, kw, strides, padding), dk


# ------------------------------------------------------------------------------------------- variants (forward)
def reparam_conv2d(x, mu, rho, eps, bias=None, act=None, strides=(1, 1), padding="valid"):
    """convolutional.py:88-98: one kernel sample per call, shared by the batch."""
    w = np.asarray(mu, dtype=np.float64) + od.stddev(rho) * np.asarray(eps, dtype=np.float64)
    y = conv2d(x, w, strides, padding)
    if bias is not None:
        y = y + np.asarray(bias, dtype=np.float64)
    return od.activation(y, act), w


def flipout_conv2d(x, mu, rho, eps, sign_in, sign_out, bias=None, act=None, strides=(1, 1), padding="valid"):
    """convolutional.py:225-249: sign_in [B, Cin], sign_out [B, Cout] (+-1), broadcast over the spatial axes."""
    x = np.asarray(x, dtype=np.float64)
    delta = od.stddev(rho) * np.asarray(eps, dtype=np.float64)
    y = conv2d(x, mu, strides, padding)
    y = y + conv2d(x * np.asarray(sign_in, dtype=np.float64)[:, None, None, :], delta, strides, padding) * \
        np.asarray(sign_out, dtype=np.float64)[:, None, None, :]
    if bias is not None:
        y = y + np.asarray(bias, dtype=np.float64)
    return od.activation(y, act), delta


def batch_ensemble_conv2d(x, kernel, alpha, gamma, bias=None, act=None, strides=(1, 1), padding="valid"):
    """convolutional.py:681-699 (rank 1): member-major batch [E * n, H, W, C]."""
    x = np.asarray(x, dtype=np.float64)
    E = alpha.shape[0]
    n = x.shape[0] // E
    a = np.repeat(np.asarray(alpha, dtype=np.float64), n, axis=0)[:, None, None, :]
    g = np.repeat(np.asarray(gamma, dtype=np.float64), n, axis=0)[:, None, None, :]
    y = conv2d(x * a, kernel, strides, padding) * g
    if bias is not None:
        y = y + np.repeat(np.asarray(bias, dtype=np.float64), n, axis=0)[:, None, None, :]
    return od.activation(y, act)


def rank1_conv2d(x, kernel, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias=None, act=None, strides=(1, 1),
                 padding="valid"):
    """convolutional.py:1968-2020, multiplicative perturbations: r [B, Cin], s [B, Cout] drawn per example
    (oracle/dense.py:rank1_factors, incl. the [-10, 10] clip) and broadcast over the spatial axes."""
    x = np.asarray(x, dtype=np.float64)
    E = mu_r.shape[0]
    n = x.shape[0] // E
    r, _, _ = od.rank1_factors(mu_r, rho_r, eps_r, E)
    s, _, _ = od.rank1_factors(mu_s, rho_s, eps_s, E)
    y = conv2d(x * r[:, None, None, :], kernel, strides, padding) * s[:, None, None, :]
    if bias is not None:
        y = y + np.repeat(np.asarray(bias, dtype=np.float64), n, axis=0)[:, None, None, :]
    return od.activation(y, act), r, s


# ------------------------------------------------------------------------------------------- backward (closed form)
def reparam_conv2d_bwd(x, mu, rho, eps, bias, act, gy, strides=(1, 1), padding="valid"):
    y, w = reparam_conv2d(x, mu, rho, eps, bias, act, strides, padding)
    gp = od.activation_grad(y, np.asarray(gy, dtype=np.float64), act)
    dx, dw = conv2d_bwd(x, w, gp, strides, padding)
    return dict(dx=dx, dmu=dw, drho=dw * np.asarray(eps, dtype=np.float64) * od.sigmoid(rho), dbias=gp.sum((0, 1, 2)))


def batch_ensemble_conv2d_bwd(x, kernel, alpha, gamma, bias, act, gy, strides=(1, 1), padding="valid"):
    x = np.asarray(x, dtype=np.float64)
    E = alpha.shape[0]
    n = x.shape[0] // E
    a = np.repeat(np.asarray(alpha, dtype=np.float64), n, axis=0)[:, None, None, :]
    g = np.repeat(np.asarray(gamma, dtype=np.float64), n, axis=0)[:, None, None, :]
    z = conv2d(x * a, kernel, strides, padding)
    y = batch_ensemble_conv2d(x, kernel, alpha, gamma, bias, act, strides, padding)
    gp = od.activation_grad(y, np.asarray(gy, dtype=np.float64), act)
    dxa, dk = conv2d_bwd(x * a, kernel, gp * g, strides, padding)
    per_member = lambda t: t.sum((1, 2)).reshape(E, n, -1).sum(1)  # noqa: E731
    return dict(dx=dxa * a, dkernel=dk, dalpha=per_member(dxa * x), dgamma=per_member(gp * z), dbias=per_member(gp))


def flipout_conv2d_bwd(x, mu, rho, eps, sign_in, sign_out, bias, act, gy, strides=(1, 1), padding="valid", y_mask=None):
    """Closed-form gradients of flipout_conv2d (convolutional.py:225-249): the mean and the perturbation convolutions are
    separate linear maps of x; Delta = sigma∘eps gives drho = dDelta ∘ eps ∘ sigmoid(rho)."""
    x = np.asarray(x, dtype=np.float64)
    si = np.asarray(sign_in, dtype=np.float64)[:, None, None, :]
    so = np.asarray(sign_out, dtype=np.float64)[:, None, None, :]
    y, delta = flipout_conv2d(x, mu, rho, eps, sign_in, sign_out, bias, act, strides, padding)
    gp = od.activation_grad(y if y_mask is None else y_mask, np.asarray(gy, dtype=np.float64), act)
    dx1, dmean = conv2d_bwd(x, mu, gp, strides, padding)
    dxs, ddelta = conv2d_bwd(x * si, delta, gp * so, strides, padding)
    return dict(dx=dx1 + dxs * si, dmu=dmean, drho=ddelta * np.asarray(eps, dtype=np.float64) * od.sigmoid(rho),
                dbias=gp.sum((0, 1, 2)))


def rank1_conv2d_bwd(x, kernel, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, act, gy, strides=(1, 1), padding="valid",
                     y_mask=None):
    """Closed-form gradients of rank1_conv2d (multiplicative; convolutional.py:1968-2020): the dense rank-1 algebra
    (oracle/dense.py:rank1_bwd) with the per-example factors broadcast over, and their gradients summed over, the
    spatial axes."""
    x = np.asarray(x, dtype=np.float64)
    E = mu_r.shape[0]
    n = x.shape[0] // E
    r, mask_r, er = od.rank1_factors(mu_r, rho_r, eps_r, E)
    s, mask_s, es = od.rank1_factors(mu_s, rho_s, eps_s, E)
    z = conv2d(x * r[:, None, None, :], kernel, strides, padding)
    y, _, _ = rank1_conv2d(x, kernel, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, act, strides, padding)
    gp = od.activation_grad(y if y_mask is None else y_mask, np.asarray(gy, dtype=np.float64), act)
    dxr, dk = conv2d_bwd(x * r[:, None, None, :], kernel, gp * s[:, None, None, :], strides, padding)
    ds = (gp * z).sum((1, 2)) * mask_s
    dr = (dxr * x).sum((1, 2)) * mask_r
    member = lambda t: t.reshape(E, n, -1).sum(1)  # noqa: E731
    return dict(dx=dxr * r[:, None, None, :], dkernel=dk, dmu_r=member(dr),
                drho_r=None if rho_r is None else member(dr * er) * od.sigmoid(rho_r), dmu_s=member(ds),
                drho_s=None if rho_s is None else member(ds * es) * od.sigmoid(rho_s), dbias=member(gp.sum((1, 2))))
