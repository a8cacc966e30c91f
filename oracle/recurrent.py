"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — float64 NumPy restatement of LSTMCellReparameterization
(edward2/tensorflow/layers/recurrent.py:30-183; the step arithmetic is Keras LSTMCell.call with implementation 2 and
_compute_carry_and_output :163-183, gate order i | f | c | o): one kernel / recurrent-kernel sample W = mu + sigma∘eps per
SEQUENCE (:138-161), reused by every step; backward = backpropagation through time in closed form."""
from __future__ import annotations

import numpy as np

from . import dense as od


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-np.asarray(x, dtype=np.float64)))


def lstm_step(x, h, c, W, U, b):
    H = h.shape[1]
    z = x @ W + h @ U + (0.0 if b is None else b)
    i, f, g, o = sigmoid(z[:, :H]), sigmoid(z[:, H:2 * H]), np.tanh(z[:, 2 * H:3 * H]), sigmoid(z[:, 3 * H:])
    c_new = f * c + i * g
    return o * np.tanh(c_new), c_new, (i, f, g, o)


def lstm_sequence(xs, h0, c0, W, U, b):
    """xs [T, B, I] -> (hs [T, B, H], c_T, caches)."""
    h, c, hs, caches = h0, c0, [], []
    for x in xs:
        h_new, c_new, gates = lstm_step(x, h, c, W, U, b)
        caches.append((x, h, c, c_new, gates))
        h, c = h_new, c_new
        hs.append(h)
    return np.stack(hs), c, caches


def lstm_sequence_bwd(caches, W, U, dhs, dc_last=None):
    """Gradients of sum_t <dhs[t], h_t> (+ <dc_last, c_T>): dxs, dh0, dc0, dW, dU, db."""
    T = len(caches)
    dW, dU = np.zeros_like(W), np.zeros_like(U)
    db = np.zeros(W.shape[1])
    dh_next = np.zeros_like(dhs[0])
    dc_next = np.zeros_like(dhs[0]) if dc_last is None else np.asarray(dc_last, dtype=np.float64)
    dxs = [None] * T
    for t in reversed(range(T)):
        x, h_prev, c_prev, c_new, (i, f, g, o) = caches[t]
        dh = dhs[t] + dh_next
        tc = np.tanh(c_new)
        dct = dc_next + dh * o * (1 - tc * tc)
        dz = np.concatenate([dct * g * i * (1 - i), dct * c_prev * f * (1 - f), dct * i * (1 - g * g), dh * tc * o * (1 - o)], 1)
        dW += x.T @ dz
        dU += h_prev.T @ dz
        db += dz.sum(0)
        dxs[t] = dz @ W.T
        dh_next = dz @ U.T
        dc_next = dct * f
    return dict(dxs=np.stack(dxs), dh0=dh_next, dc0=dc_next, dW=dW, dU=dU, db=db)


def reparam_kernel(mu, rho, eps):
    return np.asarray(mu, dtype=np.float64) + od.stddev(rho) * np.asarray(eps, dtype=np.float64)
