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


# ---- two-branch cells: LSTMCellFlipout (recurrent.py:185-358) and LSTMCellRank1 (:360-830) -----------------------------
def gates(z, c_prev):
    """(h, c, cache) from the summed pre-activations z [B, 4H] (gate order i | f | c | o)."""
    H = c_prev.shape[1]
    i, f, g, o = sigmoid(z[:, :H]), sigmoid(z[:, H:2 * H]), np.tanh(z[:, 2 * H:3 * H]), sigmoid(z[:, 3 * H:])
    c = f * c_prev + i * g
    return o * np.tanh(c), c, (i, f, g, o)


def gates_bwd(c_prev, c, cache, dh, dc):
    """(dz, dc_prev) of `gates`."""
    i, f, g, o = cache
    tc = np.tanh(c)
    dct = dc + dh * o * (1 - tc * tc)
    dz = np.concatenate([dct * g * i * (1 - i), dct * c_prev * f * (1 - f), dct * i * (1 - g * g), dh * tc * o * (1 - o)], 1)
    return dz, dct * f


def two_branch_sequence(xs, h0, c0, branch_x, branch_h, bwd_x, bwd_h, dhs):
    """Unrolled cell whose pre-activation is branch_x(x) + branch_h(h): forward over xs [T, B, I] and backpropagation
    through time of sum_t <dhs[t], h_t>.  branch_* (inp) -> z;  bwd_* (inp, dz) -> dict with 'dx' and parameter gradients
    (summed over the steps here).  Returns (hs, dxs, grads_x, grads_h)."""
    h, c, hs, tape = h0, c0, [], []
    for x in xs:
        z = branch_x(x) + branch_h(h)
        h_new, c_new, cache = gates(z, c)
        tape.append((x, h, c, c_new, cache))
        h, c = h_new, c_new
        hs.append(h)
    dh_next, dc_next = np.zeros_like(h0), np.zeros_like(c0)
    dxs, gx, gh = [None] * len(xs), {}, {}
    for t in reversed(range(len(xs))):
        x, h_prev, c_prev, c_new, cache = tape[t]
        dz, dc_next = gates_bwd(c_prev, c_new, cache, dhs[t] + dh_next, dc_next)
        bx, bh = bwd_x(x, dz), bwd_h(h_prev, dz)
        dxs[t], dh_next = bx.pop("dx"), bh.pop("dx")
        for acc, b in ((gx, bx), (gh, bh)):
            for k, v in b.items():
                if v is not None:
                    acc[k] = acc.get(k, 0.0) + v
    return np.stack(hs), np.stack(dxs), gx, gh
