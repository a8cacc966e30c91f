"""Bayesian LSTM cell on libed2b200: LSTMCellReparameterization (edward2/tensorflow/layers/recurrent.py:30-183).

The kernel and the recurrent kernel have Normal posteriors whose noise is drawn ONCE per sequence (`call_weights()` /
`get_initial_state()`, recurrent.py:138-161) and reused by every time step; a step is two tcgen05 GEMMs (the second
accumulates onto the first) and one pointwise gate kernel, its backward one pointwise kernel and four GEMMs, and the raw
weight gradients of all steps are summed on the tensor cores before ONE finishing pass per kernel regenerates the noise
and adds the KL gradient (functional.SequenceKernel).  `LSTMCellFlipout` (:185) and `LSTMCellRank1` (:360) keep the input
and the recurrent branch as two stochastic-weight products of the library (Flipout / rank-1 dense) whose pre-activations
meet in the gate kernel."""
from __future__ import annotations

import torch
from torch import nn

from .. import functional as F, initializers, regularizers
from .base import Layer
from .dense import _is_trainable, _make_weight

S_LSTM_KERNEL, S_LSTM_RECURRENT = 8, 9


class LSTMCellReparameterization(Layer):
    """call(inputs [B, input_dim], states [h, c]) -> (h, [h, c]); `.losses` = KL of the kernel and the recurrent kernel."""

    def __init__(self, units, activation="tanh", recurrent_activation="sigmoid", use_bias=True,
                 kernel_initializer="trainable_normal", recurrent_initializer="trainable_normal", bias_initializer="zeros",
                 unit_forget_bias=True, kernel_regularizer="normal_kl_divergence",
                 recurrent_regularizer="normal_kl_divergence", bias_regularizer=None, kernel_constraint=None,
                 recurrent_constraint=None, bias_constraint=None, dropout=0.0, recurrent_dropout=0.0, implementation=2,
                 **kwargs):
        super().__init__(**kwargs)
        if activation != "tanh" or recurrent_activation != "sigmoid":
            raise NotImplementedError("the fused gate kernel implements activation='tanh', recurrent_activation='sigmoid'")
        if dropout or recurrent_dropout:
            raise NotImplementedError("dropout / recurrent_dropout are not built")
        self.units = int(units)
        self.use_bias, self.unit_forget_bias = use_bias, unit_forget_bias
        self.kernel_initializer = initializers.get(kernel_initializer)
        self.recurrent_initializer = initializers.get(recurrent_initializer)
        self.bias_initializer = initializers.get(bias_initializer)
        self.kernel_regularizer = regularizers.get(kernel_regularizer)
        self.recurrent_regularizer = regularizers.get(recurrent_regularizer)
        self.bias_regularizer = regularizers.get(bias_regularizer)
        self.state_size = [self.units, self.units]
        self.output_size = self.units
        self.called_weights = False
        self._kw = self._ku = None

    def build(self, input_shape):
        dev = self._device
        H = self.units
        self.kernel = _make_weight(self, "kernel", (input_shape[-1], 4 * H), self.kernel_initializer, dev)
        self.recurrent_kernel = _make_weight(self, "recurrent_kernel", (H, 4 * H), self.recurrent_initializer, dev)
        if _is_trainable(self.bias_initializer):
            raise NotImplementedError("a trainable bias initializer is not built for the LSTM cell")
        if self.use_bias:
            if self.unit_forget_bias:  # recurrent.py:116-124: forget-gate bias starts at one
                b = torch.cat([self.bias_initializer((H,)), torch.ones(H), self.bias_initializer((2 * H,))])
            else:
                b = self.bias_initializer((4 * H,))
            self.bias = nn.Parameter(b.to(device=dev, dtype=torch.float32))
        else:
            self.bias = None
        self.built = True

    # -- weights -------------------------------------------------------------------------------------------------
    def _sample(self, init, param, reg, stream):
        if not _is_trainable(init):
            return None, param
        mu, rho = init.params()
        if rho is None:
            return None, mu
        scale, pm, ps = ((reg.scale_factor, reg.mean, reg.stddev) if isinstance(reg, regularizers.NormalKLDivergence)
                         else (0.0, 0.0, 1.0))
        k = F.SequenceKernel(mu, rho, self._noise(f"eps_{stream}", stream), self.compute_dtype, scale, pm, ps)
        return k, k.w_lp

    def call_weights(self):
        """Draw a new sample of the kernel and the recurrent kernel; every following step reuses it (recurrent.py:149-161)."""
        if not self.built:
            raise ValueError("call_weights() needs a built cell: call it on an input first or use get_initial_state(inputs)")
        self._kw, self._w = self._sample(self.kernel_initializer, self.kernel, self.kernel_regularizer, S_LSTM_KERNEL)
        self._ku, self._u = self._sample(self.recurrent_initializer, self.recurrent_kernel, self.recurrent_regularizer,
                                         S_LSTM_RECURRENT)
        self._sample_calls = self._calls
        self.called_weights = True

    def get_initial_state(self, inputs=None, batch_size=None, dtype=None):
        """Zero states; side effect: a fresh weight sample for the sequence that starts here (recurrent.py:156-161)."""
        if self.built:
            self.call_weights()
        if inputs is not None:
            batch_size = inputs.shape[0]
        dev = self._device
        return [torch.zeros(batch_size, self.units, device=dev, dtype=self.compute_dtype),
                torch.zeros(batch_size, self.units, device=dev, dtype=torch.float32)]

    def call(self, inputs, states, training=None):
        if inputs.dim() != 2:
            raise ValueError("LSTM cells take [batch, input_dim] inputs")
        if not self.called_weights:
            self.call_weights()
        h_prev, c_prev = states
        tok = lambda k: k.token if k is not None else None  # noqa: E731
        h, c = F.LSTMStep.apply(inputs, h_prev, c_prev, self._w, self._u, self.bias, tok(self._kw), tok(self._ku), self._kw,
                                self._ku, self.compute_dtype)
        return h, [h, c]

    @property
    def losses(self):
        out = []
        for k, init, reg in ((self._kw, self.kernel_initializer, self.kernel_regularizer),
                             (self._ku, self.recurrent_initializer, self.recurrent_regularizer)):
            if reg is None or not _is_trainable(init):
                continue
            out.append(k.kl if (k is not None and isinstance(reg, regularizers.NormalKLDivergence)) else reg(init))
        if self.bias_regularizer is not None and self.bias is not None:
            out.append(self.bias_regularizer(self.bias))
        return out

    def get_config(self):
        cfg = super().get_config()
        cfg.update(units=self.units, activation="tanh", recurrent_activation="sigmoid", use_bias=self.use_bias,
                   unit_forget_bias=self.unit_forget_bias,
                   kernel_initializer=initializers.serialize(self.kernel_initializer),
                   recurrent_initializer=initializers.serialize(self.recurrent_initializer),
                   kernel_regularizer=regularizers.serialize(self.kernel_regularizer),
                   recurrent_regularizer=regularizers.serialize(self.recurrent_regularizer))
        return cfg


S_LSTM_SIGN_IN, S_LSTM_SIGN_OUT, S_LSTM_RSIGN_IN, S_LSTM_RSIGN_OUT = 10, 11, 12, 13
S_LSTM_ALPHA, S_LSTM_GAMMA, S_LSTM_RALPHA, S_LSTM_RGAMMA = 14, 15, 16, 17
_NO_CLIP = (-3.0e38, 3.0e38)  # the LSTM rank-1 cell does not clip its fast weights (recurrent.py:641-673)


class LSTMCellFlipout(LSTMCellReparameterization):
    """Bayesian LSTM cell estimated via Flipout (recurrent.py:185-358): with Delta = sigma∘eps of each kernel,
        z = x·mean_W + ((x∘S_in)·Delta_W)∘S_out + h·mean_U + ((h∘S_rin)·Delta_U)∘S_rout + b,
    the kernel noise AND the four sign tensors drawn once per sequence (`call_weights` / `_call_sign_flips`).  Each branch
    is the library's Flipout product (the dual-accumulator tcgen05 launches in bf16 for batches > 128, the two-GEMM call
    otherwise); the gates take the two pre-activation tensors (include/ed2b200.h: ed2_lstm_gates_fwd)."""

    def call_weights(self):
        if not self.built:
            raise ValueError("call_weights() needs a built cell: call it on an input first or use get_initial_state(inputs)")
        # one Philox key for the whole sequence: kernel noise and sign tensors are functions of (key, stream, index)
        from .base import GOLDEN, MASK64

        self._seq_key = (self.seed + self._calls * GOLDEN) & MASK64 if self.step_counter is None else self.seed
        self.called_weights = True

    def _rnd(self, name, stream, per_example=None):
        if name in self._injected and self._injected[name] is not None:
            return F.Randomness(self._injected[name])
        return F.Randomness(seed=self._seq_key, stream=stream, step=self.step_counter, rows=self._row_map(per_example))

    def _branch(self, x, init, bias, streams, names):
        from .. import ops

        mu, rho = init.params()
        B = x.shape[0]
        eps = self._rnd(names[0], streams[0])
        s_in, s_out = self._rnd(names[1], streams[1], B), self._rnd(names[2], streams[2], B)
        # kl_scale = 0: the KL of the two kernels is taken once per sequence by `.losses`, not once per step
        y, _ = F.FlipoutDense.apply(x, mu, rho, bias, ops.ACT_NONE, self.compute_dtype, eps, s_in, s_out, 0.0, 0.0, 1.0)
        return y

    def call(self, inputs, states, training=None):
        if not (_is_trainable(self.kernel_initializer) and _is_trainable(self.recurrent_initializer)
                and self.kernel_initializer.params()[1] is not None and self.recurrent_initializer.params()[1] is not None):
            return LSTMCellReparameterization.call(self, inputs, states, training)  # recurrent.py:275-277
        if inputs.dim() != 2:
            raise ValueError("LSTM cells take [batch, input_dim] inputs")
        if not self.called_weights:
            self.call_weights()
        h_prev, c_prev = states
        z_x = self._branch(inputs, self.kernel_initializer, self.bias, (S_LSTM_KERNEL, S_LSTM_SIGN_IN, S_LSTM_SIGN_OUT),
                           ("eps_kernel", "sign_input", "sign_output"))
        z_h = self._branch(h_prev, self.recurrent_initializer, None,
                           (S_LSTM_RECURRENT, S_LSTM_RSIGN_IN, S_LSTM_RSIGN_OUT),
                           ("eps_recurrent", "recurrent_sign_input", "recurrent_sign_output"))
        h, c = F.LSTMGates.apply(z_x, z_h, c_prev, self.compute_dtype)
        return h, [h, c]

    @property
    def losses(self):
        out = []
        for init, reg in ((self.kernel_initializer, self.kernel_regularizer),
                          (self.recurrent_initializer, self.recurrent_regularizer)):
            if reg is not None and _is_trainable(init):
                out.append(reg(init))
        if self.bias_regularizer is not None and self.bias is not None:
            out.append(self.bias_regularizer(self.bias))
        return out


class LSTMCellRank1(Layer):
    """A rank-1 Bayesian neural net LSTM cell (recurrent.py:360-830): member-major batch [ensemble_size * n, input_dim],
        z = ((x∘alpha)·W)∘gamma + ((h∘alpha_rec)·U)∘gamma_rec + bias[e]
    with per-example alpha / gamma / recurrent alpha / recurrent gamma drawn once per sequence (`_sample_weights`,
    :641-673; unclipped) and reused by every step.  Each branch is the library's rank-1 dense product."""

    def __init__(self, units, activation="tanh", recurrent_activation="sigmoid", use_bias=True,
                 alpha_initializer="trainable_normal", gamma_initializer="trainable_normal",
                 kernel_initializer="glorot_uniform", recurrent_alpha_initializer="trainable_normal",
                 recurrent_gamma_initializer="trainable_normal", recurrent_initializer="orthogonal",
                 bias_initializer="zeros", unit_forget_bias=True, alpha_regularizer="normal_kl_divergence",
                 gamma_regularizer="normal_kl_divergence", kernel_regularizer=None,
                 recurrent_alpha_regularizer="normal_kl_divergence", recurrent_gamma_regularizer="normal_kl_divergence",
                 recurrent_regularizer=None, bias_regularizer=None, alpha_constraint=None, gamma_constraint=None,
                 kernel_constraint=None, recurrent_alpha_constraint=None, recurrent_gamma_constraint=None,
                 recurrent_constraint=None, bias_constraint=None, dropout=0.0, recurrent_dropout=0.0, implementation=2,
                 use_additive_perturbation=False, ensemble_size=1, **kwargs):
        super().__init__(**kwargs)
        if activation != "tanh" or recurrent_activation != "sigmoid":
            raise NotImplementedError("the fused gate kernel implements activation='tanh', recurrent_activation='sigmoid'")
        if dropout or recurrent_dropout:
            raise NotImplementedError("dropout / recurrent_dropout are not built")
        if use_additive_perturbation:
            raise NotImplementedError("use_additive_perturbation (recurrent.py:743-746) is not built for the LSTM cell")
        self.units, self.ensemble_size = int(units), int(ensemble_size)
        self.use_bias, self.unit_forget_bias = use_bias, unit_forget_bias
        g = initializers.get
        self.alpha_initializer, self.gamma_initializer = g(alpha_initializer), g(gamma_initializer)
        self.recurrent_alpha_initializer = g(recurrent_alpha_initializer)
        self.recurrent_gamma_initializer = g(recurrent_gamma_initializer)
        self.kernel_initializer, self.recurrent_initializer = g(kernel_initializer), g(recurrent_initializer)
        self.bias_initializer = g(bias_initializer)
        r = regularizers.get
        self.alpha_regularizer, self.gamma_regularizer = r(alpha_regularizer), r(gamma_regularizer)
        self.recurrent_alpha_regularizer = r(recurrent_alpha_regularizer)
        self.recurrent_gamma_regularizer = r(recurrent_gamma_regularizer)
        self.kernel_regularizer, self.recurrent_regularizer = r(kernel_regularizer), r(recurrent_regularizer)
        self.bias_regularizer = r(bias_regularizer)
        self.state_size = [self.units, self.units]
        self.output_size = self.units
        self.sampled_weights = False

    def build(self, input_shape):
        dev = self._device
        I, H, E = input_shape[-1], self.units, self.ensemble_size
        self.alpha = _make_weight(self, "alpha", (E, I), self.alpha_initializer, dev)
        self.gamma = _make_weight(self, "gamma", (E, 4 * H), self.gamma_initializer, dev)
        self.kernel = nn.Parameter(self.kernel_initializer((I, 4 * H)).to(dev))
        self.recurrent_alpha = _make_weight(self, "recurrent_alpha", (E, H), self.recurrent_alpha_initializer, dev)
        self.recurrent_gamma = _make_weight(self, "recurrent_gamma", (E, 4 * H), self.recurrent_gamma_initializer, dev)
        self.recurrent_kernel = nn.Parameter(self.recurrent_initializer((H, 4 * H)).to(dev))
        if _is_trainable(self.bias_initializer):
            raise NotImplementedError("a trainable bias initializer is not built for the LSTM cell")
        if self.use_bias:
            if self.unit_forget_bias:  # recurrent.py:589-600
                b = torch.cat([self.bias_initializer((E, H)), torch.ones(E, H), self.bias_initializer((E, 2 * H))], dim=1)
            else:
                b = self.bias_initializer((E, 4 * H))
            self.bias = nn.Parameter(b.to(device=dev, dtype=torch.float32))
        else:
            self.bias = None
        self.built = True

    def _sample_weights(self, inputs=None, batch_size=None, dtype=None):
        """A new draw of the four fast-weight tensors for the sequence that starts here (recurrent.py:641-673)."""
        from .base import GOLDEN, MASK64

        self._seq_key = (self.seed + self._calls * GOLDEN) & MASK64 if self.step_counter is None else self.seed
        self.sampled_weights = True

    def get_initial_state(self, inputs=None, batch_size=None, dtype=None):
        if self.built:
            self._sample_weights(inputs=inputs, batch_size=batch_size)
        if inputs is not None:
            batch_size = inputs.shape[0]
        dev = self._device
        return [torch.zeros(batch_size, self.units, device=dev, dtype=self.compute_dtype),
                torch.zeros(batch_size, self.units, device=dev, dtype=torch.float32)]

    def _rnd(self, name, stream, per_example):
        if name in self._injected and self._injected[name] is not None:
            return F.Randomness(self._injected[name])
        return F.Randomness(seed=self._seq_key, stream=stream, step=self.step_counter, rows=self._row_map(per_example))

    def _fast(self, init, param):
        return init.params() if _is_trainable(init) else (param, None)

    def _branch(self, x, kernel, a_init, a_param, g_init, g_param, bias, streams, names):
        from .. import ops

        per = x.shape[0] // self.ensemble_size
        mu_r, rho_r = self._fast(a_init, a_param)
        mu_s, rho_s = self._fast(g_init, g_param)
        return F.Rank1Dense.apply(x, kernel, mu_r, rho_r, mu_s, rho_s, bias, ops.ACT_NONE, self.ensemble_size,
                                  self.compute_dtype, self._rnd(names[0], streams[0], per), self._rnd(names[1], streams[1], per),
                                  False, _NO_CLIP)

    def call(self, inputs, states, training=None):
        if inputs.dim() != 2:
            raise ValueError("LSTMCellRank1 takes [ensemble_size * examples_per_model, input_dim]")
        if inputs.shape[0] % self.ensemble_size != 0:
            raise ValueError(f"batch size {inputs.shape[0]} is not divisible by ensemble_size {self.ensemble_size}")
        if not self.sampled_weights:
            self._sample_weights(inputs=inputs)
        h_prev, c_prev = states
        z_x = self._branch(inputs, self.kernel, self.alpha_initializer, self.alpha, self.gamma_initializer, self.gamma,
                           self.bias, (S_LSTM_ALPHA, S_LSTM_GAMMA), ("eps_alpha", "eps_gamma"))
        z_h = self._branch(h_prev, self.recurrent_kernel, self.recurrent_alpha_initializer, self.recurrent_alpha,
                           self.recurrent_gamma_initializer, self.recurrent_gamma, None, (S_LSTM_RALPHA, S_LSTM_RGAMMA),
                           ("eps_recurrent_alpha", "eps_recurrent_gamma"))
        h, c = F.LSTMGates.apply(z_x, z_h, c_prev, self.compute_dtype)
        return h, [h, c]

    @property
    def losses(self):
        out = []
        for init, reg in ((self.alpha_initializer, self.alpha_regularizer), (self.gamma_initializer, self.gamma_regularizer),
                          (self.recurrent_alpha_initializer, self.recurrent_alpha_regularizer),
                          (self.recurrent_gamma_initializer, self.recurrent_gamma_regularizer)):
            if reg is not None and _is_trainable(init):
                out.append(reg(init))
        if self.kernel_regularizer is not None:
            out.append(self.kernel_regularizer(self.kernel))
        if self.recurrent_regularizer is not None:
            out.append(self.recurrent_regularizer(self.recurrent_kernel))
        if self.bias_regularizer is not None and self.bias is not None:
            out.append(self.bias_regularizer(self.bias))
        return out

    def get_config(self):
        cfg = super().get_config()
        cfg.update(units=self.units, ensemble_size=self.ensemble_size, use_bias=self.use_bias,
                   unit_forget_bias=self.unit_forget_bias, use_additive_perturbation=False)
        return cfg
