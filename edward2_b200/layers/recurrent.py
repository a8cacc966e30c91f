"""Bayesian LSTM cell on libed2b200: LSTMCellReparameterization (edward2/tensorflow/layers/recurrent.py:30-183).

The kernel and the recurrent kernel have Normal posteriors whose noise is drawn ONCE per sequence (`call_weights()` /
`get_initial_state()`, recurrent.py:138-161) and reused by every time step; a step is two tcgen05 GEMMs (the second
accumulates onto the first) and one pointwise gate kernel, its backward one pointwise kernel and four GEMMs, and the raw
weight gradients of all steps are summed on the tensor cores before ONE finishing pass per kernel regenerates the noise
and adds the KL gradient (functional.SequenceKernel).  `LSTMCellFlipout` (:185) and `LSTMCellRank1` (:360) are not built."""
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
