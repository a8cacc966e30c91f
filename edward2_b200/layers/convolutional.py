"""`ed.layers` Conv2D stochastic-weight layers on libed2b200 (edward2/tensorflow/layers/convolutional.py:
Conv2DReparameterization :32, Conv2DFlipout :167, Conv2DBatchEnsemble :560, Conv2DRank1 :2056).

Every variant runs on the patch matrix cols [B*Ho*Wo, kh*kw*Cin] (include/ed2b200.h: ed2_im2col / ed2_col2im) through
the dense kernels of the same family: the per-example factors of the conv layers are channel vectors broadcast over the
spatial axes, i.e. a dense BatchEnsemble-style product with rows-per-member = Ho*Wo and one "member" per example.
NHWC ("channels_last") inputs, dilation 1, padding 'valid' | 'same' — other settings raise."""
from __future__ import annotations

import torch
from torch import nn

from .. import functional as F, initializers, ops, regularizers
from .base import Layer
from .dense import (S_B, S_KERNEL, S_R, S_S, S_SIGN_IN, S_SIGN_OUT, _FusedKL, _is_trainable,  # noqa: F401
                    _make_weight)


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


class _Conv2DBase(Layer):
    def _init_conv(self, filters, kernel_size, strides, padding, data_format, dilation_rate, activation):
        self.filters = int(filters)
        self.kernel_size = _pair(kernel_size)
        self.strides = _pair(strides)
        self.padding = str(padding).lower()
        if self.padding not in ("valid", "same"):
            raise ValueError(f"padding must be 'valid' or 'same', got {padding!r}")
        if data_format not in (None, "channels_last"):
            raise NotImplementedError("only data_format='channels_last' (NHWC) is built")
        if _pair(dilation_rate) != (1, 1):
            raise NotImplementedError("dilation_rate != 1 is not built")
        self.data_format = "channels_last"
        self.activation = activation
        self._act = ops.act_id(activation)

    def _check(self, inputs):
        if inputs.dim() != 4:
            raise ValueError("Conv2D layers take [batch, height, width, channels] inputs")

    def _patches(self, x):
        cols = F.Im2Col.apply(x, self.kernel_size, self.strides, self.padding, self.compute_dtype)
        g = F.conv_geometry(x.shape, self.kernel_size, self.strides, self.padding)
        ho, wo = F.conv_out_size(g)
        return cols, ho, wo

    def _out_hw(self, x):
        return F.conv_out_size(F.conv_geometry(x.shape, self.kernel_size, self.strides, self.padding))

    def _conv_config(self):
        return dict(filters=self.filters, kernel_size=self.kernel_size, strides=self.strides, padding=self.padding,
                    data_format=self.data_format, dilation_rate=(1, 1), activation=self.activation)


class Conv2DReparameterization(_Conv2DBase):
    """2D convolution with a Normal kernel posterior estimated via reparameterization (convolutional.py:32-98): one kernel
    sample per call, y = act(conv(x, mu + sigma∘eps) + b); `.losses` holds the KL regularizer."""

    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", data_format=None, dilation_rate=(1, 1),
                 activation=None, use_bias=True, kernel_initializer="trainable_normal", bias_initializer="zeros",
                 kernel_regularizer="normal_kl_divergence", bias_regularizer=None, activity_regularizer=None,
                 kernel_constraint=None, bias_constraint=None, **kwargs):
        super().__init__(**kwargs)
        self._init_conv(filters, kernel_size, strides, padding, data_format, dilation_rate, activation)
        self.use_bias = use_bias
        self.kernel_initializer = initializers.get(kernel_initializer)
        self.bias_initializer = initializers.get(bias_initializer)
        self.kernel_regularizer = regularizers.get(kernel_regularizer)
        self.bias_regularizer = regularizers.get(bias_regularizer)
        self._kl = _FusedKL()

    def build(self, input_shape):
        dev = self._device
        kh, kw = self.kernel_size
        self.kernel = _make_weight(self, "kernel", (kh, kw, input_shape[-1], self.filters), self.kernel_initializer, dev)
        self.bias = _make_weight(self, "bias", (self.filters,), self.bias_initializer, dev) if self.use_bias else None
        if _is_trainable(self.bias_initializer):
            raise NotImplementedError("a trainable bias initializer is not built for the Conv2D layers")
        self.built = True

    def _kernel_params(self):
        if _is_trainable(self.kernel_initializer):
            return self.kernel_initializer.params()
        return self.kernel, None

    def _fused_kl_args(self):
        r = self.kernel_regularizer
        if isinstance(r, regularizers.NormalKLDivergence):
            return r.scale_factor, r.mean, r.stddev
        return 0.0, 0.0, 1.0

    def _finish(self, y2d, x, ho, wo):
        return y2d.reshape(x.shape[0], ho, wo, self.filters)

    def call(self, inputs, training=None):
        self._check(inputs)
        mean, rho = self._kernel_params()
        cols, ho, wo = self._patches(inputs)
        K = cols.shape[1]
        if rho is None:
            y = F.Dense.apply(cols, mean.reshape(K, self.filters), self.bias, self._act, self.compute_dtype, None)
            return self._finish(y, inputs, ho, wo)
        kl_scale, pm, ps = self._fused_kl_args()
        y, kl = F.ReparamDense.apply(cols, mean.reshape(K, self.filters), rho.reshape(K, self.filters), self.bias,
                                     self._act, self.compute_dtype, self._noise("eps", S_KERNEL), kl_scale, pm, ps,
                                     None, None, None)
        if kl_scale != 0.0:
            self._kl.set(kl, mean, rho)
        return self._finish(y, inputs, ho, wo)

    @property
    def losses(self):
        out = []
        if self.kernel_regularizer is not None and _is_trainable(self.kernel_initializer):
            mean, rho = self._kernel_params()
            fused = self._kl.get(mean, rho) if rho is not None else None
            out.append(fused if fused is not None else self.kernel_regularizer(self.kernel_initializer))
        if self.bias_regularizer is not None and self.use_bias:
            out.append(self.bias_regularizer(self.bias))
        return out

    def get_config(self):
        cfg = super().get_config()
        cfg.update(self._conv_config())
        cfg.update(use_bias=self.use_bias, kernel_initializer=initializers.serialize(self.kernel_initializer),
                   bias_initializer=initializers.serialize(self.bias_initializer),
                   kernel_regularizer=regularizers.serialize(self.kernel_regularizer),
                   bias_regularizer=regularizers.serialize(self.bias_regularizer))
        return cfg


class Conv2DFlipout(Conv2DReparameterization):
    """2D convolution estimated via Flipout (convolutional.py:167-249): y = act(conv(x, mean) + conv(x∘S, Delta)∘T + b)
    with per-example sign vectors S [B,1,1,Cin], T [B,1,1,Cout] (±1) and Delta = sigma∘eps."""

    def call(self, inputs, training=None):
        mean, rho = self._kernel_params()
        if rho is None:
            return super().call(inputs)
        self._check(inputs)
        B, cin = inputs.shape[0], inputs.shape[-1]
        kh, kw = self.kernel_size
        cols, ho, wo = self._patches(inputs)
        K = cols.shape[1]
        s_in = self._sign_table("sign_input", S_SIGN_IN, B, cin)
        s_out = self._sign_table("sign_output", S_SIGN_OUT, B, self.filters)
        s_tab = F.ExpandTable.apply(s_in, B, 1, kh * kw)
        kl_scale, pm, ps = self._fused_kl_args()
        y, kl = F.FlipoutPatches.apply(cols, mean.reshape(K, self.filters), rho.reshape(K, self.filters), self.bias,
                                       s_tab, s_out, ho * wo, self._act, self.compute_dtype,
                                       self._noise("eps", S_KERNEL), kl_scale, pm, ps)
        if kl_scale != 0.0:
            self._kl.set(kl, mean, rho)
        return self._finish(y, inputs, ho, wo)

    def _sign_table(self, name, stream, batch, channels):
        """[batch, channels] fp32 ±1 vectors of this call (element (b, c) = element global_row(b) * channels + c of the
        layer's sign stream), or the injected tensor."""
        from .. import _lib

        nz = self._noise(name, stream, per_example=batch)
        if nz.injected is not None:
            return nz.injected.reshape(batch, channels)
        out = torch.empty(batch, channels, dtype=torch.float32, device=self._device)
        _lib.check(_lib.load().ed2_noise_table(1, nz.c(), out.data_ptr(), _lib.F32, channels, batch, channels,
                                               _lib.stream_ptr()), "ed2_noise_table")
        return out


class Conv2DBatchEnsemble(_Conv2DBase):
    """A batch ensemble convolutional layer (convolutional.py:560-733, rank 1): member-major batch [E*n, H, W, C],
    y[e] = act(conv(x[e]∘alpha[e], W)∘gamma[e] + bias[e])."""

    def __init__(self, filters, kernel_size, rank=1, ensemble_size=4, alpha_initializer="ones", gamma_initializer="ones",
                 strides=(1, 1), padding="valid", data_format="channels_last", dilation_rate=(1, 1), activation=None,
                 use_bias=True, kernel_initializer="glorot_uniform", bias_initializer="zeros", kernel_regularizer=None,
                 bias_regularizer=None, activity_regularizer=None, kernel_constraint=None, bias_constraint=None, **kwargs):
        super().__init__(**kwargs)
        if rank != 1:
            raise NotImplementedError("rank > 1 fast weights (convolutional.py:651-679) are not built")
        self._init_conv(filters, kernel_size, strides, padding, data_format, dilation_rate, activation)
        self.rank, self.ensemble_size = rank, int(ensemble_size)
        self.use_bias = use_bias
        self.alpha_initializer = initializers.get(alpha_initializer)
        self.gamma_initializer = initializers.get(gamma_initializer)
        self.kernel_initializer = initializers.get(kernel_initializer)
        self.bias_initializer = initializers.get(bias_initializer)
        self.kernel_regularizer = regularizers.get(kernel_regularizer)
        self.bias_regularizer = regularizers.get(bias_regularizer)

    def build(self, input_shape):
        dev = self._device
        kh, kw = self.kernel_size
        cin, E = input_shape[-1], self.ensemble_size
        self.kernel = nn.Parameter(self.kernel_initializer((kh, kw, cin, self.filters)).to(dev))
        self.alpha = nn.Parameter(self.alpha_initializer((E, cin)).to(dev))
        self.gamma = nn.Parameter(self.gamma_initializer((E, self.filters)).to(dev))
        self.ensemble_bias = nn.Parameter(self.bias_initializer((E, self.filters)).to(dev)) if self.use_bias else None
        self.built = True

    def call(self, inputs, training=None):
        self._check(inputs)
        B, E = inputs.shape[0], self.ensemble_size
        if B % E != 0:
            raise ValueError(f"batch size {B} is not divisible by ensemble_size {E}")
        n = B // E
        kh, kw = self.kernel_size
        a_tab = F.ExpandTable.apply(self.alpha, B, n, 1)
        g_tab = F.ExpandTable.apply(self.gamma, B, n, 1)
        b_tab = F.ExpandTable.apply(self.ensemble_bias, B, n, 1) if self.ensemble_bias is not None else None
        y = F.ConvFactorized.apply(inputs, self.kernel.reshape(kh * kw * inputs.shape[-1], self.filters), a_tab, g_tab,
                                   b_tab, self.kernel_size, self.strides, self.padding, self._act, self.compute_dtype)
        return y.reshape(B, *self._out_hw(inputs), self.filters)

    @property
    def losses(self):
        out = []
        if self.kernel_regularizer is not None:
            out.append(self.kernel_regularizer(self.kernel))
        if self.bias_regularizer is not None and self.ensemble_bias is not None:
            out.append(self.bias_regularizer(self.ensemble_bias))
        return out

    def get_config(self):
        cfg = super().get_config()
        cfg.update(self._conv_config())
        cfg.update(ensemble_size=self.ensemble_size, rank=self.rank, use_bias=self.use_bias,
                   alpha_initializer=initializers.serialize(self.alpha_initializer),
                   gamma_initializer=initializers.serialize(self.gamma_initializer))
        return cfg


class Conv2DRank1(_Conv2DBase):
    """A rank-1 Bayesian neural net 2D convolution (convolutional.py:2056-2265): per-example fast weights
    r = clip(mu_r[e] + sigma_r[e]∘eps_r) [B, Cin], s likewise [B, Cout], broadcast over the spatial axes;
    y = act(conv(x∘r, W)∘s + bias[e])  (multiplicative) or act(conv(x + r, W) + s + bias[e]) is not built."""

    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", data_format="channels_last",
                 dilation_rate=(1, 1), activation=None, use_bias=True, alpha_initializer="trainable_normal",
                 gamma_initializer="trainable_normal", kernel_initializer="glorot_uniform", bias_initializer="zeros",
                 alpha_regularizer="normal_kl_divergence", gamma_regularizer="normal_kl_divergence", kernel_regularizer=None,
                 bias_regularizer=None, activity_regularizer=None, alpha_constraint=None, gamma_constraint=None,
                 kernel_constraint=None, bias_constraint=None, use_additive_perturbation=False,
                 min_perturbation_value=-10, max_perturbation_value=10, ensemble_size=1, **kwargs):
        super().__init__(**kwargs)
        self._init_conv(filters, kernel_size, strides, padding, data_format, dilation_rate, activation)
        if use_additive_perturbation:
            raise NotImplementedError("additive perturbations (convolutional.py:2003-2005) are not built for Conv2DRank1")
        if not float(min_perturbation_value) < float(max_perturbation_value):
            raise ValueError("min_perturbation_value must be below max_perturbation_value")
        self.min_perturbation_value, self.max_perturbation_value = min_perturbation_value, max_perturbation_value
        self.ensemble_size = int(ensemble_size)
        self.use_bias = use_bias
        self.alpha_initializer = initializers.get(alpha_initializer)
        self.gamma_initializer = initializers.get(gamma_initializer)
        self.kernel_initializer = initializers.get(kernel_initializer)
        self.bias_initializer = initializers.get(bias_initializer)
        self.alpha_regularizer = regularizers.get(alpha_regularizer)
        self.gamma_regularizer = regularizers.get(gamma_regularizer)
        self.kernel_regularizer = regularizers.get(kernel_regularizer)
        self.bias_regularizer = regularizers.get(bias_regularizer)

    def build(self, input_shape):
        dev = self._device
        kh, kw = self.kernel_size
        cin, E = input_shape[-1], self.ensemble_size
        self.kernel = nn.Parameter(self.kernel_initializer((kh, kw, cin, self.filters)).to(dev))
        self.alpha = _make_weight(self, "alpha", (E, cin), self.alpha_initializer, dev)
        self.gamma = _make_weight(self, "gamma", (E, self.filters), self.gamma_initializer, dev)
        if _is_trainable(self.bias_initializer):
            raise NotImplementedError("a trainable bias initializer is not built for Conv2DRank1")
        self.ensemble_bias = nn.Parameter(self.bias_initializer((E, self.filters)).to(dev)) if self.use_bias else None
        self.built = True

    def _fast(self, init, param):
        return init.params() if _is_trainable(init) else (param, None)

    def call(self, inputs, training=None):
        self._check(inputs)
        B, E = inputs.shape[0], self.ensemble_size
        if B % E != 0:
            raise ValueError(f"batch size {B} is not divisible by ensemble_size {E}")
        n = B // E
        kh, kw = self.kernel_size
        clip = (float(self.min_perturbation_value), float(self.max_perturbation_value))
        mu_r, rho_r = self._fast(self.alpha_initializer, self.alpha)
        mu_s, rho_s = self._fast(self.gamma_initializer, self.gamma)
        r = F.Rank1Factors.apply(mu_r, rho_r, self._noise("eps_alpha", S_R, per_example=n), B, E, clip)
        s = F.Rank1Factors.apply(mu_s, rho_s, self._noise("eps_gamma", S_S, per_example=n), B, E, clip)
        b_tab = F.ExpandTable.apply(self.ensemble_bias, B, n, 1) if self.ensemble_bias is not None else None
        y = F.ConvFactorized.apply(inputs, self.kernel.reshape(kh * kw * inputs.shape[-1], self.filters), r, s, b_tab,
                                   self.kernel_size, self.strides, self.padding, self._act, self.compute_dtype)
        return y.reshape(B, *self._out_hw(inputs), self.filters)

    @property
    def losses(self):
        out = []
        if self.alpha_regularizer is not None and _is_trainable(self.alpha_initializer):
            out.append(self.alpha_regularizer(self.alpha_initializer))
        if self.gamma_regularizer is not None and _is_trainable(self.gamma_initializer):
            out.append(self.gamma_regularizer(self.gamma_initializer))
        if self.kernel_regularizer is not None:
            out.append(self.kernel_regularizer(self.kernel))
        if self.bias_regularizer is not None and self.ensemble_bias is not None:
            out.append(self.bias_regularizer(self.ensemble_bias))
        return out

    def get_config(self):
        cfg = super().get_config()
        cfg.update(self._conv_config())
        cfg.update(ensemble_size=self.ensemble_size, use_bias=self.use_bias,
                   min_perturbation_value=self.min_perturbation_value, max_perturbation_value=self.max_perturbation_value,
                   alpha_initializer=initializers.serialize(self.alpha_initializer),
                   gamma_initializer=initializers.serialize(self.gamma_initializer))
        return cfg
