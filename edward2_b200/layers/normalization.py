"""`ed.layers.SpectralNormalization` on libed2b200 (edward2/tensorflow/layers/normalization.py:284-395).

TF semantics: every call runs `iteration` power-iteration steps (when training), computes sigma = v W uᵀ and
assigns W * (c / sigma) INTO the wrapped layer's kernel when c / sigma < 1 — `self.w` *is* `layer.kernel`
(:333), so `restore_weights` (:393-395) is a no-op and the normalisation is persistent.  The matvecs are the
coalesced bandwidth-bound kernels of csrc/sngp.cu.
"""
from __future__ import annotations

import torch

from .. import ops
from .base import Layer


class SpectralNormalization(Layer):
    """Implements spectral normalization for Dense layer."""

    def __init__(self, layer, iteration=1, norm_multiplier=0.95, training=True, aggregation=None,
                 inhere_layer_name=False, **kwargs):
        if not isinstance(layer, Layer):
            raise ValueError("`layer` must be a `tf.keras.layer.Layer`. Observed `{}`".format(layer))
        name = kwargs.pop("name", None)
        if inhere_layer_name:
            name = layer.layer_name
        super().__init__(name=name, dtype=layer.compute_dtype, **kwargs)
        self.layer = layer
        self.iteration = iteration
        self.do_power_iteration = training
        self.norm_multiplier = norm_multiplier

    def build(self, input_shape):
        if not self.layer.built:
            self.layer._device = self._device
            self.layer.build(input_shape)
        self.w = self.layer.kernel
        self.w_shape = list(self.w.shape)
        in_dim = 1
        for s in self.w_shape[:-1]:
            in_dim *= s
        gen = torch.Generator(device="cpu").manual_seed(self.seed % (2**31))
        # tf.initializers.random_normal() default stddev 0.05
        self.register_buffer("v", (torch.randn(1, in_dim, generator=gen) * 0.05).to(self._device))
        self.register_buffer("u", (torch.randn(1, self.w_shape[-1], generator=gen) * 0.05).to(self._device))
        self.built = True
        self.update_weights()

    def call(self, inputs, training=None):
        training = self.do_power_iteration if training is None else training
        self.update_weights(training=training)
        output = self.layer(inputs)
        self.restore_weights()
        return output

    def update_weights(self, training=True):
        with torch.no_grad():
            w2d = self.w.data.reshape(-1, self.w_shape[-1])
            sigma, _ = ops.spectral_norm_update(w2d, self.u.reshape(-1), self.v.reshape(-1), self.iteration,
                                                self.norm_multiplier, 0, bool(training), w2d)
        self.sigma = sigma
        return self.u, self.v, self.w

    def restore_weights(self):
        """normalization.py:393-395: assigns self.w back — self.w is layer.kernel, so nothing changes."""
        return self.w

    @property
    def losses(self):
        return self.layer.losses

    def get_config(self):
        cfg = super().get_config()
        cfg.update(iteration=self.iteration, norm_multiplier=self.norm_multiplier, training=self.do_power_iteration)
        return cfg
