"""`edward2.jax.nn.MCSoftmaxDenseFA` on libed2b200 (edward2/jax/nn/heteroscedastic_lib.py:43-336): softmax and
factor-analysis approximation to heteroscedastic predictions, Monte-Carlo estimated.

Same fields, variable names (`loc_layer`, `scale_layer`, `diag_layer` under `params`) and rng collections
(`diag_noise_samples`, `standard_norm_noise_samples`) as the reference module.  The three parameter layers are Dense
layers on the tcgen05 GEMM; the MC estimate is the library's kernel (include/ed2b200.h: ed2_mc_softmax_fa_fwd / _bwd) in its
JAX form — ONE diagonal noise draw per (example, sample) shared by the classes (:186-188), where the TF class draws one
per class.  Not built (raise): parameter_efficient, tune_temperature, latent_dim, num_factors == 0."""
from __future__ import annotations

from typing import Optional

import torch

from .. import functional as F
from . import module as m
from .normalization import Dense

S_MC = 7  # Philox stream of the Monte-Carlo logit noise (same as layers.heteroscedastic)


class MCSoftmaxDenseFA(m.Module):
    num_classes: int
    num_factors: int
    temperature: float = 1.0
    parameter_efficient: bool = False
    train_mc_samples: int = 1000
    test_mc_samples: int = 1000
    share_samples_across_batch: bool = False
    logits_only: bool = False
    return_locs: bool = False
    eps: float = 1e-7
    tune_temperature: bool = False
    temperature_lower_bound: Optional[float] = None
    temperature_upper_bound: Optional[float] = None
    latent_dim: Optional[int] = None
    cov_layer_kernel_init_scale: Optional[float] = None

    def _cov_init(self):
        if self.cov_layer_kernel_init_scale is None:
            return m.lecun_normal()
        if self.cov_layer_kernel_init_scale <= 0:  # heteroscedastic_lib.py:93-97
            raise ValueError("scale_layer_kernel_init_factor must be positive;"
                             f" received instead {self.cov_layer_kernel_init_scale}.")
        return m.variance_scaling(self.cov_layer_kernel_init_scale, "fan_in", "truncated_normal")

    def _noise(self, inputs, injected=None):
        if injected is not None:
            return F.Randomness(injected)
        # the reference draws from two named jax streams; here both halves of a sample come from one Philox stream whose
        # seed folds the two names (the streams themselves are not part of the reference's contract)
        a, b = self.make_rng("diag_noise_samples"), self.make_rng("standard_norm_noise_samples")
        return F.Randomness(seed=(a.seed * 0x9E3779B97F4A7C15 + b.seed) & 0xFFFFFFFFFFFFFFFF, stream=S_MC)

    def __call__(self, inputs, training=True, mc_noise=None):
        """inputs [batch, features] -> logits if logits_only else (logits, log_probs, probs).  `mc_noise` (tests): injected
        noise [batch or 1, samples, round8(num_factors) + round8(num_classes)] fp32 — factor draws first, the diagonal
        draw at element round8(num_factors)."""
        if self.parameter_efficient or self.tune_temperature or self.latent_dim is not None or self.num_factors <= 0:
            raise NotImplementedError("parameter_efficient / tune_temperature / latent_dim / num_factors == 0 are not built")
        if self.num_classes <= 2:
            raise NotImplementedError("num_classes > 2")
        K, Fn = self.num_classes, self.num_factors
        cov_init = self._cov_init()
        loc = self.sub(Dense(K), "loc_layer")(inputs)
        fac = self.sub(Dense(K * Fn, kernel_init=cov_init), "scale_layer")(inputs)
        diag_raw = self.sub(Dense(K, kernel_init=cov_init), "diag_layer")(inputs)
        samples = self.train_mc_samples if training else self.test_mc_samples
        share = (1 if self.share_samples_across_batch else 0) | 2  # bit 1: one diagonal draw per sample (JAX form)
        logits, probs, _ = F.MCSoftmaxFA.apply(loc.float(), fac.float(), diag_raw.float(), None, 1.0, 1.0,
                                               self._noise(inputs, mc_noise), samples, Fn, share, float(self.temperature),
                                               float(self.eps), False)
        log_probs = logits
        if self.return_locs:
            logits = loc
        if self.logits_only:
            return logits
        # jnp.clip(probs_mean, min=eps) (:317): no upper clip
        return logits, log_probs, torch.clamp(probs, min=self.eps)


class MCSigmoidDenseFA(m.Module):
    """`edward2.jax.nn.MCSigmoidDenseFA` (heteroscedastic_lib.py:338-594): the multi-label form — independent sigmoid
    outputs, logits = inverse sigmoid of the Monte-Carlo mean probabilities."""
    num_outputs: int
    num_factors: int
    temperature: float = 1.0
    parameter_efficient: bool = False
    train_mc_samples: int = 1000
    test_mc_samples: int = 1000
    share_samples_across_batch: bool = False
    logits_only: bool = False
    return_locs: bool = False
    eps: float = 1e-7
    tune_temperature: bool = False
    temperature_lower_bound: Optional[float] = None
    temperature_upper_bound: Optional[float] = None
    latent_dim: Optional[int] = None

    _noise = MCSoftmaxDenseFA._noise

    def __call__(self, inputs, training=True, mc_noise=None):
        if self.parameter_efficient or self.tune_temperature or self.latent_dim is not None or self.num_factors <= 0:
            raise NotImplementedError("parameter_efficient / tune_temperature / latent_dim / num_factors == 0 are not built")
        K, Fn = self.num_outputs, self.num_factors
        loc = self.sub(Dense(K), "loc_layer")(inputs)
        fac = self.sub(Dense(K * Fn), "scale_layer")(inputs)
        diag_raw = self.sub(Dense(K), "diag_layer")(inputs)
        samples = self.train_mc_samples if training else self.test_mc_samples
        flags = (1 if self.share_samples_across_batch else 0) | 2 | 4  # JAX diagonal noise, sigmoid link
        logits, probs, _ = F.MCSoftmaxFA.apply(loc.float(), fac.float(), diag_raw.float(), None, 1.0, 1.0,
                                               self._noise(inputs, mc_noise), samples, Fn, flags, float(self.temperature),
                                               float(self.eps), False)
        probs_lo = torch.clamp(probs, min=self.eps)                      # :575
        log_probs = torch.log(probs_lo)
        if self.return_locs:
            logits = loc
        if self.logits_only:
            return logits
        return logits, log_probs, torch.clamp(probs_lo, self.eps, 1.0 - self.eps)  # :578: the returned probs carry both clips
