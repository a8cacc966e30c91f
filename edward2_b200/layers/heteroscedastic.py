"""Heteroscedastic output heads on libed2b200: MCSoftmaxDenseFA (edward2/tensorflow/layers/heteroscedastic.py:508) and
HeteroscedasticSNGPLayer (edward2/tensorflow/layers/hetsngp.py:25) — the direct consumers of the random-feature GP head.

The three parameter layers (loc, factor loadings, diagonal scale) are Dense layers on the tcgen05 GEMM; the Monte-Carlo
estimate of the softmax predictive (1000 samples per example by default) is ONE kernel per direction that draws its
noise from Philox per (example, sample) and never materialises a [batch, samples, classes] tensor
(include/ed2b200.h: ed2_mc_softmax_fa_fwd / _bwd).  num_classes > 2 (the binary / sigmoid branch and the
parameter-efficient variant raise)."""
from __future__ import annotations

import torch

from .. import functional as F
from .base import GOLDEN, MASK64, Layer
from .dense import Dense
from .random_feature import RandomFeatureGaussianProcess

S_MC = 7  # Philox stream of the Monte-Carlo logit noise


class MCSoftmaxDenseFA(Layer):
    """Softmax and factor-analysis approximation to heteroscedastic predictions (heteroscedastic.py:508-806).

    __call__(inputs, training=True, seed=None) -> logits, or (logits, log_probs, probs, predictive_variance) when
    logits_only=False."""

    def __init__(self, num_classes, num_factors, temperature=1.0, parameter_efficient=False, train_mc_samples=1000,
                 test_mc_samples=1000, compute_pred_variance=False, share_samples_across_batch=False, logits_only=False,
                 eps=1e-7, dtype=None, kernel_regularizer=None, bias_regularizer=None, name="MCSoftmaxDenseFA", **kwargs):
        super().__init__(dtype=dtype, name=name, **kwargs)
        if num_classes <= 2:
            raise NotImplementedError("the binary (sigmoid) branch of the Monte-Carlo heads is not built: num_classes > 2")
        if parameter_efficient:
            raise NotImplementedError("parameter_efficient=True (heteroscedastic.py:621-630) is not built")
        self._num_classes, self._num_factors = int(num_classes), int(num_factors)
        self._temperature = float(temperature)
        self._train_mc_samples, self._test_mc_samples = int(train_mc_samples), int(test_mc_samples)
        self._compute_pred_variance = bool(compute_pred_variance)
        self._share_samples_across_batch = bool(share_samples_across_batch)
        self._logits_only = bool(logits_only)
        self._eps = float(eps)
        mk = lambda units, nm: Dense(units, kernel_regularizer=kernel_regularizer, bias_regularizer=bias_regularizer,  # noqa: E731
                                     dtype=dtype, name=f"{name}_{nm}")
        self._scale_layer = mk(self._num_classes * self._num_factors, "scale_layer")
        self._loc_layer = mk(self._num_classes, "loc_layer")
        self._diag_layer = mk(self._num_classes, "diag_layer")  # softplus + MIN_SCALE are applied inside the MC kernel

    def build(self, input_shape):
        self.built = True

    # -- the reference's hooks ---------------------------------------------------------------------------------------
    def _compute_loc_param(self, inputs, training=None):
        return self._loc_layer(inputs), None

    def _scale_params(self, inputs):
        """(factor loadings [B, K*F], PRE-softplus diagonal [B, K])."""
        return self._scale_layer(inputs), self._diag_layer(inputs)

    def _variance_mixing(self, covmat, training):
        return None, 1.0, 1.0

    def _mc_noise(self, batch, seed):
        if "mc_noise" in self._injected and self._injected["mc_noise"] is not None:
            return F.Randomness(self._injected["mc_noise"])
        if seed is not None:
            return F.Randomness(seed=int(seed) & MASK64, stream=S_MC, rows=self._row_map(batch))
        if self.step_counter is not None:
            return F.Randomness(seed=self.seed, stream=S_MC, step=self.step_counter, rows=self._row_map(batch))
        return F.Randomness(seed=(self.seed + self._calls * GOLDEN) & MASK64, stream=S_MC, rows=self._row_map(batch))

    def call(self, inputs, training=True, seed=None):
        if inputs.dim() != 2:
            raise ValueError("the Monte-Carlo heads take [batch, features] inputs")
        locs, covmat = self._compute_loc_param(inputs, training)
        fac, diag_raw = self._scale_params(inputs)
        sngp_var, w_het, w_sngp = self._variance_mixing(covmat, training)
        samples = self._train_mc_samples if training else self._test_mc_samples
        logits, probs, var = F.MCSoftmaxFA.apply(locs.float(), fac.float(), diag_raw.float(), sngp_var, w_het, w_sngp,
                                                 self._mc_noise(inputs.shape[0], seed), samples, self._num_factors,
                                                 self._share_samples_across_batch, self._temperature, self._eps,
                                                 self._compute_pred_variance)
        if self._logits_only:
            return logits
        # num_classes > 2: logits = log_probs (hetsngp.py:224-232)
        return logits, logits, torch.clamp(probs, self._eps, 1.0), (var if self._compute_pred_variance else None)

    @property
    def losses(self):
        return self._scale_layer.losses + self._loc_layer.losses + self._diag_layer.losses

    def get_config(self):
        cfg = super().get_config()
        cfg.update(num_classes=self._num_classes, num_factors=self._num_factors, temperature=self._temperature,
                   train_mc_samples=self._train_mc_samples, test_mc_samples=self._test_mc_samples,
                   compute_pred_variance=self._compute_pred_variance,
                   share_samples_across_batch=self._share_samples_across_batch, logits_only=self._logits_only,
                   eps=self._eps, parameter_efficient=False)
        return cfg


class MCSigmoidDenseFA(MCSoftmaxDenseFA):
    """Sigmoid and factor-analysis approximation to heteroscedastic (multi-label) predictions
    (edward2/tensorflow/layers/heteroscedastic.py:1362-1680): independent sigmoid outputs, logits = inverse sigmoid of the
    Monte-Carlo mean probabilities.  __call__ -> logits, or (logits, log_probs, probs, predictive_variance)."""

    def __init__(self, num_outputs, num_factors=0, temperature=1.0, parameter_efficient=False, train_mc_samples=1000,
                 test_mc_samples=1000, compute_pred_variance=False, share_samples_across_batch=False, logits_only=False,
                 eps=1e-7, dtype=None, kernel_regularizer=None, bias_regularizer=None, return_unaveraged_logits=False,
                 tune_temperature=False, temperature_lower_bound=None, temperature_upper_bound=None,
                 name="MCSigmoidDenseFA", **kwargs):
        if num_factors <= 0:
            raise NotImplementedError("num_factors = 0 (the diagonal method) is not built: num_factors >= 1")
        if return_unaveraged_logits or tune_temperature:
            raise NotImplementedError("return_unaveraged_logits / tune_temperature are not built")
        # the softmax head's constructor rejects num_classes <= 2; the sigmoid head has no such restriction
        super().__init__(num_classes=max(int(num_outputs), 3), num_factors=num_factors, temperature=temperature,
                         parameter_efficient=parameter_efficient, train_mc_samples=train_mc_samples,
                         test_mc_samples=test_mc_samples, compute_pred_variance=compute_pred_variance,
                         share_samples_across_batch=share_samples_across_batch, logits_only=logits_only, eps=eps,
                         dtype=dtype, kernel_regularizer=kernel_regularizer, bias_regularizer=bias_regularizer, name=name,
                         **kwargs)
        if int(num_outputs) != self._num_classes:  # rebuild the parameter layers at the true width
            from .dense import Dense

            self._num_classes = int(num_outputs)
            mk = lambda units, nm: Dense(units, kernel_regularizer=kernel_regularizer, bias_regularizer=bias_regularizer,  # noqa: E731
                                         dtype=dtype, name=f"{name}_{nm}")
            self._scale_layer = mk(self._num_classes * self._num_factors, "scale_layer")
            self._loc_layer = mk(self._num_classes, "loc_layer")
            self._diag_layer = mk(self._num_classes, "diag_layer")

    def call(self, inputs, training=True, seed=None):
        if inputs.dim() != 2:
            raise ValueError("the Monte-Carlo heads take [batch, features] inputs")
        locs, _ = self._compute_loc_param(inputs, training)
        fac, diag_raw = self._scale_params(inputs)
        samples = self._train_mc_samples if training else self._test_mc_samples
        flags = (1 if self._share_samples_across_batch else 0) | 4  # sigmoid link; one diagonal draw per class (TF)
        logits, probs, var = F.MCSoftmaxFA.apply(locs.float(), fac.float(), diag_raw.float(), None, 1.0, 1.0,
                                                 self._mc_noise(inputs.shape[0], seed), samples, self._num_factors, flags,
                                                 self._temperature, self._eps, self._compute_pred_variance)
        if self._logits_only:
            return logits
        probs_lo = torch.clamp(probs, min=self._eps)
        return (logits, torch.log(probs_lo), torch.clamp(probs_lo, self._eps, 1.0 - self._eps),
                (var if self._compute_pred_variance else None))


class HeteroscedasticSNGPLayer(MCSoftmaxDenseFA):
    """MC estimation of the softmax approximation to heteroscedastic SNGP predictions (hetsngp.py:25-245): the location
    parameter is the random-feature GP's output, and at evaluation the diagonal scale mixes in the SNGP marginal variance
    sqrt(het_var_weight diag² + sngp_var_weight diag(cov))."""

    def __init__(self, num_classes, num_factors=10, temperature=1.0, train_mc_samples=1000, test_mc_samples=1000,
                 compute_pred_variance=False, share_samples_across_batch=False, logits_only=True, eps=1e-7, dtype=None,
                 kernel_regularizer=None, bias_regularizer=None, num_inducing=1024, gp_kernel_type="gaussian",
                 gp_kernel_scale=1.0, gp_output_bias=0.0, normalize_input=False, gp_kernel_scale_trainable=False,
                 gp_output_bias_trainable=False, gp_cov_momentum=-1.0, gp_cov_ridge_penalty=1.0,
                 scale_random_features=True, use_custom_random_features=True, custom_random_features_initializer=None,
                 custom_random_features_activation=None, l2_regularization=1e-6, gp_cov_likelihood="gaussian",
                 return_gp_cov=True, return_random_features=False, sngp_var_weight=1.0, het_var_weight=1.0,
                 name="HeteroscedasticSNGPLayer", **kwargs):
        super().__init__(num_classes=num_classes, num_factors=num_factors, temperature=temperature,
                         train_mc_samples=train_mc_samples, test_mc_samples=test_mc_samples,
                         compute_pred_variance=compute_pred_variance,
                         share_samples_across_batch=share_samples_across_batch, logits_only=logits_only, eps=eps,
                         dtype=dtype, kernel_regularizer=kernel_regularizer, bias_regularizer=bias_regularizer, name=name)
        if not return_gp_cov or return_random_features:
            raise NotImplementedError("HeteroscedasticSNGPLayer unpacks (logits, covmat) from the GP layer (hetsngp.py:212)")
        self.sngp_layer = RandomFeatureGaussianProcess(
            units=num_classes, num_inducing=num_inducing, gp_kernel_type=gp_kernel_type, gp_kernel_scale=gp_kernel_scale,
            gp_output_bias=gp_output_bias, normalize_input=normalize_input,
            gp_kernel_scale_trainable=gp_kernel_scale_trainable, gp_output_bias_trainable=gp_output_bias_trainable,
            gp_cov_momentum=gp_cov_momentum, gp_cov_ridge_penalty=gp_cov_ridge_penalty,
            scale_random_features=scale_random_features, use_custom_random_features=use_custom_random_features,
            custom_random_features_initializer=custom_random_features_initializer,
            custom_random_features_activation=custom_random_features_activation, l2_regularization=l2_regularization,
            gp_cov_likelihood=gp_cov_likelihood, return_gp_cov=True, return_random_features=False, dtype=dtype,
            name="SNGP_layer", **kwargs)
        self.sngp_var_weight, self.het_var_weight = float(sngp_var_weight), float(het_var_weight)

    def _compute_loc_param(self, inputs, training=None):
        locs, covmat = self.sngp_layer(inputs, training=bool(training))
        return locs, covmat

    def _variance_mixing(self, covmat, training):
        if training:
            return None, 1.0, 1.0
        return torch.diagonal(covmat).contiguous(), self.het_var_weight, self.sngp_var_weight

    def reset_covariance_matrix(self):
        self.sngp_layer.reset_covariance_matrix()

    @property
    def losses(self):
        return super().losses + self.sngp_layer.losses
