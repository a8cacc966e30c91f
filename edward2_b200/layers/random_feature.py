"""`ed.layers` SNGP head on libed2b200: RandomFeatureGaussianProcess and LaplaceRandomFeatureCovariance.

Mirrors /root/reference/edward2/tensorflow/layers/random_feature.py (:33-281, :284-549): constructor arguments,
the `[logits, covmat(, features)]` list return, the training / inference state machine of the covariance layer and
its ValueErrors.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .. import functional as F, initializers, ops, regularizers
from .base import Layer, flatten_batch
from .dense import Dense

_SUPPORTED_LIKELIHOOD = ("binary_logistic", "poisson", "gaussian")


class LaplaceRandomFeatureCovariance(Layer):
    """Computes the Gaussian Process covariance using the Laplace method (random_feature.py:284-549).

    Training: precision <- momentum * precision + (1 - momentum) * ΦᵀΦ / B  (momentum > 0), else precision + ΦᵀΦ;
    returns the identity.  Inference: lazily inverts (ridge I + precision) once per update and returns
    ridge * Φ Σ Φᵀ.  `data_parallel=True` all-reduces ΦᵀΦ over the default process group (NCCL) before the blend,
    which reproduces the single-device global-batch result.  `data_parallel="deferred"` gives the same
    result without any per-step communication: every rank blends its own ΦᵀΦ / B_global into a rank-local matrix and
    the sum over ranks is taken once, when the covariance is needed (`dist.DeferredSum`; the update is linear in ΦᵀΦ).
    In that mode `precision_matrix` holds a rank-local partial during training: call `synchronize_precision()` before
    reading it directly (the layer does so itself before inverting, resetting or evaluating).
    """

    def __init__(self, momentum=0.999, ridge_penalty=1e-6, likelihood="gaussian", dtype=None,
                 name="laplace_covariance", use_on_read_synchronization_for_single_replica_vars=False,
                 data_parallel=False, **kwargs):
        if likelihood not in _SUPPORTED_LIKELIHOOD:
            raise ValueError(f'"likelihood" must be one of {_SUPPORTED_LIKELIHOOD}, got {likelihood}.')
        super().__init__(dtype=dtype, name=name, **kwargs)
        self.ridge_penalty = ridge_penalty
        self.momentum = momentum
        self.likelihood = likelihood
        self.data_parallel = data_parallel

    def build(self, input_shape):
        d = input_shape[-1]
        dev = self._device
        self.register_buffer("precision_matrix", torch.zeros(d, d, dtype=torch.float32, device=dev))
        self.register_buffer("covariance_matrix", torch.zeros(d, d, dtype=torch.float32, device=dev))
        self.register_buffer("if_update_covariance", torch.zeros((), dtype=torch.bool, device=dev))
        self._needs_inverse = False  # host mirror of if_update_covariance (avoids a device sync per call)
        self._flag_version = self.if_update_covariance._version
        self._cov_lp, self._cov_lp_version = None, None
        self._deferred = None
        if self.data_parallel == "deferred":
            from ..dist import DeferredSum

            self._deferred = DeferredSum(self.precision_matrix)
        self.built = True

    def _set_flag(self, value: bool):
        self._needs_inverse = value
        self.if_update_covariance.fill_(value)
        self._flag_version = self.if_update_covariance._version

    def _sync_flag(self):
        """`if_update_covariance` written from outside (load_state_dict / checkpoint.load_reference_variables copy_
        into the buffer and bump its version): re-derive the host mirror from the restored value — one device read,
        only on that occasion."""
        if self.if_update_covariance._version != self._flag_version:
            self._needs_inverse = bool(self.if_update_covariance.item())
            self._flag_version = self.if_update_covariance._version

    def synchronize_precision(self):
        """Collective no-op unless `data_parallel="deferred"` accumulated rank-local updates: sums them over ranks."""
        if self._deferred is not None:
            self._deferred.synchronize()
        return self.precision_matrix

    # -- helpers -----------------------------------------------------------------------------------------------
    def _adjusted_features(self, gp_feature, logits):
        """random_feature.py:384-406: non-Gaussian likelihoods scale Φ rows by sqrt(prob multiplier)."""
        if self.likelihood == "gaussian":
            return gp_feature
        if logits is None:
            raise ValueError(f'"logits" cannot be None when likelihood={self.likelihood}')
        if logits.shape[-1] != 1:
            raise ValueError(f"likelihood={self.likelihood} only support univariate logits."
                             f"Got logits dimension: {logits.shape[-1]}")
        lg = logits.detach().float().reshape(-1, 1)
        if self.likelihood == "binary_logistic":
            p = torch.sigmoid(lg)
            mult = p * (1.0 - p)
        else:
            mult = torch.exp(lg)
        return (torch.sqrt(mult) * gp_feature.float()).to(gp_feature.dtype)

    def update_feature_precision_matrix(self, gp_feature, logits=None):
        """In-place ΦᵀΦ contraction on the tcgen05 GEMM with the momentum blend in its epilogue."""
        phi = F.as_compute(self._adjusted_features(gp_feature.detach(), logits), self.compute_dtype)
        B = phi.shape[0]
        multi = torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
        if self._deferred is not None and multi:
            self._deferred.begin_local_update()
            ops.laplace_precision_update(phi, self.precision_matrix, self.momentum,
                                         batch_total=B * torch.distributed.get_world_size())
        elif self.data_parallel and multi:
            partial = torch.empty_like(self.precision_matrix)
            ops.laplace_precision_update(phi, self.precision_matrix, self.momentum, partial_out=partial)
            torch.distributed.all_reduce(partial)
            ops.precision_blend(self.precision_matrix, partial, self.momentum, B * torch.distributed.get_world_size())
        else:
            ops.laplace_precision_update(phi, self.precision_matrix, self.momentum, batch_total=B)
        return self.precision_matrix

    def reset_precision_matrix(self):
        if self._deferred is not None:
            self._deferred.open = False  # pending rank-local partials are discarded with the matrix
        self.precision_matrix.zero_()

    def update_feature_covariance_matrix(self):
        """random_feature.py:435-459: Sigma = (ridge I + P)^-1, computed on the GPU by the library's own block Gauss-Jordan
        kernel chain (include/ed2b200.h: ed2_spd_inverse; no cuSOLVER / torch.linalg call).  One-off per evaluation phase,
        outside the per-step figure."""
        self._sync_flag()
        if self._needs_inverse:
            self.synchronize_precision()
            a = self.precision_matrix.clone()
            a.diagonal().add_(self.ridge_penalty)
            sigma, _status = ops.spd_inverse(a)
            self.covariance_matrix.copy_(sigma)
        return self.covariance_matrix

    def _covariance_lp(self):
        if self.compute_dtype == torch.float32:
            return self.covariance_matrix
        # keyed on the buffer's version: the inverse above, load_state_dict and any other in-place write bump it
        if self._cov_lp is None or self._cov_lp_version != self.covariance_matrix._version:
            self._cov_lp = ops.cast(self.covariance_matrix, self.compute_dtype)
            self._cov_lp_version = self.covariance_matrix._version
        return self._cov_lp

    _EYE_CACHE = {}

    @classmethod
    def _identity(cls, n, device):
        """random_feature.py:531 returns tf.eye(batch_size) from every training call.  The matrix is constant: a dense
        copy is cached per (n, device) up to 1 GiB (no per-step write of B*B*4 bytes); beyond that (B = 65536 would be
        17 GB, SURVEY.md §7 hard part 9) a sparse COO identity with the same shape / dtype is returned."""
        key = (int(n), str(device))
        if key not in cls._EYE_CACHE:
            if 4 * n * n <= (1 << 30):
                cls._EYE_CACHE[key] = torch.eye(n, dtype=torch.float32, device=device)
            else:
                idx = torch.arange(n, device=device)
                cls._EYE_CACHE[key] = torch.sparse_coo_tensor(torch.stack([idx, idx]), torch.ones(n, device=device),
                                                              (n, n)).coalesce()
        return cls._EYE_CACHE[key]

    def compute_predictive_covariance(self, gp_feature):
        """random_feature.py:461-485: Φ (Σ Φᵀ ridge), full [B, B] — two GEMMs."""
        phi = F.as_compute(gp_feature.detach(), self.compute_dtype)
        B, D = phi.shape
        cov = self._covariance_lp()
        t = ops.gemm(phi, cov, B, D, D, major_a=ops.MAJOR_K, major_b=ops.MAJOR_K, out_dtype=self.compute_dtype)
        return ops.gemm(t, phi, B, B, D, major_a=ops.MAJOR_K, major_b=ops.MAJOR_K, out_dtype=torch.float32,
                        alpha=self.ridge_penalty)

    def compute_predictive_variance(self, gp_feature, logits=None, mean_field_factor=-1.0, likelihood="logistic"):
        """diag of compute_predictive_covariance without forming [B, B] (one GEMM with a row-sum epilogue), with
        optional fused mean-field logits.  Returns (variance [B], adjusted logits or None)."""
        self.update_feature_covariance_matrix()
        if self._needs_inverse:
            self._set_flag(False)
        phi = F.as_compute(gp_feature.detach(), self.compute_dtype)
        lk = 1 if likelihood == "poisson" else 0
        return ops.predictive_variance(phi, self._covariance_lp(), self.ridge_penalty, logits, mean_field_factor, lk)

    def _get_training_value(self, training=None):
        if training is None:
            training = self.training
        if isinstance(training, int):
            training = bool(training)
        return training

    def call(self, inputs, logits=None, training=None):
        batch_size = inputs.shape[0]
        training = self._get_training_value(training)
        if training:
            self.update_feature_precision_matrix(gp_feature=inputs, logits=logits)
            if not self._needs_inverse or self.if_update_covariance._version != self._flag_version:
                self._set_flag(True)
            return self._identity(batch_size, inputs.device)
        self.update_feature_covariance_matrix()
        if self._needs_inverse:
            self._set_flag(False)
        return self.compute_predictive_covariance(gp_feature=inputs)

    def get_config(self):
        cfg = super().get_config()
        cfg.update(momentum=self.momentum, ridge_penalty=self.ridge_penalty, likelihood=self.likelihood)
        return cfg


class _LayerNormalization(Layer):
    """tf.keras.layers.LayerNormalization defaults: axis=-1, epsilon=1e-3, trainable gamma / beta."""

    def __init__(self, epsilon=1e-3, **kwargs):
        super().__init__(**kwargs)
        self.epsilon = epsilon

    def build(self, input_shape):
        d = input_shape[-1]
        self.gamma = nn.Parameter(torch.ones(d, device=self._device))
        self.beta = nn.Parameter(torch.zeros(d, device=self._device))
        self.built = True

    def call(self, inputs):
        return F.LayerNorm.apply(inputs, self.gamma, self.beta, self.epsilon, self.compute_dtype)


class _RandomFeature(Layer):
    """Dense(num_inducing, activation=cos, trainable=False) of random_feature.py:228-235, fused with the
    sqrt(2/D) feature scale (:260-266) in the GEMM epilogue.

    phase_precision: 'extended' (default) feeds the bf16 tensor cores with split operands x = h + m, W = h + m (16
    mantissa bits, one GEMM with K' = 3K); 'extended3' adds a third term (24 bits, K' = 6K); 'native' uses the compute
    dtype as is.  With the TF defaults (LayerNorm inputs, orthogonal features of stddev 1) the phase is O(sqrt(d))
    radians, so bf16 operands alone put ~0.1 rad of error into the cosine (SURVEY.md §7 hard part 6).  Measured on
    B200 at d = 1024: native 1.4e-1 of max|phi|, extended 1.9e-3, extended3 1.9e-3 — beyond 16 operand bits the error is
    set by the tensor core's fp32 accumulation (it truncates; an FFMA reference is at ~3e-5), so 'extended' is the
    default and fp32-exact features need dtype='float32' (FFMA path)."""

    def __init__(self, units, kernel_initializer, bias_initializer, scale, phase_precision="extended", **kwargs):
        super().__init__(**kwargs)
        self.units, self.scale = units, scale
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer
        if phase_precision not in ("extended", "extended3", "native"):
            raise ValueError(f"phase_precision must be 'extended', 'extended3' or 'native', got {phase_precision!r}")
        self.phase_precision = phase_precision

    def build(self, input_shape):
        dev = self._device
        k = self.kernel_initializer((input_shape[-1], self.units)).to(dev)
        self.register_buffer("kernel", k)
        self.register_buffer("bias", self.bias_initializer((self.units,)).to(dev))
        self._refresh()
        self.built = True

    def _cache_key(self):
        return (self.kernel.data_ptr(), self.kernel._version, self.kernel.dtype)

    def _refresh(self):
        k = self.kernel
        self._key = self._cache_key()
        d = k.shape[0]
        nsplit = {"extended": 2, "extended3": 3, "native": 0}[self.phase_precision]
        # the exact-fp32 FFMA path is kept for small fp32 problems; split operands need a pitch multiple of 8
        if nsplit and ((6 if nsplit == 3 else 3) * d) % 8 != 0:
            nsplit = 0
        if nsplit and self.compute_dtype == torch.float32 and d * self.units < (1 << 18):
            nsplit = 0
        self._nsplit = nsplit
        kt = k.t().contiguous()
        if nsplit:
            self._wt = ops.split_bf16(kt, True, nsplit)
            self._w = ops.cast(k, torch.bfloat16)
        else:
            self._w = k if self.compute_dtype == torch.float32 else ops.cast(k, self.compute_dtype)
            self._wt = kt if self.compute_dtype == torch.float32 else ops.cast(kt, self.compute_dtype)

    def call(self, inputs):
        if self._key != self._cache_key():  # kernel restored from a checkpoint: rebuild the operand copies
            self._refresh()
        return F.RandomFourierFeatures.apply(inputs, self._wt, self._w, self.bias, self.scale, self.compute_dtype,
                                             self.compute_dtype, self._nsplit)


class RandomFeatureGaussianProcess(Layer):
    """Gaussian process layer with random feature approximation (random_feature.py:33-281).

    call(inputs, global_step=None, training=None) -> [gp_output, gp_covmat(, gp_feature)].
    """

    def __init__(self, units, num_inducing=1024, gp_kernel_type="gaussian", gp_kernel_scale=1.0, gp_output_bias=0.0,
                 normalize_input=True, gp_kernel_scale_trainable=False, gp_output_bias_trainable=False,
                 gp_cov_momentum=0.999, gp_cov_ridge_penalty=1e-6, scale_random_features=True,
                 use_custom_random_features=True, custom_random_features_initializer=None,
                 custom_random_features_activation=None, l2_regularization=0.0, gp_cov_likelihood="gaussian",
                 return_gp_cov=True, return_random_features=False, dtype=None,
                 name="random_feature_gaussian_process", data_parallel=False, phase_precision="extended",
                 **gp_output_kwargs):
        super().__init__(dtype=dtype, name=name)
        self.units, self.num_inducing = units, num_inducing
        self.normalize_input = normalize_input
        self.gp_input_scale = 1.0 / math.sqrt(gp_kernel_scale) if gp_kernel_scale is not None else None
        self.gp_feature_scale = math.sqrt(2.0 / float(num_inducing))
        self.scale_random_features = scale_random_features
        self.return_random_features = return_random_features
        self.return_gp_cov = return_gp_cov
        self.gp_kernel_type = gp_kernel_type
        self.gp_kernel_scale = gp_kernel_scale
        self.gp_output_bias = gp_output_bias
        self.gp_kernel_scale_trainable = gp_kernel_scale_trainable
        self.gp_output_bias_trainable = gp_output_bias_trainable
        self.use_custom_random_features = use_custom_random_features
        if not use_custom_random_features:
            raise NotImplementedError("tf.keras RandomFourierFeatures (use_custom_random_features=False) is not built")
        self._linear = gp_kernel_type.lower() == "linear"   # identity features (random_feature.py:221-223)
        if custom_random_features_activation is not None and getattr(custom_random_features_activation, "__name__", "") != "cos":
            raise NotImplementedError("only the cosine random-feature activation is fused")
        self.custom_random_features_initializer = (
            initializers.get(custom_random_features_initializer) if custom_random_features_initializer is not None
            else initializers.OrthogonalRandomFeatures(stddev=1.0, seed=0))
        self.random_features_bias_initializer = initializers.RandomUniform(minval=0.0, maxval=2.0 * math.pi, seed=0)
        self.l2_regularization = l2_regularization
        self.gp_output_kwargs = gp_output_kwargs
        self.gp_cov_momentum = gp_cov_momentum
        self.gp_cov_ridge_penalty = gp_cov_ridge_penalty
        self.gp_cov_likelihood = gp_cov_likelihood
        self.data_parallel = data_parallel
        self.phase_precision = phase_precision

    def build(self, input_shape):
        dt = self.compute_dtype
        # the normalised input stays fp32 when the phase is evaluated at fp32 grade (it is split on the fly)
        norm_dt = torch.float32 if (self.phase_precision != "native" and not self._linear) else dt
        if self.normalize_input:
            self._input_norm_layer = _LayerNormalization(dtype=norm_dt, name="gp_input_normalization")
        scale = self.gp_feature_scale if self.scale_random_features else 1.0
        self._random_feature = None if self._linear else _RandomFeature(
            self.num_inducing, self.custom_random_features_initializer, self.random_features_bias_initializer, scale,
            phase_precision=self.phase_precision, dtype=dt, name="gp_random_feature")
        if self.return_gp_cov:
            self._gp_cov_layer = LaplaceRandomFeatureCovariance(
                momentum=self.gp_cov_momentum, ridge_penalty=self.gp_cov_ridge_penalty,
                likelihood=self.gp_cov_likelihood, dtype=dt, name="gp_covariance", data_parallel=self.data_parallel)
        reg = regularizers.L2(self.l2_regularization) if self.l2_regularization else None
        self._gp_output_layer = Dense(self.units, use_bias=False, kernel_regularizer=reg, dtype=dt,
                                      name="gp_output_weights", **self.gp_output_kwargs)
        bias = torch.full((self.units,), float(self.gp_output_bias), device=self._device)
        if self.gp_output_bias_trainable:
            self._gp_output_bias = nn.Parameter(bias)
        else:
            self.register_buffer("_gp_output_bias", bias)
        self.built = True

    def reset_covariance_matrix(self):
        self._gp_cov_layer.reset_precision_matrix()

    def features(self, x):
        """Random features of a 2-D batch: input normalisation / scaling + cos(xW + b) * scale (random_feature.py:250-266)."""
        gp_inputs = x
        rf = self._random_feature
        if self._linear:
            # gp_kernel_type='linear': the features are the (normalised / rescaled) inputs themselves, times
            # sqrt(2 / num_inducing) when scale_random_features (:250-266 with the Lambda layer of :221-223)
            if self.normalize_input:
                gp_inputs = self._input_norm_layer(gp_inputs)
            elif self.gp_input_scale is not None:
                gp_inputs = gp_inputs * self.gp_input_scale
            if self.scale_random_features:
                gp_inputs = gp_inputs * self.gp_feature_scale
            return gp_inputs.to(self.compute_dtype)
        if (self.normalize_input and self._input_norm_layer.built and rf.built and rf._nsplit == 2
                and x.shape[-1] % 8 == 0 and x.dtype in (torch.float32, torch.bfloat16)):
            # fused unit: LayerNorm writes the split operand, the GEMM epilogue writes phi and -scale*sin(phase)
            if rf._key != rf._cache_key():
                rf._refresh()
            ln = self._input_norm_layer
            return F.NormalizedRFF.apply(x, ln.gamma, ln.beta, ln.epsilon, rf._wt, rf._w, rf.bias, rf.scale, rf.compute_dtype)
        if self.normalize_input:
            gp_inputs = self._input_norm_layer(gp_inputs)
        elif self.gp_input_scale is not None:
            gp_inputs = ops.input_normalize(F.as_compute(gp_inputs.detach(), self.compute_dtype), self.compute_dtype,
                                            scale_only=True, scale=self.gp_input_scale) \
                if not gp_inputs.requires_grad else gp_inputs * self.gp_input_scale
        return self._random_feature(gp_inputs)

    def call(self, inputs, global_step=None, training=None):
        if training is None:
            training = self.training  # sub-layers are built lazily: they must follow THIS module's train / eval mode
        x, lead = flatten_batch(inputs)
        gp_feature = self.features(x)
        gp_output = self._gp_output_layer(gp_feature).float() + self._gp_output_bias
        model_output = [gp_output.reshape(*lead, self.units)]
        if self.return_gp_cov:
            model_output.append(self._gp_cov_layer(gp_feature, gp_output, training))
        if self.return_random_features:
            model_output.append(gp_feature)
        return model_output

    @property
    def losses(self):
        return self._gp_output_layer.losses if self.built else []

    def get_config(self):
        cfg = super().get_config()
        cfg.update(units=self.units, num_inducing=self.num_inducing, gp_kernel_scale=self.gp_kernel_scale,
                   normalize_input=self.normalize_input, gp_cov_momentum=self.gp_cov_momentum,
                   gp_cov_ridge_penalty=self.gp_cov_ridge_penalty, scale_random_features=self.scale_random_features,
                   l2_regularization=self.l2_regularization, gp_cov_likelihood=self.gp_cov_likelihood,
                   return_gp_cov=self.return_gp_cov, return_random_features=self.return_random_features)
        return cfg
