"""`edward2.jax.nn` SNGP head on libed2b200: RandomFourierFeatures (:143-214), LaplaceRandomFeatureCovariance
(:217-404) and RandomFeatureGaussianProcess (:72-140) of edward2/jax/nn/random_feature.py."""
from __future__ import annotations

import math
from typing import Any, Callable, Mapping, Optional

import torch

from .. import functional as F, ops
from ..layers.base import resolve_dtype
from . import module as m
from .normalization import Dense

SUPPORTED_LIKELIHOOD = ("binary_logistic", "poisson", "gaussian")

default_rbf_bias_init = m.uniform(scale=2.0 * math.pi)
default_rbf_kernel_init = m.variance_scaling(2.0, "fan_in", "normal")


def _dt(dtype, inputs):
    if dtype is not None:
        return resolve_dtype(dtype)
    return inputs.dtype if inputs.dtype in (torch.float32, torch.bfloat16) else torch.float32


class RandomFourierFeatures(m.Module):
    """f(x) = feature_scale * cos(x @ kernel + bias), kernel / bias frozen in the `random_features` collection."""
    features: int
    feature_scale: Optional[float] = 1.0
    activation: Optional[Callable] = None
    kernel_init: Callable = default_rbf_kernel_init
    bias_init: Callable = default_rbf_bias_init
    seed: int = 0
    dtype: Optional[Any] = None
    collection_name: str = "random_features"
    phase_precision: str = "extended"   # 'extended' | 'extended3' | 'native' (see layers.random_feature._RandomFeature)

    def setup(self):
        if self.activation is not None and getattr(self.activation, "__name__", "") != "cos":
            raise NotImplementedError("only the cosine random-feature activation is fused")
        if self.phase_precision not in ("extended", "extended3", "native"):
            raise ValueError(f"phase_precision must be 'extended', 'extended3' or 'native', got {self.phase_precision!r}")
        self._feature_scale = math.sqrt(2.0 / self.features) if self.feature_scale is None else float(self.feature_scale)
        self._operand_cache = {}

    def _operands(self, k, dtype):
        """Frozen-kernel operand copies (transposed / split / cast), cached per kernel tensor."""
        key = (k.data_ptr(), k._version, dtype)
        hit = self._operand_cache.get(key)
        if hit is not None:
            return hit
        d = k.shape[0]
        nsplit = {"extended": 2, "extended3": 3, "native": 0}[self.phase_precision]
        if nsplit and ((6 if nsplit == 3 else 3) * d) % 8 != 0:
            nsplit = 0
        if nsplit and dtype == torch.float32 and d * self.features < (1 << 18):
            nsplit = 0   # small fp32 problems stay on the exact-fp32 FFMA path
        kt = k.t().contiguous()
        if nsplit:
            wt, w = ops.split_bf16(kt, True, nsplit), ops.cast(k, torch.bfloat16)
        else:
            w = k if dtype == torch.float32 else ops.cast(k, dtype)
            wt = kt if dtype == torch.float32 else ops.cast(kt, dtype)
        self._operand_cache = {key: (wt, w, nsplit)}
        return wt, w, nsplit

    def __call__(self, inputs):
        dtype = _dt(self.dtype, inputs)
        d = inputs.shape[-1]
        rng = m.PRNG(self.seed)  # fixed seed overrides external rngs (random_feature.py:169,175)
        kernel = self.variable(self.collection_name, "kernel", lambda: self.kernel_init(rng.fold("kernel"), (d, self.features)))
        bias = self.variable(self.collection_name, "bias", lambda: self.bias_init(rng.fold("bias"), (self.features,)))
        wt, w, nsplit = self._operands(kernel.value, dtype)
        x2 = inputs.reshape(-1, d)
        phi = F.RandomFourierFeatures.apply(x2, wt, w, bias.value, self._feature_scale, dtype, dtype, nsplit)
        return phi.reshape(*inputs.shape[:-1], self.features)


class LaplaceRandomFeatureCovariance(m.Module):
    hidden_features: int
    ridge_penalty: float = 1.0
    momentum: Optional[float] = None
    likelihood: str = "gaussian"
    collection_name: str = "laplace_covariance"
    dtype: Optional[Any] = None

    def setup(self):
        if self.momentum is not None and (self.momentum < 0.0 or self.momentum > 1.0):
            raise ValueError(f"`momentum` must be between (0, 1). Got {self.momentum}.")
        if self.likelihood not in SUPPORTED_LIKELIHOOD:
            raise ValueError(f'"likelihood" must be one of {SUPPORTED_LIKELIHOOD}, got {self.likelihood}.')

    def initial_precision_matrix(self):
        return torch.eye(self.hidden_features) * self.ridge_penalty

    def __call__(self, gp_features, gp_logits=None, diagonal_only: bool = True):
        dtype = _dt(self.dtype, gp_features)
        phi = gp_features.detach().reshape(-1, self.hidden_features)
        if gp_logits is not None:
            gp_logits = gp_logits.detach().reshape(phi.shape[0], -1)
        precision = self.variable(self.collection_name, "precision_matrix", self.initial_precision_matrix)
        initializing = self.is_initializing()
        training = self.is_mutable_collection(self.collection_name)
        if training and not initializing:
            precision.value = self.update_precision_matrix(phi, gp_logits, precision.value, dtype)
        if not training:
            return self.compute_predictive_covariance(phi, precision, diagonal_only, dtype)
        return None

    def update_precision_matrix(self, gp_features, gp_logits, precision_matrix, dtype=torch.float32):
        if self.likelihood != "gaussian":
            if gp_logits is None:
                raise ValueError(f"`gp_logits` cannot be None when likelihood=`{self.likelihood}`")
            if gp_logits.dim() > 1 and gp_logits.shape[-1] != 1:
                raise ValueError(f"likelihood `{self.likelihood}` only support univariate logits. "
                                 f"Got logits dimension: {gp_logits.shape[-1]}")
            lg = gp_logits.float().reshape(-1, 1)
            mult = torch.sigmoid(lg) * (1 - torch.sigmoid(lg)) if self.likelihood == "binary_logistic" else torch.exp(lg)
            gp_features = (torch.sqrt(mult) * gp_features.float()).to(gp_features.dtype)
        phi = F.as_compute(gp_features, dtype)
        new = precision_matrix.detach().clone().contiguous()
        # momentum None -> exact sum (kernel convention: momentum <= 0)
        ops.laplace_precision_update(phi, new, -1.0 if self.momentum is None else float(self.momentum))
        return new

    def compute_predictive_covariance(self, gp_features, precision_matrix, diagonal_only, dtype=torch.float32):
        """ridge * Φ P^-1 Φᵀ (:393-404).  The reference factorises P by Cholesky on every call; here P^-1 comes from
        the library's own on-GPU block Gauss-Jordan inverse (ed2_spd_inverse) and the contraction with
        Φ is the fused GEMM (row-sum epilogue for the diagonal)."""
        p = precision_matrix.value if hasattr(precision_matrix, "value") else precision_matrix
        sigma, _status = ops.spd_inverse(p.detach().to(torch.float32))
        phi = F.as_compute(gp_features, dtype)
        sig = sigma if dtype == torch.float32 else ops.cast(sigma, dtype)
        if diagonal_only:
            var, _ = ops.predictive_variance(phi, sig, float(self.ridge_penalty))
            return var
        B, D = phi.shape
        t = ops.gemm(phi, sig, B, D, D, major_a=ops.MAJOR_K, major_b=ops.MAJOR_K, out_dtype=dtype)
        return ops.gemm(t, phi, B, B, D, major_a=ops.MAJOR_K, major_b=ops.MAJOR_K, out_dtype=torch.float32,
                        alpha=float(self.ridge_penalty))


class _LayerNorm(m.Module):
    """flax nn.LayerNorm(use_bias=False, use_scale=False): epsilon 1e-6, no parameters."""
    epsilon: float = 1e-6
    dtype: Optional[Any] = None
    keep_fp32: bool = False

    def __call__(self, inputs):
        dtype = _dt(self.dtype, inputs)
        x2 = inputs.reshape(-1, inputs.shape[-1])
        out_dtype = torch.float32 if self.keep_fp32 else dtype
        return F.LayerNorm.apply(x2, None, None, float(self.epsilon), out_dtype).reshape(inputs.shape)


class RandomFeatureGaussianProcess(m.Module):
    features: int
    hidden_features: int = 1024
    normalize_input: bool = True
    norm_kwargs: Mapping[str, Any] = None
    hidden_kwargs: Mapping[str, Any] = None
    output_kwargs: Mapping[str, Any] = None
    covmat_kwargs: Mapping[str, Any] = None

    def setup(self):
        self.hidden_layer = RandomFourierFeatures(features=self.hidden_features, **(self.hidden_kwargs or {}))
        if self.normalize_input:
            self.norm_layer = _LayerNorm(keep_fp32=self.hidden_layer.phase_precision != "native",
                                         **(self.norm_kwargs or {}))
        self.output_layer = Dense(features=self.features, **(self.output_kwargs or {}))
        self.covmat_layer = LaplaceRandomFeatureCovariance(hidden_features=self.hidden_features,
                                                           **(self.covmat_kwargs or {}))

    def __call__(self, inputs, return_full_covmat: bool = False, return_random_features: bool = False):
        gp_inputs = self.sub(self.norm_layer, "norm_layer")(inputs) if self.normalize_input else inputs
        gp_features = self.sub(self.hidden_layer, "hidden_layer")(gp_inputs)
        gp_logits = self.sub(self.output_layer, "output_layer")(gp_features)
        gp_covmat = self.sub(self.covmat_layer, "covmat_layer")(gp_features, gp_logits,
                                                                diagonal_only=not return_full_covmat)
        if return_random_features:
            return gp_logits, gp_covmat, gp_features
        return gp_logits, gp_covmat


class MCSigmoidDenseFASNGP(m.Module):
    """`edward2.jax.nn.MCSigmoidDenseFASNGP` (edward2/jax/nn/random_feature.py:407-695): heteroscedastic SNGP head with
    sigmoid outputs.  The location is the RandomFeatureGaussianProcess head (`loc_layer`), the factor loadings and
    the diagonal scale are Dense layers (`scale_layer`, `diag_layer`); at evaluation the diagonal scale becomes
    sqrt(het_var_weight·diag² + sngp_var_weight·var_GP(x)) (:531-536) inside the Monte-Carlo kernel
    (include/ed2b200.h: ed2_mc_softmax_fa_fwd, sigmoid link, JAX noise form).  Returns (logits, covmat_sngp) if
    logits_only else (logits, covmat_sngp, log_probs, probs).  Not built (raise): parameter_efficient, num_factors == 0."""
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
    het_var_weight: float = 1.0
    sngp_var_weight: float = 0.0
    hidden_features: int = 1024
    normalize_input: bool = True
    norm_kwargs: Mapping[str, Any] = None
    hidden_kwargs: Mapping[str, Any] = None
    output_kwargs: Mapping[str, Any] = None
    covmat_kwargs: Mapping[str, Any] = None

    def _dense(self, features):
        return Dense(features)

    def setup(self):
        K = self.num_outputs
        self._loc_layer = RandomFeatureGaussianProcess(features=K, hidden_features=self.hidden_features,
                                                       normalize_input=self.normalize_input, norm_kwargs=self.norm_kwargs,
                                                       hidden_kwargs=self.hidden_kwargs, output_kwargs=self.output_kwargs,
                                                       covmat_kwargs=self.covmat_kwargs)
        self._scale_layer = self._dense(K * max(self.num_factors, 1))
        self._diag_layer = self._dense(K)

    def _noise(self, injected=None):
        if injected is not None:
            return F.Randomness(injected)
        a, b = self.make_rng("diag_noise_samples"), self.make_rng("standard_norm_noise_samples")
        return F.Randomness(seed=(a.seed * 0x9E3779B97F4A7C15 + b.seed) & 0xFFFFFFFFFFFFFFFF, stream=7)

    def __call__(self, inputs, training=True, mc_noise=None):
        """`mc_noise` (tests): injected noise [batch or 1, samples, round8(num_factors) + round8(num_outputs)] fp32 —
        factor draws first, the diagonal draw at element round8(num_factors)."""
        if self.parameter_efficient or self.num_factors <= 0:
            raise NotImplementedError("parameter_efficient / num_factors == 0 are not built")
        K, Fn = self.num_outputs, self.num_factors
        locs, covmat_sngp = self.sub(self._loc_layer, "loc_layer")(inputs)
        fac = self.sub(self._scale_layer, "scale_layer")(inputs)
        diag_raw = self.sub(self._diag_layer, "diag_layer")(inputs)
        initializing = self.is_mutable_collection("params")
        mix = not (training or initializing)                                  # :528-536
        if mix and covmat_sngp is None:
            raise ValueError("evaluation needs the GP variance: apply without `laplace_covariance` in `mutable`")
        samples = self.train_mc_samples if training else self.test_mc_samples
        flags = (1 if self.share_samples_across_batch else 0) | 2 | 4         # JAX diagonal noise, sigmoid link
        logits, probs, _ = F.MCSoftmaxFA.apply(locs.float(), fac.float(), diag_raw.float(),
                                               covmat_sngp.reshape(-1) if mix else None, float(self.het_var_weight),
                                               float(self.sngp_var_weight), self._noise(mc_noise), samples, Fn, flags,
                                               float(self.temperature), float(self.eps), False)
        probs_lo = torch.clamp(probs, min=self.eps)                           # :680
        log_probs = torch.log(probs_lo)
        if self.return_locs:
            logits = locs
        if self.logits_only:
            return logits, covmat_sngp
        return logits, covmat_sngp, log_probs, torch.clamp(probs_lo, self.eps, 1.0 - self.eps)


class MCSigmoidDenseFASNGPBE(MCSigmoidDenseFASNGP):
    """`edward2.jax.nn.MCSigmoidDenseFASNGPBE` (:698-735): the same head with BatchEnsemble scale / diagonal layers
    (`DenseBatchEnsemble(ens_size)`, batch rows member-major)."""
    ens_size: int = 1

    def _dense(self, features):
        from .dense import DenseBatchEnsemble
        return DenseBatchEnsemble(features, ens_size=self.ens_size)
