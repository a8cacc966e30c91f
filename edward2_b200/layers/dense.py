"""`ed.layers` dense layers on libed2b200: DenseReparameterization, DenseFlipout, DenseBatchEnsemble, DenseRank1 and
the deterministic Dense that SpectralNormalization wraps.

Constructor arguments, call signatures, error behaviour and `.losses` follow
/root/reference/edward2/tensorflow/layers/dense.py (:33-83, :227-285, :470-607, :991-1175).  Parameters are fp32
masters; `dtype` selects the compute dtype ('float32' or 'bfloat16').
"""
from __future__ import annotations

import torch
from torch import nn

from .. import constraints, functional as F, initializers, ops, regularizers
from .base import Layer, flatten_batch

# Philox stream ids of the noise tensors of one layer call
S_KERNEL, S_BIAS, S_SIGN_IN, S_SIGN_OUT, S_R, S_S, S_B = 0, 1, 2, 3, 4, 5, 6


def _is_trainable(init) -> bool:
    return isinstance(init, initializers.TrainableInitializer)


def _make_weight(layer: nn.Module, name: str, shape, initializer, device):
    """`add_weight` semantics of edward2/tensorflow/layers/utils.py:38-75: a trainable initializer becomes the
    weight's variational posterior (it owns mean / stddev); a plain initializer yields a Parameter."""
    if _is_trainable(initializer):
        if not initializer.built:
            initializer.build(tuple(shape), device=device)
        return None  # the initializer module (already an attribute of the layer) owns <name>/mean, <name>/stddev
    p = nn.Parameter(initializer(tuple(shape)).to(device=device, dtype=torch.float32))
    layer.register_parameter(name, p)
    return p


class _FusedKL:
    """KL tensor produced by the sampling pass of the last forward, valid while (mean, rho) are unchanged."""

    def __init__(self):
        self.value = None
        self.versions = None

    def set(self, kl, mean, rho):
        self.value, self.versions = kl, (mean._version, rho._version)

    def get(self, mean, rho):
        if self.value is not None and self.versions == (mean._version, rho._version):
            return self.value
        return None


class Dense(Layer):
    """Deterministic densely-connected layer (the tf.keras.layers.Dense the reference wraps / subclasses)."""

    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None,
                 kernel_constraint=None, bias_constraint=None, **kwargs):
        super().__init__(**kwargs)
        self.units = int(units)
        self.activation = activation
        self._act = ops.act_id(activation)
        self.use_bias = use_bias
        self.kernel_initializer = initializers.get(kernel_initializer)
        self.bias_initializer = initializers.get(bias_initializer)
        self.kernel_regularizer = regularizers.get(kernel_regularizer)
        self.bias_regularizer = regularizers.get(bias_regularizer)
        self._w_scale = None  # set by the JAX-style SpectralNormalization (stop-gradient factor)

    def build(self, input_shape):
        dev = self._device
        self.kernel = nn.Parameter(self.kernel_initializer((input_shape[-1], self.units)).to(dev))
        self.bias = nn.Parameter(self.bias_initializer((self.units,)).to(dev)) if self.use_bias else None
        self.built = True

    def call(self, inputs, training=None):
        x, lead = flatten_batch(inputs)
        y = F.Dense.apply(x, self.kernel, self.bias, self._act, self.compute_dtype, None)
        return y.reshape(*lead, self.units)

    @property
    def losses(self):
        out = []
        if self.kernel_regularizer is not None:
            out.append(self.kernel_regularizer(self.kernel))
        if self.bias_regularizer is not None and self.bias is not None:
            out.append(self.bias_regularizer(self.bias))
        return out

    def get_config(self):
        cfg = super().get_config()
        cfg.update(units=self.units, activation=self.activation, use_bias=self.use_bias,
                   kernel_initializer=initializers.serialize(self.kernel_initializer),
                   bias_initializer=initializers.serialize(self.bias_initializer))
        return cfg


class DenseReparameterization(Layer):
    """Bayesian dense layer estimated via reparameterization (dense.py:33-83): y = act(x (mu + sigma∘eps) + b) with
    one kernel sample per call, drawn in-kernel; `.losses` holds the KL regularizer (fused into the sampling pass)."""

    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="trainable_normal",
                 bias_initializer="zero", kernel_regularizer="normal_kl_divergence", bias_regularizer=None,
                 activity_regularizer=None, **kwargs):
        super().__init__(**kwargs)
        self.units = int(units)
        self.activation = activation
        self._act = ops.act_id(activation)
        self.use_bias = use_bias
        self.kernel_initializer = initializers.get(kernel_initializer)
        self.bias_initializer = initializers.get(bias_initializer)
        self.kernel_regularizer = regularizers.get(kernel_regularizer)
        self.bias_regularizer = regularizers.get(bias_regularizer)
        if activity_regularizer is not None:
            raise NotImplementedError("activity_regularizer is outside the hot path")
        self._kl = _FusedKL()

    def build(self, input_shape):
        dev = self._device
        in_dim = input_shape[-1]
        self.kernel = _make_weight(self, "kernel", (in_dim, self.units), self.kernel_initializer, dev)
        self.bias = _make_weight(self, "bias", (self.units,), self.bias_initializer, dev) if self.use_bias else None
        self.built = True

    # -- weights ---------------------------------------------------------------------------------------------
    def _kernel_params(self):
        if _is_trainable(self.kernel_initializer):
            return self.kernel_initializer.params()
        return self.kernel, None

    def _bias_value(self):
        """dense.py:73-78: a trainable bias initializer is re-sampled on every call (small vector: host-side
        composition of the Philox draw with autograd through mean / rho)."""
        if not self.use_bias:
            return None
        if _is_trainable(self.bias_initializer):
            mean, rho = self.bias_initializer.params()
            if rho is None:
                return mean
            nz = self._noise("bias_eps", S_BIAS)
            eps = nz.injected if nz.injected is not None else ops.philox("normal", mean.numel(), nz.seed, nz.stream, mean.device)
            return mean + (torch.nn.functional.softplus(rho) + constraints.EPSILON) * eps.reshape(mean.shape)
        return self.bias

    def _fused_kl_args(self):
        r = self.kernel_regularizer
        if isinstance(r, regularizers.NormalKLDivergence):
            return r.scale_factor, r.mean, r.stddev
        return 0.0, 0.0, 1.0

    def presample(self, after=None):
        """Optional: draw this call's kernel sample (and its KL) NOW on a side stream, so that the issue-bound sampling
        pass overlaps whatever the main stream is doing (typically the previous layer's GEMM).  The next `call` uses
        it.  Purely a scheduling hint: the draw is the same Philox stream the un-hinted call would use.
        `after`: another layer whose presample must finish first (so that the first layer's sample, which the first
        GEMM is waiting for, is not slowed by the later layers' sampling)."""
        if not self.built or not _is_trainable(self.kernel_initializer):
            return
        mean, rho = self._kernel_params()
        if rho is None:
            return
        from .. import _lib
        import ctypes as C

        kl_scale, pm, ps = self._fused_kl_args()
        dev = mean.device
        I, O = mean.shape
        if not hasattr(self, "_side_stream"):
            self._side_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        w_lp = _lib.empty_padded(I, O, self.compute_dtype, dev)
        kl64 = torch.zeros((), dtype=torch.float64, device=dev) if kl_scale != 0 else None
        self._side_stream.wait_stream(main)
        if after is not None and getattr(after, "_side_stream", None) is not None:
            self._side_stream.wait_stream(after._side_stream)
        with torch.cuda.stream(self._side_stream):
            nz = self._noise("eps", S_KERNEL)
            lib = _lib.load()
            import os

            split = (self.compute_dtype == torch.bfloat16 and kl64 is not None
                     and os.environ.get("ED2_KL_FUSED", "1") == "0")
            # bf16: the lean sampling kernel reduces the KL itself; ED2_KL_FUSED=0: separate accurate KL pass after it
            _lib.check(lib.ed2_sample_normal_weights(
                mean.data_ptr(), rho.data_ptr(), nz.c(), I, O, 0, w_lp.data_ptr(), _lib.dt_of(w_lp), _lib.ld(w_lp),
                None, 0, None if (kl64 is None or split) else kl64.data_ptr(), kl_scale, pm, ps, _lib.stream_ptr()),
                "ed2_sample_normal_weights")
            if split:
                _lib.check(lib.ed2_normal_kl_fwd(mean.data_ptr(), rho.data_ptr(), I * O, pm, ps, kl_scale,
                                                 kl64.data_ptr(), _lib.stream_ptr()), "ed2_normal_kl_fwd")
        self._presampled = (w_lp, kl64, nz)

    def call(self, inputs, training=None):
        in_link = F.take_link(inputs)
        x, lead = flatten_batch(inputs)
        mean, rho = self._kernel_params()
        bias = self._bias_value()
        if rho is None:  # deterministic kernel: ordinary dense layer
            y = F.Dense.apply(x, mean, bias, self._act, self.compute_dtype, None)
            return y.reshape(*lead, self.units)
        kl_scale, pm, ps = self._fused_kl_args()
        fusable = len(lead) == 1 and not _is_trainable(self.bias_initializer)
        out_link = F.ReluLink() if (fusable and self._act == ops.ACT_RELU) else None
        pre = getattr(self, "_presampled", None)
        self._presampled = None
        if pre is not None:
            w_lp, kl64, nz = pre
            torch.cuda.current_stream().wait_stream(self._side_stream)
            y, kl = F.ReparamDense.apply(x, mean, rho, bias, self._act, self.compute_dtype, nz, kl_scale, pm, ps,
                                         in_link if fusable else None, out_link, (w_lp, kl64))
        else:
            y, kl = F.ReparamDense.apply(x, mean, rho, bias, self._act, self.compute_dtype,
                                         self._noise("eps", S_KERNEL), kl_scale, pm, ps,
                                         in_link if fusable else None, out_link, None)
        if kl_scale != 0.0:
            self._kl.set(kl, mean, rho)
        if len(lead) == 1:
            return F.attach_link(y, out_link)
        return y.reshape(*lead, self.units)

    @property
    def losses(self):
        out = []
        if self.kernel_regularizer is not None and _is_trainable(self.kernel_initializer):
            mean, rho = self._kernel_params()
            fused = self._kl.get(mean, rho) if rho is not None else None
            out.append(fused if fused is not None else self.kernel_regularizer(self.kernel_initializer))
        if self.bias_regularizer is not None and self.use_bias and _is_trainable(self.bias_initializer):
            out.append(self.bias_regularizer(self.bias_initializer))
        return out

    def get_config(self):
        cfg = super().get_config()
        cfg.update(units=self.units, activation=self.activation, use_bias=self.use_bias,
                   kernel_initializer=initializers.serialize(self.kernel_initializer),
                   bias_initializer=initializers.serialize(self.bias_initializer),
                   kernel_regularizer=regularizers.serialize(self.kernel_regularizer),
                   bias_regularizer=regularizers.serialize(self.bias_regularizer))
        return cfg


class DenseFlipout(DenseReparameterization):
    """Bayesian dense layer estimated via Flipout (dense.py:227-285):
    y = act(x·mean + ((x∘S)·Delta)∘T + b), Delta = sigma∘eps, per-example sign tensors S [.., I] and T [.., O]."""

    def presample(self, after=None):
        """Optional scheduling hint (fused bf16 path only): draw this call's mean / perturbation operands and the KL NOW on
        a side stream, beside whatever the main stream is doing.  Same Philox stream as the un-hinted call."""
        if not self.built or not _is_trainable(self.kernel_initializer) or self.compute_dtype != torch.bfloat16:
            return
        mean, rho = self._kernel_params()
        if rho is None or "eps" in self._injected:
            return
        from .. import _lib

        kl_scale, pm, ps = self._fused_kl_args()
        dev = mean.device
        I, O = mean.shape
        if not hasattr(self, "_side_stream"):
            self._side_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        mu_lp = _lib.empty_padded(I, O, self.compute_dtype, dev)
        delta_lp = _lib.empty_padded(I, O, self.compute_dtype, dev)
        kl64 = torch.zeros((), dtype=torch.float64, device=dev) if kl_scale != 0 else None
        self._side_stream.wait_stream(main)
        if after is not None and getattr(after, "_side_stream", None) is not None:
            self._side_stream.wait_stream(after._side_stream)
        with torch.cuda.stream(self._side_stream):
            nz = self._noise("eps", S_KERNEL)
            _lib.check(_lib.load().ed2_sample_normal_weights(
                mean.data_ptr(), rho.data_ptr(), nz.c(), I, O, 1, delta_lp.data_ptr(), _lib.dt_of(delta_lp),
                _lib.ld(delta_lp), mu_lp.data_ptr(), _lib.ld(mu_lp), None if kl64 is None else kl64.data_ptr(),
                kl_scale, pm, ps, _lib.stream_ptr()), "ed2_sample_normal_weights")
        self._presampled = (mu_lp, delta_lp, kl64, nz)

    def feeds(self, consumer):
        """Scheduling hint for stacked Flipout layers: this layer's output goes straight into `consumer` (another
        DenseFlipout).  This layer's GEMM epilogue then also writes the consumer's x∘S — signs from the consumer's own
        sign_input stream — and the consumer's dX epilogue runs this layer's output side (activation mask, sign_output,
        bias gradient), so neither costs a pass over HBM (functional.FlipoutLink).  Results are unchanged; any other use
        of the output falls back.  Returns `consumer` so chains read l1.feeds(l2).feeds(l3)."""
        import weakref

        self._consumer = weakref.ref(consumer) if consumer is not None else None
        return consumer

    def _next_sign(self, rows):
        ref = getattr(self, "_consumer", None)
        nxt = ref() if ref is not None else None
        if (not isinstance(nxt, DenseFlipout) or not nxt.built or nxt.compute_dtype != self.compute_dtype
                or "sign_input" in nxt._injected or nxt._kernel_params()[0].shape[0] != self.units):
            return None
        return nxt._noise("sign_input", S_SIGN_IN, per_example=rows)

    def call(self, inputs, training=None):
        mean, rho = self._kernel_params()
        if rho is None:  # dense.py:256-257: not a random variable -> parent behaviour
            return super().call(inputs)
        in_link = getattr(inputs, "_ed2_flipout_link", None)
        x, lead = flatten_batch(inputs)
        bias = self._bias_value()
        kl_scale, pm, ps = self._fused_kl_args()
        eps = self._noise("eps", S_KERNEL)
        sign_in = self._noise("sign_input", S_SIGN_IN, per_example=x.shape[0])
        sign_out = self._noise("sign_output", S_SIGN_OUT, per_example=x.shape[0])
        pre = getattr(self, "_presampled", None)
        self._presampled = None
        if (eps.injected is None and not _is_trainable(self.bias_initializer)
                and F.flipout_fusable(x, mean, self.compute_dtype, sign_in, sign_out)):
            chain = len(lead) == 1
            out_link = F.FlipoutLink() if chain else None
            presampled = None
            if pre is not None:
                torch.cuda.current_stream().wait_stream(self._side_stream)
                presampled, eps = pre[:3], pre[3]
            y, kl = F.FlipoutDenseFused.apply(x, mean, rho, bias, self._act, eps, sign_in, sign_out, kl_scale, pm, ps,
                                              in_link if chain else None, out_link,
                                              self._next_sign(x.shape[0]) if chain else None, presampled)
            if kl_scale != 0.0:
                self._kl.set(kl, mean, rho)
            if chain:
                y._ed2_flipout_link = out_link
                return y
            return y.reshape(*lead, self.units)
        if pre is not None:  # a hint that cannot be honoured on the two-GEMM path: wait for it, then draw as usual
            torch.cuda.current_stream().wait_stream(self._side_stream)
        y, kl = F.FlipoutDense.apply(x, mean, rho, bias, self._act, self.compute_dtype, eps, sign_in, sign_out,
                                     kl_scale, pm, ps)
        if kl_scale != 0.0:
            self._kl.set(kl, mean, rho)
        return y.reshape(*lead, self.units)


class DenseBatchEnsemble(Layer):
    """A batch ensemble dense layer (dense.py:470-607): y[e] = act(((x[e]∘alpha[e]) W)∘gamma[e] + bias[e]),
    batch rows member-major [ensemble_size * examples_per_model, features]."""

    def __init__(self, units, rank=1, ensemble_size=4, activation=None, use_bias=True, alpha_initializer="ones",
                 gamma_initializer="ones", kernel_initializer="glorot_uniform", bias_initializer="zeros",
                 kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None, kernel_constraint=None,
                 bias_constraint=None, **kwargs):
        super().__init__(**kwargs)
        if int(rank) < 1:
            raise ValueError("rank must be at least 1")
        self.units, self.rank, self.ensemble_size = int(units), int(rank), int(ensemble_size)
        self.ensemble_activation = activation
        self._act = ops.act_id(activation)
        if self.rank > 1 and self._act not in (ops.ACT_NONE, ops.ACT_RELU):
            raise NotImplementedError("rank > 1 fast weights take activation None or 'relu'")
        self.use_ensemble_bias = use_bias
        self.alpha_initializer = initializers.get(alpha_initializer)
        self.gamma_initializer = initializers.get(gamma_initializer)
        self.kernel_initializer = initializers.get(kernel_initializer)
        self.ensemble_bias_initializer = initializers.get(bias_initializer)
        self.kernel_regularizer = regularizers.get(kernel_regularizer)
        self.ensemble_bias_regularizer = regularizers.get(bias_regularizer)

    def build(self, input_shape):
        dev = self._device
        in_dim, E = input_shape[-1], self.ensemble_size
        self.kernel = nn.Parameter(self.kernel_initializer((in_dim, self.units)).to(dev))
        lead = (self.rank,) if self.rank > 1 else ()   # dense.py:518-523: [rank, ensemble_size, dim] when rank > 1
        self.alpha = nn.Parameter(self.alpha_initializer(lead + (E, in_dim)).to(dev))
        self.gamma = nn.Parameter(self.gamma_initializer(lead + (E, self.units)).to(dev))
        self.ensemble_bias = (nn.Parameter(self.ensemble_bias_initializer((E, self.units)).to(dev))
                              if self.use_ensemble_bias else None)
        self.built = True

    def call(self, inputs, training=None):
        if inputs.dim() != 2:
            raise ValueError("DenseBatchEnsemble (TF signature) takes [ensemble_size * examples_per_model, features]")
        if inputs.shape[0] % self.ensemble_size != 0:
            raise ValueError(f"batch size {inputs.shape[0]} is not divisible by ensemble_size {self.ensemble_size}")
        if self.rank > 1:
            return self._call_rank_r(inputs)
        return F.BatchEnsembleDense.apply(inputs, self.kernel, self.alpha, self.gamma, self.ensemble_bias, self._act,
                                          self.ensemble_size, self.compute_dtype)

    def _call_rank_r(self, inputs):
        """dense.py:560-570: y = act(sum_r ((x∘alpha[r, e]) W)∘gamma[r, e] + bias[e]).  The rank axis is stacked onto the
        batch — ONE BatchEnsemble launch over rank·ensemble_size members ([rank·B, I] rows, member (r, e) owns rows
        r·B + e·n .. + n, exactly the reference's tile / reshape) — and the sum over rank, the bias and the activation
        follow as elementwise tensor ops (this branch is off the measured path)."""
        R, E = self.rank, self.ensemble_size
        B, I = inputs.shape
        stacked = inputs.unsqueeze(0).expand(R, B, I).reshape(R * B, I)
        y = F.BatchEnsembleDense.apply(stacked, self.kernel, self.alpha.reshape(R * E, I),
                                       self.gamma.reshape(R * E, self.units), None, ops.ACT_NONE, R * E, self.compute_dtype)
        y = y.view(R, B, self.units).float().sum(0)
        if self.ensemble_bias is not None:
            y = (y.view(E, B // E, self.units) + self.ensemble_bias[:, None, :]).view(B, self.units)
        if self._act == ops.ACT_RELU:
            y = torch.relu(y)
        return y.to(self.compute_dtype)

    @property
    def losses(self):
        out = []
        if self.kernel_regularizer is not None:
            out.append(self.kernel_regularizer(self.kernel))
        if self.ensemble_bias_regularizer is not None and self.ensemble_bias is not None:
            out.append(self.ensemble_bias_regularizer(self.ensemble_bias))
        return out

    def get_config(self):
        cfg = super().get_config()
        cfg.update(units=self.units, rank=self.rank, ensemble_size=self.ensemble_size,
                   ensemble_activation=self.ensemble_activation, use_ensemble_bias=self.use_ensemble_bias,
                   alpha_initializer=initializers.serialize(self.alpha_initializer),
                   gamma_initializer=initializers.serialize(self.gamma_initializer),
                   ensemble_bias_initializer=initializers.serialize(self.ensemble_bias_initializer),
                   ensemble_bias_regularizer=regularizers.serialize(self.ensemble_bias_regularizer))
        return cfg


class DenseRank1(Layer):
    """A rank-1 Bayesian neural net dense layer (dense.py:991-1175): per-example fast weights
    r = clip(mu_r[e] + sigma_r[e]∘eps_r, min, max), s likewise; y = act(((x∘r) W)∘s + bias[e])."""

    def __init__(self, units, activation=None, use_bias=True, alpha_initializer="trainable_normal",
                 gamma_initializer="trainable_normal", kernel_initializer="glorot_uniform", bias_initializer="zeros",
                 alpha_regularizer="normal_kl_divergence", gamma_regularizer="normal_kl_divergence",
                 kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None, alpha_constraint=None,
                 gamma_constraint=None, kernel_constraint=None, bias_constraint=None, use_additive_perturbation=False,
                 min_perturbation_value=-10, max_perturbation_value=10, ensemble_size=1, **kwargs):
        super().__init__(**kwargs)
        if not float(min_perturbation_value) < float(max_perturbation_value):
            raise ValueError("min_perturbation_value must be below max_perturbation_value")
        self.use_additive_perturbation = bool(use_additive_perturbation)
        self.min_perturbation_value = min_perturbation_value
        self.max_perturbation_value = max_perturbation_value
        self.units, self.ensemble_size = int(units), int(ensemble_size)
        self.ensemble_activation = activation
        self._act = ops.act_id(activation)
        self.use_ensemble_bias = use_bias
        self.alpha_initializer = initializers.get(alpha_initializer)
        self.gamma_initializer = initializers.get(gamma_initializer)
        self.kernel_initializer = initializers.get(kernel_initializer)
        self.ensemble_bias_initializer = initializers.get(bias_initializer)
        self.alpha_regularizer = regularizers.get(alpha_regularizer)
        self.gamma_regularizer = regularizers.get(gamma_regularizer)
        self.kernel_regularizer = regularizers.get(kernel_regularizer)
        self.ensemble_bias_regularizer = regularizers.get(bias_regularizer)

    def build(self, input_shape):
        dev = self._device
        in_dim, E = input_shape[-1], self.ensemble_size
        self.kernel = nn.Parameter(self.kernel_initializer((in_dim, self.units)).to(dev))
        self.alpha = _make_weight(self, "alpha", (E, in_dim), self.alpha_initializer, dev)
        self.gamma = _make_weight(self, "gamma", (E, self.units), self.gamma_initializer, dev)
        self.ensemble_bias = (_make_weight(self, "ensemble_bias", (E, self.units), self.ensemble_bias_initializer, dev)
                              if self.use_ensemble_bias else None)
        self.built = True

    def _fast(self, init, param):
        return init.params() if _is_trainable(init) else (param, None)

    def feeds(self, consumer):
        """Scheduling hint for stacked rank-1 layers: this layer's output goes straight into `consumer` (another
        DenseRank1 with the same ensemble_size).  The GEMM epilogue of this layer then also writes the consumer's
        x∘r — drawn from the consumer's own Philox stream — and the consumer's backward runs this layer's output side in
        its dX epilogue, so neither costs a pass over HBM (functional.Rank1Link).  Results are unchanged; any other use
        of the output falls back to the unfused kernels.  Returns `consumer` so chains read l1.feeds(l2).feeds(l3)."""
        import weakref

        self._consumer = weakref.ref(consumer) if consumer is not None else None
        return consumer

    def precast(self, after=None, ready=None):
        """Optional scheduling hint: make the compute-dtype copy of the kernel NOW on a side stream (it depends on the
        parameter only), so the bandwidth-bound cast runs beside the previous layers' GEMMs instead of preceding this
        layer's.  `after`: a layer whose precast must finish first.  `ready`: a CUDA event after which the parameter
        holds its current value (e.g. recorded at the start of the step); without it the copy waits for everything
        queued on the current stream so far.  The next `call` uses the copy."""
        if not self.built or self.compute_dtype == torch.float32:
            return
        dev = self.kernel.device
        if not hasattr(self, "_cast_stream"):
            self._cast_stream = torch.cuda.Stream(device=dev)
        if ready is not None:
            self._cast_stream.wait_event(ready)
        else:
            self._cast_stream.wait_stream(torch.cuda.current_stream(dev))
        if after is not None and getattr(after, "_cast_stream", None) is not None:
            self._cast_stream.wait_stream(after._cast_stream)
        with torch.cuda.stream(self._cast_stream):
            self._precast = (ops.cast(self.kernel.detach(), self.compute_dtype), self.kernel._version)

    def gemm_ready(self):
        """A CUDA event that the next `call` records right before its GEMM is enqueued (fused bf16 path).  Side-stream work
        waiting on it — typically `next_layer.precast(ready=layer.gemm_ready())` — becomes runnable together with the GEMM,
        whose CTAs (higher-priority stream) are placed first; the side work then runs on the SMs the GEMM grid leaves free
        instead of occupying every SM before the GEMM can start."""
        ev = torch.cuda.Event()
        ev.record()  # materialises the handle; re-recorded by the library
        self._ready_ev = ev
        return ev

    def _gemm_ready_event(self):
        ev, self._ready_ev = getattr(self, "_ready_ev", None), None
        return ev

    def _next_fast(self, per):
        ref = getattr(self, "_consumer", None)
        nxt = ref() if ref is not None else None
        if (not isinstance(nxt, DenseRank1) or not nxt.built or nxt.ensemble_size != self.ensemble_size
                or nxt.use_additive_perturbation or nxt.compute_dtype != self.compute_dtype
                or nxt.kernel.shape[0] != self.units):
            return None
        mu_r, rho_r = nxt._fast(nxt.alpha_initializer, nxt.alpha)
        return mu_r, rho_r, nxt._noise("eps_alpha", S_R, per_example=per)

    def call(self, inputs, training=None):
        if inputs.dim() != 2:  # dense.py:1100: restricted to 2-D inputs
            raise ValueError("DenseRank1 takes [ensemble_size * examples_per_model, features]")
        if inputs.shape[0] % self.ensemble_size != 0:
            raise ValueError(f"batch size {inputs.shape[0]} is not divisible by ensemble_size {self.ensemble_size}")
        mu_r, rho_r = self._fast(self.alpha_initializer, self.alpha)
        mu_s, rho_s = self._fast(self.gamma_initializer, self.gamma)
        bias, rho_b = (self._fast(self.ensemble_bias_initializer, self.ensemble_bias)
                       if self.use_ensemble_bias else (None, None))
        per = inputs.shape[0] // self.ensemble_size
        in_link = getattr(inputs, "_ed2_rank1_link", None)
        out_link = F.Rank1Link()
        pre, self._precast = getattr(self, "_precast", None), None
        w_pre = None
        if pre is not None:
            torch.cuda.current_stream().wait_stream(self._cast_stream)
            if pre[1] == self.kernel._version:  # the parameter was not updated since the copy was made
                w_pre = pre[0]
        y = F.Rank1Dense.apply(inputs, self.kernel, mu_r, rho_r, mu_s, rho_s, bias, self._act,
                               self.ensemble_size, self.compute_dtype,
                               self._noise("eps_alpha", S_R, per_example=per),
                               self._noise("eps_gamma", S_S, per_example=per),
                               self.use_additive_perturbation,
                               (self.min_perturbation_value, self.max_perturbation_value), rho_b,
                               self._noise("eps_bias", S_B, per_example=per) if rho_b is not None else None,
                               in_link, out_link, self._next_fast(per), w_pre, self._gemm_ready_event())
        y._ed2_rank1_link = out_link
        return y

    @property
    def losses(self):
        out = []
        if self.alpha_regularizer is not None and _is_trainable(self.alpha_initializer):
            out.append(self.alpha_regularizer(self.alpha_initializer))
        if self.gamma_regularizer is not None and _is_trainable(self.gamma_initializer):
            out.append(self.gamma_regularizer(self.gamma_initializer))
        if self.kernel_regularizer is not None:
            out.append(self.kernel_regularizer(self.kernel))
        if self.ensemble_bias_regularizer is not None and self.use_ensemble_bias:
            out.append(self.ensemble_bias_regularizer(
                self.ensemble_bias_initializer if _is_trainable(self.ensemble_bias_initializer) else self.ensemble_bias))
        return out

    def get_config(self):
        cfg = super().get_config()
        cfg.update(units=self.units, ensemble_size=self.ensemble_size, ensemble_activation=self.ensemble_activation,
                   use_ensemble_bias=self.use_ensemble_bias,
                   alpha_initializer=initializers.serialize(self.alpha_initializer),
                   gamma_initializer=initializers.serialize(self.gamma_initializer),
                   ensemble_bias_initializer=initializers.serialize(self.ensemble_bias_initializer),
                   alpha_regularizer=regularizers.serialize(self.alpha_regularizer),
                   gamma_regularizer=regularizers.serialize(self.gamma_regularizer),
                   ensemble_bias_regularizer=regularizers.serialize(self.ensemble_bias_regularizer),
                   use_additive_perturbation=self.use_additive_perturbation,
                   min_perturbation_value=self.min_perturbation_value,
                   max_perturbation_value=self.max_perturbation_value)
        return cfg
