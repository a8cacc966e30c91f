"""torch-CPU fp32 restatement of the reference's CPU path for the benchmark workloads (bench.py `cpu_baseline` and
`--impl reference`).  TEST / BASELINE INFRASTRUCTURE — never imported by `edward2_b200/`.

The reference itself cannot run here or on the GPU box (TensorFlow / TFP / JAX / Flax are not installed, no network;
SURVEY.md §0.4, §8c), so this is labelled a *port* ("restatement of the reference's CPU path"), never "the reference".
It follows the reference op-for-op — including what the reference materialises: the sampled kernel, the separate sign
tensors, the regulariser evaluated on its own — with autograd for the backward pass, fp32, all host threads.
"""
import os

import torch


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def _softplus(x):
    return torch.nn.functional.softplus(x) + 1e-7


def _kl(mu, rho, scale):
    """edward2/tensorflow/regularizers.py:175-186 via TFP Normal.kl_divergence (3p form)."""
    sigma = _softplus(rho)
    return scale * (-torch.log(sigma) + 0.5 * (sigma * sigma + mu * mu) - 0.5).sum()


class BayesianMLP:
    """cfg2: DenseReparameterization / DenseFlipout MLP (edward2/tensorflow/layers/dense.py:73-83, 255-285)."""

    def __init__(self, dims, flipout=False, kl_scale=1.0 / 50000, seed=42):
        g = torch.Generator().manual_seed(seed)
        self.flipout, self.kl_scale = flipout, kl_scale
        self.params = []
        for i, o in zip(dims[:-1], dims[1:]):
            mu = (torch.randn(i, o, generator=g) * (1.0 / i ** 0.5)).requires_grad_(True)
            rho = (torch.randn(i, o, generator=g) * 0.1 - 3.0).requires_grad_(True)
            b = torch.zeros(o, requires_grad=True)
            self.params.append((mu, rho, b))

    def step(self, x, labels):
        h = x
        kl = 0.0
        n = len(self.params)
        for li, (mu, rho, b) in enumerate(self.params):
            eps = torch.randn_like(mu)                       # tfp Normal.sample
            if self.flipout:
                delta = _softplus(rho) * eps                 # kernel - kernel_mean, materialised
                s_in = torch.randint(0, 2, h.shape).to(h.dtype) * 2 - 1
                s_out = torch.randint(0, 2, (h.shape[0], mu.shape[1])).to(h.dtype) * 2 - 1
                z = h @ mu + ((h * s_in) @ delta) * s_out + b
            else:
                w = mu + _softplus(rho) * eps                # sampled kernel, materialised
                z = h @ w + b
            h = torch.relu(z) if li < n - 1 else z
            kl = kl + _kl(mu, rho, self.kl_scale)            # regulariser, evaluated separately
        loss = torch.nn.functional.cross_entropy(h, labels) + kl
        grads = torch.autograd.grad(loss, [p for t in self.params for p in t])
        return loss.detach(), grads


def time_bayesian_mlp(dims, batch, flipout, steps, warmup=1, threads=None):
    """Returns (samples/s, seconds per step, threads)."""
    import time

    threads = threads or host_threads()
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, dims[0], generator=g)
    labels = torch.randint(0, dims[-1], (batch,), generator=g)
    model = BayesianMLP(dims, flipout)
    for _ in range(warmup):
        model.step(x, labels)
    t0 = time.perf_counter()
    for _ in range(steps):
        model.step(x, labels)
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt, threads


# ------------------------------------------------------------------------------------------------------------------
def _time(step, batch, steps, warmup, threads=None):
    import time

    threads = threads or host_threads()
    torch.set_num_threads(threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt, threads


class Rank1MLP:
    """cfg4: DenseRank1 MLP (edward2/tensorflow/layers/dense.py:1096-1144) as the reference computes it:
    `distribution.sample(examples_per_model)` materialises [n, E, I] / [n, E, O] fast-weight samples, clip, transpose to
    member-major, reshape, one matmul per layer; KL of alpha / gamma against the N(1, 0.1) prior evaluated separately."""

    def __init__(self, dims, ensemble_size=4, kl_scale=1.0 / 50000, seed=42):
        g = torch.Generator().manual_seed(seed)
        self.E, self.kl_scale, self.params = ensemble_size, kl_scale, []
        rho0 = float(torch.log(torch.expm1(torch.tensor((1e-3 / (1 - 1e-3)) ** 0.5))))
        for i, o in zip(dims[:-1], dims[1:]):
            w = (torch.randn(i, o, generator=g) * (1.0 / i ** 0.5)).requires_grad_(True)
            sign = lambda shape: (torch.randint(0, 2, shape, generator=g).float() * 2 - 1)  # noqa: E731
            mr, ms = sign((ensemble_size, i)).requires_grad_(True), sign((ensemble_size, o)).requires_grad_(True)
            rr = torch.full((ensemble_size, i), rho0, requires_grad=True)
            rs = torch.full((ensemble_size, o), rho0, requires_grad=True)
            b = torch.zeros(ensemble_size, o, requires_grad=True)
            self.params.append((w, mr, rr, ms, rs, b))

    @staticmethod
    def _kl(mu, rho, scale, pm=1.0, ps=0.1):
        sigma = _softplus(rho)
        return scale * (torch.log(ps / sigma) + (sigma * sigma + (mu - pm) ** 2) / (2 * ps * ps) - 0.5).sum()

    def step(self, x, labels):
        E = self.E
        n = x.shape[0] // E
        h, kl, L = x, 0.0, len(self.params)
        for li, (w, mr, rr, ms, rs, b) in enumerate(self.params):
            hin = h.reshape(E, n, -1)
            alpha = torch.clamp(mr + _softplus(rr) * torch.randn(n, E, mr.shape[1]), -10, 10).transpose(0, 1)
            gamma = torch.clamp(ms + _softplus(rs) * torch.randn(n, E, ms.shape[1]), -10, 10).transpose(0, 1)
            z = ((hin * alpha) @ w) * gamma + b.unsqueeze(1)
            z = z.reshape(E * n, -1)
            h = torch.relu(z) if li < L - 1 else z
            kl = kl + self._kl(mr, rr, self.kl_scale) + self._kl(ms, rs, self.kl_scale)
        loss = torch.nn.functional.cross_entropy(h, labels) + kl
        grads = torch.autograd.grad(loss, [p for t in self.params for p in t])
        return loss.detach(), grads


def time_rank1_mlp(dims, batch, steps, warmup=1, ensemble_size=4, threads=None):
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, dims[0], generator=g)
    labels = torch.randint(0, dims[-1], (batch,), generator=g)
    model = Rank1MLP(dims, ensemble_size)
    return _time(lambda: model.step(x, labels), batch, steps, warmup, threads)


def time_rank1_conv(batch, height, width, channels, n_layers, steps, warmup=1, ensemble_size=4, threads=None):
    """Stacked Conv2DRank1 3x3 'same' layers as the reference computes them (convolutional.py:1968-2020): per-example
    alpha / gamma samples [B, C] broadcast over the spatial axes, x * alpha -> conv2d -> * gamma + bias."""
    g = torch.Generator().manual_seed(1234)
    E, n = ensemble_size, batch // ensemble_size
    x = torch.randn(batch, channels, height, width, generator=g)
    labels = torch.randint(0, channels, (batch,), generator=g)
    rho0 = float(torch.log(torch.expm1(torch.tensor((1e-3 / (1 - 1e-3)) ** 0.5))))
    params = []
    for _ in range(n_layers):
        w = (torch.randn(channels, channels, 3, 3, generator=g) * (1.0 / (9 * channels) ** 0.5)).requires_grad_(True)
        sign = lambda shape: (torch.randint(0, 2, shape, generator=g).float() * 2 - 1)  # noqa: E731
        params.append((w, sign((E, channels)).requires_grad_(True), torch.full((E, channels), rho0, requires_grad=True),
                       sign((E, channels)).requires_grad_(True), torch.full((E, channels), rho0, requires_grad=True),
                       torch.zeros(E, channels, requires_grad=True)))

    def step():
        h, kl = x, 0.0
        for li, (w, mr, rr, ms, rs, b) in enumerate(params):
            rep = lambda t: t.repeat_interleave(n, 0)  # noqa: E731
            alpha = torch.clamp(rep(mr) + rep(_softplus(rr)) * torch.randn(batch, channels), -10, 10)[:, :, None, None]
            gamma = torch.clamp(rep(ms) + rep(_softplus(rs)) * torch.randn(batch, channels), -10, 10)[:, :, None, None]
            z = torch.nn.functional.conv2d(h * alpha, w, padding=1) * gamma + rep(b)[:, :, None, None]
            h = torch.relu(z) if li < n_layers - 1 else z
            kl = kl + Rank1MLP._kl(mr, rr, 1.0 / 50000) + Rank1MLP._kl(ms, rs, 1.0 / 50000)
        loss = torch.nn.functional.cross_entropy(h.mean(dim=(2, 3)), labels) + kl
        return loss.detach(), torch.autograd.grad(loss, [p for t in params for p in t])

    return _time(step, batch, steps, warmup, threads)


def time_hetsngp(in_dim, num_inducing, classes, factors, mc_samples, batch, steps, warmup=1, threads=None):
    """HeteroscedasticSNGPLayer train step as the reference computes it (hetsngp.py:186-240, heteroscedastic.py:178-226,
    749-785): materialised [S, B, F] / [S, B, K] noise, einsum to [B, S, K] latents, softmax, mean over samples."""
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, in_dim, generator=g)
    labels = torch.randint(0, classes, (batch,), generator=g)
    w_rf = torch.randn(in_dim, num_inducing, generator=g)
    b_rf = torch.rand(num_inducing, generator=g) * 6.283185307179586
    mk = lambda o: ((torch.randn(in_dim, o, generator=g) / in_dim ** 0.5).requires_grad_(True), torch.zeros(o, requires_grad=True))  # noqa: E731
    beta = (torch.randn(num_inducing, classes, generator=g) * 0.05).requires_grad_(True)
    (ws, bs), (wd, bd) = mk(classes * factors), mk(classes)
    ln_g, ln_b = torch.ones(in_dim, requires_grad=True), torch.zeros(in_dim, requires_grad=True)
    prec = torch.zeros(num_inducing, num_inducing)
    params = [beta, ws, bs, wd, bd, ln_g, ln_b]

    def step():
        h = torch.nn.functional.layer_norm(x, (in_dim,), ln_g, ln_b, 1e-3)
        phi = torch.cos(h @ w_rf + b_rf) * (2.0 / num_inducing) ** 0.5
        loc = phi @ beta
        with torch.no_grad():
            prec.mul_(0.999).add_(phi.T @ phi, alpha=0.001 / batch)
        fac = (x @ ws + bs).reshape(batch, classes, factors)
        diag = torch.nn.functional.softplus(x @ wd + bd) + 1e-3
        z = torch.randn(mc_samples, batch, factors).transpose(0, 1)
        e = torch.randn(mc_samples, batch, classes).transpose(0, 1)
        lat = loc[:, None, :] + torch.einsum("ijk,iak->iaj", fac, z) + e * diag[:, None, :]
        pm = torch.softmax(lat, -1).mean(1)
        loss = torch.nn.functional.nll_loss(torch.log(torch.clamp(pm, 1e-7, 1.0)), labels)
        return loss.detach(), torch.autograd.grad(loss, params)

    return _time(step, batch, steps, warmup, threads)


class BatchEnsembleMLP:
    """cfg1: DenseBatchEnsemble MLP (dense.py:551-584, rank-1 branch): reshape to [E, n, I], x * alpha, matmul,
    * gamma, + bias, activation, reshape back — each a separate eager op."""

    def __init__(self, dims, ensemble_size=4, seed=42):
        g = torch.Generator().manual_seed(seed)
        self.E, self.params = ensemble_size, []
        for i, o in zip(dims[:-1], dims[1:]):
            lim = (6.0 / (i + o)) ** 0.5
            w = ((torch.rand(i, o, generator=g) * 2 - 1) * lim).requires_grad_(True)
            sign = lambda shape: (torch.randint(0, 2, shape, generator=g).float() * 2 - 1)  # noqa: E731
            self.params.append((w, sign((ensemble_size, i)).requires_grad_(True),
                                sign((ensemble_size, o)).requires_grad_(True),
                                torch.zeros(ensemble_size, o, requires_grad=True)))

    def step(self, x, labels):
        E = self.E
        n = x.shape[0] // E
        h, L = x, len(self.params)
        for li, (w, a, gm, b) in enumerate(self.params):
            z = ((h.reshape(E, n, -1) * a.unsqueeze(1)) @ w) * gm.unsqueeze(1) + b.unsqueeze(1)
            z = z.reshape(E * n, -1)
            h = torch.relu(z) if li < L - 1 else z
        loss = torch.nn.functional.cross_entropy(h, labels)
        return loss.detach(), torch.autograd.grad(loss, [p for t in self.params for p in t])


def time_batch_ensemble_mlp(dims, batch, steps, warmup=2, ensemble_size=4, threads=None):
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, dims[0], generator=g)
    labels = torch.randint(0, dims[-1], (batch,), generator=g)
    model = BatchEnsembleMLP(dims, ensemble_size)
    return _time(lambda: model.step(x, labels), batch, steps, warmup, threads)


class SNGPHead:
    """cfg3 / cfg5: [SpectralNormalization(Dense)]* -> RandomFeatureGaussianProcess with the Laplace covariance
    (edward2/tensorflow/layers/normalization.py:369-391; random_feature.py:247-281, 394-423, 447-485), TF semantics:
    one power iteration per call, LayerNorm, cos features * sqrt(2/D), precision EMA (momentum > 0) or exact sum
    (momentum < 0); eval = explicit matrix inverse (done once, outside the timed step) + Phi Sigma Phi^T diagonal
    + mean-field logits (utils.py:400-424)."""

    def __init__(self, in_dim, backbone, num_inducing, classes, momentum, ridge, seed=42):
        g = torch.Generator().manual_seed(seed)
        self.momentum, self.ridge, self.D = momentum, ridge, num_inducing
        self.backbone = []
        d = in_dim
        for width in backbone:
            w = (torch.randn(d, width, generator=g) / d ** 0.5).requires_grad_(True)
            b = torch.zeros(width, requires_grad=True)
            self.backbone.append([w, b, torch.randn(1, width, generator=g) * 0.05, torch.randn(1, d, generator=g) * 0.05])
            d = width
        self.ln_g, self.ln_b = torch.ones(d, requires_grad=True), torch.zeros(d, requires_grad=True)
        self.W = torch.randn(d, num_inducing, generator=g)
        self.b = torch.rand(num_inducing, generator=g) * 6.283185307179586
        self.beta = (torch.randn(num_inducing, classes, generator=g) * 0.05).requires_grad_(True)
        self.P = torch.zeros(num_inducing, num_inducing)
        self.Sigma = torch.eye(num_inducing) / max(ridge, 1e-6)

    def _spectral(self, layer, training):
        w, _, u, v = layer
        with torch.no_grad():
            if training:  # normalization.py:369-378
                v_hat = torch.nn.functional.normalize(u @ w.t(), dim=-1)
                u_hat = torch.nn.functional.normalize(v_hat @ w, dim=-1)
                layer[2], layer[3] = u_hat, v_hat
            sigma = (layer[3] @ w @ layer[2].t()).reshape(())
            if 0.95 / sigma < 1:
                w.mul_(0.95 / sigma)

    def features(self, x, training):
        h = x
        for layer in self.backbone:
            self._spectral(layer, training)
            h = torch.relu(h @ layer[0] + layer[1])
        h = torch.nn.functional.layer_norm(h, h.shape[-1:], self.ln_g, self.ln_b, 1e-3)
        return torch.cos(h @ self.W + self.b) * (2.0 / self.D) ** 0.5

    def train_step(self, x, labels):
        phi = self.features(x, True)
        logits = phi @ self.beta
        with torch.no_grad():
            pm = phi.t() @ phi  # random_feature.py:406-421
            if self.momentum > 0:
                self.P.mul_(self.momentum).add_(pm / x.shape[0], alpha=1 - self.momentum)
            else:
                self.P.add_(pm)
            cov = torch.eye(x.shape[0]) if x.shape[0] <= 8192 else None  # random_feature.py:531
        loss = torch.nn.functional.cross_entropy(logits, labels)
        params = [p for l in self.backbone for p in l[:2]] + [self.ln_g, self.ln_b, self.beta]
        params = [p for p in params if p.requires_grad]
        return loss.detach(), torch.autograd.grad(loss, params, allow_unused=True), cov

    def invert(self):
        self.Sigma = torch.linalg.inv(self.ridge * torch.eye(self.D) + self.P)  # random_feature.py:447-459

    def eval_step(self, x):
        with torch.no_grad():
            phi = self.features(x, False)
            logits = phi @ self.beta
            var = ((phi @ self.Sigma) * phi).sum(-1) * self.ridge  # diag of random_feature.py:482-485
            return logits / torch.sqrt(1.0 + var.unsqueeze(-1) * (3.141592653589793 / 8.0))  # utils.py:400-424


def time_sngp(in_dim, backbone, num_inducing, classes, momentum, ridge, batch, steps, warmup=1, threads=None):
    """Returns {'train': (sps, dt, threads), 'eval': (...)}."""
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, in_dim, generator=g)
    labels = torch.randint(0, classes, (batch,), generator=g)
    m = SNGPHead(in_dim, backbone, num_inducing, classes, momentum, ridge)
    out = {"train": _time(lambda: m.train_step(x, labels), batch, steps, warmup, threads)}
    out["eval"] = _time(lambda: m.eval_step(x), batch, steps, warmup, threads)
    return out
