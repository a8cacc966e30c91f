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
