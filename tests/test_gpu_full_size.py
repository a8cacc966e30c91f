"""GPU parity at BASELINE.json's FULL sizes, through the oracle on row samples and size-independent properties
(linearity, symmetry, trace == squared Frobenius norm, diag == row-sum form)."""
import numpy as np
import pytest
import torch

from oracle import dense as od, sngp as og
from tests.util import rel_err, to_np

pytestmark = pytest.mark.gpu


def _rows(n, k, seed):
    return np.sort(np.random.default_rng(seed).choice(n, k, replace=False))


def test_cfg1_batch_ensemble_mlp_fp32_full(cuda):
    """cfg1: BatchEnsemble MLP 784-512-512-10, E=4, B=128, fp32 — whole forward + backward vs the oracle (1e-5)."""
    import edward2_b200 as ed

    rng = np.random.default_rng(0)
    E, B = 4, 128
    x = torch.tensor(rng.standard_normal((B, 784)), dtype=torch.float32, device=cuda)
    layers = [ed.layers.DenseBatchEnsemble(u, ensemble_size=E, activation=a, alpha_initializer=ed.initializers.RandomSign(0.5, seed=s),
                                           gamma_initializer=ed.initializers.RandomSign(0.5, seed=s + 10))
              for u, a, s in ((512, "relu", 1), (512, "relu", 2), (10, None, 3))]
    h = x
    for l in layers:
        h = l(h)
    gy = torch.tensor(rng.standard_normal((B, 10)), dtype=torch.float32, device=cuda)
    h.backward(gy)
    torch.cuda.synchronize()
    # oracle chain
    acts, cur = [], to_np(x)
    for l in layers:
        y, _ = od.batch_ensemble_fwd(cur, to_np(l.kernel), to_np(l.alpha), to_np(l.gamma), to_np(l.ensemble_bias), l.ensemble_activation)
        acts.append((cur, y))
        cur = y
    assert rel_err(to_np(h), cur) < 1e-5
    g = to_np(gy)
    for l, (xin, y) in zip(reversed(layers), reversed(acts)):
        r = od.batch_ensemble_bwd(xin, to_np(l.kernel), to_np(l.alpha), to_np(l.gamma), to_np(l.ensemble_bias), l.ensemble_activation, g)
        assert rel_err(to_np(l.kernel.grad), r["dw"]) < 1e-5
        assert rel_err(to_np(l.alpha.grad), r["dalpha"]) < 1e-5
        assert rel_err(to_np(l.gamma.grad), r["dgamma"]) < 1e-5
        g = r["dx"]


def test_cfg2_reparam_layer_full_size(cuda):
    """cfg2 layer 4096 -> 4096 at batch 4096, bf16, in-kernel Philox: oracle on 64 sampled rows (with the host
    restatement of the same noise stream) + linearity in x."""
    from edward2_b200 import functional as F
    from oracle import philox as op

    B, I, O = 4096, 4096, 4096
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(B, I, generator=gen, device=cuda).to(torch.bfloat16)
    mu = (torch.randn(I, O, generator=gen, device=cuda) / 64).requires_grad_(True)
    rho = torch.full((I, O), -3.0, device=cuda, requires_grad=True)
    rnd = F.Randomness(seed=99, stream=0)
    y, kl = F.ReparamDense.apply(x, mu, rho, None, 0, torch.bfloat16, rnd, 1.0, 0.0, 1.0)
    y2, _ = F.ReparamDense.apply(x * 2, mu, rho, None, 0, torch.bfloat16, rnd, 0.0, 0.0, 1.0)
    torch.cuda.synchronize()
    assert rel_err(to_np(y2), 2 * to_np(y)) < 1e-2            # linearity (same sampled kernel)
    rows = _rows(B, 64, 1)
    eps = op.normal(99, 0, I * O).reshape(I, O)
    w = torch.tensor(to_np(mu) + od.stddev(to_np(rho)) * eps, dtype=torch.float32).to(torch.bfloat16).double().numpy()
    assert rel_err(to_np(y)[rows], to_np(x)[rows] @ w) < 2e-2
    kl_ref = od.normal_kl(to_np(mu), to_np(rho))
    assert abs(float(kl.detach()) - kl_ref) <= 1e-5 * kl_ref


def test_cfg4_rank1_layer_full_size(cuda):
    """cfg4 layer 2048 -> 8192 at the per-GPU batch 2048 (E=4), bf16, injected per-example noise: oracle on sampled rows."""
    from edward2_b200 import functional as F

    rng = np.random.default_rng(4)
    E, B, I, O = 4, 2048, 2048, 8192
    x = torch.tensor(rng.standard_normal((B, I)), dtype=torch.float32, device=cuda).to(torch.bfloat16)
    w = torch.tensor(rng.standard_normal((I, O)) / np.sqrt(I), dtype=torch.float32, device=cuda)
    mu_r = torch.tensor(rng.choice([-1.0, 1.0], (E, I)), dtype=torch.float32, device=cuda)
    mu_s = torch.tensor(rng.choice([-1.0, 1.0], (E, O)), dtype=torch.float32, device=cuda)
    rho = float(np.log(np.expm1(np.sqrt(1e-3 / (1 - 1e-3)))))   # experimental/rank1_bnns/utils.py:47-53
    rho_r, rho_s = torch.full((E, I), rho, device=cuda), torch.full((E, O), rho, device=cuda)
    eps_r = torch.tensor(rng.standard_normal((B, I)), dtype=torch.float32, device=cuda)
    eps_s = torch.tensor(rng.standard_normal((B, O)), dtype=torch.float32, device=cuda)
    y = F.Rank1Dense.apply(x, w, mu_r, rho_r, mu_s, rho_s, None, 1, E, torch.bfloat16, F.Randomness(eps_r), F.Randomness(eps_s))
    torch.cuda.synchronize()
    rows = _rows(B, 32, 2)
    wb = to_np(w.to(torch.bfloat16))
    n = B // E
    for r in rows[:8]:
        e = r // n
        rr = np.clip(to_np(mu_r)[e] + od.stddev(to_np(rho_r))[e] * to_np(eps_r)[r], -10, 10)
        ss = np.clip(to_np(mu_s)[e] + od.stddev(to_np(rho_s))[e] * to_np(eps_s)[r], -10, 10)
        ref = np.maximum(((to_np(x)[r] * rr) @ wb) * ss, 0)
        assert rel_err(to_np(y)[r], ref) < 2e-2


def test_cfg3_cfg5_sngp_full_size(cuda):
    """cfg3 (B=8192, d=512, D=1024) in full vs the oracle; cfg5 per-GPU shard (B=8192, d=4096, D=8192) through
    properties: symmetry of the precision, trace(P) == ||Phi||_F^2, variance == row-sum form on sampled rows."""
    from edward2_b200 import ops

    gen = torch.Generator(device="cuda").manual_seed(5)
    # ---- cfg3, fp32-grade path for the features, bf16 for the contractions
    B, d, D = 8192, 512, 1024
    x = torch.randn(B, d, generator=gen, device=cuda)
    w = torch.randn(d, D, generator=gen, device=cuda)
    b = torch.rand(D, generator=gen, device=cuda) * 6.2831853
    xn = ops.input_normalize(x, torch.float32, eps=1e-3)
    phi, _ = ops.rff_fwd(xn, w.t().contiguous(), b, float(np.sqrt(2.0 / D)), torch.float32, False)
    rows = _rows(B, 64, 3)
    phi_ref, _ = og.rff(og.layer_norm(to_np(x)[rows]), to_np(w), to_np(b), np.sqrt(2.0 / D))
    assert rel_err(to_np(phi)[rows], phi_ref) < 1e-4     # fp32 accumulation of a ~sqrt(d)-radian phase
    phib = ops.cast(phi, torch.bfloat16)
    P = torch.zeros(D, D, device=cuda)
    ops.laplace_precision_update(phib, P, 0.999)
    torch.cuda.synchronize()
    pb = to_np(phib)
    assert rel_err(to_np(P), 0.001 * pb.T @ pb / B) < 1e-3
    # ---- cfg5 shard
    B, d, D = 8192, 4096, 8192
    x = torch.randn(B, d, generator=gen, device=cuda).to(torch.bfloat16)
    wt = (torch.randn(D, d, generator=gen, device=cuda) * np.sqrt(2.0 / d)).to(torch.bfloat16)
    b = torch.rand(D, generator=gen, device=cuda) * 6.2831853
    phi, _ = ops.rff_fwd(x, wt, b, float(np.sqrt(2.0 / D)), torch.bfloat16, False)
    P = torch.zeros(D, D, device=cuda)
    ops.laplace_precision_update(phi, P, -1.0)
    torch.cuda.synchronize()
    assert rel_err(to_np(P), to_np(P).T) < 1e-6                       # symmetric
    fro = float((phi.double() ** 2).sum())
    assert abs(float(torch.trace(P.double())) - fro) <= 1e-3 * fro     # checksum: trace(PhiT Phi) == ||Phi||_F^2
    rows = _rows(B, 16, 4)
    pr = to_np(phi)[rows]
    assert rel_err(to_np(phi)[rows], np.sqrt(2.0 / D) * np.cos(to_np(x)[rows] @ to_np(wt).T + to_np(b))) < 2e-2
    sigma = (torch.eye(D, device=cuda) * 0.5 + 0.01).to(torch.bfloat16)
    var, _ = ops.predictive_variance(phi, sigma, 1e-3)
    torch.cuda.synchronize()
    assert rel_err(to_np(var)[rows], og.predictive_variance(pr, to_np(sigma), 1e-3)) < 2e-2
