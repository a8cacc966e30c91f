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


def _bf(a):
    return torch.tensor(np.asarray(a), dtype=torch.float32).to(torch.bfloat16).double().numpy()


@pytest.mark.parametrize("flipout", [False, True])
def test_cfg2_layer_full_size_forward_and_backward(cuda, flipout):
    """cfg2 layer 4096 -> 4096 at batch 4096, bf16, in-kernel Philox — forward AND backward against the oracle on
    sampled rows / columns (the oracle is evaluated with the kernel's own ReLU pattern and bf16-rounded operands):
    y[rows], dx[rows], dmu[rows_in, :], drho[rows_in, :], dbias."""
    import edward2_b200 as ed
    from edward2_b200.layers import dense as ld
    from oracle import philox as op

    B, I, O = 4096, 4096, 4096
    cls = ed.layers.DenseFlipout if flipout else ed.layers.DenseReparameterization
    layer = cls(O, activation="relu", dtype="bfloat16", seed=404, kernel_regularizer=None,
                kernel_initializer=ed.initializers.TrainableNormal(
                    mean_initializer=ed.initializers.RandomNormal(stddev=1.0 / 64, seed=1),
                    stddev_initializer=ed.initializers.TruncatedNormal(-4.0, 0.1, seed=2)))
    gen = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randn(B, I, generator=gen, device=cuda).to(torch.bfloat16).requires_grad_(True)
    gy = torch.randn(B, O, generator=gen, device=cuda).to(torch.bfloat16)
    with torch.no_grad():
        layer(x[:8])  # build
    layer._calls = 0
    y = layer(x)
    y.backward(gy)
    torch.cuda.synchronize()
    mu, rho = (to_np(t) for t in layer.kernel_initializer.params())
    key = layer.seed
    eps = op.normal(key, ld.S_KERNEL, I * O).reshape(I, O)
    sigma = od.stddev(rho)
    rows, cin = _rows(B, 48, 7), _rows(I, 48, 8)
    xs, yk, gyn = to_np(x), to_np(y), to_np(gy)
    gp = gyn * (yk > 0)                                   # the kernel's own activation pattern
    bias = to_np(layer.bias)
    if flipout:
        s_in = op.sign(key, ld.S_SIGN_IN, B * I).reshape(B, I)
        s_out = op.sign(key, ld.S_SIGN_OUT, B * O).reshape(B, O)
        mub, dlt = _bf(mu), _bf(sigma * eps)
        y_ref = np.maximum(xs[rows] @ mub + ((xs[rows] * s_in[rows]) @ dlt) * s_out[rows] + bias, 0)
        dx_ref = gp[rows] @ mub.T + ((gp[rows] * s_out[rows]) @ dlt.T) * s_in[rows]
        dmu_ref = xs[:, cin].T @ gp
        ddelta = (xs[:, cin] * s_in[:, cin]).T @ (gp * s_out)
        drho_ref = ddelta * eps[cin] * od.sigmoid(rho[cin])
    else:
        w = _bf(mu + sigma * eps)
        y_ref = np.maximum(xs[rows] @ w + bias, 0)
        dx_ref = gp[rows] @ w.T
        dmu_ref = xs[:, cin].T @ gp
        drho_ref = dmu_ref * eps[cin] * od.sigmoid(rho[cin])
    gmu, grho = layer.kernel_initializer.params()[0].grad, layer.kernel_initializer.params()[1].grad
    assert rel_err(yk[rows], y_ref) < 2e-2
    assert rel_err(to_np(x.grad)[rows], dx_ref) < 2e-2
    assert rel_err(to_np(gmu)[cin], dmu_ref) < 2e-2
    assert rel_err(to_np(grho)[cin], drho_ref) < 2e-2
    assert rel_err(to_np(layer.bias.grad), gp.sum(0)) < 2e-2


def test_cfg4_rank1_chain_full_size_forward_and_backward(cuda):
    """cfg4 layers 2 and 3 (8192 -> 8192 relu -> 1000) at the per-GPU batch 2048, E = 4, bf16, in-kernel Philox with
    the fused epilogues (feeds hint): forward and backward against the oracle on sampled rows / columns."""
    import edward2_b200 as ed
    from edward2_b200.layers import dense as ld
    from oracle import philox as op

    E, B, I, H, O = 4, 2048, 8192, 8192, 1000
    n = B // E
    mk = lambda u, act, s: ed.layers.DenseRank1(  # noqa: E731
        u, activation=act, ensemble_size=E, dtype="bfloat16", seed=500 + s,
        kernel_initializer=ed.initializers.RandomNormal(stddev=1.0 / 90, seed=s),
        alpha_initializer=ed.initializers.TrainableNormal(mean_initializer=ed.initializers.RandomSign(seed=10 + s),
                                                          stddev_initializer=ed.initializers.Constant(-2.5)),
        gamma_initializer=ed.initializers.TrainableNormal(mean_initializer=ed.initializers.RandomSign(seed=20 + s),
                                                          stddev_initializer=ed.initializers.Constant(-2.5)))
    l2, l3 = mk(H, "relu", 1), mk(O, None, 2)
    gen = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(B, I, generator=gen, device=cuda).abs().to(torch.bfloat16).requires_grad_(True)
    gy = torch.randn(B, O, generator=gen, device=cuda).to(torch.bfloat16)
    with torch.no_grad():
        l3(l2(x[:E * 8]))
    l2.feeds(l3)
    l2._calls = l3._calls = 0
    h = l2(x)
    h.retain_grad()
    y = l3(h)
    y.backward(gy)
    torch.cuda.synchronize()
    assert h._ed2_rank1_link.fused
    rows = _rows(B, 24, 11)
    for layer, inp, out, gin, nin, nout, act in ((l3, h, y, gy, H, O, None), (l2, x, h, h.grad, I, H, "relu")):
        key = layer.seed
        cin = _rows(nin, 24, 12)
        eps_r = op.normal(key, ld.S_R, B * nin).reshape(B, nin)
        eps_s = op.normal(key, ld.S_S, B * nout).reshape(B, nout)
        mu_r, rho_r = (to_np(t) for t in layer.alpha_initializer.params())
        mu_s, rho_s = (to_np(t) for t in layer.gamma_initializer.params())
        wb = to_np(layer.kernel.detach().to(torch.bfloat16))
        member = np.arange(B) // n
        r = np.clip(mu_r[member] + od.stddev(rho_r)[member] * eps_r, -10, 10)
        s = np.clip(mu_s[member] + od.stddev(rho_s)[member] * eps_s, -10, 10)
        xin, yk, g = to_np(inp), to_np(out), to_np(gin)
        xr = _bf(xin * r)
        bias = to_np(layer.ensemble_bias)[member]
        pre = (xr[rows] @ wb) * s[rows] + bias[rows]
        assert rel_err(yk[rows], np.maximum(pre, 0) if act else pre) < 2e-2
        gp = g * (yk > 0) if act else g
        dz = _bf(gp * s)
        assert rel_err(to_np(layer.kernel.grad)[cin], xr[:, cin].T @ dz) < 2e-2
        dxr = dz[rows] @ wb.T
        assert rel_err(to_np(inp.grad)[rows], dxr * r[rows]) < 2e-2
        # fast-weight gradients on the sampled input columns: needs dxr for ALL rows on those columns
        dxr_c = dz @ wb[cin].T
        dr = dxr_c * xin[:, cin]
        dmu_r = dr.reshape(E, n, -1).sum(1)
        drho_r = (dr * eps_r[:, cin]).reshape(E, n, -1).sum(1) * od.sigmoid(rho_r[:, cin])
        assert rel_err(to_np(layer.alpha_initializer.params()[0].grad)[:, cin], dmu_r) < 2e-2
        assert rel_err(to_np(layer.alpha_initializer.params()[1].grad)[:, cin], drho_r) < 2e-2
        z = xr @ wb[:, :256]                                   # output side on the first 256 output columns
        ds = gp[:, :256] * _bf(z)
        assert rel_err(to_np(layer.gamma_initializer.params()[0].grad)[:, :256], ds.reshape(E, n, -1).sum(1)) < 2e-2
        assert rel_err(to_np(layer.gamma_initializer.params()[1].grad)[:, :256],
                       (ds * eps_s[:, :256]).reshape(E, n, -1).sum(1) * od.sigmoid(rho_s[:, :256])) < 2e-2
        assert rel_err(to_np(layer.ensemble_bias.grad), gp.reshape(E, n, -1).sum(1)) < 2e-2
