"""GPU parity: layer-level C entry points (through torch.autograd bindings) vs the float64 oracle on identical,
injected random tensors.  Tolerances: 1e-5 (fp32), 2e-2 (bf16), tensor-scale relative error."""
import numpy as np
import pytest
import torch

from oracle import dense as od
from tests.util import TOL, dev, rel_err, round_to, to_np

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.bfloat16]
ACTS = {None: 0, "relu": 1}


def _mask(y, dtype):
    """fp32: the oracle's own relu pattern; bf16: the kernel's (see oracle/dense.py *_bwd y_mask)."""
    return None if dtype == torch.float32 else to_np(y)


def _check(name, got, want, tol):
    e = rel_err(to_np(got), want)
    assert e <= tol, f"{name}: rel err {e:.3e} > {tol:.1e}"


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("act", [None, "relu"])
@pytest.mark.parametrize("shape", [(4, 8, 40, 24), (3, 12, 784, 512), (4, 32, 512, 10)])
def test_batch_ensemble(cuda, dtype, act, shape):
    from edward2_b200 import functional as F

    E, n, I, O = shape
    rng = np.random.default_rng(0)
    x = round_to(rng.standard_normal((E * n, I)), dtype)
    w = round_to(rng.standard_normal((I, O)) / np.sqrt(I), dtype)
    alpha = np.float32(rng.standard_normal((E, I)) * 0.5 + 1.0).astype(np.float64)
    gamma = np.float32(rng.standard_normal((E, O)) * 0.5 + 1.0).astype(np.float64)
    bias = np.float32(rng.standard_normal((E, O)) * 0.1).astype(np.float64)
    gy = round_to(rng.standard_normal((E * n, O)), dtype)
    tol = TOL[dtype]

    xt = dev(x, dtype, cuda, True)
    wt = dev(w, torch.float32, cuda, True)
    at, gt, bt = dev(alpha, torch.float32, cuda, True), dev(gamma, torch.float32, cuda, True), dev(bias, torch.float32, cuda, True)
    y = F.BatchEnsembleDense.apply(xt, wt, at, gt, bt, ACTS[act], E, dtype)
    y.backward(dev(gy, dtype, cuda))
    torch.cuda.synchronize()

    y_ref, _ = od.batch_ensemble_fwd(x, w, alpha, gamma, bias, act)
    _check("y", y, y_ref, tol)
    # bf16: relu' is conditioned on the kernel's own activation pattern (elements within rounding noise of 0 flip)
    g = od.batch_ensemble_bwd(x, w, alpha, gamma, bias, act, gy, y_mask=_mask(y, dtype))
    _check("dx", xt.grad, g["dx"], tol)
    _check("dw", wt.grad, g["dw"], tol)
    _check("dalpha", at.grad, g["dalpha"], tol)
    _check("dgamma", gt.grad, g["dgamma"], tol)
    _check("dbias", bt.grad, g["dbias"], tol)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("sampled", [True, False])
@pytest.mark.parametrize("shape", [(4, 8, 40, 24), (2, 16, 264, 136)])
def test_rank1(cuda, dtype, sampled, shape):
    from edward2_b200 import functional as F

    E, n, I, O = shape
    B = E * n
    rng = np.random.default_rng(1)
    x = round_to(rng.standard_normal((B, I)), dtype)
    w = round_to(rng.standard_normal((I, O)) / np.sqrt(I), dtype)
    f32 = lambda a: np.float32(a).astype(np.float64)  # noqa: E731
    mu_r, mu_s = f32(rng.choice([-1.0, 1.0], (E, I))), f32(rng.choice([-1.0, 1.0], (E, O)))
    rho_r = f32(rng.standard_normal((E, I)) * 0.3 - 1.0) if sampled else None
    rho_s = f32(rng.standard_normal((E, O)) * 0.3 - 1.0) if sampled else None
    eps_r, eps_s = f32(rng.standard_normal((B, I))), f32(rng.standard_normal((B, O)))
    # make a few draws clip at +-10 so the clip mask is exercised
    if sampled:
        eps_r[0, :3] = 80.0
        eps_s[1, :3] = -80.0
    bias = f32(rng.standard_normal((E, O)) * 0.1)
    gy = round_to(rng.standard_normal((B, O)), dtype)
    tol = TOL[dtype]

    xt, wt = dev(x, dtype, cuda, True), dev(w, torch.float32, cuda, True)
    mr, ms = dev(mu_r, torch.float32, cuda, True), dev(mu_s, torch.float32, cuda, True)
    rr = dev(rho_r, torch.float32, cuda, True) if sampled else None
    rs = dev(rho_s, torch.float32, cuda, True) if sampled else None
    bt = dev(bias, torch.float32, cuda, True)
    nr = F.Randomness(dev(eps_r, torch.float32, cuda))
    ns = F.Randomness(dev(eps_s, torch.float32, cuda))
    y = F.Rank1Dense.apply(xt, wt, mr, rr, ms, rs, bt, ACTS["relu"], E, dtype, nr, ns)
    y.backward(dev(gy, dtype, cuda))
    torch.cuda.synchronize()

    y_ref, _, _, s_ref = od.rank1_fwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, "relu")
    # in bf16 the kernel stores s rounded to bf16; tolerance covers it
    _check("y", y, y_ref, tol)
    g = od.rank1_bwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, "relu", gy, y_mask=_mask(y, dtype))
    _check("dx", xt.grad, g["dx"], tol)
    _check("dw", wt.grad, g["dw"], tol)
    _check("dmu_r", mr.grad, g["dmu_r"], tol)
    _check("dmu_s", ms.grad, g["dmu_s"], tol)
    _check("dbias", bt.grad, g["dbias"], tol)
    if sampled:
        _check("drho_r", rr.grad, g["drho_r"], tol)
        _check("drho_s", rs.grad, g["drho_s"], tol)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(24, 40, 24), (64, 264, 136), (128, 512, 1000)])
def test_reparam(cuda, dtype, shape):
    from edward2_b200 import functional as F

    B, I, O = shape
    rng = np.random.default_rng(2)
    f32 = lambda a: np.float32(a).astype(np.float64)  # noqa: E731
    x = round_to(rng.standard_normal((B, I)), dtype)
    mu = f32(rng.standard_normal((I, O)) / np.sqrt(I))
    rho = f32(rng.standard_normal((I, O)) * 0.1 - 3.0)
    eps = f32(rng.standard_normal((I, O)))
    bias = f32(rng.standard_normal(O) * 0.1)
    gy = round_to(rng.standard_normal((B, O)), dtype)
    tol = TOL[dtype]
    kl_scale = 1.0 / 50000

    xt = dev(x, dtype, cuda, True)
    mt, rt, bt = dev(mu, torch.float32, cuda, True), dev(rho, torch.float32, cuda, True), dev(bias, torch.float32, cuda, True)
    y, kl = F.ReparamDense.apply(xt, mt, rt, bt, ACTS["relu"], dtype, F.Randomness(dev(eps, torch.float32, cuda)), kl_scale, 0.0, 1.0)
    (y * dev(gy, dtype, cuda)).sum().float().add(kl * 3.0).backward()
    torch.cuda.synchronize()

    if dtype == torch.bfloat16:
        # the kernel rounds the sampled W to bf16 before the GEMM; the oracle sees the same rounded W through
        # an equivalent (mu', eps=0) parameterisation for the forward check
        w_rounded = round_to(mu + od.stddev(rho) * eps, dtype)
        y_ref = od.activation(x @ w_rounded + bias, "relu")
    else:
        y_ref, _ = od.reparam_fwd(x, mu, rho, eps, bias, "relu")
    _check("y", y, y_ref, tol)
    kl_ref = od.normal_kl(mu, rho, 0.0, 1.0, kl_scale)
    assert abs(kl.item() - kl_ref) <= 1e-5 * abs(kl_ref)
    g = od.reparam_bwd(x, mu, rho, eps, bias, "relu", gy, y_mask=_mask(y, dtype))
    kmu, krho = od.normal_kl_grad(mu, rho, 0.0, 1.0, kl_scale * 3.0)
    _check("dx", xt.grad, g["dx"], tol)
    _check("dmu", mt.grad, g["dmu"] + kmu, tol)
    _check("drho", rt.grad, g["drho"] + krho, tol)
    _check("dbias", bt.grad, g["dbias"], tol)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("uniform_sign_out", [False, True])
@pytest.mark.parametrize("shape", [(24, 40, 24), (64, 264, 136)])
def test_flipout(cuda, dtype, uniform_sign_out, shape):
    from edward2_b200 import functional as F

    B, I, O = shape
    rng = np.random.default_rng(3)
    f32 = lambda a: np.float32(a).astype(np.float64)  # noqa: E731
    x = round_to(rng.standard_normal((B, I)), dtype)
    mu = f32(rng.standard_normal((I, O)) / np.sqrt(I))
    rho = f32(rng.standard_normal((I, O)) * 0.1 - 1.0)
    eps = f32(rng.standard_normal((I, O)))
    s_in = f32(rng.choice([-1.0, 1.0], (B, I)))
    # the reference draws sign_output ~ Uniform[-1, 3) for float inputs (dense.py:266-270); +-1 is the intent
    s_out = f32(rng.uniform(-1, 3, (B, O))) if uniform_sign_out else f32(rng.choice([-1.0, 1.0], (B, O)))
    bias = f32(rng.standard_normal(O) * 0.1)
    gy = round_to(rng.standard_normal((B, O)), dtype)
    tol = TOL[dtype]

    xt = dev(x, dtype, cuda, True)
    mt, rt, bt = dev(mu, torch.float32, cuda, True), dev(rho, torch.float32, cuda, True), dev(bias, torch.float32, cuda, True)
    R = F.Randomness
    y, kl = F.FlipoutDense.apply(xt, mt, rt, bt, ACTS["relu"], dtype, R(dev(eps, torch.float32, cuda)),
                                 R(dev(s_in, torch.float32, cuda)), R(dev(s_out, torch.float32, cuda)), 1.0, 0.0, 1.0)
    y.backward(dev(gy, dtype, cuda))
    torch.cuda.synchronize()

    y_ref, _ = od.flipout_fwd(x, mu, rho, eps, s_in, s_out, bias, "relu")
    _check("y", y, y_ref, tol)
    kl_ref = od.normal_kl(mu, rho)
    assert abs(kl.item() - kl_ref) <= 1e-5 * abs(kl_ref)
    g = od.flipout_bwd(x, mu, rho, eps, s_in, s_out, bias, "relu", gy, y_mask=_mask(y, dtype))
    _check("dx", xt.grad, g["dx"], tol)
    _check("dmu", mt.grad, g["dmu"], tol)
    _check("drho", rt.grad, g["drho"], tol)
    _check("dbias", bt.grad, g["dbias"], tol)


def test_philox_stream_invariance(cuda):
    """In-kernel Philox: forward, backward and a different batch tiling regenerate identical noise."""
    from edward2_b200 import functional as F, ops
    from oracle import philox as op

    n = 4096 + 3
    z = ops.philox("normal", n, 2026, 7)
    raw = ops.philox("raw", 64, 2026, 7)
    torch.cuda.synchronize()
    want_raw = op.blocks(2026, 7, 64).reshape(-1)
    assert np.array_equal(raw.cpu().numpy().view(np.uint32), want_raw)
    assert rel_err(z.cpu().numpy(), op.normal(2026, 7, n)) < 1e-5
    s = ops.philox("sign", n, 5, 9)
    assert np.array_equal(s.cpu().numpy(), op.sign(5, 9, n))
    u = ops.philox("uniform", n, 5, 9)
    assert np.abs(u.cpu().numpy() - op.uniform(5, 9, n)).max() < 1e-7

    # reparam layer with in-kernel eps == same layer with the host-restated Philox eps injected
    B, I, O = 32, 72, 40
    rng = np.random.default_rng(0)
    x = torch.tensor(rng.standard_normal((B, I)), dtype=torch.float32, device=cuda)
    mu = torch.tensor(rng.standard_normal((I, O)) * 0.1, dtype=torch.float32, device=cuda, requires_grad=True)
    rho = torch.full((I, O), -1.0, device=cuda, requires_grad=True)
    y1, _ = F.ReparamDense.apply(x, mu, rho, None, 0, torch.float32, F.Randomness(seed=11, stream=3), 0.0, 0.0, 1.0)
    y1.sum().backward()
    g1 = rho.grad.clone()
    rho.grad = None
    eps = torch.tensor(op.normal(11, 3, I * O).reshape(I, O), dtype=torch.float32, device=cuda)
    y2, _ = F.ReparamDense.apply(x, mu, rho, None, 0, torch.float32, F.Randomness(eps), 0.0, 0.0, 1.0)
    y2.sum().backward()
    torch.cuda.synchronize()
    assert rel_err(to_np(y1), to_np(y2)) < 1e-5
    assert rel_err(to_np(g1), to_np(rho.grad)) < 1e-5


def test_reparam_bf16_native_sampler(cuda):
    """bf16 path with the in-kernel Philox sampler (hardware-approximate Box-Muller / softplus, KL on the side stream)
    vs the oracle fed with the host restatement of the same stream."""
    from edward2_b200 import functional as F
    from oracle import philox as op

    B, I, O = 136, 264, 520
    rng = np.random.default_rng(5)
    f32 = lambda a: np.float32(a).astype(np.float64)  # noqa: E731
    x = round_to(rng.standard_normal((B, I)), torch.bfloat16)
    mu = f32(rng.standard_normal((I, O)) / np.sqrt(I))
    rho = f32(rng.standard_normal((I, O)) * 0.5 - 3.0)
    gy = round_to(rng.standard_normal((B, O)), torch.bfloat16)
    eps = op.normal(77, 0, I * O).reshape(I, O)
    xt = dev(x, torch.bfloat16, cuda, True)
    mt, rt = dev(mu, torch.float32, cuda, True), dev(rho, torch.float32, cuda, True)
    y, kl = F.ReparamDense.apply(xt, mt, rt, None, ACTS["relu"], torch.bfloat16, F.Randomness(seed=77, stream=0), 0.5, 0.0, 1.0)
    (y.float() * dev(gy, torch.float32, cuda)).sum().backward()
    torch.cuda.synchronize()
    w_rounded = round_to(mu + od.stddev(rho) * eps, torch.bfloat16)
    y_ref = od.activation(x @ w_rounded, "relu")
    _check("y", y, y_ref, 2e-2)
    kl_ref = od.normal_kl(mu, rho, 0.0, 1.0, 0.5)
    assert abs(kl.item() - kl_ref) <= 1e-5 * abs(kl_ref)
    g = od.reparam_bwd(x, mu, rho, eps, None, "relu", gy, y_mask=to_np(y))
    _check("dmu", mt.grad, g["dmu"], 2e-2)
    _check("drho", rt.grad, g["drho"], 2e-2)
    _check("dx", xt.grad, g["dx"], 2e-2)


def test_rank1_bf16_native_sampler(cuda):
    """bf16 rank-1 layer with the in-kernel Philox sampler (lean kernels: 8 columns per thread, hardware-approximate
    Box-Muller) vs the oracle fed with the host restatement of the same per-example noise streams."""
    from edward2_b200 import functional as F
    from oracle import philox as op

    E, n, I, O = 4, 96, 264, 520
    B = E * n
    rng = np.random.default_rng(6)
    f32 = lambda a: np.float32(a).astype(np.float64)  # noqa: E731
    x = round_to(rng.standard_normal((B, I)), torch.bfloat16)
    w = round_to(rng.standard_normal((I, O)) / np.sqrt(I), torch.bfloat16)
    mu_r, mu_s = f32(rng.choice([-1.0, 1.0], (E, I))), f32(rng.choice([-1.0, 1.0], (E, O)))
    rho_r, rho_s = f32(rng.standard_normal((E, I)) * 0.3 - 1.0), f32(rng.standard_normal((E, O)) * 0.3 - 1.0)
    gy = round_to(rng.standard_normal((B, O)), torch.bfloat16)
    eps_r = op.normal(31, 4, B * I).reshape(B, I)
    eps_s = op.normal(31, 5, B * O).reshape(B, O)
    xt, wt = dev(x, torch.bfloat16, cuda, True), dev(w, torch.float32, cuda, True)
    mr, ms = dev(mu_r, torch.float32, cuda, True), dev(mu_s, torch.float32, cuda, True)
    rr, rs = dev(rho_r, torch.float32, cuda, True), dev(rho_s, torch.float32, cuda, True)
    y = F.Rank1Dense.apply(xt, wt, mr, rr, ms, rs, None, ACTS["relu"], E, torch.bfloat16,
                           F.Randomness(seed=31, stream=4), F.Randomness(seed=31, stream=5))
    y.backward(dev(gy, torch.bfloat16, cuda))
    torch.cuda.synchronize()
    y_ref, _, _, _ = od.rank1_fwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, None, "relu")
    _check("y", y, y_ref, 2e-2)
    g = od.rank1_bwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, None, "relu", gy, y_mask=to_np(y))
    for name, got in (("dx", xt.grad), ("dw", wt.grad), ("dmu_r", mr.grad), ("drho_r", rr.grad), ("dmu_s", ms.grad),
                      ("drho_s", rs.grad)):
        _check(name, got, g[name], 2e-2)
