"""GPU parity: SNGP-head entry points vs the float64 oracle (oracle/sngp.py) on identical inputs."""
import numpy as np
import pytest
import torch

from oracle import sngp as og
from tests.util import TOL, dev, rel_err, round_to, to_np

pytestmark = pytest.mark.gpu
DTYPES = [torch.float32, torch.bfloat16]


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("shape", [(64, 48), (512, 512), (200, 1000)])
@pytest.mark.parametrize("training", [True, False])
def test_spectral_norm(cuda, mode, shape, training):
    from edward2_b200 import ops

    I, O = shape
    rng = np.random.default_rng(0)
    w = np.float32(rng.standard_normal((I, O))).astype(np.float64)
    u = np.float32(rng.standard_normal(O) * 0.05).astype(np.float64)
    v = np.float32(rng.standard_normal(I) * 0.05).astype(np.float64)
    c = 0.95
    wt, ut, vt = dev(w, torch.float32, cuda), dev(u, torch.float32, cuda), dev(v, torch.float32, cuda)
    out = torch.empty_like(wt)
    lp = torch.empty(I, (O + 7) // 8 * 8, dtype=torch.bfloat16, device=cuda)[:, :O]
    sigma, _ = ops.spectral_norm_update(wt, ut, vt, 3, c, mode, training, out, lp)
    torch.cuda.synchronize()
    fn = og.spectral_norm_tf if mode == 0 else og.spectral_norm_jax
    w_ref, u_ref, v_ref, s_ref = fn(w, u, v, iteration=3, norm_multiplier=c, training=training)
    assert abs(sigma.item() - s_ref) <= 1e-5 * abs(s_ref)
    assert rel_err(to_np(out), w_ref) < 1e-5
    assert rel_err(to_np(ut), u_ref.reshape(-1)) < 1e-5
    assert rel_err(to_np(vt), v_ref.reshape(-1)) < 1e-5
    assert rel_err(to_np(lp), w_ref) < 1e-2


def test_spectral_norm_converges(cuda):
    """normalization_test.py:68-102 / jax normalization_test.py:52-104: after many iterations the spectral norm of the
    normalised kernel (via SVD) is norm_multiplier at atol 1e-2."""
    from edward2_b200 import ops

    rng = np.random.default_rng(1)
    w = torch.tensor(rng.standard_normal((96, 80)) * 0.5, dtype=torch.float32, device=cuda)
    u = torch.tensor(rng.standard_normal(80), dtype=torch.float32, device=cuda)
    v = torch.tensor(rng.standard_normal(96), dtype=torch.float32, device=cuda)
    out = torch.empty_like(w)
    for _ in range(10):
        ops.spectral_norm_update(w, u, v, 100, 0.95, 0, True, out)
    s = np.linalg.svd(to_np(out), compute_uv=False)[0]
    assert abs(s - 0.95) < 1e-2


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("affine", [True, False])
@pytest.mark.parametrize("d", [264, 128, 512, 1024])   # 128 / 256 / 512 / 1024: register-resident backward kernel
def test_layernorm_fwd_bwd(cuda, dtype, affine, d):
    from edward2_b200 import functional as F

    rng = np.random.default_rng(2)
    B = 200 if d == 264 else 3000   # several rows per warp on the register-resident path
    x = round_to(rng.standard_normal((B, d)) * 2 + 0.5, dtype)
    g = np.float32(rng.standard_normal(d) * 0.2 + 1).astype(np.float64) if affine else None
    b = np.float32(rng.standard_normal(d) * 0.2).astype(np.float64) if affine else None
    dy = round_to(rng.standard_normal((B, d)), dtype)
    eps = 1e-3
    xt = dev(x, dtype, cuda, True)
    gt = dev(g, torch.float32, cuda, True) if affine else None
    bt = dev(b, torch.float32, cuda, True) if affine else None
    y = F.LayerNorm.apply(xt, gt, bt, eps, dtype)
    y.backward(dev(dy, dtype, cuda))
    torch.cuda.synchronize()
    assert rel_err(to_np(y), og.layer_norm(x, g, b, eps)) < TOL[dtype]
    # float64 autograd reference for the backward
    x64 = torch.tensor(x, requires_grad=True)
    g64 = torch.tensor(g, requires_grad=True) if affine else None
    b64 = torch.tensor(b, requires_grad=True) if affine else None
    torch.nn.functional.layer_norm(x64, (d,), g64, b64, eps).backward(torch.tensor(dy))
    assert rel_err(to_np(xt.grad), x64.grad.numpy()) < TOL[dtype]
    if affine:
        assert rel_err(to_np(gt.grad), g64.grad.numpy()) < TOL[dtype]
        assert rel_err(to_np(bt.grad), b64.grad.numpy()) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
def test_rff_fwd_bwd(cuda, dtype):
    from edward2_b200 import functional as F, ops

    rng = np.random.default_rng(3)
    B, d, D = 136, 72, 264
    x = round_to(rng.standard_normal((B, d)), dtype)
    w = round_to(rng.standard_normal((d, D)) * np.sqrt(2.0 / d), dtype)  # JAX default scale (benign phase)
    b = np.float32(rng.uniform(0, 2 * np.pi, D)).astype(np.float64)
    scale = float(np.sqrt(2.0 / D))
    dphi = round_to(rng.standard_normal((B, D)), dtype)
    xt = dev(x, dtype, cuda, True)
    wt_ = dev(w, dtype, cuda)
    w_t = ops.cast(wt_.float().t().contiguous(), dtype)
    w_c = ops.cast(wt_.float().contiguous(), dtype)
    bt = dev(b, torch.float32, cuda)
    phi = F.RandomFourierFeatures.apply(xt, w_t, w_c, bt, scale, dtype, dtype)
    phi.backward(dev(dphi, dtype, cuda))
    torch.cuda.synchronize()
    phi_ref, _ = og.rff(x, w, b, scale)
    assert rel_err(to_np(phi), phi_ref) < TOL[dtype]
    assert rel_err(to_np(xt.grad), og.rff_bwd(x, w, b, scale, dphi)) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("momentum", [0.999, -1.0])
def test_precision_update_and_variance(cuda, dtype, momentum):
    from edward2_b200 import ops

    rng = np.random.default_rng(4)
    B, D, K = 328, 136, 10
    phi = round_to(rng.standard_normal((B, D)) * 0.3, dtype)
    p0 = np.float32(np.eye(D) * 0.5 + 0.01).astype(np.float64)
    pt = dev(p0, torch.float32, cuda)
    phit = dev(phi, dtype, cuda)
    ops.laplace_precision_update(phit, pt, momentum)
    # data-parallel form: raw partial + blend must agree with the fused in-place form
    p2 = dev(p0, torch.float32, cuda)
    partial = torch.empty_like(p2)
    ops.laplace_precision_update(phit, p2, momentum, partial_out=partial)
    ops.precision_blend(p2, partial, momentum, B)
    torch.cuda.synchronize()
    p_ref = og.precision_update_tf(p0, phi, momentum)
    assert rel_err(to_np(pt), p_ref) < (1e-5 if dtype == torch.float32 else 1e-3)
    assert rel_err(to_np(p2), p_ref) < (1e-5 if dtype == torch.float32 else 1e-3)

    ridge = 1e-3
    sigma = og.covariance_tf(p_ref, ridge)
    sig_in = round_to(sigma, dtype)
    var, lg = ops.predictive_variance(phit, dev(sig_in, dtype, cuda), ridge,
                                      dev(rng.standard_normal((B, K)), torch.float32, cuda), 0.5, 0)
    torch.cuda.synchronize()
    var_ref = og.predictive_variance(phi, sig_in, ridge)
    assert rel_err(to_np(var), var_ref) < TOL[dtype]
    full = og.predictive_covariance_tf(phi, sig_in, ridge)
    assert rel_err(var_ref, np.diag(full)) < 1e-10


def test_mean_field_logits_known_answers(cuda):
    """utils_test.py:201-241 / jax utils_test.py:28-68."""
    import edward2_b200 as ed

    rng = np.random.RandomState(0)
    logits = rng.randn(10, 12)
    lt = torch.tensor(logits, dtype=torch.float32, device=cuda)
    covmat = torch.diag(torch.full((10,), 1.5, device=cuda))
    a = ed.layers.utils.mean_field_logits(lt, covmat, mean_field_factor=2.0)
    b = ed.layers.utils.mean_field_logits(lt, covmat, mean_field_factor=2.0, likelihood="poisson")
    c = ed.layers.utils.mean_field_logits(lt, covmat=None, mean_field_factor=-1)
    d = ed.layers.utils.mean_field_logits(lt, covmat=None, mean_field_factor=3.0)
    np.testing.assert_allclose(to_np(a), logits / 2.0, atol=1e-4)
    np.testing.assert_allclose(to_np(b), logits * np.exp(1.5), atol=1e-4)
    np.testing.assert_allclose(to_np(c), logits, atol=1e-4)
    np.testing.assert_allclose(to_np(d), logits / 2.0, atol=1e-4)
    with pytest.raises(ValueError):
        ed.layers.utils.mean_field_logits(lt, covmat, likelihood="gaussian")


@pytest.mark.parametrize("nsplit,tol", [(3, 3e-3), (2, 3e-3)])
def test_rff_split_bf16_phase_precision(cuda, nsplit, tol):
    """fp32-grade random-feature phase on the bf16 tensor cores (split operands, K' = 6K / 3K), at the TF defaults where
    the phase is ~sqrt(d) radians.  Split operands remove the operand-rounding error (plain bf16 is off by >1e-2);
    what remains (~2e-3 of max|phi| at d = 1024, the same for 16 and 24 operand bits) is the tensor core's truncating
    fp32 accumulation of ~1000 products into partial sums of ~30 — well inside the bf16 tolerance of 2e-2."""
    from edward2_b200 import ops

    rng = np.random.default_rng(9)
    B, d, D = 264, 1024, 520
    x = np.float32(og.layer_norm(rng.standard_normal((B, d)))).astype(np.float64)        # LayerNorm'd inputs
    w = np.float32(rng.standard_normal((d, D))).astype(np.float64)                        # stddev-1 features
    b = np.float32(rng.uniform(0, 2 * np.pi, D)).astype(np.float64)
    scale = float(np.sqrt(2.0 / D))
    xt, wt, bt = dev(x, torch.float32, cuda), dev(w, torch.float32, cuda), dev(b, torch.float32, cuda)
    phi_ref, pre_ref = og.rff(x, w, b, scale)
    w_split = ops.split_bf16(wt.t().contiguous(), True, nsplit)
    phi, pre = ops.rff_fwd_split(xt, w_split, bt, scale, D, torch.float32, True, nsplit)
    torch.cuda.synchronize()
    err = rel_err(to_np(phi), phi_ref)
    assert err < tol, f"phi rel err {err:.3e}"
    # second epilogue output: the derivative factor -scale * sin(phase) in bf16 (the backward never re-evaluates sin)
    nsin_ref = -scale * np.sin(pre_ref)
    serr = np.abs(to_np(pre) - nsin_ref).max() / scale
    assert pre.dtype == torch.bfloat16 and serr < 1e-2, f"-sin(phase) abs err {serr:.3e} (in units of scale)"
    # plain bf16 operands: visibly wrong phase at these scales (why the split path is the default)
    phi_bf16, _ = ops.rff_fwd(xt.to(torch.bfloat16), ops.cast(wt.t().contiguous(), torch.bfloat16), bt, scale,
                              torch.float32, False)
    torch.cuda.synchronize()
    assert rel_err(to_np(phi_bf16), phi_ref) > 1e-2


@pytest.mark.parametrize("shape", [(8192, 1024), (4096, 520), (2048, 2048)])
@pytest.mark.parametrize("momentum", [0.999, -1.0])
def test_precision_update_syrk_split_k(cuda, shape, momentum):
    """The bf16 precision update computes only the lower triangle (SYRK) — in K slices with fp32 atomics when the
    output has few tiles — and mirrors it: result equals the oracle update, exactly symmetric, over repeated steps."""
    from edward2_b200 import ops

    B, D = shape
    rng = np.random.default_rng(5)
    P = torch.tensor(rng.standard_normal((D, D)), dtype=torch.float32, device=cuda)
    P = (P + P.t()).contiguous()
    ref = to_np(P)
    for step in range(2):
        phi = torch.tensor(rng.standard_normal((B, D)) * 0.05, dtype=torch.float32, device=cuda).to(torch.bfloat16)
        ops.laplace_precision_update(phi, P, momentum, batch_total=B)
        pn = to_np(phi)
        ref = (momentum * ref + (1 - momentum) * (pn.T @ pn) / B) if momentum > 0 else ref + pn.T @ pn
    torch.cuda.synchronize()
    assert rel_err(to_np(P), ref) < 1e-4
    assert torch.equal(P, P.t())


def test_gemm_accumulate_split_k(cuda):
    """ed2_gemm accumulate: out += alpha * A^T B on a narrow output with a long K (split-K atomics) and on a wide one
    (read-modify-write epilogue)."""
    from edward2_b200 import ops

    rng = np.random.default_rng(6)
    for (K, M, N) in ((8192, 512, 512), (4096, 264, 136), (512, 2048, 2048)):
        a = torch.tensor(rng.standard_normal((K, M)), dtype=torch.float32, device=cuda).to(torch.bfloat16)
        b = torch.tensor(rng.standard_normal((K, N)), dtype=torch.float32, device=cuda).to(torch.bfloat16)
        out = torch.tensor(rng.standard_normal((M, N)), dtype=torch.float32, device=cuda)
        want = to_np(out) + 0.5 * to_np(a).T @ to_np(b)
        ops.gemm(a, b, M, N, K, major_a=ops.MAJOR_MN, major_b=ops.MAJOR_MN, out=out, alpha=0.5, accumulate=True)
        torch.cuda.synchronize()
        assert rel_err(to_np(out), want) < 1e-5, (K, M, N)
