"""Fused DenseFlipout (one tcgen05 launch with two TMEM accumulators, include/ed2b200.h: ed2_flipout_fused_fwd / _bwd)
against the float64 oracle (oracle/dense.py: flipout_fwd / flipout_bwd, following edward2/tensorflow/layers/dense.py:
255-285) and against the two-GEMM kernels, single layers with ragged tiles and a linked three-layer stack."""
import numpy as np
import pytest
import torch

from oracle import dense as od
from oracle import philox as op
from tests.util import rel_err, to_np

pytestmark = pytest.mark.gpu
TOL = 2e-2  # bf16


def _layer(ed, units, act, seed, std=0.05):
    return ed.layers.DenseFlipout(units, activation=act, dtype="bfloat16", seed=seed, kernel_regularizer=None,
                                  kernel_initializer=ed.initializers.TrainableNormal(
                                      mean_initializer=ed.initializers.RandomNormal(stddev=std, seed=seed + 1),
                                      stddev_initializer=ed.initializers.TruncatedNormal(-3.0, 0.2, seed=seed + 2)),
                                  bias_initializer=ed.initializers.RandomNormal(stddev=0.1, seed=seed + 3))


def _bf(a):
    return torch.tensor(np.asarray(a), dtype=torch.float32).to(torch.bfloat16).double().numpy()


def _oracle_layer(layer, x, gy, y_kernel, act):
    """Oracle forward / backward of one layer on the kernel's own bf16 operands and activation pattern."""
    from edward2_b200.layers import dense as ld

    B, I = x.shape
    mu, rho = (to_np(t) for t in layer.kernel_initializer.params())
    O = mu.shape[1]
    key = layer.seed
    eps = op.normal(key, ld.S_KERNEL, I * O).reshape(I, O)
    s_in = op.sign(key, ld.S_SIGN_IN, B * I).reshape(B, I)
    s_out = op.sign(key, ld.S_SIGN_OUT, B * O).reshape(B, O)
    bias = to_np(layer.bias)
    mub, dlt = _bf(mu), _bf(od.stddev(rho) * eps)
    z = x @ mub + ((x * s_in) @ dlt) * s_out + bias
    y = np.maximum(z, 0) if act == "relu" else z
    gp = gy * (y_kernel > 0) if act == "relu" else gy
    dx = gp @ mub.T + ((gp * s_out) @ dlt.T) * s_in
    dmu = x.T @ gp
    drho = ((x * s_in).T @ (gp * s_out)) * eps * od.sigmoid(rho)
    return dict(y=y, dx=dx, dmu=dmu, drho=drho, dbias=gp.sum(0))


@pytest.mark.parametrize("shape", [(384, 264, 200), (256, 64, 136), (520, 1000, 72)])
@pytest.mark.parametrize("act", ["relu", None])
def test_fused_layer_matches_oracle(cuda, shape, act):
    """Single layer: ragged M / N / K tiles (rows not a multiple of 256, columns not a multiple of 128, K not a
    multiple of 64), forward and backward, in-kernel Philox signs."""
    import edward2_b200 as ed
    from edward2_b200 import functional as F

    B, I, O = shape
    assert F.FLIPOUT_FUSION
    layer = _layer(ed, O, act, seed=77)
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(B, I, generator=gen, device=cuda).to(torch.bfloat16).requires_grad_(True)
    gy = torch.randn(B, O, generator=gen, device=cuda).to(torch.bfloat16)
    with torch.no_grad():
        layer(x[:8])
    layer._calls = 0
    y = layer(x)
    assert getattr(y, "_ed2_flipout_link", None) is not None, "the fused path did not run"
    y.backward(gy)
    torch.cuda.synchronize()
    ref = _oracle_layer(layer, to_np(x), to_np(gy), to_np(y), act)
    gmu, grho = (p.grad for p in layer.kernel_initializer.params())
    assert rel_err(to_np(y), ref["y"]) < TOL
    assert rel_err(to_np(x.grad), ref["dx"]) < TOL
    assert rel_err(to_np(gmu), ref["dmu"]) < TOL
    assert rel_err(to_np(grho), ref["drho"]) < TOL
    assert rel_err(to_np(layer.bias.grad), ref["dbias"]) < TOL


def _run_stack(ed, F, cuda, dims, B, fused, feeds, presample):
    F.FLIPOUT_FUSION = fused
    try:
        layers = [_layer(ed, u, "relu" if i < len(dims) - 2 else None, seed=100 + 10 * i) for i, u in enumerate(dims[1:])]
        gen = torch.Generator(device="cuda").manual_seed(11)
        x = torch.randn(B, dims[0], generator=gen, device=cuda).to(torch.bfloat16).requires_grad_(True)
        gy = torch.randn(B, dims[-1], generator=gen, device=cuda).to(torch.bfloat16)
        with torch.no_grad():
            h = x[:8]
            for l in layers:
                h = l(h)
        for l in layers:
            l._calls = 0
        if feeds:
            for a, b in zip(layers[:-1], layers[1:]):
                a.feeds(b)
        if presample:
            for l in layers:
                l.presample()
        hs = [x]
        for l in layers:
            hs.append(l(hs[-1]))
        links = [getattr(h, "_ed2_flipout_link", None) for h in hs[1:]]
        for h in hs[1:-1]:
            h.retain_grad()
        hs[-1].backward(gy)
        torch.cuda.synchronize()
        return layers, hs, gy, links
    finally:
        F.FLIPOUT_FUSION = True


@pytest.mark.parametrize("presample", [False, True])
def test_fused_stack_matches_oracle_and_two_gemm_path(cuda, presample):
    """Three linked layers (l1.feeds(l2).feeds(l3)): the producer's epilogue writes the consumer's x∘S, the consumer's dX
    epilogue writes the producer's gp / gp∘T / dbias.  Every layer is checked against the oracle evaluated on the
    kernel's own bf16 activations, and the whole stack against the two-GEMM kernels (same Philox streams)."""
    import edward2_b200 as ed
    from edward2_b200 import functional as F

    dims, B = (264, 520, 392, 200), 384
    layers, hs, gy, links = _run_stack(ed, F, cuda, dims, B, fused=True, feeds=True, presample=presample)
    assert links[0].fused and links[1].fused, "the producers did not take the linked backward"
    # per-layer oracle: inputs = the kernel's activations, upstream gradient = what the layer received
    for i, l in enumerate(layers):
        act = "relu" if i < 2 else None
        g_in = to_np(gy) if i == 2 else to_np(hs[i + 1].grad)  # premasked for i < 2 (mask is idempotent)
        ref = _oracle_layer(l, to_np(hs[i]), g_in, to_np(hs[i + 1]), act)
        gmu, grho = (p.grad for p in l.kernel_initializer.params())
        assert rel_err(to_np(hs[i + 1]), ref["y"]) < TOL, f"y{i}"
        assert rel_err(to_np(gmu), ref["dmu"]) < TOL, f"dmu{i}"
        assert rel_err(to_np(grho), ref["drho"]) < TOL, f"drho{i}"
        assert rel_err(to_np(l.bias.grad), ref["dbias"]) < TOL, f"dbias{i}"
        dx_k = to_np(hs[i].grad)
        dx_ref = ref["dx"] * (to_np(hs[i]) > 0) if i > 0 else ref["dx"]  # linked layers return the masked gradient
        assert rel_err(dx_k, dx_ref) < TOL, f"dx{i}"
    # whole stack vs the two-GEMM kernels
    layers_u, hs_u, _, _ = _run_stack(ed, F, cuda, dims, B, fused=False, feeds=False, presample=False)
    assert rel_err(to_np(hs[-1]), to_np(hs_u[-1])) < TOL
    assert rel_err(to_np(hs[0].grad), to_np(hs_u[0].grad)) < TOL
    for l, lu in zip(layers, layers_u):
        for p, pu in zip(l.kernel_initializer.params(), lu.kernel_initializer.params()):
            assert rel_err(to_np(p.grad), to_np(pu.grad)) < TOL
        assert rel_err(to_np(l.bias.grad), to_np(lu.bias.grad)) < TOL


def test_fused_stack_fan_out_falls_back(cuda):
    """A second consumer of a linked activation: autograd hands the producer a SUMMED gradient (a new tensor), the
    link does not match and the producer runs its own output side — gradients equal the two-GEMM path's."""
    import edward2_b200 as ed
    from edward2_b200 import functional as F

    B, dims = 256, (128, 136, 72)

    def run(fused):
        F.FLIPOUT_FUSION = fused
        try:
            l1, l2 = _layer(ed, dims[1], "relu", 300), _layer(ed, dims[2], None, 310)
            gen = torch.Generator(device="cuda").manual_seed(5)
            x = torch.randn(B, dims[0], generator=gen, device=cuda).to(torch.bfloat16).requires_grad_(True)
            with torch.no_grad():
                l2(l1(x[:8]))
            l1._calls = l2._calls = 0
            if fused:
                l1.feeds(l2)
            h = l1(x)
            out = l2(h).float().sum() + (h.float() * 0.5).sum()  # skip branch
            out.backward()
            torch.cuda.synchronize()
            return l1, getattr(h, "_ed2_flipout_link", None), x.grad
        finally:
            F.FLIPOUT_FUSION = True

    l1f, link, dxf = run(True)
    l1u, _, dxu = run(False)
    assert link is not None and not link.fused
    assert rel_err(to_np(dxf), to_np(dxu)) < TOL
    for p, pu in zip(l1f.kernel_initializer.params(), l1u.kernel_initializer.params()):
        assert rel_err(to_np(p.grad), to_np(pu.grad)) < TOL
    assert rel_err(to_np(l1f.bias.grad), to_np(l1u.bias.grad)) < TOL
