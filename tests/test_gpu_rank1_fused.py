"""GPU parity of the FUSED DenseRank1 chain (functional.Rank1Link; gemm_sm100.cuh kEpiRank1Fwd / kEpiRank1Bwd): s and
the next layer's r are drawn inside the forward GEMM epilogue, the dX GEMM epilogue applies r, reduces dmu_r / drho_r and
runs the producer layer's output side.  Every layer is checked against the float64 oracle fed with the host
restatement of the same Philox streams, on the kernel's own bf16 activations / incoming gradients (per-layer 2e-2),
and the whole step against the unfused kernels."""
import numpy as np
import pytest
import torch

from oracle import dense as od
from oracle import philox as op
from tests.util import dev, rel_err, to_np

pytestmark = pytest.mark.gpu
GOLDEN, MASK64 = 0x9E3779B97F4A7C15, (1 << 64) - 1


def _build(ed, dims, E, cuda, hint):
    mk = lambda u, act, li: ed.layers.DenseRank1(  # noqa: E731
        u, activation=act, ensemble_size=E, dtype="bfloat16", seed=900 + li,
        kernel_initializer=ed.initializers.RandomNormal(stddev=1.0 / np.sqrt(dims[li]), seed=50 + li),
        alpha_initializer=ed.initializers.TrainableNormal(
            mean_initializer=ed.initializers.RandomSign(seed=10 + li),
            stddev_initializer=ed.initializers.TruncatedNormal(-2.0, 0.2, seed=30 + li)),
        gamma_initializer=ed.initializers.TrainableNormal(
            mean_initializer=ed.initializers.RandomSign(seed=20 + li),
            stddev_initializer=ed.initializers.TruncatedNormal(-2.0, 0.2, seed=40 + li)))
    layers = [mk(u, "relu" if li < len(dims) - 2 else None, li) for li, u in enumerate(dims[1:])]
    with torch.no_grad():
        h = torch.zeros(E * 32, dims[0], dtype=torch.bfloat16, device=cuda)
        for l in layers:
            h = l(h)
    if hint:
        for a, b in zip(layers[:-1], layers[1:]):
            a.feeds(b)
    return layers


def _run(layers, x, gy, shard=None):
    for l in layers:
        l._calls = 1
        l.data_shard = shard
        for p in l.parameters():
            p.grad = None
    hs, h = [], x
    for l in layers:
        h = l(h)
        h.retain_grad()
        hs.append(h)
    h.backward(gy)
    torch.cuda.synchronize()
    return hs


def _check_chain_against_oracle(layers, dims, x, hs, B, tag):
    from edward2_b200.layers import dense as ld

    inp = x
    for li, (l, h) in enumerate(zip(layers, hs)):
        I, O = dims[li], dims[li + 1]
        key = (l.seed + 1 * GOLDEN) & MASK64
        eps_r = op.normal(key, ld.S_R, B * I).reshape(B, I)
        eps_s = op.normal(key, ld.S_S, B * O).reshape(B, O)
        mu_r, rho_r = (to_np(t) for t in l.alpha_initializer.params())
        mu_s, rho_s = (to_np(t) for t in l.gamma_initializer.params())
        w = to_np(l.kernel.detach().to(torch.bfloat16))
        act = "relu" if li < len(layers) - 1 else None
        bias = to_np(l.ensemble_bias)
        y_ref, _, _, _ = od.rank1_fwd(to_np(inp), w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, act)
        assert rel_err(to_np(h), y_ref) <= 2e-2, f"{tag} layer {li} y"
        g = od.rank1_bwd(to_np(inp), w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, act, to_np(h.grad), y_mask=to_np(h))
        got = {"dx": inp.grad, "dw": l.kernel.grad, "dmu_r": l.alpha_initializer.params()[0].grad,
               "drho_r": l.alpha_initializer.params()[1].grad, "dmu_s": l.gamma_initializer.params()[0].grad,
               "drho_s": l.gamma_initializer.params()[1].grad, "dbias": l.ensemble_bias.grad}
        for name, t in got.items():
            e = rel_err(to_np(t), g[name])
            assert e <= 2e-2, f"{tag} layer {li} {name}: rel err {e:.3e}"
        inp = h


@pytest.mark.parametrize("dims,E,n", [([264, 520, 392, 136], 4, 128), ([128, 1000, 256], 2, 128)])
def test_fused_chain_matches_oracle_and_unfused(cuda, dims, E, n):
    import edward2_b200 as ed
    from edward2_b200 import functional as F

    B = E * n
    rng = np.random.default_rng(7)
    x = dev(rng.standard_normal((B, dims[0])), torch.bfloat16, cuda, True)
    gy = dev(rng.standard_normal((B, dims[-1])), torch.bfloat16, cuda)
    layers = _build(ed, dims, E, cuda, hint=True)
    hs = _run(layers, x, gy)
    links = [getattr(h, "_ed2_rank1_link", None) for h in hs]
    assert all(lk is not None for lk in links)
    assert all(lk.fused for lk in links[:-1]), "the consumer's dX epilogue did not run the producer's output side"
    _check_chain_against_oracle(layers, dims, x, hs, B, "fused")
    y_fused = to_np(hs[-1])

    # the same step on the unfused kernels (same Philox streams): same per-layer oracle check
    F.RANK1_FUSION = False
    try:
        x2 = x.detach().clone().requires_grad_(True)
        hs2 = _run(layers, x2, gy)
        assert not any(h._ed2_rank1_link.fused for h in hs2)
        _check_chain_against_oracle(layers, dims, x2, hs2, B, "unfused")
    finally:
        F.RANK1_FUSION = True
    assert rel_err(to_np(hs2[-1]), y_fused) <= 2e-2


def test_fused_chain_without_hint_and_fan_out(cuda):
    """No feeds() hint: the forward scale pass runs, the backward fusion still applies.  With a skip connection the
    producer receives a summed gradient and must fall back to its own output-side pass."""
    import edward2_b200 as ed

    dims, E, n = [136, 264, 264], 2, 128
    B = E * n
    rng = np.random.default_rng(8)
    x = dev(rng.standard_normal((B, dims[0])), torch.bfloat16, cuda, True)
    gy = dev(rng.standard_normal((B, dims[-1])), torch.bfloat16, cuda)
    layers = _build(ed, dims, E, cuda, hint=False)
    hs = _run(layers, x, gy)
    assert hs[0]._ed2_rank1_link.fused
    ref = {id(p): to_np(p.grad) for l in layers for p in l.parameters()}
    # skip connection around layer 2: d h1 = dx_2 + gy
    for l in layers:
        l._calls = 1
        for p in l.parameters():
            p.grad = None
    h1 = layers[0](x)
    out = layers[1](h1) + h1
    out.backward(gy)
    torch.cuda.synchronize()
    assert not h1._ed2_rank1_link.fused
    # layer 2's own parameter gradients are unchanged by the skip; layer 1's differ (they see the extra gy term)
    for p in layers[1].parameters():
        assert rel_err(to_np(p.grad), ref[id(p)]) <= 1e-2
    assert rel_err(to_np(layers[0].kernel.grad), ref[id(layers[0].kernel)]) > 1e-2


@pytest.mark.parametrize("world", [2, 4])
def test_fused_chain_sharded_equals_global(cuda, world):
    """Global-row noise offsets inside the fused epilogues: `world` emulated ranks reproduce the single-device step.
    world = 2 keeps 128 rows per member locally (fused path on both sides: identical draws and per-row arithmetic ->
    bit-equal outputs); world = 4 leaves 64, which run the unfused kernels (s rounded to bf16): bf16 tolerance."""
    import edward2_b200 as ed

    dims, E, n = [136, 264, 136], 2, 256
    B = E * n
    rng = np.random.default_rng(9)
    xg = rng.standard_normal((B, dims[0]))
    gyg = rng.standard_normal((B, dims[-1]))
    layers = _build(ed, dims, E, cuda, hint=True)
    hs = _run(layers, dev(xg, torch.bfloat16, cuda, True), dev(gyg, torch.bfloat16, cuda))
    y_full = to_np(hs[-1])
    g_full = {id(p): to_np(p.grad) for l in layers for p in l.parameters()}
    per = n // world
    ys, acc = [], {}
    for rank in range(world):
        sl = lambda a: a.reshape(E, n, -1)[:, rank * per:(rank + 1) * per].reshape(E * per, -1)  # noqa: E731
        h = _run(layers, dev(sl(xg), torch.bfloat16, cuda, True), dev(sl(gyg), torch.bfloat16, cuda), shard=(rank, world))
        ys.append(to_np(h[-1]).reshape(E, per, -1))
        for l in layers:
            for p in l.parameters():
                acc[id(p)] = acc.get(id(p), 0) + to_np(p.grad)
    y = np.concatenate(ys, axis=1).reshape(B, -1)
    assert rel_err(y, y_full) <= (1e-6 if world == 2 else 2e-2)
    for k, v in acc.items():
        assert rel_err(v, g_full[k]) <= 1e-2  # bf16 GEMM accumulation order differs across the split
