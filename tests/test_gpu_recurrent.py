"""LSTMCellReparameterization on the CUDA path (edward2_b200/layers/recurrent.py; include/ed2b200.h:
ed2_lstm_pointwise_fwd / _bwd + the tcgen05 GEMMs) against the float64 oracle (oracle/recurrent.py, following
edward2/tensorflow/layers/recurrent.py:100-183): an unrolled sequence, forward and backpropagation through time, with the
kernel / recurrent-kernel noise drawn once per sequence from the native Philox stream."""
import numpy as np
import pytest
import torch

from oracle import dense as od
from oracle import philox as op
from oracle import recurrent as orc
from tests.util import TOL, rel_err, round_to, to_np

pytestmark = pytest.mark.gpu


def _init(ed, std, seed):
    return ed.initializers.TrainableNormal(mean_initializer=ed.initializers.RandomNormal(stddev=std, seed=seed),
                                           stddev_initializer=ed.initializers.TruncatedNormal(-3.0, 0.2, seed=seed + 1))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_lstm_cell_reparameterization_sequence(cuda, dtype):
    import edward2_b200 as ed
    from edward2_b200.layers import recurrent as lr

    T, B, I, H = 5, 16, 24, 32
    rng = np.random.default_rng(5)
    xs = round_to(rng.standard_normal((T, B, I)), dtype)
    cell = ed.layers.LSTMCellReparameterization(H, dtype="bfloat16" if dtype == torch.bfloat16 else "float32", seed=17,
                                                kernel_initializer=_init(ed, 0.3, 1), recurrent_initializer=_init(ed, 0.3, 3),
                                                kernel_regularizer=None, recurrent_regularizer=None)
    xt = torch.tensor(xs, dtype=torch.float32, device=cuda).to(dtype).requires_grad_(True)
    with torch.no_grad():
        cell(xt[0], [torch.zeros(B, H, device=cuda, dtype=dtype), torch.zeros(B, H, device=cuda)])  # build
    cell._calls = 0
    states = cell.get_initial_state(batch_size=B)      # draws the sequence's weight sample
    hs = []
    for t in range(T):
        h, states = cell(xt[t], states)
        hs.append(h)
    dhs = round_to(rng.standard_normal((T, B, H)), dtype)
    loss = sum((h.float() * torch.tensor(dhs[t], dtype=torch.float32, device=cuda)).sum() for t, h in enumerate(hs))
    loss.backward()
    torch.cuda.synchronize()
    # oracle on the same sample: one (W, U) for the whole sequence
    key = cell.seed  # _calls == 0 at get_initial_state
    mu_w, rho_w = (to_np(p) for p in cell.kernel_initializer.params())
    mu_u, rho_u = (to_np(p) for p in cell.recurrent_initializer.params())
    e_w = op.normal(key, lr.S_LSTM_KERNEL, mu_w.size).reshape(mu_w.shape)
    e_u = op.normal(key, lr.S_LSTM_RECURRENT, mu_u.size).reshape(mu_u.shape)
    W, U = round_to(orc.reparam_kernel(mu_w, rho_w, e_w), dtype), round_to(orc.reparam_kernel(mu_u, rho_u, e_u), dtype)
    b = to_np(cell.bias)
    assert b[H:2 * H].min() == 1.0 and abs(b[:H]).max() == 0.0  # unit_forget_bias
    hs_ref, _, caches = orc.lstm_sequence(xs, np.zeros((B, H)), np.zeros((B, H)), W, U, b)
    g = orc.lstm_sequence_bwd(caches, W, U, dhs)
    tol = TOL[dtype] if dtype == torch.float32 else 3e-2  # bf16: T chained roundings of h
    assert rel_err(np.stack([to_np(h) for h in hs]), hs_ref) < tol
    assert rel_err(to_np(xt.grad), g["dxs"]) < tol
    assert rel_err(to_np(cell.bias.grad), g["db"]) < tol
    pw, pu = cell.kernel_initializer.params(), cell.recurrent_initializer.params()
    assert rel_err(to_np(pw[0].grad), g["dW"]) < tol
    assert rel_err(to_np(pw[1].grad), g["dW"] * e_w * od.sigmoid(rho_w)) < tol
    assert rel_err(to_np(pu[0].grad), g["dU"]) < tol
    assert rel_err(to_np(pu[1].grad), g["dU"] * e_u * od.sigmoid(rho_u)) < tol


def test_lstm_cell_weight_sample_is_per_sequence_and_losses(cuda):
    """The noise is reused across steps and redrawn by get_initial_state / call_weights (recurrent.py:138-161); `.losses`
    holds the two KL terms; unsupported options raise."""
    import edward2_b200 as ed

    B, I, H = 8, 16, 16
    cell = ed.layers.LSTMCellReparameterization(H, seed=3)
    x = torch.randn(B, I, device=cuda)
    s0 = [torch.zeros(B, H, device=cuda), torch.zeros(B, H, device=cuda)]
    with torch.no_grad():
        a, _ = cell(x, s0)
        b_, _ = cell(x, s0)                 # same sample: identical output
        cell.get_initial_state(inputs=x)    # new sample
        c, _ = cell(x, s0)
    assert torch.equal(a, b_) and not torch.equal(a, c)
    assert len(cell.losses) == 2 and all(float(l) >= 0 for l in cell.losses)
    with pytest.raises(NotImplementedError):
        ed.layers.LSTMCellReparameterization(4, activation="relu")
    with pytest.raises(NotImplementedError):
        ed.layers.LSTMCellReparameterization(4, dropout=0.1)


def test_lstm_cell_flipout_sequence(cuda):
    """LSTMCellFlipout (recurrent.py:185-358), fp32: kernel noise and the four sign tensors are drawn once per sequence;
    forward and BPTT against the oracle composed of oracle/dense.py's Flipout algebra and oracle/recurrent.py's gates."""
    import edward2_b200 as ed
    from edward2_b200.layers import recurrent as lr

    T, B, I, H = 4, 12, 20, 16
    rng = np.random.default_rng(7)
    xs = rng.standard_normal((T, B, I)).astype(np.float32).astype(np.float64)
    cell = ed.layers.LSTMCellFlipout(H, seed=23, kernel_initializer=_init(ed, 0.3, 1), recurrent_initializer=_init(ed, 0.3, 3))
    xt = torch.tensor(xs, dtype=torch.float32, device=cuda).requires_grad_(True)
    with torch.no_grad():
        cell(xt[0], [torch.zeros(B, H, device=cuda), torch.zeros(B, H, device=cuda)])
    cell._calls = 0
    states = cell.get_initial_state(batch_size=B)
    hs = []
    for t in range(T):
        h, states = cell(xt[t], states)
        hs.append(h)
    dhs = rng.standard_normal((T, B, H)).astype(np.float32).astype(np.float64)
    sum((h * torch.tensor(dhs[t], dtype=torch.float32, device=cuda)).sum() for t, h in enumerate(hs)).backward()
    torch.cuda.synchronize()
    key = cell.seed
    pw, pu = cell.kernel_initializer.params(), cell.recurrent_initializer.params()
    mu_w, rho_w, mu_u, rho_u = (to_np(p) for p in (*pw, *pu))
    e_w = op.normal(key, lr.S_LSTM_KERNEL, mu_w.size).reshape(mu_w.shape)
    e_u = op.normal(key, lr.S_LSTM_RECURRENT, mu_u.size).reshape(mu_u.shape)
    sg = lambda stream, n: op.sign(key, stream, B * n).reshape(B, n)  # noqa: E731
    s_in, s_out, r_in, r_out = sg(lr.S_LSTM_SIGN_IN, I), sg(lr.S_LSTM_SIGN_OUT, 4 * H), sg(lr.S_LSTM_RSIGN_IN, H), sg(lr.S_LSTM_RSIGN_OUT, 4 * H)
    b = to_np(cell.bias)
    bx = lambda x: od.flipout_fwd(x, mu_w, rho_w, e_w, s_in, s_out, b, None)[0]  # noqa: E731
    bh = lambda h: od.flipout_fwd(h, mu_u, rho_u, e_u, r_in, r_out, None, None)[0]  # noqa: E731
    gx_ = lambda x, dz: od.flipout_bwd(x, mu_w, rho_w, e_w, s_in, s_out, b, None, dz)  # noqa: E731
    gh_ = lambda h, dz: od.flipout_bwd(h, mu_u, rho_u, e_u, r_in, r_out, None, None, dz)  # noqa: E731
    hs_ref, dxs, gx, gh = orc.two_branch_sequence(xs, np.zeros((B, H)), np.zeros((B, H)), bx, bh, gx_, gh_, dhs)
    tol = 2e-5
    assert rel_err(np.stack([to_np(h) for h in hs]), hs_ref) < tol
    assert rel_err(to_np(xt.grad), dxs) < tol
    assert rel_err(to_np(pw[0].grad), gx["dmu"]) < tol and rel_err(to_np(pw[1].grad), gx["drho"]) < tol
    assert rel_err(to_np(pu[0].grad), gh["dmu"]) < tol and rel_err(to_np(pu[1].grad), gh["drho"]) < tol
    assert rel_err(to_np(cell.bias.grad), gx["dbias"]) < tol
    assert len(cell.losses) == 2


def test_lstm_cell_rank1_sequence(cuda):
    """LSTMCellRank1 (recurrent.py:360-830), fp32, ensemble_size 2: per-example alpha / gamma / recurrent alpha / gamma
    drawn once per sequence (unclipped), forward and BPTT against the oracle's rank-1 algebra + gates."""
    import edward2_b200 as ed
    from edward2_b200.layers import recurrent as lr

    T, E, n, I, H = 3, 2, 6, 20, 16
    B = E * n
    rng = np.random.default_rng(8)
    xs = rng.standard_normal((T, B, I)).astype(np.float32).astype(np.float64)
    fast = lambda seed: ed.initializers.TrainableNormal(  # noqa: E731
        mean_initializer=ed.initializers.RandomNormal(1.0, 0.2, seed=seed),
        stddev_initializer=ed.initializers.TruncatedNormal(-2.0, 0.2, seed=seed + 1))
    cell = ed.layers.LSTMCellRank1(H, ensemble_size=E, seed=29, alpha_initializer=fast(1), gamma_initializer=fast(3),
                                   recurrent_alpha_initializer=fast(5), recurrent_gamma_initializer=fast(7),
                                   kernel_initializer=ed.initializers.RandomNormal(stddev=0.3, seed=9),
                                   recurrent_initializer=ed.initializers.RandomNormal(stddev=0.3, seed=10))
    xt = torch.tensor(xs, dtype=torch.float32, device=cuda).requires_grad_(True)
    with torch.no_grad():
        cell(xt[0], [torch.zeros(B, H, device=cuda), torch.zeros(B, H, device=cuda)])
    cell._calls = 0
    states = cell.get_initial_state(batch_size=B)
    hs = []
    for t in range(T):
        h, states = cell(xt[t], states)
        hs.append(h)
    dhs = rng.standard_normal((T, B, H)).astype(np.float32).astype(np.float64)
    sum((h * torch.tensor(dhs[t], dtype=torch.float32, device=cuda)).sum() for t, h in enumerate(hs)).backward()
    torch.cuda.synchronize()
    key = cell.seed
    noise = lambda stream, d: op.per_example("normal", key, stream, B, d)  # noqa: E731
    e_a, e_g, e_ra, e_rg = noise(lr.S_LSTM_ALPHA, I), noise(lr.S_LSTM_GAMMA, 4 * H), noise(lr.S_LSTM_RALPHA, H), noise(lr.S_LSTM_RGAMMA, 4 * H)
    P = lambda init: tuple(to_np(p) for p in init.params())  # noqa: E731
    (ma, ra), (mg, rg) = P(cell.alpha_initializer), P(cell.gamma_initializer)
    (mra, rra), (mrg, rrg) = P(cell.recurrent_alpha_initializer), P(cell.recurrent_gamma_initializer)
    W, U, b = to_np(cell.kernel), to_np(cell.recurrent_kernel), to_np(cell.bias)
    noclip = (-np.inf, np.inf)
    bx = lambda x: od.rank1_fwd(x, W, ma, ra, e_a, mg, rg, e_g, b, None, clip=noclip)[0]  # noqa: E731
    bh = lambda h: od.rank1_fwd(h, U, mra, rra, e_ra, mrg, rrg, e_rg, None, None, clip=noclip)[0]  # noqa: E731
    gx_ = lambda x, dz: od.rank1_bwd(x, W, ma, ra, e_a, mg, rg, e_g, b, None, dz, clip=noclip)  # noqa: E731
    gh_ = lambda h, dz: od.rank1_bwd(h, U, mra, rra, e_ra, mrg, rrg, e_rg, None, None, dz, clip=noclip)  # noqa: E731
    hs_ref, dxs, gx, gh = orc.two_branch_sequence(xs, np.zeros((B, H)), np.zeros((B, H)), bx, bh, gx_, gh_, dhs)
    tol = 2e-5
    assert rel_err(np.stack([to_np(h) for h in hs]), hs_ref) < tol
    assert rel_err(to_np(xt.grad), dxs) < tol
    assert rel_err(to_np(cell.kernel.grad), gx["dw"]) < tol and rel_err(to_np(cell.recurrent_kernel.grad), gh["dw"]) < tol
    pa, pg = cell.alpha_initializer.params(), cell.gamma_initializer.params()
    pra, prg = cell.recurrent_alpha_initializer.params(), cell.recurrent_gamma_initializer.params()
    for got, want in ((pa[0], gx["dmu_r"]), (pa[1], gx["drho_r"]), (pg[0], gx["dmu_s"]), (pg[1], gx["drho_s"]),
                      (pra[0], gh["dmu_r"]), (pra[1], gh["drho_r"]), (prg[0], gh["dmu_s"]), (prg[1], gh["drho_s"])):
        assert rel_err(to_np(got.grad), want) < tol
    assert rel_err(to_np(cell.bias.grad), gx["dbias"]) < tol
    assert b[:, H:2 * H].min() == 1.0 and len(cell.losses) == 4
