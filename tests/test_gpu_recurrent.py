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
