"""Data-parallel parity against the ORACLE on the GLOBAL batch, on one GPU: the per-example noise of
DenseRank1 / DenseFlipout is a function of the global row (ed2_noise.row0 / rows_per_group_*), so running the layer
once per emulated rank on that rank's member-major shard — exactly what each process of a `world`-rank job executes —
must reproduce the single-device global-batch result of the float64 oracle fed with the host restatement of the same
Philox streams: outputs row for row, parameter gradients after the cross-rank sum.  The multi-process version of the
same check (real NCCL ranks) is tests/test_gpu_dp.py::test_dp_stochastic_layers_match_oracle."""
import numpy as np
import pytest
import torch

from oracle import dense as od
from oracle import philox as op
from tests.util import TOL, dev, rel_err, round_to, to_np

pytestmark = pytest.mark.gpu


def _check(name, got, want, tol):
    e = rel_err(to_np(got) if isinstance(got, torch.Tensor) else got, want)
    assert e <= tol, f"{name}: rel err {e:.3e} > {tol:.1e}"


def _shard(a, E, rank, world):
    n = a.shape[0] // E
    per = n // world
    return a.reshape(E, n, -1)[:, rank * per:(rank + 1) * per].reshape(E * per, -1)


def _unshard(parts, E):
    per = parts[0].shape[0] // E
    return np.concatenate([p.reshape(E, per, -1) for p in parts], axis=1).reshape(E * per * len(parts), -1)


def _run_sharded(layer, x, gy, E, world, dtype, cuda):
    """One forward + backward per emulated rank; parameter gradients accumulate (= the all-reduce sum)."""
    ys, dxs = [], []
    for rank in range(world):
        layer._calls = 0  # every rank of a real job is at the same call count
        layer.data_shard = (rank, world)
        xt = dev(_shard(x, E, rank, world), dtype, cuda, True)
        y = layer(xt)
        y.backward(dev(_shard(gy, E, rank, world), dtype, cuda))
        ys.append(to_np(y))
        dxs.append(to_np(xt.grad))
    torch.cuda.synchronize()
    return _unshard(ys, E), _unshard(dxs, E)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("world", [2, 8])
def test_rank1_sharded_equals_global_oracle(cuda, dtype, world):
    import edward2_b200 as ed
    from edward2_b200.layers import dense as ld

    E, n, I, O = 4, 64, 72, 136  # 64 examples per member -> 32 / 8 per rank
    B = E * n
    rng = np.random.default_rng(21)
    x = round_to(rng.standard_normal((B, I)), dtype)
    gy = round_to(rng.standard_normal((B, O)), dtype)
    layer = ed.layers.DenseRank1(O, activation="relu", ensemble_size=E, seed=1234,
                                 dtype="bfloat16" if dtype == torch.bfloat16 else "float32",
                                 alpha_initializer=ed.initializers.TrainableNormal(
                                     mean_initializer=ed.initializers.RandomSign(seed=1),
                                     stddev_initializer=ed.initializers.TruncatedNormal(-1.0, 0.2, seed=2)),
                                 gamma_initializer=ed.initializers.TrainableNormal(
                                     mean_initializer=ed.initializers.RandomSign(seed=3),
                                     stddev_initializer=ed.initializers.TruncatedNormal(-1.0, 0.2, seed=4)))
    y, dx = _run_sharded(layer, x, gy, E, world, dtype, cuda)

    key = layer.seed
    eps_r = op.normal(key, ld.S_R, B * I).reshape(B, I)
    eps_s = op.normal(key, ld.S_S, B * O).reshape(B, O)
    w = round_to(to_np(layer.kernel), dtype)
    mu_r, rho_r = (to_np(t) for t in layer.alpha_initializer.params())
    mu_s, rho_s = (to_np(t) for t in layer.gamma_initializer.params())
    bias = to_np(layer.ensemble_bias)
    tol = TOL[dtype]
    y_ref, _, _, _ = od.rank1_fwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, "relu")
    _check("y", y, y_ref, tol)
    g = od.rank1_bwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, "relu", gy,
                     y_mask=None if dtype == torch.float32 else y)
    _check("dx", dx, g["dx"], tol)
    _check("dw", layer.kernel.grad, g["dw"], tol)
    _check("dmu_r", layer.alpha_initializer.params()[0].grad, g["dmu_r"], tol)
    _check("drho_r", layer.alpha_initializer.params()[1].grad, g["drho_r"], tol)
    _check("dmu_s", layer.gamma_initializer.params()[0].grad, g["dmu_s"], tol)
    _check("drho_s", layer.gamma_initializer.params()[1].grad, g["drho_s"], tol)
    _check("dbias", layer.ensemble_bias.grad, g["dbias"], tol)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("world", [2, 8])
def test_flipout_sharded_equals_global_oracle(cuda, dtype, world):
    import edward2_b200 as ed
    from edward2_b200.layers import dense as ld

    B, I, O = 128, 72, 136
    rng = np.random.default_rng(22)
    x = round_to(rng.standard_normal((B, I)), dtype)
    gy = round_to(rng.standard_normal((B, O)), dtype)
    layer = ed.layers.DenseFlipout(O, activation="relu", seed=4321, kernel_regularizer=None,
                                   dtype="bfloat16" if dtype == torch.bfloat16 else "float32",
                                   kernel_initializer=ed.initializers.TrainableNormal(
                                       mean_initializer=ed.initializers.RandomNormal(stddev=0.1, seed=5),
                                       stddev_initializer=ed.initializers.TruncatedNormal(-1.5, 0.2, seed=6)))
    y, dx = _run_sharded(layer, x, gy, 1, world, dtype, cuda)

    key = layer.seed
    eps = op.normal(key, ld.S_KERNEL, I * O).reshape(I, O)
    s_in = op.sign(key, ld.S_SIGN_IN, B * I).reshape(B, I)
    s_out = op.sign(key, ld.S_SIGN_OUT, B * O).reshape(B, O)
    mu, rho = (to_np(t) for t in layer.kernel_initializer.params())
    bias = to_np(layer.bias)
    tol = TOL[dtype]
    y_ref, _ = od.flipout_fwd(x, mu, rho, eps, s_in, s_out, bias, "relu")
    _check("y", y, y_ref, tol)
    g = od.flipout_bwd(x, mu, rho, eps, s_in, s_out, bias, "relu", gy, y_mask=None if dtype == torch.float32 else y)
    _check("dx", dx, g["dx"], tol)
    _check("dmu", layer.kernel_initializer.params()[0].grad, g["dmu"], tol)
    _check("drho", layer.kernel_initializer.params()[1].grad, g["drho"], tol)
    _check("dbias", layer.bias.grad, g["dbias"], tol)


@pytest.mark.parametrize("world", [2, 4])
def test_batch_ensemble_member_major_shard_equals_global_oracle(cuda, world):
    import edward2_b200 as ed

    E, n, I, O = 4, 32, 784, 512
    B = E * n
    rng = np.random.default_rng(23)
    x = round_to(rng.standard_normal((B, I)), torch.float32)
    gy = round_to(rng.standard_normal((B, O)), torch.float32)
    layer = ed.layers.DenseBatchEnsemble(O, ensemble_size=E, activation="relu",
                                         alpha_initializer=ed.initializers.RandomSign(seed=1),
                                         gamma_initializer=ed.initializers.RandomSign(seed=2))
    y, dx = _run_sharded(layer, x, gy, E, world, torch.float32, cuda)
    w, alpha, gamma, bias = (to_np(t) for t in (layer.kernel, layer.alpha, layer.gamma, layer.ensemble_bias))
    y_ref, _ = od.batch_ensemble_fwd(x, w, alpha, gamma, bias, "relu")
    _check("y", y, y_ref, 1e-5)
    g = od.batch_ensemble_bwd(x, w, alpha, gamma, bias, "relu", gy)
    _check("dx", dx, g["dx"], 1e-5)
    for name, p in (("dw", layer.kernel), ("dalpha", layer.alpha), ("dgamma", layer.gamma), ("dbias", layer.ensemble_bias)):
        _check(name, p.grad, g[name], 1e-5)
