"""2-GPU NCCL test (skipped with fewer than 2 GPUs): a data-parallel step of a 2-layer reparameterisation MLP with the
early raw-dW all-reduce equals the single-device global-batch step (same seeds => same eps on every rank)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(ed, dev):
    mk = lambda u, act, seed: ed.layers.DenseReparameterization(  # noqa: E731
        u, activation=act, dtype="bfloat16", seed=seed,
        kernel_regularizer=ed.regularizers.NormalKLDivergence(scale_factor=1e-4),
        kernel_initializer=ed.initializers.TrainableNormal(
            mean_initializer=ed.initializers.RandomNormal(stddev=0.05, seed=seed),
            stddev_initializer=ed.initializers.TruncatedNormal(-3.0, 0.1, seed=seed + 100)))
    return [mk(512, "relu", 1), mk(264, None, 2)]


def _step(layers, x, y, global_batch, F):
    h = x
    for l in layers:
        h = l(h)
    loss = F.SoftmaxCrossEntropy.apply(h, y, global_batch, *[l.losses[0] for l in layers])
    F.backward(loss)
    return loss


def _worker(rank, world, port, out, exchange="peer", steps=1):
    sys.path.insert(0, ROOT)
    if exchange == "peer-disabled":  # the arena cannot be built: every rank must agree to use the NCCL path
        os.environ["ED2_PEER_DISABLE"] = "1"
        exchange = "peer"
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import edward2_b200 as ed
    from edward2_b200 import functional as F
    from edward2_b200.dist import GradientAllReduce

    g = torch.Generator().manual_seed(0)
    B = 256
    x = torch.randn(B, 384, generator=g).to(torch.bfloat16)
    y = torch.randint(0, 264, (B,), generator=g)
    per = B // world
    layers = _build(ed, dev)
    xs, ys = x[rank * per:(rank + 1) * per].to(dev), y[rank * per:(rank + 1) * per].to(dev)
    with torch.no_grad():
        h = xs
        for l in layers:
            h = l(h)
    for l in layers:
        l._calls = 7  # same Philox offset as the single-device run below
    params = [p for l in layers for p in l.parameters()]
    if exchange in ("peer-hooks", "dma-hooks"):  # generic hook path: fp32 dmu / drho cast to bf16, summed through peer
        # memory by SM pulls ("peer") or by the copy engines ("dma")
        red = GradientAllReduce(params, compress=torch.bfloat16, exchange="peer", early=False,
                                hook_exchange=exchange.split("-")[0])
    elif exchange == "fp32-hooks":  # per-parameter fp32 all-reduce: every rank adds its KL gradient / world before the sum
        red = GradientAllReduce(params, compress=None)
    else:
        red = GradientAllReduce(params, compress=torch.bfloat16, exchange=exchange)
    for _ in range(steps):  # repeated steps re-use the exchange buffers (flag epochs)
        for p in params:
            p.grad = None
        for l in layers:
            l._calls = 7
        _step(layers, xs, ys, B, F)
        red.wait()
    torch.cuda.synchronize()
    grads = {n: p.grad.float().cpu() for l_i, l in enumerate(layers) for n, p in ((f"{l_i}.{k}", v) for k, v in l.named_parameters())}
    red.remove()
    if rank == 0:
        # single-device global batch with the same parameters and the same Philox offset
        ref_layers = _build(ed, dev)
        xd, yd = x.to(dev), y.to(dev)
        with torch.no_grad():
            h = xd
            for l in ref_layers:
                h = l(h)
        for l in ref_layers:
            l._calls = 7
        _step(ref_layers, xd, yd, B, F)
        torch.cuda.synchronize()
        ok = True
        worst = 0.0
        for l_i, l in enumerate(ref_layers):
            for k, v in l.named_parameters():
                a, b = grads[f"{l_i}.{k}"], v.grad.float().cpu()
                err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
                worst = max(worst, err)
                ok &= err < 2e-2
        out["ok"], out["worst"] = bool(ok), worst
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange,steps,world", [("peer", 1, 2), ("peer", 3, 2), ("nccl", 1, 2), ("peer-disabled", 2, 2), ("fp32-hooks", 1, 2), ("peer-hooks", 2, 2),
                                                  ("dma-hooks", 3, 2), ("peer", 2, 0), ("dma-hooks", 2, 0)])
def test_data_parallel_matches_global_batch(exchange, steps, world):
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs 2 GPUs")
    if world == 0:  # every GPU of the box (power of two, <= 8)
        world = 8 if n_dev >= 8 else 4 if n_dev >= 4 else 2
        if world == 2:
            pytest.skip("needs more than 2 GPUs")
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out, exchange, steps), nprocs=world, join=True)
    assert out.get("ok"), f"DP gradients differ from the global-batch step (worst rel err {out.get('worst')})"
    print(f"[{exchange} x{steps}] worst rel err {out.get('worst'):.3e}")


def _precision_worker(rank, world, port, out):
    """LaplaceRandomFeatureCovariance over `world` ranks: per-step all-reduce (data_parallel=True) and the deferred sum
    (data_parallel="deferred") both reproduce the single-device global-batch precision matrix and variance."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import edward2_b200 as ed

    B, D, steps = 512, 256, 3
    g = torch.Generator().manual_seed(1)
    phis = [torch.randn(B, D, generator=g).to(torch.bfloat16) for _ in range(steps + 1)]
    per = B // world
    res = {}
    for momentum in (0.9, -1.0):
        mats = {}
        for mode in (True, "deferred"):
            cov = ed.layers.LaplaceRandomFeatureCovariance(momentum=momentum, ridge_penalty=1e-3, dtype="bfloat16",
                                                           data_parallel=mode)
            cov._device = dev
            for s in range(steps):
                cov(phis[s][rank * per:(rank + 1) * per].to(dev), training=True)
            var, _ = cov.compute_predictive_variance(phis[steps][:per].to(dev))
            mats[mode] = (cov.precision_matrix.float().cpu(), var.float().cpu())
        if rank == 0:
            ref = ed.layers.LaplaceRandomFeatureCovariance(momentum=momentum, ridge_penalty=1e-3, dtype="bfloat16")
            ref._device = dev
            for s in range(steps):
                ref(phis[s].to(dev), training=True)
            rvar, _ = ref.compute_predictive_variance(phis[steps][:per].to(dev))
            rp = ref.precision_matrix.float().cpu()
            for mode, (pm, var) in mats.items():
                res[f"{momentum}/{mode}/P"] = float((pm - rp).abs().max() / rp.abs().max())
                res[f"{momentum}/{mode}/var"] = float((var - rvar.float().cpu()).abs().max() / rvar.abs().max().cpu())
    if rank == 0:
        out["res"] = res
    dist.destroy_process_group()


def test_precision_matrix_data_parallel_and_deferred():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_precision_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = dict(out["res"])
    print(res)
    assert res and all(v < 1e-4 for k, v in res.items() if k.endswith("/P")), res  # bf16 operands, fp32 accumulate
    assert all(v < 2e-2 for k, v in res.items() if k.endswith("/var")), res


def _stochastic_worker(rank, world, port, out):
    """One DenseRank1 / DenseFlipout / DenseBatchEnsemble layer per test case, every rank on its member-major shard with
    global-row noise offsets (dist.shard_noise) and the copy-engine gradient exchange; rank 0 checks the gathered outputs
    and the summed gradients against the float64 ORACLE on the GLOBAL batch."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import edward2_b200 as ed
    from edward2_b200.dist import GradientAllReduce, member_major_shard, shard_noise
    from edward2_b200.layers import dense as ld
    from oracle import dense as od
    from oracle import philox as op
    from tests.util import rel_err, round_to, to_np

    E, n, I, O = 4, 256, 264, 520  # 128 rows per member per rank at world 2: the fused rank-1 epilogues run
    B = E * n
    rng = np.random.default_rng(31)
    x = round_to(rng.standard_normal((B, I)), torch.bfloat16)
    gy = round_to(rng.standard_normal((B, O)), torch.bfloat16)
    tn = lambda s: ed.initializers.TrainableNormal(  # noqa: E731
        mean_initializer=ed.initializers.RandomSign(seed=s), stddev_initializer=ed.initializers.TruncatedNormal(-2.0, 0.2, seed=s + 1))
    cases = {
        "rank1": lambda: ed.layers.DenseRank1(O, activation="relu", ensemble_size=E, dtype="bfloat16", seed=11,
                                              alpha_initializer=tn(1), gamma_initializer=tn(3),
                                              kernel_initializer=ed.initializers.RandomNormal(stddev=0.06, seed=5)),
        "flipout": lambda: ed.layers.DenseFlipout(O, activation="relu", dtype="bfloat16", seed=12, kernel_regularizer=None,
                                                  kernel_initializer=ed.initializers.TrainableNormal(
                                                      mean_initializer=ed.initializers.RandomNormal(stddev=0.06, seed=6),
                                                      stddev_initializer=ed.initializers.TruncatedNormal(-3.0, 0.2, seed=7))),
        "be": lambda: ed.layers.DenseBatchEnsemble(O, ensemble_size=E, activation="relu", dtype="bfloat16",
                                                   alpha_initializer=ed.initializers.RandomSign(seed=8),
                                                   gamma_initializer=ed.initializers.RandomSign(seed=9),
                                                   kernel_initializer=ed.initializers.RandomNormal(stddev=0.06, seed=10)),
    }
    res = {}
    for name, make in cases.items():
        layer = make()
        groups = 1 if name == "flipout" else E
        xs = member_major_shard(torch.tensor(x, dtype=torch.float32), groups, rank, world).to(dev).to(torch.bfloat16)
        gs = member_major_shard(torch.tensor(gy, dtype=torch.float32), groups, rank, world).to(dev).to(torch.bfloat16)
        with torch.no_grad():
            layer(xs)  # build
        layer._calls = 1
        shard_noise(layer, rank, world)
        params = list(layer.parameters())
        red = GradientAllReduce(params, compress=torch.bfloat16, hook_exchange="dma")
        for _ in range(2):  # second step re-uses the exchange buffers
            for p in params:
                p.grad = None
            layer._calls = 1
            y = layer(xs)
            y.backward(gs)
            red.wait()
        torch.cuda.synchronize()
        ys = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(ys, y.detach().contiguous())
        red.remove()
        if rank == 0:
            per = ys[0].shape[0] // groups
            yg = torch.cat([t.reshape(groups, per, -1) for t in ys], dim=1).reshape(B, -1)
            key = (layer.seed + ld.base.GOLDEN) & ld.base.MASK64 if hasattr(ld, "base") else (layer.seed + 0x9E3779B97F4A7C15) & ((1 << 64) - 1)
            if name == "rank1":
                eps_r = op.normal(key, ld.S_R, B * I).reshape(B, I)
                eps_s = op.normal(key, ld.S_S, B * O).reshape(B, O)
                mu_r, rho_r = (to_np(t) for t in layer.alpha_initializer.params())
                mu_s, rho_s = (to_np(t) for t in layer.gamma_initializer.params())
                w = to_np(layer.kernel.detach().to(torch.bfloat16))
                bias = to_np(layer.ensemble_bias)
                y_ref, _, _, _ = od.rank1_fwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, "relu")
                g = od.rank1_bwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, "relu", gy, y_mask=to_np(yg))
                got = {"dw": layer.kernel.grad, "dmu_r": layer.alpha_initializer.params()[0].grad,
                       "drho_r": layer.alpha_initializer.params()[1].grad, "dmu_s": layer.gamma_initializer.params()[0].grad,
                       "drho_s": layer.gamma_initializer.params()[1].grad, "dbias": layer.ensemble_bias.grad}
            elif name == "flipout":
                eps = op.normal(key, ld.S_KERNEL, I * O).reshape(I, O)
                s_in = op.sign(key, ld.S_SIGN_IN, B * I).reshape(B, I)
                s_out = op.sign(key, ld.S_SIGN_OUT, B * O).reshape(B, O)
                mu, rho = (to_np(t) for t in layer.kernel_initializer.params())
                bias = to_np(layer.bias)
                y_ref, _ = od.flipout_fwd(x, mu, rho, eps, s_in, s_out, bias, "relu")
                g = od.flipout_bwd(x, mu, rho, eps, s_in, s_out, bias, "relu", gy, y_mask=to_np(yg))
                got = {"dmu": layer.kernel_initializer.params()[0].grad, "drho": layer.kernel_initializer.params()[1].grad,
                       "dbias": layer.bias.grad}
            else:
                w = to_np(layer.kernel.detach().to(torch.bfloat16))
                alpha, gamma, bias = (to_np(t) for t in (layer.alpha, layer.gamma, layer.ensemble_bias))
                y_ref, _ = od.batch_ensemble_fwd(x, w, alpha, gamma, bias, "relu")
                g = od.batch_ensemble_bwd(x, w, alpha, gamma, bias, "relu", gy, y_mask=to_np(yg))
                got = {"dw": layer.kernel.grad, "dalpha": layer.alpha.grad, "dgamma": layer.gamma.grad,
                       "dbias": layer.ensemble_bias.grad}
            res[f"{name}/y"] = rel_err(to_np(yg), y_ref)
            for k, t in got.items():
                res[f"{name}/{k}"] = rel_err(to_np(t), g[k])
        dist.barrier()
    if rank == 0:
        out["res"] = res
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 0])
def test_dp_stochastic_layers_match_oracle(world):
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs 2 GPUs")
    if world == 0:
        world = 8 if n_dev >= 8 else 4 if n_dev >= 4 else 2
        if world == 2:
            pytest.skip("needs more than 2 GPUs")
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_stochastic_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    res = dict(out["res"])
    print(res)
    bad = {k: v for k, v in res.items() if not v <= 2e-2}
    assert res and not bad, bad
