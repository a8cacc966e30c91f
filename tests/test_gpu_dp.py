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
    if exchange == "peer-hooks":  # generic hook path: fp32 dmu / drho cast to bf16, summed through peer memory
        red = GradientAllReduce(params, compress=torch.bfloat16, exchange="peer", early=False, hook_exchange="peer")
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
                                                  ("peer", 2, 0)])
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
