"""World-size-2 gloo tests (CPU) of the data-parallel host logic: member-major sharding, the gradient all-reduce
hooks, and that the all-reduced precision partial + blend equals the single-device global-batch update."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from edward2_b200 import dist as ed_dist
    from oracle import dense as od, sngp as og

    rng = np.random.default_rng(0)  # same stream on every rank: global tensors
    E, n, I, O, D = 2, 8, 6, 5, 7
    x = rng.standard_normal((E * n, I))
    w, alpha, gamma, bias = rng.standard_normal((I, O)), rng.standard_normal((E, I)), rng.standard_normal((E, O)), rng.standard_normal((E, O))
    gy = rng.standard_normal((E * n, O))
    xs = ed_dist.member_major_shard(torch.tensor(x), E, rank, world).numpy()
    gys = ed_dist.member_major_shard(torch.tensor(gy), E, rank, world).numpy()

    # local forward/backward (oracle arithmetic on the shard), gradients summed with the hook-driven all-reduce
    params = [torch.nn.Parameter(torch.tensor(t)) for t in (w, alpha, gamma, bias)]
    red = ed_dist.GradientAllReduce(params)
    g = od.batch_ensemble_bwd(xs, w, alpha, gamma, bias, "relu", gys)
    loss = sum((p * torch.tensor(gr)).sum() for p, gr in zip(params, (g["dw"], g["dalpha"], g["dgamma"], g["dbias"])))
    loss.backward()   # grad of each parameter == its local gradient; hooks fire and all-reduce
    red.wait()
    full = od.batch_ensemble_bwd(x, w, alpha, gamma, bias, "relu", gy)
    ok = all(np.allclose(p.grad.numpy(), full[k], rtol=1e-10, atol=1e-10)
             for p, k in zip(params, ("dw", "dalpha", "dgamma", "dbias")))

    # forward outputs re-assembled in member-major order equal the global forward
    y_local, _ = od.batch_ensemble_fwd(xs, w, alpha, gamma, bias, "relu")
    gathered = [torch.zeros_like(torch.tensor(y_local)) for _ in range(world)]
    dist.all_gather(gathered, torch.tensor(y_local))
    y_full, _ = od.batch_ensemble_fwd(x, w, alpha, gamma, bias, "relu")
    ok &= np.allclose(ed_dist.member_major_unshard(gathered, E).numpy(), y_full)

    # precision: sum of local phiᵀphi partials, then ONE blend with the global batch size
    phi = rng.standard_normal((E * n, D))
    phis = ed_dist.member_major_shard(torch.tensor(phi), E, rank, world)
    partial = ed_dist.allreduce_precision_partial(phis.t() @ phis)
    p0 = np.eye(D) * 0.3
    for m in (0.9, -1.0):
        blended = m * p0 + (1 - m) * partial.numpy() / (E * n) if m > 0 else p0 + partial.numpy()
        ok &= np.allclose(blended, og.precision_update_tf(p0, phi, m))
    # deferred sum (dist.DeferredSum): rank-local linear updates over several steps, ONE all-reduce at the end, equals
    # the per-step global-batch recursion — with and without momentum, across two accumulation windows
    for m in (0.9, -1.0):
        p_ref = np.eye(D) * 0.3
        p_loc = torch.tensor(np.eye(D) * 0.3)
        ds = ed_dist.DeferredSum(p_loc)
        for window in range(2):
            for step in range(3):
                phi_t = np.random.default_rng(100 * window + step).standard_normal((E * n, D))
                p_ref = og.precision_update_tf(p_ref, phi_t, m)
                ph = ed_dist.member_major_shard(torch.tensor(phi_t), E, rank, world)
                ds.begin_local_update()
                if m > 0:
                    p_loc.mul_(m).add_((1 - m) / (E * n) * (ph.t() @ ph))
                else:
                    p_loc.add_(ph.t() @ ph)
            ds.synchronize()
            ok &= np.allclose(p_loc.numpy(), p_ref, rtol=1e-12, atol=1e-12)
        ds.synchronize()  # nothing pending: must not double-count
        ok &= np.allclose(p_loc.numpy(), p_ref, rtol=1e-12, atol=1e-12)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_data_parallel_host_logic_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_member_major_shard_errors():
    import pytest

    from edward2_b200 import dist as ed_dist

    with pytest.raises(ValueError):
        ed_dist.member_major_shard(torch.zeros(10, 3), 4, 0, 2)
    with pytest.raises(ValueError):
        ed_dist.member_major_shard(torch.zeros(12, 3), 4, 0, 2)
    x = torch.arange(16.0).reshape(8, 2)
    parts = [ed_dist.member_major_shard(x, 2, r, 2) for r in range(2)]
    assert torch.equal(ed_dist.member_major_unshard(parts, 2), x)
    assert ed_dist.global_row_offset(3, 128) == 384
