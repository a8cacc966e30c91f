"""Data-parallel host logic (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The path shards naturally over batch rows.  Only two collectives exist:
  * sum-all-reduce of parameter gradients, launched as each gradient becomes final so it overlaps the rest of backward;
  * sum-all-reduce of the local ΦᵀΦ partial before the momentum blend (LaplaceRandomFeatureCovariance(data_parallel=True)),
    which reproduces the single-device global-batch result (the reference's ONLY_FIRST_REPLICA aggregation keeps
    replica 0's minibatch only, random_feature.py:353 — not reproduced).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def member_major_shard(x: torch.Tensor, ensemble_size: int, rank: int, world: int) -> torch.Tensor:
    """Shard a member-major batch [E * n, ...] so that every rank holds n / world examples OF EVERY MEMBER, still
    member-major: alpha / gamma / r / s indexing by row // (n / world) stays valid locally."""
    B = x.shape[0]
    if B % ensemble_size != 0:
        raise ValueError(f"batch {B} is not divisible by ensemble_size {ensemble_size}")
    n = B // ensemble_size
    if n % world != 0:
        raise ValueError(f"examples per member {n} is not divisible by world size {world}")
    per = n // world
    xe = x.reshape(ensemble_size, n, *x.shape[1:])
    return xe[:, rank * per:(rank + 1) * per].reshape(ensemble_size * per, *x.shape[1:])


def member_major_unshard(parts, ensemble_size: int) -> torch.Tensor:
    """Inverse of member_major_shard given the per-rank outputs in rank order."""
    per = parts[0].shape[0] // ensemble_size
    stacked = [p.reshape(ensemble_size, per, *p.shape[1:]) for p in parts]
    return torch.cat(stacked, dim=1).reshape(-1, *parts[0].shape[1:])


def global_row_offset(rank: int, rows_per_rank: int) -> int:
    """Philox offsets are functions of the GLOBAL row index so results do not depend on the number of ranks."""
    return rank * rows_per_rank


class GradientAllReduce:
    """Overlapped gradient all-reduce: a post-accumulate hook per parameter launches an async sum-all-reduce on the
    process group's own stream the moment that gradient is final; `wait()` joins them before the optimizer step."""

    def __init__(self, params, group=None, average: bool = False):
        self.handles, self.group, self.average = [], group, average
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._hooks = []
        if self.world > 1:
            for p in params:
                if p.requires_grad:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._hook))

    def _hook(self, p):
        if self.average:
            p.grad.div_(self.world)
        self.handles.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self):
        for h in self.handles:
            h.wait()
        self.handles.clear()

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks.clear()


def allreduce_precision_partial(partial: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the local ΦᵀΦ partials over ranks (in place); the blend P <- m P + (1-m) sum / B_global follows."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial
