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
    """First global row (within a member's block) of the rows a rank holds: Philox offsets are functions of the GLOBAL
    row index so results do not depend on the number of ranks (used by Layer._row_map)."""
    return rank * rows_per_rank


def shard_noise(module, rank: int | None = None, world: int | None = None, group=None):
    """Tell every layer under `module` which slice of the global batch this process holds (the member_major_shard
    layout: rows [rank * n/world, (rank + 1) * n/world) of every ensemble member; for layers without members, of the
    whole batch).  From then on the per-example noise tensors of DenseRank1 (eps_alpha / eps_gamma / eps_bias) and
    DenseFlipout (sign_input / sign_output) are drawn for the GLOBAL row — ed2_noise.row0 / rows_per_group_* — so
    `world` ranks reproduce the single-device global-batch draw bit for bit.  Per-weight noise (one eps per kernel
    element) is identical on every rank by construction.  Also sets a common Philox seed discipline check: the layers'
    seeds must be equal across ranks (they derive from construction order); verified here when a process group exists."""
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    from .layers.base import Layer

    layers = [m for m in module.modules() if isinstance(m, Layer)]
    for m in layers:
        m.data_shard = (int(rank), int(world))
    if dist.is_initialized() and dist.get_world_size(group) > 1 and layers:
        check_noise_keys(layers, group)
    return module


def check_noise_keys(layers, group=None):
    """Raise if the (seed, call count) pairs that key the layers' Philox streams differ between ranks: the raw-dW
    exchange finishes drho = dW_sum∘eps∘sigmoid(rho) locally and is only correct when every rank drew the same eps."""
    keys = [(m.seed & 0x7FFFFFFFFFFFFFFF, m._calls) for m in layers]
    flat = torch.tensor([v for k in keys for v in k], dtype=torch.int64)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    lo, hi = flat.to(dev), flat.to(dev)
    lo, hi = lo.clone(), hi.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    if not torch.equal(lo, hi):
        bad = [layers[i // 2].layer_name for i in torch.nonzero(lo != hi).flatten().tolist()]
        raise RuntimeError(f"Philox keys differ across data-parallel ranks for layers {sorted(set(bad))}: build the "
                           f"model identically on every rank and call it the same number of times (or pass a shared "
                           f"step_counter)")


class PeerArena:
    """One cudaMalloc'ed arena on this rank, IPC-mapped into every peer of `group` (one NVLink / NVSwitch box).
    `ptrs[r]` is the address of rank r's arena in THIS process.  Construction is collective."""

    def __init__(self, nbytes: int, group=None):
        import ctypes as C

        from . import _lib

        self.lib, self.group = _lib.load(), group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.ptrs, self.nbytes, err = [], nbytes, None
        local, handle = C.c_void_p(), C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        try:
            import os

            if os.environ.get("ED2_PEER_DISABLE") == "1":  # test hook for the fallback below
                raise RuntimeError("disabled by ED2_PEER_DISABLE")
            _lib.check(self.lib.ed2_peer_alloc(nbytes, C.byref(local), handle), "ed2_peer_alloc")
        except Exception as e:  # noqa: BLE001  (agreed on collectively below: no rank may skip a collective)
            err, local = e, C.c_void_p()
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw if err is None else None, group=group)
        opened = []
        if err is None and all(h is not None for h in handles):
            try:
                for r, h in enumerate(handles):
                    if r == self.rank:
                        opened.append(local.value)
                    else:
                        p = C.c_void_p()
                        _lib.check(self.lib.ed2_peer_open(C.create_string_buffer(h, _lib.IPC_HANDLE_BYTES), C.byref(p)),
                                   "ed2_peer_open")
                        opened.append(p.value)
            except Exception as e:  # noqa: BLE001
                err = e
        elif err is None:
            err = RuntimeError("a peer could not allocate its arena")
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # also: every arena is zeroed and mapped from here on
        if int(ok.item()) == 0:
            for r, p in enumerate(opened):
                if r != self.rank:
                    self.lib.ed2_peer_close(p)
            dist.barrier(group=group)
            if local.value:
                self.lib.ed2_peer_free(local.value)
            raise RuntimeError(f"peer-memory arena unavailable on this box: {err or 'a peer failed'}")
        self.ptrs = opened

    def close(self):
        if not self.ptrs:
            return
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                self.lib.ed2_peer_close(p)
        dist.barrier(group=self.group)
        self.lib.ed2_peer_free(self.ptrs[self.rank])
        self.ptrs = []


def peer_layout(n: int, small_n: int, world: int) -> dict:
    """Byte offsets of one exchange arena: [bucket n bf16 | reduced n bf16 | small 2 x m fp32 | flags], every region
    256-byte aligned; `chunk` = elements per owner (multiple of 8 so that a 16-byte vector never straddles owners).
    `small` is double-buffered by epoch parity: a lagging peer still reads epoch e while this rank publishes e + 1."""
    if n <= 0 or n % 8 != 0:
        raise ValueError(f"peer exchange needs n % 8 == 0 (got {n})")
    if not 2 <= world <= 8:
        raise ValueError(f"peer exchange needs 2 <= world <= 8 (got {world})")
    al = lambda b: (b + 255) // 256 * 256  # noqa: E731
    off_reduced = al(2 * n)
    off_small = off_reduced + al(2 * n)
    off_flags = off_small + al(2 * 4 * max(small_n, 1))
    return dict(bucket=0, reduced=off_reduced, small=off_small, flags=off_flags,
                nbytes=off_flags + al(8 * (2 * world + 3)), chunk=((n + world - 1) // world + 7) // 8 * 8)


class PeerExchange:
    """The peer-memory exchange of ONE weight gradient (include/ed2b200.h: ed2_peer_*); arena layout: peer_layout()."""

    def __init__(self, n: int, small_n: int, group=None):
        import ctypes as C

        from . import _lib

        self.lib = _lib.load()
        world = dist.get_world_size(group)
        lay = peer_layout(n, small_n, world)
        self.arena = PeerArena(lay["nbytes"], group)
        c = _lib.PeerExchange()
        c.world, c.rank, c.n, c.chunk, c.small_n = world, self.arena.rank, n, lay["chunk"], small_n
        for r, base in enumerate(self.arena.ptrs):
            c.bucket[r], c.reduced[r] = base + lay["bucket"], base + lay["reduced"]
            c.small_vec[r], c.flags[r] = base + lay["small"], base + lay["flags"]
        self.c, self._ref = c, C.byref(c)
        self.bucket_ptr = self.arena.ptrs[self.arena.rank]
        self.n, self.small_n = n, small_n
        self._pending = False  # set by signal(), cleared when the reducer has joined the exchange (wait())

    def signal(self, small_src=None):
        from . import _lib

        if self._pending:
            raise RuntimeError("PeerExchange.signal() called again before the previous exchange was joined "
                               "(GradientAllReduce.wait()): two backward passes per step (gradient accumulation, a "
                               "shared layer) would overwrite a bucket peers may still be pulling — use exchange='nccl'")
        self._pending = True
        _lib.check(self.lib.ed2_peer_signal(self._ref, None if small_src is None else small_src.data_ptr(),
                                            _lib.stream_ptr()), "ed2_peer_signal")

    def reduce_slice(self, light: bool = False):
        from . import _lib

        _lib.check(self.lib.ed2_peer_reduce_slice(self._ref, int(light), _lib.stream_ptr()), "ed2_peer_reduce_slice")

    def grad_finish(self, mu, rho, noise, klg, kl_upstream, prior_mean, prior_std, dmu, drho, small_out, fast,
                    blocks_per_sm=5):
        from . import _lib

        _lib.check(self.lib.ed2_peer_grad_finish(
            self._ref, mu.data_ptr(), rho.data_ptr(), noise, klg, None if kl_upstream is None else kl_upstream.data_ptr(),
            prior_mean, prior_std, dmu.data_ptr(), drho.data_ptr(), None if small_out is None else small_out.data_ptr(),
            int(fast), int(blocks_per_sm), _lib.stream_ptr()), "ed2_peer_grad_finish")

    def gather_f32(self, out):
        from . import _lib

        _lib.check(self.lib.ed2_peer_gather_f32(self._ref, out.data_ptr(), _lib.stream_ptr()), "ed2_peer_gather_f32")

    def cast_into_bucket(self, grad2d):
        """bf16(grad) -> this rank's bucket (the library's cast kernel, current stream)."""
        from . import _lib

        rows, cols = grad2d.shape
        _lib.check(self.lib.ed2_cast(grad2d.data_ptr(), _lib.dt_of(grad2d), grad2d.stride(0), self.bucket_ptr, _lib.BF16,
                                     cols, rows, cols, _lib.stream_ptr()), "ed2_cast")

    def close(self):
        self.arena.close()


def dma_layout(n: int, world: int) -> dict:
    """Byte offsets of one copy-engine exchange arena:
    [bucket W*chunk bf16 | recv W*chunk bf16 | gathered W*chunk bf16 | flags | scratch flags], 256-byte aligned."""
    if n <= 0 or n % 8 != 0:
        raise ValueError(f"dma exchange needs n % 8 == 0 (got {n})")
    if not 2 <= world <= 8:
        raise ValueError(f"dma exchange needs 2 <= world <= 8 (got {world})")
    al = lambda b: (b + 255) // 256 * 256  # noqa: E731
    chunk = ((n + world - 1) // world + 7) // 8 * 8
    region = al(2 * chunk * world)
    fl = al(8 * (2 * world + 3))
    return dict(bucket=0, recv=region, gathered=2 * region, flags=3 * region, scratch=3 * region + fl,
                nbytes=3 * region + 2 * fl, chunk=chunk)


class DmaExchange:
    """Sum of ONE bf16 gradient over the ranks of a box with the data moved by the COPY ENGINES (include/ed2b200.h:
    ed2_dma_copy / ed2_peer_flag): reduce-scatter = every rank DMA-pushes slice q of its bucket into rank q's receive
    slots, the owner sums its slots from LOCAL memory; all-gather = the owner DMA-pushes the reduced slice into every
    rank's gathered buffer, which is converted to fp32 locally.  No SM takes part in the transfers, so they overlap the
    persistent GEMMs (which own every SM of the fused rank-1 step) instead of queueing behind them."""

    def __init__(self, n: int, group=None):
        import ctypes as C

        from . import _lib

        self.lib = _lib.load()
        world = dist.get_world_size(group)
        lay = dma_layout(n, world)
        self.arena = PeerArena(lay["nbytes"], group)
        rank, ptrs = self.arena.rank, self.arena.ptrs
        self.n, self.small_n, self.world, self.rank, self.chunk, self.lay = n, 0, world, rank, lay["chunk"], lay
        me = ptrs[rank]
        lo = rank * lay["chunk"]
        real, local = _lib.PeerExchange(), _lib.PeerExchange()
        for c in (real, local):
            c.world, c.rank, c.n, c.chunk, c.small_n = world, rank, n, lay["chunk"], 0
        for r, base in enumerate(ptrs):
            real.bucket[r], real.reduced[r], real.flags[r] = base + lay["bucket"], base + lay["gathered"], base + lay["flags"]
            real.small_vec[r] = base + lay["flags"]
            # local view: index i of "rank r's bucket" (i in my owned slice) -> my receive slot r
            local.bucket[r] = (me + lay["bucket"]) if r == rank else (me + lay["recv"] + 2 * (r * lay["chunk"] - lo))
            local.reduced[r] = me + lay["gathered"]
            local.flags[r] = (me + lay["flags"]) if r == rank else (me + lay["scratch"])
            local.small_vec[r] = me + lay["flags"]
        self.real, self.local = real, local
        self._real, self._local = C.byref(real), C.byref(local)
        self.bucket_ptr = me + lay["bucket"]
        self._pending = False
        self._pool = None

    def cast_into_bucket(self, grad2d):
        from . import _lib

        rows, cols = grad2d.shape
        _lib.check(self.lib.ed2_cast(grad2d.data_ptr(), _lib.dt_of(grad2d), grad2d.stride(0), self.bucket_ptr, _lib.BF16,
                                     cols, rows, cols, _lib.stream_ptr()), "ed2_cast")

    def _slice(self, q):
        lo = q * self.chunk
        return lo, max(0, min(self.n, lo + self.chunk) - lo)

    N_COPY_STREAMS = 4

    def _fan_out(self, copies):
        """Issue the (dst, src, bytes) DMA copies over several streams forked from / joined back into the current one:
        independent copies then run on different copy engines (a single stream would serialise them on one)."""
        from . import _lib

        copies = [c for c in copies if c[2] > 0]
        if not copies:
            return
        # (measured on B200: one engine at a time serves this GPU's peer writes at ~0.47 TB/s — cutting a copy into pieces
        # on several streams does not add bandwidth, profiles/r02_trace_n2_dma.txt — so copies are not split; the streams
        # only keep pushes to DIFFERENT peers independent of each other)
        cur = torch.cuda.current_stream()
        if len(copies) == 1:
            _lib.check(self.lib.ed2_dma_copy(*copies[0], cur.cuda_stream), "ed2_dma_copy")
            return
        if self._pool is None:  # per exchange: the chains of different layers never wait on each other's copies
            self._pool = [torch.cuda.Stream(priority=-1) for _ in range(self.N_COPY_STREAMS)]
        pool = self._pool
        fork = torch.cuda.Event()
        fork.record(cur)
        used = pool[:min(len(pool), len(copies))]
        for st in used:
            st.wait_event(fork)
        for i, (dst, src, nbytes) in enumerate(copies):
            _lib.check(self.lib.ed2_dma_copy(dst, src, nbytes, used[i % len(used)].cuda_stream), "ed2_dma_copy")
        for st in used:
            ev = torch.cuda.Event()
            ev.record(st)
            cur.wait_event(ev)

    def scatter_reduce(self):
        """Current stream: push my bucket's slices to their owners, publish, sum my slice, push it to everyone."""
        from . import _lib

        if self._pending:
            raise RuntimeError("DmaExchange started again before the previous exchange was joined (GradientAllReduce.wait())")
        self._pending = True
        lay, me = self.lay, self.arena.ptrs[self.rank]
        pushes = []
        for d in range(1, self.world):  # staggered peers: no two ranks target the same receiver at the same time
            q = (self.rank + d) % self.world
            lo, cnt = self._slice(q)
            pushes.append((self.arena.ptrs[q] + lay["recv"] + 2 * self.rank * self.chunk, me + lay["bucket"] + 2 * lo, 2 * cnt))
        self._fan_out(pushes)
        st = _lib.stream_ptr()
        _lib.check(self.lib.ed2_peer_signal(self._real, None, st), "ed2_peer_signal")
        _lib.check(self.lib.ed2_peer_wait(self._real, 0, st), "ed2_peer_wait")  # one warp: holds no SM while peers push
        _lib.check(self.lib.ed2_peer_reduce_slice(self._local, 0, st), "ed2_peer_reduce_slice")
        lo, cnt = self._slice(self.rank)
        self._fan_out([(self.arena.ptrs[(self.rank + d) % self.world] + lay["gathered"] + 2 * lo, me + lay["gathered"] + 2 * lo,
                        2 * cnt) for d in range(1, self.world)])
        _lib.check(self.lib.ed2_peer_flag(self._real, 1, _lib.stream_ptr()), "ed2_peer_flag")

    def gather_f32(self, out):
        from . import _lib

        _lib.check(self.lib.ed2_peer_wait(self._real, 1, _lib.stream_ptr()), "ed2_peer_wait")
        _lib.check(self.lib.ed2_peer_gather_f32(self._local, out.data_ptr(), _lib.stream_ptr()), "ed2_peer_gather_f32")

    def close(self):
        self.arena.close()


class GradientAllReduce:
    """Overlapped gradient all-reduce: a post-accumulate hook per parameter launches an async sum-all-reduce on the
    process group's own stream the moment that gradient is final; `wait()` joins them before the optimizer step.

    compress=torch.bfloat16 sends bf16 buckets (half the NVLink bytes): the hook casts the fp32
    gradient into a persistent bf16 bucket with the library's cast kernel, all-reduces the bucket, and `wait()` casts
    the sum back into `p.grad`."""

    def __init__(self, params, group=None, average: bool = False, compress=None, early: bool = True,
                 exchange: str = "peer", hook_exchange: str = "nccl"):
        """exchange='peer' (default when compress=bf16 and the ranks share one box): the raw bf16 weight gradients of
        DenseReparameterization kernels and their bias gradients are exchanged through IPC-mapped peer memory by this
        library's own kernels (PeerExchange) — no NCCL call on that path; 'nccl': async NCCL all-reduce of the bucket."""
        self.handles, self.group, self.average, self.compress = [], group, average, compress
        self.exchange_mode, self._exchanges, self._side = exchange, {}, []
        # hook_exchange='peer': large fp32 gradients that reach the per-parameter hook (deterministic kernels of
        # DenseBatchEnsemble / DenseRank1 / Dense, Flipout's dmu / drho) are cast to bf16 into a peer arena and summed by
        # the same signal / owner-pass kernels, then gathered back as fp32 — no NCCL call for them either.
        self.hook_exchange = hook_exchange
        # gradients below 64 K elements are packed and summed by ONE all-reduce at wait() instead of one NCCL call each
        # (a rank-1 MLP has 15 of them per step; each call is a kernel that must find SMs between persistent GEMMs)
        self.flat_small, self._small, self._flat = True, [], {}
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._hooks, self._buckets, self._early_done = [], {}, set()
        self._ids = set()
        if self.world > 1:
            for p in params:
                if p.requires_grad:
                    self._ids.add(id(p))
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._hook))
            from . import functional as F

            if not average:
                F.GradSync.replicas = self.world  # KL gradients are added once, not once per rank (functional.GradSync)
            if early and compress is not None:
                F.GradSync.current = self

    # -- early path: called by the layer's backward between its parameter-gradient phase and its dX GEMM -------------
    def wants(self, p) -> bool:
        # average=True is applied by the per-parameter hook only: the raw-dW paths (which finish dmu / drho after the
        # exchange) are sum-only, so an averaging reducer keeps every parameter on the hook path
        return (self.world > 1 and self.compress is not None and not self.average and id(p) in self._ids
                and p.numel() >= 65536)

    def wants_peer(self, p) -> bool:
        return (self.exchange_mode == "peer" and self.wants(p) and not self.average and p.numel() % 8 == 0
                and p.shape[-1] % 8 == 0 and self.world <= 8)  # (also the gate of the hook-path peer / dma exchanges)

    def peer_exchange(self, p, small_n: int) -> PeerExchange:
        """The (lazily built, collective) exchange of parameter `p`; every rank reaches this in the same order."""
        ex = self._exchanges.get(id(p))
        if ex is None or ex.small_n != small_n:
            try:
                ex = PeerExchange(p.numel(), small_n, self.group)
            except RuntimeError as e:  # raised on EVERY rank (PeerArena agrees collectively): NCCL path from now on
                import sys

                sys.stderr.write(f"[edward2_b200.dist] {e}; falling back to the NCCL all-reduce of the bf16 buckets\n")
                self.exchange_mode = "nccl"
                return None
            self._exchanges[id(p)] = ex
        return ex

    def dma_exchange(self, p):
        """The (lazily built, collective) copy-engine exchange of parameter `p`."""
        ex = self._exchanges.get(id(p))
        if ex is None:
            try:
                ex = DmaExchange(p.numel(), self.group)
            except RuntimeError as e:  # raised on EVERY rank (PeerArena agrees collectively): NCCL path from now on
                import sys

                sys.stderr.write(f"[edward2_b200.dist] {e}; falling back to the NCCL all-reduce of the bf16 buckets\n")
                self.hook_exchange = "nccl"
                return None
            self._exchanges[id(p)] = ex
        return ex

    def wants_dma(self, p) -> bool:
        return self.hook_exchange == "dma" and self.wants_peer(p)

    def exchange_bucket(self, p, ex, after=None):
        """`ex`'s bucket already holds this rank's bf16 gradient of `p` (written by the layer's dW GEMM): run the
        exchange on the side streams and install the summed fp32 gradient as p.grad (bypassing the hook).  `after`:
        event recorded right behind the dW GEMM — the exchange then overlaps whatever the layer enqueued after it."""
        g = torch.empty(p.shape, dtype=torch.float32, device=p.device)
        # every exchange has its own pair of side streams: the chain of a later (smaller) layer does not queue behind an
        # earlier layer's second DMA phase
        lane = 2 + 2 * (list(self._exchanges).index(id(p)) % 3)
        reduced_ev = self.on_side_stream(ex.scatter_reduce, which=lane, fork=after)
        done_ev = self.on_side_stream(lambda: ex.gather_f32(g), which=lane + 1, after=reduced_ev)
        self.defer(None, after=done_ev, keep=(g,))
        self.mark_done(p)
        if p.grad is None:
            p.grad = g
        else:  # accumulate only once the sum has landed
            self.defer(lambda: p.grad.add_(g), after=done_ev, keep=(g,))

    def defer(self, finish, after=None, keep=None):
        """At wait(): make the current stream wait for the CUDA event `after`, then run `finish` (if any).  `keep` holds
        tensors that side-stream kernels read, alive until the streams have joined."""
        self.handles.append((after, finish, keep))

    def mark_done(self, *params):
        """These parameters' gradients are reduced and installed by the caller this step: their hooks must not
        all-reduce them again."""
        for p in params:
            if p is not None:
                self._early_done.add(id(p))

    def on_side_stream(self, fn, which: int = 0, after=None, fork=None):
        """Run `fn` (kernel launches) on side stream `which` of the reducer, ordered after everything enqueued so far on
        the current stream (and after the event `after`); returns the event that marks its completion.  Capturable in a
        CUDA graph as long as wait() joins the returned event back."""
        main = torch.cuda.current_stream()
        while len(self._side) <= which:
            self._side.append(torch.cuda.Stream(priority=-1))
        side = self._side[which]
        if fork is None:  # ordered after everything enqueued so far on the current stream
            fork = torch.cuda.Event()
            fork.record(main)
        with torch.cuda.stream(side):
            side.wait_event(fork)
            if after is not None:
                side.wait_event(after)
            fn()
            done = torch.cuda.Event()
            done.record(side)
        return done

    def bucket(self, p):
        b = self._buckets.get(id(p))
        if b is None:
            b = torch.empty(p.shape, dtype=self.compress, device=p.device)
            self._buckets[id(p)] = b
        return b

    def early(self, p, grad, bucket):
        """`bucket` already holds the bf16 copy of `grad` (written by the gradient-finishing kernel)."""
        h = dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        g2 = grad.reshape(-1, grad.shape[-1]) if grad.dim() > 1 else grad.reshape(1, -1)
        self.handles.append((h, g2, bucket.reshape(g2.shape)))
        self._early_done.add(id(p))

    def early_raw(self, params, bucket, finish):
        """All-reduce a raw bf16 weight-gradient bucket shared by several parameters (mean and stddev of one kernel);
        `finish()` turns the reduced bucket into their fp32 gradients at wait()."""
        h = dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.handles.append((h, finish, None))

    def _hook(self, p):
        if id(p) in self._early_done:  # its all-reduce is already in flight
            self._early_done.discard(id(p))
            return
        g = p.grad
        if g is None:  # this parameter's gradient is installed later by a deferred finishing pass (early_raw)
            return
        if self.average:
            g.div_(self.world)
        if (self.hook_exchange == "dma" and self.wants_peer(p) and g.is_cuda and g.dtype == torch.float32
                and g.is_contiguous() and g.data_ptr() % 16 == 0):
            ex = self.dma_exchange(p)
            if ex is not None:
                g2 = g.reshape(-1, g.shape[-1]) if g.dim() > 1 else g.reshape(1, -1)
                ex.cast_into_bucket(g2)
                reduced_ev = self.on_side_stream(ex.scatter_reduce, which=0)
                done_ev = self.on_side_stream(lambda ex=ex, g=g: ex.gather_f32(g), which=1, after=reduced_ev)
                self.defer(None, after=done_ev, keep=(g,))
                return
        if (self.hook_exchange == "peer" and self.wants_peer(p) and g.is_cuda and g.dtype == torch.float32
                and g.is_contiguous() and g.data_ptr() % 16 == 0):
            ex = self.peer_exchange(p, 0)
            if ex is not None:
                g2 = g.reshape(-1, g.shape[-1]) if g.dim() > 1 else g.reshape(1, -1)
                ex.cast_into_bucket(g2)

                def exchange(ex=ex):
                    ex.signal(None)
                    ex.reduce_slice()

                reduced_ev = self.on_side_stream(exchange, which=0)
                done_ev = self.on_side_stream(lambda ex=ex, g=g: ex.gather_f32(g), which=1, after=reduced_ev)
                self.defer(None, after=done_ev, keep=(g,))
                return
        if g.is_cuda and g.dtype == torch.float32 and g.numel() < 65536 and self.flat_small:
            # small gradients (fast weights, biases): packed into ONE flat fp32 buffer, one all-reduce at wait()
            self._small.append((p, g))
            return
        if self.compress is not None and g.is_cuda and g.dtype == torch.float32 and g.numel() >= 65536:
            from . import ops

            g2 = g.reshape(-1, g.shape[-1]) if g.dim() > 1 else g.reshape(1, -1)
            b = self.bucket(p).reshape(g2.shape)
            ops.cast(g2, self.compress, out=b)
            h = dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.handles.append((h, g2, b))
        else:
            self.handles.append((dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group, async_op=True), None, None))

    def _reduce_small(self):
        if not self._small:
            return
        total = sum(g.numel() for _, g in self._small)
        flat = self._flat.get(total)
        if flat is None:
            flat = self._flat[total] = torch.empty(total, dtype=torch.float32, device=self._small[0][1].device)
        views, o = [], 0
        for _, g in self._small:
            views.append(flat[o:o + g.numel()].view(g.shape))
            o += g.numel()
        torch._foreach_copy_(views, [g for _, g in self._small])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        for (p, _), v in zip(self._small, views):
            p.grad = v  # the summed gradient lives in the flat buffer until the next step's reduction
        self._small.clear()

    def wait(self):
        self._reduce_small()
        for h, g2, b in self.handles:
            if isinstance(h, torch.cuda.Event):  # side-stream work of the peer exchange
                torch.cuda.current_stream().wait_event(h)
                if callable(g2):
                    g2()
                continue
            h.wait()
            if callable(g2):
                g2()  # deferred gradient-finishing pass on the reduced bucket
            elif b is not None:
                from . import ops

                ops.cast(b, torch.float32, out=g2)
        self.handles.clear()
        for ex in self._exchanges.values():
            ex._pending = False

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks.clear()
        from . import functional as F

        if F.GradSync.current is self:
            F.GradSync.current = None
        if self.world > 1:
            F.GradSync.replicas = 1
        for ex in self._exchanges.values():
            ex.close()
        self._exchanges.clear()


class DeferredSum:
    """A replicated tensor that receives rank-local LINEAR updates  x <- a * x + u_r  (the Laplace precision matrix:
    P <- m P + (1-m) Phi_r^T Phi_r / B_global, or P + Phi_r^T Phi_r) and whose true value is the sum over ranks of what
    each rank WOULD have added.  Instead of one all-reduce per training step, the ranks accumulate locally and the sum
    is taken once, when the value is needed (the inverse at the train -> eval transition).  Linearity makes the result
    identical (up to fp32 summation order) to all-reducing every step: rank 0 carries the value from before the window,
    the other ranks start the window from zero, and  sum_r x_r  after any number of local updates is the global value.
    Removes the per-step 268 MB exchange of the D = 8192 head from the training step."""

    def __init__(self, tensor: torch.Tensor, group=None):
        self.tensor, self.group, self.open = tensor, group, False

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def begin_local_update(self):
        """Call before every rank-local update of the tensor."""
        if self._world() > 1 and not self.open:
            if dist.get_rank(self.group) != 0:
                self.tensor.zero_()
            self.open = True

    def synchronize(self):
        """Collective: after this every rank holds the global value.  No-op when nothing was accumulated."""
        if self.open:
            dist.all_reduce(self.tensor, op=dist.ReduceOp.SUM, group=self.group)
            self.open = False
        return self.tensor


def allreduce_precision_partial(partial: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the local ΦᵀΦ partials over ranks (in place); the blend P <- m P + (1-m) sum / B_global follows."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial
