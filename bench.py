#!/usr/bin/env python
"""bench.py — samples/s for one forward + backward step of the Edward2 stochastic-weight hot path on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` (under torchrun for N > 1)
prints ONE JSON line from rank 0.  Workload at N=1 = BASELINE.json configs[1]: Bayesian MLP 1024-4096-4096-1000,
batch 4096, bf16 compute, DenseReparameterization layers with the analytic KL regularizer (the DenseFlipout variant
of the same config is measured in the same run and reported under "variants").  `--impl reference` times the
torch-CPU restatement of the reference's CPU path (oracle/torch_cpu.py; the reference itself cannot run here).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = [1024, 4096, 4096, 1000]
BATCH = 4096
KL_SCALE = 1.0 / 50000
METRIC = "samples/sec fwd+bwd"
NCU_GEMM_DRAM_BYTES_PER_LAUNCH = 67.7e6  # (464.7 MB read + 76.7 MB written) / 8 launches, profiles/r01_ncu_summary.csv
WORKLOAD = "cfg2: DenseReparameterization Bayesian MLP 1024-4096-4096-1000 + KL, batch 4096/GPU, bf16"


def flops_per_step(batch: int, flipout: bool) -> float:
    """fwd + dW for every layer, dX for layers 2..L; Flipout doubles every product."""
    f = 0.0
    for li, (i, o) in enumerate(zip(DIMS[:-1], DIMS[1:])):
        f += 2.0 * batch * i * o * 2          # forward + dW
        if li > 0:
            f += 2.0 * batch * i * o          # dX
    return f * (2.0 if flipout else 1.0)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=d.get("bf16_tflops_sustained", 1400.0), bf16_burst=d.get("bf16_tflops", 1590.0),
                    hbm=d.get("hbm_gbs", 6650.0), src="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(bf16=1400.0, bf16_burst=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples nvidia-smi SM clocks and throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """`--impl reference`: the reference's CPU path, restated (kind "port"), on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import torch_cpu

    # the config's own batch when the run stays within a few minutes (~1 s per 4096-row step on the GPU box's host
    # cores), a 1024-row sample of it for long runs
    sample_batch = BATCH if args.steps + args.warmup <= 40 else 1024
    sps, dt, threads = torch_cpu.time_bayesian_mlp(DIMS, sample_batch, False, max(1, args.steps), max(1, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "torch-CPU fp32 restatement of the reference's CPU path "
                   f"(TF/TFP/JAX not installable here); each step = {sample_batch} rows of the 4096-row batch"},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{sample_batch}-row batches of the cfg2 MLP, {args.steps} fwd+bwd steps"},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------------
def build_model(ed, flipout: bool, device, step_counter):
    import torch

    cls = ed.layers.DenseFlipout if flipout else ed.layers.DenseReparameterization
    layers = []
    for li, units in enumerate(DIMS[1:]):
        reg = ed.regularizers.NormalKLDivergence(scale_factor=KL_SCALE)
        layers.append(cls(units, activation="relu" if li < len(DIMS) - 2 else None, kernel_regularizer=reg,
                          dtype="bfloat16", seed=2026 + li))
    x0 = torch.zeros(8, DIMS[0], dtype=torch.bfloat16, device=device)
    h = x0
    with torch.no_grad():
        for l in layers:
            h = l(h)
    for l in layers:
        l.step_counter = step_counter
    return layers


def make_step(layers, global_batch):
    from edward2_b200 import functional as F

    def step(x, labels):
        # scheduling hint: draw the kernel samples of the later layers on side streams now, so the issue-bound sampling
        # passes overlap the first layers' GEMMs (same Philox streams as without the hint)
        # (the first layer's sample goes first and alone: the first GEMM waits for it; the others follow beside that GEMM)
        if hasattr(layers[0], "presample"):
            layers[0].presample()
        for l in layers[1:]:
            if hasattr(l, "presample"):
                l.presample(after=layers[0])
        h = x
        for l in layers:
            h = l(h)
        # cross-entropy + the layers' KL regularisers, summed by one kernel
        loss = F.SoftmaxCrossEntropy.apply(h, labels, global_batch, *[l.losses[0] for l in layers])
        F.backward(loss)
        return loss.detach()

    return step


def params_of(layers):
    return [p for l in layers for p in l.parameters()]


def run_ours(args):
    import torch
    import torch.distributed as dist

    import edward2_b200 as ed
    from edward2_b200 import _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and args.gemm_ctas > 0:
        # leave SMs for the NCCL all-reduce kernels (read by libed2b200 at its first GEMM launch)
        os.environ.setdefault("ED2_GEMM_MAX_CTAS", str(args.gemm_ctas))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    lib = _lib.load()
    _lib.check(lib.ed2_check_device(local), "ed2_check_device")
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    global_batch = BATCH * world
    pk = peaks()

    gen = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(BATCH, DIMS[0], generator=gen).to(torch.bfloat16).pin_memory()
    y_host = torch.randint(0, DIMS[-1], (BATCH,), generator=gen).pin_memory()
    x_dev, y_dev = x_host.to(device), y_host.to(device)
    step_counter = torch.zeros((), dtype=torch.int64, device=device)

    def measure(flipout: bool, steps: int, warmup: int, want_e2e: bool):
        layers = build_model(ed, flipout, device, step_counter)
        params = params_of(layers)
        reducer = None
        if world > 1 and not args.no_allreduce:
            # gradient all-reduce launched as soon as each parameter's gradient is final, overlapping the rest of
            # the backward pass (NCCL runs on its own stream); bf16 buckets halve the NVLink bytes
            from edward2_b200.dist import GradientAllReduce

            reducer = GradientAllReduce(params, compress=None if args.fp32_allreduce else torch.bfloat16,
                                        exchange=args.grad_exchange)
        raw_step = make_step(layers, global_batch)

        def zero():
            for p in params:
                p.grad = None

        use_graph = args.graph in ("on", "auto")
        graph, static_loss = None, None

        def eager_step(x, y):
            zero()
            step_counter.add_(1)
            loss = raw_step(x, y)
            if reducer is not None:
                reducer.wait()
            return loss

        x_static, y_static = x_dev.clone(), y_dev.clone()
        # every pre-capture step runs on the capture-side stream so that autograd's AccumulateGrad nodes are bound
        # to it (a node created on the legacy stream would invalidate the capture)
        # high priority: when an SM has room, a GEMM CTA of the main stream is placed before side-stream blocks
        side = torch.cuda.Stream(priority=-1) if use_graph else torch.cuda.current_stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            eager_step(x_static, y_static)  # builds workspaces, sets kernel attributes
            torch.cuda.synchronize()
            a0 = _lib.launch_count()
            eager_step(x_static, y_static)
            launches_per_step = _lib.launch_count() - a0
            eager_step(x_static, y_static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if use_graph:
            zero()
            try:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    step_counter.add_(1)
                    static_loss = raw_step(x_static, y_static)
                    if reducer is not None:  # join the NCCL all-reduces into the captured step
                        reducer.wait()
            except Exception as e:  # noqa: BLE001  (e.g. a collective that cannot be captured) -> eager steps
                sys.stderr.write(f"[bench] CUDA-graph capture failed, running eagerly: {str(e)[:300]}\n")
                graph, static_loss = None, None
                if reducer is not None:
                    reducer.handles.clear()
                torch.cuda.synchronize()

        def run_step():
            if graph is not None:
                graph.replay()
                return static_loss
            return eager_step(x_static, y_static)

        def sync_all():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()

        for _ in range(max(warmup, 3)):
            run_step()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            run_step()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1) / steps
        # kernels of libed2b200 launched in the timed region (a graph replay re-launches the captured ones)
        launches = launches_per_step * steps
        t = torch.tensor([ms], device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

        if args.trace and not flipout:
            # kernel timeline of 3 steps (CUPTI through torch.profiler) for offline analysis; never a bench number
            import contextlib

            from torch.profiler import ProfilerActivity, profile

            with (profile(activities=[ProfilerActivity.CUDA]) if rank == 0 else contextlib.nullcontext()) as prof:
                for _ in range(3):
                    run_step()
                sync_all()
            if rank == 0:
                prof.export_chrome_trace(args.trace)
            sync_all()

        e2e_ms = None
        if want_e2e:
            # end to end through the public layer API with HOST buffers: every step copies ITS batch from pinned host
            # memory (H2D on a copy stream into one of two staging slots, overlapping the previous step's compute),
            # runs the step and reads the loss back (D2H)
            loss_host = torch.empty((), dtype=torch.float64).pin_memory()
            copy_stream = torch.cuda.Stream()
            xs, ys = [x_dev.clone(), x_dev.clone()], [y_dev.clone(), y_dev.clone()]
            ev_h2d = [torch.cuda.Event(), torch.cuda.Event()]
            ev_used = [torch.cuda.Event(), torch.cuda.Event()]
            main = torch.cuda.current_stream()
            for ev in ev_used:
                ev.record(main)

            def prefetch(i):
                slot = i % 2
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(ev_used[slot])
                    xs[slot].copy_(x_host, non_blocking=True)
                    ys[slot].copy_(y_host, non_blocking=True)
                    ev_h2d[slot].record(copy_stream)

            def e2e_loop(n):
                prefetch(0)
                for i in range(n):
                    prefetch(i + 1)
                    slot = i % 2
                    main.wait_event(ev_h2d[slot])
                    x_static.copy_(xs[slot], non_blocking=True)
                    y_static.copy_(ys[slot], non_blocking=True)
                    ev_used[slot].record(main)
                    loss_host.copy_(run_step().to(torch.float64), non_blocking=True)

            e2e_loop(2)
            sync_all()
            e0.record()
            e2e_loop(steps)
            e1.record()
            sync_all()
            t = torch.tensor([e0.elapsed_time(e1) / steps], device=device)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        loss_val = float(run_step().item())
        exchange = None
        if reducer is not None:
            exchange = reducer.exchange_mode  # "nccl" if the peer-memory arena could not be built on this box
            reducer.remove()
        return dict(ms=ms, e2e_ms=e2e_ms, launches=launches, loss=loss_val, graph=graph is not None, layers=layers,
                    exchange=exchange)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    main = measure(False, args.steps, args.warmup, True)
    clocks = sampler.stop() if sampler else None
    del main["layers"]
    torch.cuda.empty_cache()
    flip = None
    if not args.skip_variants:
        try:
            flip = measure(True, max(3, args.steps // 2), 3, False)
            del flip["layers"]
        except Exception as e:  # noqa: BLE001
            flip = {"error": str(e)[:200]}
        torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel (tcgen05 GEMM): the 8 GEMM launches of one step, timed alone
    roof = None
    try:
        shapes = []
        for li, (i, o) in enumerate(zip(DIMS[:-1], DIMS[1:])):
            shapes.append(("fwd", BATCH, o, i, ops.MAJOR_K, ops.MAJOR_MN))
            shapes.append(("dW", i, o, BATCH, ops.MAJOR_MN, ops.MAJOR_MN))
            if li > 0:
                shapes.append(("dX", BATCH, i, o, ops.MAJOR_K, ops.MAJOR_K))
        bufs = []
        for (_, m, n, k, ma, mb) in shapes:
            a = torch.randn((m, k) if ma == ops.MAJOR_K else (k, m), device=device).to(torch.bfloat16)
            b = torch.randn((n, k) if mb == ops.MAJOR_K else (k, n), device=device).to(torch.bfloat16)
            out = _lib.empty_padded(m, n, torch.float32 if _ == "dW" else torch.bfloat16, device)
            bufs.append((a, b, out))
        def gemms():
            for (_, m, n, k, ma, mb), (a, b, out) in zip(shapes, bufs):
                ops.gemm(a, b, m, n, k, major_a=ma, major_b=mb, out=out)
        for _ in range(3):
            gemms()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            gemms()
        e1.record()
        torch.cuda.synchronize()
        gemm_ms = e0.elapsed_time(e1) / reps
        fl = flops_per_step(BATCH, False)
        achieved = fl / (gemm_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "ed2::tc::gemm_bf16_kernel (8 launches per step)", "achieved": achieved,
                "peak": pk["bf16"], "unit": "TFLOP/s", "frac": achieved / pk["bf16"],
                # DRAM bytes per GEMM launch (read + write, mean of the step's 8 launches) from the committed
                # `ncu --set full` capture of this workload (profiles/r01_ncu_summary.csv); algorithmic operand + output
                # bytes are 83.7 MB per launch, the difference is operands still resident in the 126 MB L2
                "traffic": NCU_GEMM_DRAM_BYTES_PER_LAUNCH, "traffic_unit": "bytes/launch (ncu capture, profiles/)",
                "algorithmic_bytes_per_launch": 83.7e6,
                "peak_source": pk["src"], "gemm_ms_per_step": gemm_ms, "algorithmic_flops_per_step": fl,
                "gemm_share_of_step": gemm_ms / main["ms"]}
        del bufs
    except Exception as e:  # noqa: BLE001
        roof = {"error": str(e)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.skip_cpu:
        try:
            from oracle import torch_cpu

            sps, dt, threads = torch_cpu.time_bayesian_mlp(DIMS, BATCH, False, steps=4, warmup=1)
            cpu = {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port",
                   "sample": f"4 fwd+bwd steps on {BATCH}-row batches of the cfg2 MLP (torch-CPU fp32 restatement)",
                   "s_per_step": dt}
        except Exception as e:  # noqa: BLE001
            cpu = {"error": str(e)[:200]}

    value = global_batch / (main["ms"] * 1e-3)
    fl = flops_per_step(BATCH, False)
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": main["ms"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": global_batch, "parallelism": f"dp{world}",
                   "l2": "inputs > L2: one step touches ~0.6 GB (fp32 mu/rho + grads 400 MB, bf16 W 50 MB, activations)",
                   "cuda_graph": main["graph"], "kl_scale": KL_SCALE,
                   "grad_allreduce": None if world == 1 else (
                       "fp32 NCCL all-reduce per parameter, overlapped with backward" if args.fp32_allreduce else
                       "raw bf16 dW through IPC peer memory: pull reduce-scatter + all-gather fused into the finishing "
                       "pass (own kernels, no NCCL on the data path)" if main.get("exchange") == "peer" else
                       "bf16 buckets, NCCL all-reduce per parameter, overlapped with backward"),
                   "gemm_ctas": int(os.environ.get("ED2_GEMM_MAX_CTAS", 148))},
        "e2e": {"value": global_batch / (main["e2e_ms"] * 1e-3), "unit": "samples/s",
                "h2d_bytes_per_step": (x_host.numel() * 2 + y_host.numel() * 8) * world,
                "d2h_bytes_per_step": 8 * world, "ms_per_step": main["e2e_ms"]},
        "gpu_launches": main["launches"],
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "step_tflops": fl * world / (main["ms"] * 1e-3) / 1e12,
        "step_frac_of_tensor_peak": fl / (main["ms"] * 1e-3) / 1e12 / pk["bf16"],
        "loss": main["loss"],
        "variants": {"flipout": None if flip is None else (
            flip if "error" in flip else {"value": global_batch / (flip["ms"] * 1e-3), "ms_per_step": flip["ms"],
                                          "step_tflops": flops_per_step(BATCH, True) * world / (flip["ms"] * 1e-3) / 1e12,
                                          "loss": flip["loss"]})},
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else libraries print (e.g. NCCL's version banner) was
    diverted to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # stray prints of native libraries must not pollute the one-line contract
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"])
    ap.add_argument("--trace", default="", help="write a chrome trace of 3 steps on rank 0 (diagnostics)")
    ap.add_argument("--fp32-allreduce", action="store_true", help="all-reduce fp32 gradients instead of bf16 buckets")
    ap.add_argument("--grad-exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how the raw bf16 weight gradients are summed over ranks")
    ap.add_argument("--gemm-ctas", type=int, default=0, help="persistent GEMM CTAs when N > 1 (rest: NCCL)")
    ap.add_argument("--no-allreduce", action="store_true", help="diagnostic only: skip the gradient all-reduce")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-variants", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
