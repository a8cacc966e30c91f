#!/usr/bin/env python
"""bench.py — samples/s for one forward + backward step of the Edward2 stochastic-weight / SNGP hot path on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` (under torchrun for N > 1)
prints ONE JSON line from rank 0.

Headline workload = BASELINE.json configs[3], the BatchEnsemble-family config that carries the scaling target:
  cfg4: DenseRank1 MLP 2048-8192-8192-1000, ensemble_size 4, sampled r/s (in-kernel Philox, global-row offsets),
        2048 rows per GPU (weak scaling: 16384 rows on 8 GPUs), bf16 compute, KL(alpha, gamma || N(1, 0.1)).
The same run also measures, at N = 1, every other configuration of BASELINE.json (cfg1 BatchEnsemble MLP, cfg2
Reparameterization and Flipout MLPs, cfg3 SNGP head train/eval, cfg5 per-GPU shard train/eval) and reports each
under "configs" with its own ms_per_step, roofline and CPU baseline.

`--impl reference` times the torch-CPU restatement of the reference's CPU path for the headline workload
(oracle/torch_cpu.py; TensorFlow / TFP / JAX are not installable here, so the reference itself cannot run).
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

KL_SCALE = 1.0 / 50000
METRIC = "samples/sec fwd+bwd"
HEADLINE = "cfg4"
# DRAM bytes per tcgen05 GEMM launch from the committed `ncu --set full` captures (profiles/); None = not captured
NCU_TRAFFIC = {"cfg2_reparam": 67.7e6,   # profiles/r01_ncu_summary.csv (8 launches; 83.7 MB algorithmic per launch)
               "cfg4": 164.7e6}          # profiles/r02_ncu_rank1_fused_summary.csv (9 launches; 168 MB algorithmic)
NCU_ALGO_BYTES = {"cfg2_reparam": 83.7e6, "cfg4": 168.0e6}


def mlp_flops(dims, batch, mult=1.0):
    """fwd + dW for every layer, dX for layers 2..L."""
    f = 0.0
    for li, (i, o) in enumerate(zip(dims[:-1], dims[1:])):
        f += 2.0 * batch * i * o * (3 if li > 0 else 2)
    return f * mult


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_sustained=d.get("bf16_tflops_sustained", 1400.0), bf16=d.get("bf16_tflops", 1590.0),
                    hbm=d.get("hbm_gbs", 6650.0), src="measured (MEASURED_PEAKS.json)")
    return dict(bf16_sustained=1400.0, bf16=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


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
# Workloads.  Each builds its layers on `device` and returns a dict:
#   step(x, y) -> loss tensor (one forward + backward, or one eval pass), params, batch (rows per GPU), in_dim,
#   x_dtype, classes, flops (algorithmic FLOPs per GPU per step), cpu(steps) -> (samples/s, s/step,
#   threads, sample description), gemm_kernel (name fragment of the dominant kernel), dtype
# ------------------------------------------------------------------------------------------------------------------
CFG4_DIMS = [2048, 8192, 8192, 1000]
CFG2_DIMS = [1024, 4096, 4096, 1000]
CFG1_DIMS = [784, 512, 512, 10]


def _warm_build(layers, x0):
    import torch

    h = x0
    with torch.no_grad():
        for l in layers:
            h = l(h)


def wl_cfg4(ed, device, step_counter, rank, world):
    import torch

    from edward2_b200 import functional as F

    E, B = 4, 2048
    mk = lambda u, act, li: ed.layers.DenseRank1(  # noqa: E731
        u, activation=act, ensemble_size=E, dtype="bfloat16", seed=2026 + li,
        alpha_initializer=ed.initializers.TrainableNormal(
            mean_initializer=ed.initializers.RandomSign(seed=10 + li),
            stddev_initializer=ed.initializers.Constant(-3.4402)),  # softplus^-1(sqrt(p/(1-p))), p = 1e-3
        gamma_initializer=ed.initializers.TrainableNormal(
            mean_initializer=ed.initializers.RandomSign(seed=20 + li),
            stddev_initializer=ed.initializers.Constant(-3.4402)),
        alpha_regularizer=ed.regularizers.NormalKLDivergence(mean=1.0, stddev=0.1, scale_factor=KL_SCALE),
        gamma_regularizer=ed.regularizers.NormalKLDivergence(mean=1.0, stddev=0.1, scale_factor=KL_SCALE))
    layers = [mk(8192, "relu", 0), mk(8192, "relu", 1), mk(1000, None, 2)]
    _warm_build(layers, torch.zeros(8, CFG4_DIMS[0], dtype=torch.bfloat16, device=device))
    layers[0].feeds(layers[1]).feeds(layers[2])  # stacked rank-1 layers: cross-layer epilogue fusion (Rank1Link)
    for l in layers:
        l.step_counter = step_counter
    model = torch.nn.ModuleList(layers)
    if world > 1:
        from edward2_b200.dist import shard_noise

        shard_noise(model, rank, world)  # r / s noise is drawn for the GLOBAL row: same result for any N
    gb = B * world
    side = torch.cuda.Stream(device=device)

    def step(x, y):
        # scheduling hint: the fast weights' KL regularisers depend on parameters only — evaluated on a side stream
        # (their backward then runs there too), their small kernels fill the tails of the persistent GEMMs
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            kls = [k for l in layers for k in l.losses]
        # ... and so do the bf16 working copies of the kernels: each is cast on a side stream while the PREVIOUS layer's
        # GEMM runs (queued right after that GEMM so its CTAs are placed first; the cast takes the SMs the GEMM grid
        # leaves free)
        # Measured alternative (layer.gemm_ready(): the cast becomes runnable together with the previous layer's GEMM, whose
        # CTAs are then placed first): 1.251 ms against 1.232 ms — a bandwidth-bound cast gets ~15 % of its work done on
        # the 20 SMs a 64-pair GEMM grid leaves free, so letting it run at full speed just before that GEMM is faster.
        start = torch.cuda.Event()
        start.record(main)
        layers[0].precast(ready=start)
        h = x
        for i, l in enumerate(layers):
            h = l(h)
            if i + 1 < len(layers):
                layers[i + 1].precast(after=layers[i] if i == 0 else None, ready=start)
        main.wait_stream(side)
        loss = F.SoftmaxCrossEntropy.apply(h, y, gb, *kls)
        F.backward(loss)
        return loss.detach()

    def cpu(steps):
        from oracle import torch_cpu

        sps, dt, th = torch_cpu.time_rank1_mlp(CFG4_DIMS, B, steps, warmup=1, ensemble_size=E)
        return sps, dt, th, f"{steps} fwd+bwd steps on {B}-row batches of the cfg4 rank-1 MLP (torch-CPU fp32 restatement)"

    return dict(step=step, params=list(model.parameters()), batch=B, in_dim=CFG4_DIMS[0], x_dtype=torch.bfloat16,
                classes=1000, flops=mlp_flops(CFG4_DIMS, B), cpu=cpu, dtype="bf16", modules=model,
                workload="cfg4: DenseRank1 MLP 2048-8192-8192-1000, ensemble_size 4, sampled r/s + KL, "
                         "2048 rows/GPU (16384 on 8 GPUs), bf16")


def wl_cfg2(flipout):
    def build(ed, device, step_counter, rank, world):
        import torch

        from edward2_b200 import functional as F

        B = 4096
        cls = ed.layers.DenseFlipout if flipout else ed.layers.DenseReparameterization
        layers = []
        for li, units in enumerate(CFG2_DIMS[1:]):
            reg = ed.regularizers.NormalKLDivergence(scale_factor=KL_SCALE)
            layers.append(cls(units, activation="relu" if li < 2 else None, kernel_regularizer=reg, dtype="bfloat16",
                              seed=2026 + li))
        _warm_build(layers, torch.zeros(8, CFG2_DIMS[0], dtype=torch.bfloat16, device=device))
        for l in layers:
            l.step_counter = step_counter
        if flipout:  # scheduling hint: stacked layers share epilogues (functional.FlipoutLink); results unchanged
            for a, b in zip(layers[:-1], layers[1:]):
                a.feeds(b)
        model = torch.nn.ModuleList(layers)
        if world > 1:
            from edward2_b200.dist import shard_noise

            shard_noise(model, rank, world)
        gb = B * world

        def step(x, y):
            # scheduling hint: draw the kernel samples on side streams so the issue-bound sampling passes overlap the
            # first layers' GEMMs (same Philox streams as without the hint)
            layers[0].presample()
            for l in layers[1:]:
                l.presample(after=layers[0])
            h = x
            for l in layers:
                h = l(h)
            loss = F.SoftmaxCrossEntropy.apply(h, y, gb, *[l.losses[0] for l in layers])
            F.backward(loss)
            return loss.detach()

        def cpu(steps):
            from oracle import torch_cpu

            sps, dt, th = torch_cpu.time_bayesian_mlp(CFG2_DIMS, B, flipout, steps, warmup=1)
            return sps, dt, th, f"{steps} fwd+bwd steps on {B}-row batches (torch-CPU fp32 restatement)"

        return dict(step=step, params=list(model.parameters()), batch=B, in_dim=CFG2_DIMS[0], x_dtype=torch.bfloat16,
                    classes=1000, flops=mlp_flops(CFG2_DIMS, B, 2.0 if flipout else 1.0), cpu=cpu, dtype="bf16",
                    modules=model,
                    workload=f"cfg2: {'DenseFlipout' if flipout else 'DenseReparameterization'} Bayesian MLP "
                             f"1024-4096-4096-1000 + KL, batch 4096/GPU, bf16")

    return build


def wl_cfg1(ed, device, step_counter, rank, world):
    import torch

    from edward2_b200 import functional as F

    E, B = 4, 128
    mk = lambda u, act: ed.layers.DenseBatchEnsemble(u, ensemble_size=E, activation=act,  # noqa: E731
                                                     alpha_initializer="random_sign", gamma_initializer="random_sign")
    layers = [mk(512, "relu"), mk(512, "relu"), mk(10, None)]
    _warm_build(layers, torch.zeros(8, 784, device=device))
    model = torch.nn.ModuleList(layers)

    def step(x, y):
        h = x
        for l in layers:
            h = l(h)
        loss = F.SoftmaxCrossEntropy.apply(h, y, B)
        F.backward(loss)
        return loss.detach()

    def cpu(steps):
        from oracle import torch_cpu

        sps, dt, th = torch_cpu.time_batch_ensemble_mlp(CFG1_DIMS, B, max(steps, 50), warmup=5, ensemble_size=E)
        return sps, dt, th, f"{max(steps, 50)} fwd+bwd steps on the config's own 128-row batch (torch-CPU fp32 restatement)"

    return dict(step=step, params=list(model.parameters()), batch=B, in_dim=784, x_dtype=torch.float32, classes=10,
                flops=mlp_flops(CFG1_DIMS, B), cpu=cpu, dtype="f32", modules=model, bound="launch latency",
                min_bytes=9e6,
                workload="cfg1: DenseBatchEnsemble MLP 784-512-512-10, ensemble_size 4, batch 128, fp32")


def wl_sngp(kind, train):
    """cfg3: 2 x SpectralNormalization(Dense 512, relu) + RFGP(D=1024, 10 classes), batch 8192.
    cfg5 (per-GPU shard): RFGP D=8192 on 4096-dim inputs, 8192 rows per GPU, exact-sum precision."""

    def build(ed, device, step_counter, rank, world):
        import torch

        from edward2_b200 import functional as F

        if kind == "cfg3":
            B, d, D, widths, mom = 8192, 512, 1024, [512, 512], 0.999
        else:
            B, d, D, widths, mom = 8192, 4096, 8192, [], -1
        backbone = [ed.layers.SpectralNormalization(ed.layers.Dense(w, activation="relu", dtype="bfloat16"),
                                                    norm_multiplier=0.95) for w in widths]
        gp = ed.layers.RandomFeatureGaussianProcess(
            10, num_inducing=D, gp_cov_momentum=mom, gp_cov_ridge_penalty=1e-3, dtype="bfloat16",
            data_parallel="deferred" if (world > 1 and kind == "cfg5") else False)
        model = torch.nn.ModuleList(backbone + [gp])
        x0 = torch.randn(256, d, device=device).to(torch.bfloat16)
        model.train()
        with torch.no_grad():
            h = x0
            for l in backbone:
                h = l(h, training=True)
            gp(h, training=True)
        cov = gp._gp_cov_layer
        gb = B * world

        if train:
            def step(x, y):
                h = x
                for l in backbone:
                    h = l(h, training=True)
                logits, _ = gp(h, training=True)
                loss = F.SoftmaxCrossEntropy.apply(logits, y, gb)
                F.backward(loss)
                return loss.detach()
        else:
            model.eval()  # the one-off D^3 inverse happens in the first (untimed, pre-capture) call

            def step(x, y):
                with torch.no_grad():
                    h = x
                    for l in backbone:
                        h = l(h, training=False)
                    phi = gp.features(h)
                    logits = gp._gp_output_layer(phi).float() + gp._gp_output_bias
                    var, adj = cov.compute_predictive_variance(phi, logits, mean_field_factor=3.141592653589793 / 8.0)
                return adj.sum()

        bb = sum(2.0 * B * a * b for a, b in zip([d] + widths[:-1], widths))
        rff, out, pp = 2.0 * B * (widths[-1] if widths else d) * D, 2.0 * B * D * 10, 2.0 * B * D * D
        if train:
            # fwd + dW + dX of the backbone (dX for layer 2 only), RFF fwd + dX, output layer fwd+dW+dX, full ΦᵀΦ
            n_dx = max(0, len(widths) - 1)
            # (the RFF dX product runs in both configs: LayerNormalization's gamma / beta are trainable, TF semantics)
            flops = bb * 2 + (bb / max(1, len(widths))) * n_dx + rff * 2 + out * 3 + pp
        else:
            flops = bb + rff + out + pp

        def cpu(steps):
            from oracle import torch_cpu

            rows = B if kind == "cfg3" else 1024
            r = torch_cpu.time_sngp(d, widths, D, 10, mom, 1e-3, rows, steps, warmup=1)["train" if train else "eval"]
            return r[0], r[1], r[2], (f"{steps} {'train' if train else 'eval'} steps on {rows}-row batches "
                                      f"(torch-CPU fp32 restatement{'' if rows == B else ', linear in rows'})")

        return dict(step=step, params=list(model.parameters()), batch=B, in_dim=d, x_dtype=torch.bfloat16, classes=10,
                    flops=flops, cpu=cpu, dtype="bf16", modules=model, no_grad=not train,
                    workload=(f"{kind} {'train' if train else 'eval'}: " + (
                        "2x SpectralNormalization(Dense 512, relu) + RFGP(D=1024, 10 classes) + Laplace precision, batch 8192"
                        if kind == "cfg3" else
                        "RFGP D=8192 on 4096-dim inputs, 8192 rows per GPU (65536 on 8), exact-sum precision") +
                        (", predictive variance + mean-field logits" if not train else "") + ", bf16"))

    return build


def wl_conv_rank1(ed, device, step_counter, rank, world):
    """SURVEY.md §8f-1 (next row): two stacked Conv2DRank1 3x3 'same' layers at the width / resolution of a WRN-28-10
    group (160 channels, 32x32; experimental/rank1_bnns/cifar_model.py), batch 64, ensemble_size 4, bf16."""
    import torch

    from edward2_b200 import functional as F

    E, B, H, W, Cn = 4, 64, 32, 32, 160
    mk = lambda act, li: ed.layers.Conv2DRank1(  # noqa: E731
        Cn, 3, padding="same", activation=act, ensemble_size=E, dtype="bfloat16", seed=3030 + li,
        alpha_initializer=ed.initializers.TrainableNormal(
            mean_initializer=ed.initializers.RandomSign(seed=10 + li), stddev_initializer=ed.initializers.Constant(-3.4402)),
        gamma_initializer=ed.initializers.TrainableNormal(
            mean_initializer=ed.initializers.RandomSign(seed=20 + li), stddev_initializer=ed.initializers.Constant(-3.4402)),
        alpha_regularizer=ed.regularizers.NormalKLDivergence(mean=1.0, stddev=0.1, scale_factor=KL_SCALE),
        gamma_regularizer=ed.regularizers.NormalKLDivergence(mean=1.0, stddev=0.1, scale_factor=KL_SCALE))
    layers = [mk("relu", 0), mk(None, 1)]
    with torch.no_grad():
        h = torch.zeros(E, H, W, Cn, dtype=torch.bfloat16, device=device)
        for l in layers:
            h = l(h)
    for l in layers:
        l.step_counter = step_counter
    model = torch.nn.ModuleList(layers)

    def step(x, y):
        h = x.view(B, H, W, Cn)
        for l in layers:
            h = l(h)
        logits = h.float().mean(dim=(1, 2))  # synthetic head: global average pool (harness plumbing, not the layer path)
        loss = F.SoftmaxCrossEntropy.apply(logits, y, B * world, *[k for l in layers for k in l.losses])
        F.backward(loss)
        return loss.detach()

    def cpu(steps):
        from oracle import torch_cpu

        sps, dt, th = torch_cpu.time_rank1_conv(B, H, W, Cn, 2, steps, warmup=1, ensemble_size=E)
        return sps, dt, th, f"{steps} fwd+bwd steps on {B}-image batches (torch-CPU fp32 restatement, F.conv2d)"

    per_layer = 2.0 * B * H * W * 9 * Cn * Cn
    return dict(step=step, params=list(model.parameters()), batch=B, in_dim=H * W * Cn, x_dtype=torch.bfloat16, classes=Cn,
                flops=5 * per_layer, cpu=cpu, dtype="bf16", modules=model,
                workload="conv (SURVEY §8f-1): 2 x Conv2DRank1 3x3 'same', 160 channels, 32x32, batch 64, ensemble_size 4, "
                         "bf16; im2col + dense rank-1 path; fwd + dW of both layers + dX of the second")


def wl_hetsngp(ed, device, step_counter, rank, world):
    """SURVEY.md §8f-2 (next row): HeteroscedasticSNGPLayer head (hetsngp.py:25) on 512-dim features, batch 8192, 10
    classes, 10 factors, D = 1024 random features, 1000 Monte-Carlo samples per example (the reference's default)."""
    import torch

    from edward2_b200 import functional as F

    B, d, K, Fn, S, D = 8192, 512, 10, 10, 1000, 1024
    head = ed.layers.HeteroscedasticSNGPLayer(K, num_factors=Fn, train_mc_samples=S, test_mc_samples=S, num_inducing=D,
                                              gp_cov_momentum=0.999, gp_cov_ridge_penalty=1e-3, normalize_input=True,
                                              dtype="bfloat16", seed=4040)
    head.train()
    with torch.no_grad():
        head(torch.randn(256, d, device=device).to(torch.bfloat16), training=True)
    head.step_counter = step_counter
    gb = B * world

    def step(x, y):
        logits = head(x, training=True)
        loss = F.SoftmaxCrossEntropy.apply(logits, y, gb)
        F.backward(loss)
        return loss.detach()

    def cpu(steps):
        from oracle import torch_cpu

        rows = 1024
        sps, dt, th = torch_cpu.time_hetsngp(d, D, K, Fn, S, rows, steps, warmup=1)
        return sps, dt, th, f"{steps} train steps on {rows}-row batches (torch-CPU fp32 restatement, linear in rows)"

    # tensor work: RFF fwd + dX, GP output layer, loc / scale / diag Dense layers (fwd + dW + dX), PhiT Phi
    flops = 2.0 * B * d * D * 2 + 2.0 * B * D * K * 3 + 2.0 * B * d * (K * Fn + 2 * K) * 3 + 2.0 * B * D * D
    return dict(step=step, params=list(head.parameters()), batch=B, in_dim=d, x_dtype=torch.bfloat16, classes=K, flops=flops,
                cpu=cpu, dtype="bf16/f32", modules=head,
                workload="hetsngp (SURVEY §8f-2): HeteroscedasticSNGPLayer train step, 512-dim features, batch 8192, 10 "
                         "classes, 10 factors, D=1024, 1000 MC samples per example (393 M normals per direction)")


WORKLOADS = {
    "cfg4": wl_cfg4, "conv_rank1": wl_conv_rank1, "hetsngp_train": wl_hetsngp, "cfg2_reparam": wl_cfg2(False), "cfg2_flipout": wl_cfg2(True), "cfg1": wl_cfg1,
    "cfg3_train": wl_sngp("cfg3", True), "cfg3_eval": wl_sngp("cfg3", False),
    "cfg5_train": wl_sngp("cfg5", True), "cfg5_eval": wl_sngp("cfg5", False),
}


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """`--impl reference`: the reference's CPU path for the headline workload, restated (kind "port"), on this box's
    host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import torch_cpu

    B = 2048
    steps = max(1, min(args.steps, 30))
    sps, dt, threads = torch_cpu.time_rank1_mlp(CFG4_DIMS, B, steps, warmup=max(1, min(args.warmup, 3)), ensemble_size=4)
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg4: DenseRank1 MLP 2048-8192-8192-1000, ensemble_size 4, sampled r/s + KL, "
                               "2048 rows/GPU (16384 on 8 GPUs), bf16",
                   "note": "torch-CPU fp32 restatement of the reference's CPU path (TF/TFP/JAX not installable here); "
                           f"each step = one {B}-row batch ({steps} timed steps)"},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} fwd+bwd steps on {B}-row batches of the cfg4 rank-1 MLP"},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------------
def kernel_times(run_step, steps=3):
    """Per-kernel device time (us per step) of `steps` steps, from CUPTI through torch.profiler.  Diagnostics for the
    roofline object only — never a bench value."""
    import torch
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            run_step()
        torch.cuda.synchronize()
    out = {}
    for ev in prof.key_averages():
        t = getattr(ev, "device_time_total", None)
        if t is None:
            t = getattr(ev, "cuda_time_total", 0.0)
        if t:
            out[ev.key] = (out.get(ev.key, (0.0, 0))[0] + t / steps, out.get(ev.key, (0.0, 0))[1] + ev.count / steps)
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    import edward2_b200 as ed
    from edward2_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and args.gemm_ctas > 0:
        os.environ.setdefault("ED2_GEMM_MAX_CTAS", str(args.gemm_ctas))
    if world > 1:
        # full GEMM grids under data parallelism: the exchange's SM kernels then run at full width between two GEMMs
        # instead of starting on the 20 SMs a minimal-wave grid leaves free (libed2b200: ED2_GEMM_MINWAVE)
        os.environ.setdefault("ED2_GEMM_MINWAVE", "0")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    lib = _lib.load()
    _lib.check(lib.ed2_check_device(local), "ed2_check_device")
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    pk = peaks()

    def measure(name: str, steps: int, warmup: int, want_e2e: bool, want_kernels: bool):
        step_counter = torch.zeros((), dtype=torch.int64, device=device)
        wl = WORKLOADS[name](ed, device, step_counter, rank, world)
        B, params = wl["batch"], wl["params"]
        gen = torch.Generator().manual_seed(1234 + rank)
        x_host = torch.randn(B, wl["in_dim"], generator=gen).to(wl["x_dtype"]).pin_memory()
        y_host = torch.randint(0, wl["classes"], (B,), generator=gen).pin_memory()
        x_dev, y_dev = x_host.to(device), y_host.to(device)
        reducer = None
        if world > 1 and not args.no_allreduce and not wl.get("no_grad"):
            from edward2_b200.dist import GradientAllReduce

            reducer = GradientAllReduce(params, compress=None if args.fp32_allreduce else torch.bfloat16,
                                        exchange=args.grad_exchange, hook_exchange=args.hook_exchange)
        raw_step = wl["step"]

        def zero():
            for p in params:
                p.grad = None

        def eager_step(x, y):
            zero()
            step_counter.add_(1)
            loss = raw_step(x, y)
            if reducer is not None:
                reducer.wait()
            return loss

        use_graph = args.graph in ("on", "auto")
        graph, static_loss = None, None
        x_static, y_static = x_dev.clone(), y_dev.clone()
        # every pre-capture step runs on the capture-side stream so that autograd's AccumulateGrad nodes are bound to it
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
                    if reducer is not None:
                        reducer.wait()
            except Exception as e:  # noqa: BLE001  (e.g. a collective that cannot be captured) -> eager steps
                sys.stderr.write(f"[bench] {name}: CUDA-graph capture failed, running eagerly: {str(e)[:300]}\n")
                graph, static_loss = None, None
                if reducer is not None:
                    reducer.handles.clear()
                    for ex in reducer._exchanges.values():
                        ex._pending = False
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
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        res = dict(ms=ms, launches=launches_per_step * steps, graph=graph is not None, batch=B, flops=wl["flops"],
                   workload=wl["workload"], dtype=wl["dtype"], wl=wl, e2e_ms=None, kernels=None,
                   h2d=(x_host.numel() * x_host.element_size() + y_host.numel() * 8), exchange=None)

        if want_kernels:
            try:
                res["kernels"] = kernel_times(run_step, 3)
                if args.dump_kernels and rank == 0:
                    with open(args.dump_kernels, "a") as f:
                        f.write(json.dumps({"workload": name, "ms_per_step": ms, "kernels_us_count_per_step": {
                            k: [round(v[0], 2), round(v[1], 2)] for k, v in sorted(res["kernels"].items(), key=lambda kv: -kv[1][0])}}) + "\n")
            except Exception as e:  # noqa: BLE001
                sys.stderr.write(f"[bench] {name}: kernel timeline unavailable: {str(e)[:200]}\n")
            sync_all()

        if args.trace and name == args.workload:
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
            res["e2e_ms"] = float(t.item())
        res["loss"] = float(run_step().item())
        if reducer is not None:
            res["exchange"] = {"raw_dw": reducer.exchange_mode, "hook": reducer.hook_exchange}
            reducer.remove()
        return res

    def roofline_of(name, r):
        """Dominant kernel = the tcgen05 GEMM: algorithmic FLOPs of the step ÷ the GEMM launches' summed device time
        INSIDE the measured step (CUPTI), against the burst bf16 peak (conservative; the sustained figure is given too).
        cfg1 is launch-bound fp32 (FFMA SGEMM): reported against HBM on its minimum traffic."""
        wl = r["wl"]
        k = r.get("kernels") or {}
        gemm = [(n, v) for n, v in k.items() if "gemm_bf16_kernel" in n or "gemm_flipout_kernel" in n or "gemm_f32" in n or "sgemm" in n.lower()]
        gemm_us = sum(v[0] for _, v in gemm)
        n_launch = sum(v[1] for _, v in gemm)
        total_us = sum(v[0] for v in k.values())
        if wl.get("bound") == "launch latency":
            ach = wl["min_bytes"] / (r["ms"] * 1e-3) / 1e9
            return {"bound": "hbm", "kernel": "whole step (launch-bound, one CUDA graph)", "achieved": ach,
                    "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": None,
                    "note": "minimum traffic 9 MB/step (SURVEY.md §8d); the step is launch-latency bound",
                    "kernel_us_per_step": total_us or None, "peak_source": pk["src"]}
        if not gemm_us:
            ach = wl["flops"] / (r["ms"] * 1e-3) / 1e12
            return {"bound": "tensor", "kernel": "whole step (no kernel timeline)", "achieved": ach, "peak": pk["bf16"],
                    "unit": "TFLOP/s", "frac": ach / pk["bf16"], "traffic": None, "peak_source": pk["src"]}
        ach = wl["flops"] / (gemm_us * 1e-6) / 1e12
        return {"bound": "tensor", "kernel": f"ed2::tc::gemm_bf16_kernel / gemm_flipout_kernel ({n_launch:.0f} launches per step, timed in-step)",
                "achieved": ach, "peak": pk["bf16"], "unit": "TFLOP/s", "frac": ach / pk["bf16"],
                "frac_of_sustained": ach / pk["bf16_sustained"], "traffic": NCU_TRAFFIC.get(name),
                "traffic_unit": "bytes/launch (ncu --set full capture, profiles/)" if NCU_TRAFFIC.get(name) else None,
                "algorithmic_bytes_per_launch": NCU_ALGO_BYTES.get(name),
                "gemm_us_per_step": gemm_us, "all_kernels_us_per_step": total_us,
                "gemm_share_of_kernel_time": gemm_us / total_us if total_us else None,
                "algorithmic_flops_per_step": wl["flops"], "peak_source": pk["src"] + ", burst bf16"}

    def cpu_of(r, steps):
        try:
            sps, dt, th, sample = r["wl"]["cpu"](steps)
            return {"value": sps, "unit": "samples/s", "cores": th, "kind": "port", "sample": sample, "s_per_step": dt}
        except Exception as e:  # noqa: BLE001
            return {"error": str(e)[:200]}

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    head = measure(args.workload, args.steps, args.warmup, True, True)
    clocks = sampler.stop() if sampler else None
    head_roof = roofline_of(args.workload, head)
    head_cpu = cpu_of(head, 3) if (rank == 0 and world == 1 and not args.skip_cpu) else None
    head.pop("wl")
    torch.cuda.empty_cache()

    others = {}
    if world == 1 and not args.skip_variants:
        for name in WORKLOADS:
            if name == args.workload or (args.only and name not in args.only.split(",")):
                continue
            try:
                r = measure(name, max(5, args.steps // 2), 3, False, True)
                entry = {"workload": r["workload"], "ms_per_step": r["ms"], "value": r["batch"] / (r["ms"] * 1e-3),
                         "unit": "samples/s", "dtype": r["dtype"], "cuda_graph": r["graph"],
                         "gpu_launches_per_step": r["launches"] // max(5, args.steps // 2),
                         "step_tflops": r["flops"] / (r["ms"] * 1e-3) / 1e12,
                         "step_frac_of_burst_peak": r["flops"] / (r["ms"] * 1e-3) / 1e12 / pk["bf16"],
                         "roofline": roofline_of(name, r), "loss": r["loss"]}
                if not args.skip_cpu:
                    entry["cpu_baseline"] = cpu_of(r, 2)
                others[name] = entry
                del r
            except Exception as e:  # noqa: BLE001
                others[name] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    gb = head["batch"] * world
    line = {
        "metric": METRIC, "value": gb / (head["ms"] * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": head["ms"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": head["dtype"], "data": "synthetic",
        "config": {"workload": head["workload"], "global_batch": gb, "parallelism": f"dp{world}",
                   "l2": "inputs > L2: one cfg4 step touches ~1.1 GB (fp32 kernels + grads 737 MB, bf16 kernels 184 MB, "
                         "activations) against a 126 MB L2",
                   "cuda_graph": head["graph"], "kl_scale": KL_SCALE,
                   "grad_allreduce": None if world == 1 else head["exchange"],
                   "gemm_ctas": int(os.environ.get("ED2_GEMM_MAX_CTAS", 148))},
        "e2e": {"value": gb / (head["e2e_ms"] * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": head["h2d"] * world,
                "d2h_bytes_per_step": 8 * world, "ms_per_step": head["e2e_ms"]},
        "gpu_launches": head["launches"],
        "clocks": clocks,
        "roofline": head_roof,
        "cpu_baseline": head_cpu,
        "step_tflops": head["flops"] * world / (head["ms"] * 1e-3) / 1e12,
        "step_frac_of_burst_peak": head["flops"] / (head["ms"] * 1e-3) / 1e12 / pk["bf16"],
        "loss": head["loss"],
        "configs": others,
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
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS),
                    help="headline workload of the line (default cfg4); the others are reported under 'configs' at N=1")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"])
    ap.add_argument("--fp32-allreduce", action="store_true", help="all-reduce fp32 gradients instead of bf16 buckets")
    ap.add_argument("--grad-exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how raw bf16 weight gradients (Reparameterization kernels) are summed over ranks")
    ap.add_argument("--hook-exchange", default="dma", choices=["dma", "peer", "nccl"],
                    help="N > 1: how gradients reaching the per-parameter hook (rank-1 / BatchEnsemble kernels) are summed")
    ap.add_argument("--gemm-ctas", type=int, default=0, help="persistent GEMM CTAs when N > 1 (rest: NCCL)")
    ap.add_argument("--no-allreduce", action="store_true", help="diagnostic only: skip the gradient all-reduce")
    ap.add_argument("--trace", default="", help="write a chrome trace of 3 steps of the headline workload on rank 0 (diagnostics)")
    ap.add_argument("--dump-kernels", default="", help="append per-kernel in-step device times (CUPTI) to this file")
    ap.add_argument("--only", default="", help="comma-separated subset of the N=1 'configs' block (diagnostics)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-variants", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
