"""Stand-alone timing probe of the other BASELINE.json configurations (cfg1, cfg3, cfg4-per-GPU, cfg5-per-GPU).
These are parity-test cases, not bench lines (bench.py measures cfg2); the numbers go to DESIGN.md for context.
Run under gpurun: `python tests/gpu_probe_configs.py`."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import edward2_b200 as ed
from edward2_b200 import functional as F, _lib


def timeit(fn, warmup=5, steps=20):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def graphed(step):
    """Capture `step` (which must end with everything joined) in a CUDA graph; returns a replay callable."""
    s = torch.cuda.Stream(priority=-1)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        step()
    return g.replay


def cfg1():
    """BatchEnsemble MLP 784-512-512-10, E=4, B=128, fp32 (launch-bound)."""
    dev = "cuda"
    layers = [ed.layers.DenseBatchEnsemble(512, ensemble_size=4, activation="relu", alpha_initializer="random_sign",
                                           gamma_initializer="random_sign"),
              ed.layers.DenseBatchEnsemble(512, ensemble_size=4, activation="relu", alpha_initializer="random_sign",
                                           gamma_initializer="random_sign"),
              ed.layers.DenseBatchEnsemble(10, ensemble_size=4, alpha_initializer="random_sign", gamma_initializer="random_sign")]
    x = torch.randn(128, 784, device=dev)
    y = torch.randint(0, 10, (128,), device=dev)
    params = None

    def step():
        nonlocal params
        h = x
        for l in layers:
            h = l(h)
        loss = F.SoftmaxCrossEntropy.apply(h, y, 128)
        if params is None:
            params = [p for l in layers for p in l.parameters()]
        for p in params:
            p.grad = None
        F.backward(loss)

    step()
    eager = timeit(step)
    rep = graphed(step)
    gr = timeit(rep)
    print(f"cfg1 BE MLP fp32 B=128: eager {eager*1e3:.1f} us/step, CUDA graph {gr*1e3:.1f} us/step -> {128/gr*1e3:.0f} samples/s")


def cfg3():
    """SNGP head: 2 x SpectralNormalization(Dense 512, relu) + RFGP(D=1024, 10 classes), B=8192, bf16."""
    dev = "cuda"
    B = 8192
    d1 = ed.layers.SpectralNormalization(ed.layers.Dense(512, activation="relu", dtype="bfloat16"), norm_multiplier=0.95)
    d2 = ed.layers.SpectralNormalization(ed.layers.Dense(512, activation="relu", dtype="bfloat16"), norm_multiplier=0.95)
    gp = ed.layers.RandomFeatureGaussianProcess(10, num_inducing=1024, gp_cov_momentum=0.999, gp_cov_ridge_penalty=1e-3,
                                                dtype="bfloat16")
    x = torch.randn(B, 512, device=dev).to(torch.bfloat16)
    y = torch.randint(0, 10, (B,), device=dev)
    mods = [d1, d2, gp]
    gp._eye_off = True

    def train():
        h = d2(d1(x, training=True), training=True)
        logits, _ = gp(h, training=True)
        loss = F.SoftmaxCrossEntropy.apply(logits, y, B)
        for m in mods:
            for p in m.parameters():
                p.grad = None
        F.backward(loss)

    def evaluate():
        with torch.no_grad():
            h = d2(d1(x, training=False), training=False)
            gp_in = gp._input_norm_layer(h)
            phi = gp._random_feature(gp_in)
            logits = gp._gp_output_layer(phi).float() + gp._gp_output_bias
            var, adj = gp._gp_cov_layer.compute_predictive_variance(phi, logits, mean_field_factor=1.0)
        return adj

    train()
    t = timeit(train)
    evaluate()
    e = timeit(evaluate)
    print(f"cfg3 SNGP head B=8192 bf16: train {t:.3f} ms/step -> {B/t*1e3/1e6:.2f} M samples/s "
          f"(incl. {B}x{B} eye materialisation); eval {e:.3f} ms -> {B/e*1e3/1e6:.2f} M samples/s")


def cfg4():
    """Rank-1 MLP 2048-8192-8192-1000, E=4, 2048 rows per GPU, bf16, sampled r/s."""
    dev = "cuda"
    B = 2048
    mk = lambda u, act: ed.layers.DenseRank1(u, activation=act, ensemble_size=4, dtype="bfloat16",  # noqa: E731
                                             alpha_regularizer=ed.regularizers.NormalKLDivergence(mean=1.0, stddev=0.1, scale_factor=1 / 50000),
                                             gamma_regularizer=ed.regularizers.NormalKLDivergence(mean=1.0, stddev=0.1, scale_factor=1 / 50000))
    layers = [mk(8192, "relu"), mk(8192, "relu"), mk(1000, None)]
    x = torch.randn(B, 2048, device=dev).to(torch.bfloat16)
    y = torch.randint(0, 1000, (B,), device=dev)
    params = None

    def step():
        nonlocal params
        h = x
        for l in layers:
            h = l(h)
        loss = F.SoftmaxCrossEntropy.apply(h, y, B, *[k for l in layers for k in l.losses])
        if params is None:
            params = [p for l in layers for p in l.parameters()]
        for p in params:
            p.grad = None
        F.backward(loss)

    step()
    t = timeit(step, 3, 10)
    fl = 0
    dims = [2048, 8192, 8192, 1000]
    for li, (i, o) in enumerate(zip(dims[:-1], dims[1:])):
        fl += 2 * B * i * o * (3 if li > 0 else 2)
    print(f"cfg4 rank-1 MLP per-GPU B=2048 bf16: {t:.3f} ms/step -> {B/t*1e3/1e6:.2f} M samples/s/GPU, {fl/t/1e9:.0f} TFLOP/s")


def cfg5(phase_precision="extended"):
    """RFGP D=8192 on 4096-dim inputs, 8192 rows per GPU: RFF + precision update (train) and variance (eval)."""
    dev = "cuda"
    B, d, D = 8192, 4096, 8192
    gp = ed.layers.RandomFeatureGaussianProcess(10, num_inducing=D, gp_cov_momentum=-1, gp_cov_ridge_penalty=1e-3,
                                                dtype="bfloat16", phase_precision=phase_precision,
                                                custom_random_features_initializer=ed.initializers.RandomNormal(stddev=1.0, seed=0))
    x = torch.randn(B, d, device=dev).to(torch.bfloat16)
    with torch.no_grad():
        gp._device = x.device
        gp.build(x.shape)
        h = gp._input_norm_layer(x)
        phi = gp._random_feature(h)
        cov = gp._gp_cov_layer
        cov(phi, training=True)

    def train():
        with torch.no_grad():
            h = gp._input_norm_layer(x)
            phi = gp._random_feature(h)
            cov.update_feature_precision_matrix(phi)

    t = timeit(train, 3, 10)
    fl = 2 * B * d * D + 2 * B * D * D
    print(f"cfg5 RFGP per-GPU B=8192 D=8192 phase={phase_precision}: train {t:.3f} ms -> {fl/t/1e9:.0f} TFLOP/s (RFF + full phi^T phi)")
    with torch.no_grad():
        cov.covariance_matrix.copy_(torch.eye(D, device=dev) * 1e-3)
        cov._needs_inverse = False
        cov._cov_lp = None

    def evaluate():
        with torch.no_grad():
            h = gp._input_norm_layer(x)
            phi = gp._random_feature(h)
            cov.compute_predictive_variance(phi)

    e = timeit(evaluate, 3, 10)
    print(f"cfg5 eval: {e:.3f} ms -> {fl/e/1e9:.0f} TFLOP/s (RFF + phi Sigma row-sum)")


if __name__ == "__main__":
    _lib.load()
    which = sys.argv[1:] or ["cfg1", "cfg3", "cfg4", "cfg5"]
    for w in which:
        try:
            {"cfg1": cfg1, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5, "cfg5n": lambda: cfg5("native")}[w]()
        except Exception as ex:  # noqa: BLE001
            print(f"{w}: FAILED {type(ex).__name__}: {str(ex)[:300]}")
        torch.cuda.empty_cache()
