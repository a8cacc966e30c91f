"""Does an issue-bound elementwise kernel overlap with the tcgen05 GEMM when launched on a second stream?
Prints GEMM-alone, sampler-alone and concurrent times.  `python tests/gpu_probe_overlap.py`"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from edward2_b200 import _lib, ops

lib = _lib.load()
dev = "cuda"
n = 4096
a = torch.randn(n, n, device=dev).bfloat16()
b = torch.randn(n, n, device=dev).bfloat16()
out = torch.empty(n, n, device=dev, dtype=torch.bfloat16)
mu = torch.randn(n, n, device=dev) * 0.02
rho = torch.full((n, n), -3.0, device=dev)
w = torch.empty(n, n, device=dev, dtype=torch.bfloat16)
kl = torch.zeros((), dtype=torch.float64, device=dev)


def gemm():
    ops.gemm(a, b, n, n, n, out=out)


def sample(with_kl):
    _lib.check(lib.ed2_sample_normal_weights(mu.data_ptr(), rho.data_ptr(), _lib.noise(None, 1, 2), n, n, 0, w.data_ptr(),
                                             _lib.BF16, n, None, 0, kl.data_ptr() if with_kl else None, 1.0, 0.0, 1.0,
                                             _lib.stream_ptr()), "sample")


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


hi = torch.cuda.Stream(priority=-1)
lo = torch.cuda.Stream(priority=0)


def both(order, with_kl):
    main = torch.cuda.current_stream()
    hi.wait_stream(main)
    lo.wait_stream(main)
    if order == "gemm_first":
        with torch.cuda.stream(hi):
            gemm()
        with torch.cuda.stream(lo):
            sample(with_kl)
    else:
        with torch.cuda.stream(lo):
            sample(with_kl)
        with torch.cuda.stream(hi):
            gemm()
    main.wait_stream(hi)
    main.wait_stream(lo)


print(f"side blocks/SM={os.environ.get('ED2_SIDE_BLOCKS_PER_SM', '2')} carveout={os.environ.get('ED2_GEMM_CARVEOUT', '1')}")
print(f"gemm alone            {timeit(gemm):8.1f} us")
for with_kl in (False, True):
    print(f"sample alone kl={int(with_kl)}     {timeit(lambda: sample(with_kl)):8.1f} us")
    for order in ("gemm_first", "sample_first"):
        print(f"concurrent {order:12s} kl={int(with_kl)} {timeit(lambda: both(order, with_kl)):8.1f} us")
