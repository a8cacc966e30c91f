"""Probe: operand pitch alignment of the 1000-wide Flipout launches (fwd x·W with N = 1000)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from edward2_b200 import _lib
from tests.gpu_probe_flipout import noise, timed


def main():
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    bf = torch.bfloat16
    B, I = 4096, 4096
    x = torch.randn(B, I, device=dev).to(bf)
    xs = torch.randn(B, I, device=dev).to(bf)
    for O, ld in [(1000, 1000), (1000, 1024), (1024, 1024), (1000, 1000), (1000, 1024)]:
        mu = (torch.randn(I, ld, device=dev) * 0.02).to(bf)
        dl = (torch.randn(I, ld, device=dev) * 0.02).to(bf)
        y = torch.empty(B, ld, device=dev, dtype=bf)
        a = _lib.FlipoutFusedFwd()
        a.batch, a.in_dim, a.out_dim, a.act = B, I, O, 0
        a.x, a.x_ld, a.xs, a.xs_ld, a.xs_ready = x.data_ptr(), I, xs.data_ptr(), I, 1
        a.sign_in, a.sign_out = noise(2), noise(3)
        a.mu_lp, a.mu_ld, a.delta_lp, a.delta_ld = mu.data_ptr(), ld, dl.data_ptr(), ld
        a.y, a.y_ld = y.data_ptr(), ld
        time.sleep(0.5)
        us = timed(lambda: _lib.check(lib.ed2_flipout_fused_fwd(C.byref(a), _lib.stream_ptr()), "fwd"), reps=10)
        print(f"fwd B={B} I={I} O={O} ld={ld}: {us:7.1f} us  {4.0 * B * I * O / us / 1e6:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
