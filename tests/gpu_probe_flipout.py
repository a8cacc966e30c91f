"""Timing probe of the dual-accumulator Flipout launches (CUDA events, one launch configuration at a time).
usage: python tests/gpu_probe_flipout.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from edward2_b200 import _lib


def noise(stream):
    return _lib.noise(None, 2026, stream, None, None)


def timed(fn, reps=20):
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


def main():
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    bf = torch.bfloat16
    for (B, I, O) in [(4096, 1024, 4096), (4096, 4096, 4096), (4096, 4096, 1000)]:
        x = torch.randn(B, I, device=dev).to(bf)
        xs = torch.randn(B, I, device=dev).to(bf)
        mu = (torch.randn(I, O, device=dev) * 0.02).to(bf)
        dl = (torch.randn(I, O, device=dev) * 0.02).to(bf)
        bias = torch.zeros(O, device=dev)
        y = torch.empty(B, O, device=dev, dtype=bf)
        nxs = torch.empty(B, O, device=dev, dtype=bf)
        for tag, use_bias, use_next, act in [("plain", 0, 0, 0), ("bias+relu", 1, 0, 1), ("bias+relu+out2", 1, 1, 1)]:
            a = _lib.FlipoutFusedFwd()
            a.batch, a.in_dim, a.out_dim, a.act = B, I, O, act
            a.x, a.x_ld, a.xs, a.xs_ld, a.xs_ready = x.data_ptr(), I, xs.data_ptr(), I, 1
            a.sign_in, a.sign_out = noise(2), noise(3)
            a.mu_lp, a.mu_ld, a.delta_lp, a.delta_ld = mu.data_ptr(), O, dl.data_ptr(), O
            a.bias = bias.data_ptr() if use_bias else None
            a.y, a.y_ld = y.data_ptr(), O
            if use_next:
                a.next_xs, a.next_xs_ld, a.next_sign_in = nxs.data_ptr(), O, noise(12)
            us = timed(lambda: _lib.check(lib.ed2_flipout_fused_fwd(C.byref(a), _lib.stream_ptr()), "fwd"))
            print(f"fwd  B={B} I={I} O={O} {tag:16s} {us:8.1f} us  {4.0 * B * I * O / us / 1e6:7.1f} TFLOP/s", flush=True)
        # backward: premade gp / gt, dW pair + dX with the producer-side variants
        gp = torch.randn(B, O, device=dev).to(bf)
        gt = torch.randn(B, O, device=dev).to(bf)
        muf = torch.randn(I, O, device=dev) * 0.02
        rho = torch.full((I, O), -3.0, device=dev)
        dmu, drho = torch.empty(I, O, device=dev), torch.empty(I, O, device=dev)
        dx = torch.empty(B, I, device=dev, dtype=bf)
        pgt = torch.empty(B, I, device=dev, dtype=bf)
        pdb = torch.empty(I, device=dev)
        for tag, want_dx, prev, relu, use_gt, use_db in [("dW only", 0, 0, 0, 0, 0), ("dW+dX plain", 1, 0, 0, 0, 0),
                                                          ("dW+dX prev", 1, 1, 0, 0, 0), ("dW+dX prev relu", 1, 1, 1, 0, 0),
                                                          ("dW+dX prev relu gt", 1, 1, 1, 1, 0),
                                                          ("dW+dX prev relu gt db", 1, 1, 1, 1, 1)]:
            a = _lib.FlipoutFusedBwd()
            a.batch, a.in_dim, a.out_dim, a.act, a.premade = B, I, O, 1, 1
            a.gp, a.gp_ld, a.gt, a.gt_ld = gp.data_ptr(), O, gt.data_ptr(), O
            a.x, a.x_ld, a.xs, a.xs_ld = x.data_ptr(), I, xs.data_ptr(), I
            a.mu_lp, a.mu_ld, a.delta_lp, a.delta_ld = mu.data_ptr(), O, dl.data_ptr(), O
            a.mu, a.rho, a.eps, a.sign_in, a.sign_out = muf.data_ptr(), rho.data_ptr(), noise(0), noise(2), noise(3)
            a.dmu, a.drho = dmu.data_ptr(), drho.data_ptr()
            a.kl_grad_scale, a.prior_mean, a.prior_std = 1e-5, 0.0, 1.0
            if want_dx and not prev:
                a.dx, a.dx_ld = dx.data_ptr(), I
            if prev:
                a.prev_relu = relu
                a.prev_gp, a.prev_gp_ld = dx.data_ptr(), I
                if use_gt:
                    a.prev_gt, a.prev_gt_ld, a.prev_sign_out = pgt.data_ptr(), I, noise(13)
                if use_db:
                    a.prev_dbias = pdb.data_ptr()
            us = timed(lambda: _lib.check(lib.ed2_flipout_fused_bwd(C.byref(a), _lib.stream_ptr()), "bwd"))
            print(f"bwd  B={B} I={I} O={O} {tag:22s} {us:8.1f} us", flush=True)


if __name__ == "__main__":
    main()
