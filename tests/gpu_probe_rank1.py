"""Profiling probe: a few eager steps of the cfg4 rank-1 MLP (2048 rows) — short enough to run under ncu.
usage: python tests/gpu_probe_rank1.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import edward2_b200 as ed
import bench


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    dev = torch.device("cuda", 0)
    sc = torch.zeros((), dtype=torch.int64, device=dev)
    wl = bench.wl_cfg4(ed, dev, sc, 0, 1)
    x = torch.randn(2048, 2048, device=dev).to(torch.bfloat16)
    y = torch.randint(0, 1000, (2048,), device=dev)
    for _ in range(steps):
        for p in wl["params"]:
            p.grad = None
        sc.add_(1)
        wl["step"](x, y)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
