"""Profiling probe: eager steps of the SNGP workloads (cfg3 train / eval, cfg5 shard train / eval) — short enough for ncu.
usage: python tests/gpu_probe_sngp.py [cfg3_train cfg3_eval cfg5_train cfg5_eval] [--steps N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import edward2_b200 as ed
import bench


def main():
    names = [a for a in sys.argv[1:] if a in bench.WORKLOADS] or ["cfg3_train", "cfg3_eval", "cfg5_train", "cfg5_eval"]
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 2
    dev = torch.device("cuda", 0)
    for name in names:
        sc = torch.zeros((), dtype=torch.int64, device=dev)
        wl = bench.WORKLOADS[name](ed, dev, sc, 0, 1)
        x = torch.randn(wl["batch"], wl["in_dim"], device=dev).to(wl["x_dtype"])
        y = torch.randint(0, wl["classes"], (wl["batch"],), device=dev)
        for _ in range(steps):
            for p in wl["params"]:
                p.grad = None
            wl["step"](x, y)
        torch.cuda.synchronize()
        print(name, "done")
        del wl
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
