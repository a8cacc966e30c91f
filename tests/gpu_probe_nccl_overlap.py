"""2-GPU probe: does an NCCL all-reduce overlap with the persistent tcgen05 GEMM?
torchrun --nproc-per-node 2 tests/gpu_probe_nccl_overlap.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from edward2_b200 import ops

rank = int(os.environ["RANK"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 4096
a = torch.randn(n, n, device="cuda").bfloat16()
b = torch.randn(n, n, device="cuda").bfloat16()
out = torch.empty(n, n, device="cuda", dtype=torch.bfloat16)
buf = torch.randn(33_554_432, device="cuda").bfloat16()  # 67 MB


def gemms(k=3):
    for _ in range(k):
        ops.gemm(a, b, n, n, n, out=out)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def ar():
    dist.all_reduce(buf)


def both():
    h = dist.all_reduce(buf, async_op=True)
    gemms()
    h.wait()


t_g, t_a, t_b = timeit(gemms), timeit(ar), timeit(both)
if rank == 0:
    print(f"gemm_ctas={os.environ.get('ED2_GEMM_MAX_CTAS', '148')}: 3 GEMMs {t_g:.0f} us | all-reduce 67 MB bf16 {t_a:.0f} us "
          f"({67.1 / t_a * 1e3:.0f} GB/s) | concurrent {t_b:.0f} us")
dist.destroy_process_group()
