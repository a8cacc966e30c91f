"""cfg5 (RFGP D=8192 on 4096-dim inputs, 8192 rows per GPU) training step under torchrun: per-step precision all-reduce
vs the deferred sum (LaplaceRandomFeatureCovariance(data_parallel="deferred")).  Diagnostic, not a bench:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/gpu_probe_cfg5_dp.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

import edward2_b200 as ed


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, d, D = 8192, 4096, 8192
    x = torch.randn(B, d, device=dev).to(torch.bfloat16)
    res = {}
    for mode in (False, True, "deferred"):
        gp = ed.layers.RandomFeatureGaussianProcess(10, num_inducing=D, gp_cov_momentum=-1, gp_cov_ridge_penalty=1e-3,
                                                    dtype="bfloat16", phase_precision="native", data_parallel=mode,
                                                    custom_random_features_initializer=ed.initializers.RandomNormal(stddev=1.0, seed=0))
        with torch.no_grad():
            gp._device = dev
            gp.build(x.shape)
            cov = gp._gp_cov_layer
            phi = gp._random_feature(gp._input_norm_layer(x))
            cov(phi, training=True)

            def train():
                cov.update_feature_precision_matrix(gp._random_feature(gp._input_norm_layer(x)))

            for _ in range(3):
                train()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                train()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e0.record()
            cov.synchronize_precision()
            e1.record()
            torch.cuda.synchronize()
            res[str(mode)] = (float(t.item()), e0.elapsed_time(e1))
        del gp, cov, phi
        torch.cuda.empty_cache()
    if rank == 0:
        for k, (ms, sync_ms) in res.items():
            print(f"cfg5 train step, {world} GPUs x {B} rows, data_parallel={k}: {ms:.3f} ms/step "
                  f"-> {world * B / ms * 1e3 / 1e6:.2f} M samples/s (final synchronize_precision: {sync_ms:.3f} ms)")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
