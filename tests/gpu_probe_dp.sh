# 2-GPU validation: full GPU test suite, bench at N=1 and N=2 (JSON lines kept), reference arm
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --steps 30 --warmup 5 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_n1.json | cut -c1-200
timeout 280 $TR 29517 bench.py --gpus 2 --steps 30 --warmup 5 2>gpurun_out/e2.log | tee gpurun_out/bench_r1_n2.json | cut -c1-200
timeout 280 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tee gpurun_out/bench_r1_ref.json | cut -c1-200
