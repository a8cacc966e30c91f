# 2-GPU validation: data-parallel tests, smoke, bench at N=2 (diagnostic output only)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -s 2>&1 | grep -E "worst|passed|failed|Error" | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 280 $TR 29517 bench.py --gpus 2 --steps 30 --warmup 5 --skip-variants 2>gpurun_out/e2.log | cut -c1-200
