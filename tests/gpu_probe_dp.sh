# 2-GPU diagnostics of the data-parallel step (kernel timelines through torch.profiler); not a bench
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["cuda_graph"], d["loss"])'
timeout 300 python -m pytest tests/test_gpu_dp.py -m gpu -q -s 2>&1 | grep -E "worst|passed|failed|Error|error" | head -12
echo "--- graph peer"; timeout 120 $TR 29511 bench.py --gpus 2 --steps 20 --warmup 5 --skip-variants --trace gpurun_out/trace_dp2_peer.json 2>gpurun_out/e_a.log | python -c "$P"
echo "--- graph nccl"; timeout 120 $TR 29513 bench.py --gpus 2 --steps 20 --warmup 5 --skip-variants --grad-exchange nccl 2>gpurun_out/e_c.log | python -c "$P"
gzip -f gpurun_out/trace_dp2_peer.json; tail -5 gpurun_out/e_a.log | cut -c1-300
