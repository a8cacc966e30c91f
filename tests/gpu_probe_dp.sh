# 2-GPU validation / diagnostics of the data-parallel step (run under gpurun --gpus 2); not a bench:
#   data-parallel parity tests, smoke, bench at N=2 for both gradient-exchange variants, a kernel timeline of the step
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["cuda_graph"], d["config"]["grad_allreduce"][:40])'
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -s 2>&1 | grep -E "worst|passed|failed|Error" | head -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "--- peer"; timeout 200 $TR 29517 bench.py --gpus 2 --steps 30 --warmup 5 --skip-variants --trace gpurun_out/trace_dp2_peer.json 2>gpurun_out/e2.log | python -c "$P"
echo "--- nccl"; timeout 200 $TR 29518 bench.py --gpus 2 --steps 30 --warmup 5 --skip-variants --grad-exchange nccl 2>gpurun_out/e2n.log | python -c "$P"
gzip -f gpurun_out/trace_dp2_peer.json
