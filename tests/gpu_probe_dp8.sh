# 8-GPU diagnostics of the data-parallel step; not a bench
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["cuda_graph"])'
for n in 8 4; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port"
echo "--- peer n=$n"; timeout 150 $TR 2951$n bench.py --gpus $n --steps 30 --warmup 5 --skip-variants --trace gpurun_out/trace_dp${n}_peer.json 2>gpurun_out/e_p$n.log | tee gpurun_out/bench_peer_n$n.json | python -c "$P"
done
gzip -f gpurun_out/trace_dp8_peer.json gpurun_out/trace_dp4_peer.json
timeout 200 python -m pytest tests/test_gpu_dp.py -m gpu -q -s -k "peer-2-0" 2>&1 | grep -E "worst|passed|failed|Error|error" | head -12
