python -m pytest tests/test_gpu_dp.py -q 2>&1 | tail -6 > gpurun_out/r02_gputests_8gpu.log; cat gpurun_out/r02_gputests_8gpu.log
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench_n${N}.json")); print(d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"], d["config"]["grad_allreduce"], d["config"]["cuda_graph"], d["clocks"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r02_bench_n${N}.err").read()[-1500:])
PY
done
