N=2
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/r02_bench_n${N}_$tag.json 2> gpurun_out/r02_bench_n${N}_$tag.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench_n${N}_$tag.json")); print("$tag", d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["grad_allreduce"], d["config"]["gemm_ctas"])
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/r02_bench_n${N}_$tag.err").read()[-800:])
PY
}
run base
run c128 --gemm-ctas 128
run c136 --gemm-ctas 136
run c120 --gemm-ctas 120
