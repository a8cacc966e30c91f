# A/B of the fused KL (ED2_KL_FUSED) on one GPU; not a bench
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["loss"], d["variants"]["flipout"]["ms_per_step"])'
for f in 1 0 1 0; do echo "--- ED2_KL_FUSED=$f"; ED2_KL_FUSED=$f timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu 2>gpurun_out/bench_err.log | python -c "$P"; done
