# A/B of the side-kernel grid cap (ED2_SIDE_BLOCKS_PER_SM) on one GPU; not a bench
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["loss"])'
for f in 8 5 4 3 2; do echo "--- ED2_SIDE_BLOCKS_PER_SM=$f"; ED2_SIDE_BLOCKS_PER_SM=$f timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-variants 2>gpurun_out/bench_err.log | python -c "$P"; done
