timeout 600 python -m pytest tests/test_gpu_rank1_fused.py -q 2>&1 | grep -E "^E  |passed|failed|Error" | head -30
rm -f gpurun_out/r02_kern_c.jsonl
python bench.py --steps 20 --warmup 5 --skip-cpu --dump-kernels gpurun_out/r02_kern_c.jsonl > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err
tail -3 gpurun_out/r02_bench_c.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_c.json"))
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["loss"])
for l in open("gpurun_out/r02_kern_c.jsonl"):
    d=json.loads(l); print(d["workload"], d["ms_per_step"])
    for k,v in list(d["kernels_us_count_per_step"].items())[:14]: print("   %9.1f us x%5.1f  %s" % (v[0], v[1], k[:120]))
PY
