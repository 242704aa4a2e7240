mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_svgf_gpu.py -x -q -m gpu 2>&1 | tail -n 12 > gpurun_out/r2e_svgf.log; cat gpurun_out/r2e_svgf.log
timeout 300 python bench.py --steps 10 --warmup 3 2>gpurun_out/r2e_bench.err | tail -n 1 > gpurun_out/r2e_bench.json; python - <<PY
import json
d=json.load(open("gpurun_out/r2e_bench.json")); print(d["value"], d["ms_per_step"], d["e2e"], d["roofline"]); print(d["c1_alias_table"]); print(d["cpu_baseline"]); print([(k["kernel"],k["ms_per_frame"],k["frac"]) for k in d["kernels"]])
PY
tail -5 gpurun_out/r2e_bench.err
