mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_svgf_gpu.py -x -q -m gpu 2>&1 | grep -E "Error|error|differs|passed|failed" | head -n 20 > gpurun_out/r2j_svgf.log; cat gpurun_out/r2j_svgf.log
for r in 1 2; do
  ZR_DENOISE=$r timeout 200 python tools/bench_scenes.py cornell 8 2>&1 | tail -n 1 > gpurun_out/r2j_denoise$r.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2j_denoise$r.json")); k=d["kernels_ms_per_frame"]; print("denoise radius $r", d["ms_per_frame"], " ".join("%s=%.3f"%(a,b) for a,b in k.items() if "svgf" in a))
except Exception as e: print("denoise $r FAILED", e, open("gpurun_out/r2j_denoise$r.json").read()[-800:])
PY
done
timeout 600 python -m pytest tests/test_rpt_gpu.py tests/test_sharded_1gpu.py -x -q -m gpu 2>&1 | tail -n 3
