mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_svgf_gpu.py -x -q -m gpu -k "frames" 2>&1 | grep -E "Error|error|differs|passed|failed" | head -n 20 > gpurun_out/r2f_svgf.log; cat gpurun_out/r2f_svgf.log
timeout 900 python -m pytest tests/test_rpt_gpu.py tests/test_rdi_gpu.py tests/test_zz_procedural_gpu.py tests/test_sharded_1gpu.py -x -q -m gpu 2>&1 | tail -n 12 > gpurun_out/r2f_tests.log; cat gpurun_out/r2f_tests.log
for mode in queued; do
  for scene in cornell tunnel; do
    ZETARAY_B200_SPATIAL=$mode timeout 200 python tools/bench_scenes.py $scene 6 2>&1 | tail -n 1 > gpurun_out/r2f_${mode}_$scene.json
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2f_${mode}_$scene.json")); k=d["kernels_ms_per_frame"]; print("$mode $scene", d["ms_per_frame"], " ".join("%s=%.3f"%(a,b) for a,b in k.items()))
except Exception as e: print("$mode $scene FAILED", e, open("gpurun_out/r2f_${mode}_$scene.json").read()[-600:])
PY
  done
done
