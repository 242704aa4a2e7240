# r2n: wavefront path generation -- parity, then timing against the lock-step kernel (fused mode keeps k_pathtrace)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rpt_gpu.py -x -q -m gpu 2>&1 | tail -n 3
for scene in cornell tunnel; do
  timeout 300 python tools/bench_scenes.py $scene 6 2>&1 | tail -n 1 > gpurun_out/r2n_$scene.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2n_$scene.json")); k=d["kernels_ms_per_frame"]; print("$scene", d["ms_per_frame"], " ".join("%s=%.3f"%(a,b) for a,b in list(k.items())[:8]))
except Exception as e: print("$scene FAILED", e, open("gpurun_out/r2n_$scene.json").read()[-800:])
PY
done
