# r2c: k_shift block size x register budget
mkdir -p gpurun_out
for cfg in 0 1 2 3 4; do
  for scene in cornell tunnel; do
    ZETARAY_B200_SHIFT_CFG=$cfg timeout 200 python tools/bench_scenes.py $scene 6 2>&1 | tail -n 1 > gpurun_out/r2c_cfg${cfg}_$scene.json
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2c_cfg${cfg}_$scene.json")); k=d["kernels_ms_per_frame"]; print("cfg$cfg $scene", d["ms_per_frame"], "k_shift=%.3f merge=%.3f"%(k["k_shift"],k["k_spatial_merge"]))
except Exception as e: print("cfg$cfg $scene FAILED", e)
PY
  done
done
