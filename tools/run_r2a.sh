mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 5 > gpurun_out/r2a_tests.log; cat gpurun_out/r2a_tests.log
for so in zetaray_b200/libzetaray_b200.so zetaray_b200/libzetaray_b200_regorder.so zetaray_b200/libzetaray_b200_skipzero.so; do
  v=$(basename $so .so); v=${v#libzetaray_b200}; v=${v#_}; v=${v:-default}
  for scene in cornell atrium tunnel; do
    ZETARAY_B200_LIB=$PWD/$so timeout 200 python tools/bench_scenes.py $scene 6 2>&1 | tail -n 1 > gpurun_out/r2a_${v}_$scene.json
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2a_${v}_$scene.json")); k=d["kernels_ms_per_frame"]; print("$v $scene", d["ms_per_frame"], " ".join("%s=%.2f"%(a,b) for a,b in list(k.items())[:7]))
except Exception as e: print("$v $scene FAILED", e)
PY
  done
done
timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | tail -n 1 > gpurun_out/r2a_bench.json; head -c 600 gpurun_out/r2a_bench.json
