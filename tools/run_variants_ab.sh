# On the GPU box: parity suite on every variant library, then the three workloads (Cornell / atrium / tunnel) on the default and on
# each variant. Results: gpurun_out/ab_<variant>_<scene>.json (tools/bench_scenes.py lines) and gpurun_out/ab_tests_<variant>.log
mkdir -p gpurun_out
for so in zetaray_b200/libzetaray_b200.so zetaray_b200/libzetaray_b200_*.so; do
  [ -f $so ] || continue
  v=$(basename $so .so); v=${v#libzetaray_b200}; v=${v#_}; v=${v:-default}
  ZETARAY_B200_LIB=$PWD/$so timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 3 > gpurun_out/ab_tests_$v.log; tail -n 1 gpurun_out/ab_tests_$v.log | sed "s/^/$v tests: /"
  for scene in cornell atrium tunnel; do
    ZETARAY_B200_LIB=$PWD/$so timeout 200 python tools/bench_scenes.py $scene 6 2>&1 | tail -n 1 > gpurun_out/ab_${v}_$scene.json
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_${v}_$scene.json")); k=d["kernels_ms_per_frame"]; print("$v $scene", d["ms_per_frame"], " ".join("%s=%.2f"%(a,b) for a,b in list(k.items())[:6]))
except Exception as e: print("$v $scene FAILED", e)
PY
  done
done
