# r2l: 512-thread x 2-blocks variant of the path-tracing / DI kernels vs the default 1024 x 1
mkdir -p gpurun_out
for so in zetaray_b200/libzetaray_b200.so zetaray_b200/libzetaray_b200_t512.so; do
  v=$(basename $so .so); v=${v#libzetaray_b200}; v=${v#_}; v=${v:-default}
  for scene in cornell tunnel atrium; do
    ZETARAY_B200_LIB=$PWD/$so timeout 200 python tools/bench_scenes.py $scene 6 2>&1 | tail -n 1 > gpurun_out/r2l_${v}_$scene.json
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2l_${v}_$scene.json")); k=d["kernels_ms_per_frame"]; print("$v $scene", d["ms_per_frame"], " ".join("%s=%.3f"%(a,b) for a,b in list(k.items())[:6]))
except Exception as e: print("$v $scene FAILED", e)
PY
  done
done
ZETARAY_B200_LIB=$PWD/zetaray_b200/libzetaray_b200_t512.so timeout 600 python -m pytest tests/test_rpt_gpu.py tests/test_rdi_gpu.py -x -q -m gpu 2>&1 | tail -n 2
