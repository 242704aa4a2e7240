# A/B on one box: the atrium frame (ReSTIR GI) with and without the degenerate-ray early-out, each twice (run-to-run spread)
mkdir -p gpurun_out
for rep in 1 2; do
for v in default nodegen; do
  so=zetaray_b200/libzetaray_b200.so; [ $v = nodegen ] && so=zetaray_b200/libzetaray_b200_nodegen.so
  ZETARAY_B200_LIB=$PWD/$so timeout 120 python tools/bench_scenes.py atrium 6 2>&1 | tail -n 1 > gpurun_out/r1g_atrium_${v}_$rep.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r1g_atrium_${v}_$rep.json")); print("$v rep $rep", d["ms_per_frame"], d["kernels_ms_per_frame"])
except Exception as e: print("$v rep $rep FAILED", e, open("gpurun_out/r1g_atrium_${v}_$rep.json").read()[-500:])
PY
done; done
