# A/B harness: GPU parity tests on the default library, then bench the default and every libzetaray_b200_<variant>.so.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5
for so in zetaray_b200/libzetaray_b200.so zetaray_b200/libzetaray_b200_*.so; do
  [ -f $so ] || continue
  v=$(basename $so .so); v=${v#libzetaray_b200}; v=${v#_}; v=${v:-default}
  for mode in "" "--single-stream"; do
  ZETARAY_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline $mode 2>/dev/null | tail -n 1 > gpurun_out/ab_$v.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$v.json"))
    print("$v", "$mode", d["value"], d["ms_per_step"], " ".join("%s=%.2f"%(k["kernel"],k["ms_per_frame"]) for k in d["kernels"][:6]))
except Exception as e:
    print("$v", "FAILED", e)
PY
  done
done
for t in ${ZR_AB_TEST:-}; do
  ZETARAY_B200_LIB=$PWD/zetaray_b200/libzetaray_b200_$t.so timeout 600 python -m pytest tests/test_rpt_gpu.py -m gpu -x -q 2>&1 | tail -n 3
done
true
