# A/B/C on one box: degenerate rays cut by an empty stack (default), by an early return (ret), not at all (nodegen)
mkdir -p gpurun_out
echo "== GPU suite on the default library"; timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 4 | tee gpurun_out/r1h_tests.log
for v in default ret nodegen; do
  so=zetaray_b200/libzetaray_b200.so; [ $v != default ] && so=zetaray_b200/libzetaray_b200_$v.so
  for scene in atrium tunnel; do
    [ $v = nodegen ] && [ $scene = tunnel ] && continue
    ZETARAY_B200_LIB=$PWD/$so timeout 150 python tools/bench_scenes.py $scene 4 2>&1 | tail -n 1 > gpurun_out/r1h_${scene}_$v.json
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r1h_${scene}_$v.json")); print("$v $scene", d["ms_per_frame"], d["kernels_ms_per_frame"])
except Exception as e: print("$v $scene FAILED", e, open("gpurun_out/r1h_${scene}_$v.json").read()[-500:])
PY
  done
done
echo "== headline bench (default)"; timeout 300 python bench.py 2>&1 | tail -n 1 > gpurun_out/r1h_bench.json; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r1h_bench.json')); print('ours', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks']); print(' '.join('%s=%.3f'%(k['kernel'],k['ms_per_frame']) for k in d['kernels']))
except Exception as e: print('bench parse failed', e)
PY
