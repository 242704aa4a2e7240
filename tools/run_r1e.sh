# One GPU call for the r1e changes: new procedural parity tests + affected suites first, then scene / headline benches,
# one ncu capture of the tiled firefly stencil, then the rest of the GPU suite (bounded). Results under gpurun_out/.
mkdir -p gpurun_out
echo "== new + affected tests"; timeout 480 python -m pytest tests/test_zz_procedural_gpu.py tests/test_zz_pathtracer_gpu.py tests/test_post_gpu.py tests/test_scene_gpu.py tests/test_rgi_gpu.py -q -m gpu 2>&1 | tail -n 25 | tee gpurun_out/r1e_tests_new.log
echo "== scene benches"
timeout 200 python tools/bench_scenes.py atrium 10 2>&1 | tail -n 1 | tee gpurun_out/r1e_scene_atrium.json
timeout 200 python tools/bench_scenes.py tunnel 6 2>&1 | tail -n 1 | tee gpurun_out/r1e_scene_tunnel.json
echo "== headline bench"; timeout 300 python bench.py 2>&1 | tail -n 1 > gpurun_out/r1e_bench.json; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r1e_bench.json')); print('ours', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks']); print(' '.join('%s=%.3f'%(k['kernel'],k['ms_per_frame']) for k in d['kernels']))
except Exception as e: print('bench parse failed', e); print(open('gpurun_out/r1e_bench.json').read()[-2000:])
PY
echo "== ncu firefly"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_firefly -s 3 -c 1 -o /tmp/prof_ff python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1e_ncu_ff.log 2>&1
ncu -i /tmp/prof_ff.ncu-rep --page details > gpurun_out/r1e_k_firefly_tiled_details.txt 2>&1
echo "== rest of the GPU suite"; timeout 360 python -m pytest tests -q -m gpu --deselect tests/test_zz_procedural_gpu.py --deselect tests/test_zz_pathtracer_gpu.py --deselect tests/test_post_gpu.py --deselect tests/test_scene_gpu.py --deselect tests/test_rgi_gpu.py 2>&1 | tail -n 12 | tee gpurun_out/r1e_tests_rest.log
du -sh gpurun_out
