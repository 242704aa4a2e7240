# r2p: swizzled merge tile, parallel search gathers, GI rows, alias scratch carve-out, SVGF saturate; second merge form (512 x 2) A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_alias_gpu.py tests/test_svgf_gpu.py tests/test_scene_gpu.py tests/test_rpt_gpu.py tests/test_zz_bench_resolution_gpu.py tests/test_renderer_gpu.py tests/test_rgi_gpu.py tests/test_sharded_1gpu.py -x -q -m gpu 2>&1 | tail -n 3
echo "--- merge form 2: parity"
ZETARAY_B200_MERGE=2 timeout 900 python -m pytest tests/test_rpt_gpu.py tests/test_zz_bench_resolution_gpu.py tests/test_zz_procedural_gpu.py tests/test_sharded_1gpu.py -x -q -m gpu 2>&1 | tail -n 8
for form in 1 2; do
for scene in cornell tunnel; do ZETARAY_B200_MERGE=$form timeout 300 python tools/bench_scenes.py $scene 8 2>&1 | tail -n 1 > gpurun_out/r2p_form${form}_$scene.json; python -c "import json;d=json.load(open('gpurun_out/r2p_form${form}_$scene.json'));print('form $form $scene',d['ms_per_frame'],' '.join('%s=%.3f'%(a,b) for a,b in d['kernels_ms_per_frame'].items()))"; done
done
prof() {   # name regex skip form
  ZETARAY_B200_MERGE=$4 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o /tmp/prof_$1 \
      python tools/bench_scenes.py cornell 2 > gpurun_out/r2p_ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/r2p_$1_details.txt 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/r2p_$1_raw.csv 2>&1
}
prof k_spatial_merge 'k_spatial_merge\(' 3 1
prof k_spatial_merge2 k_spatial_merge2 3 2
prof k_spatial_search k_spatial_search 3 1
ZR_DENOISE=2 timeout 200 python tools/bench_scenes.py cornell 8 2>&1 | tail -n 1 > gpurun_out/r2p_denoise2.json; python -c "import json;d=json.load(open('gpurun_out/r2p_denoise2.json'));print('svgf',' '.join('%s=%.3f'%(a,b) for a,b in d['kernels_ms_per_frame'].items() if 'svgf' in a))"
