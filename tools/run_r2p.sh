# r2p: swizzled merge tile + speculative search gathers + GI rows -- parity, timing, ncu of merge / search
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rpt_gpu.py tests/test_zz_bench_resolution_gpu.py tests/test_renderer_gpu.py tests/test_rgi_gpu.py tests/test_sharded_1gpu.py -x -q -m gpu 2>&1 | tail -n 3
for scene in cornell tunnel; do timeout 300 python tools/bench_scenes.py $scene 6 2>&1 | tail -n 1 > gpurun_out/r2p_$scene.json; python -c "import json;d=json.load(open('gpurun_out/r2p_$scene.json'));print('$scene',d['ms_per_frame'],' '.join('%s=%.3f'%(a,b) for a,b in d['kernels_ms_per_frame'].items()))"; done
prof() {   # name regex skip
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o /tmp/prof_$1 \
      python tools/bench_scenes.py cornell 2 > gpurun_out/r2p_ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/r2p_$1_details.txt 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/r2p_$1_raw.csv 2>&1
}
prof k_spatial_merge k_spatial_merge 3
prof k_spatial_search k_spatial_search 3
