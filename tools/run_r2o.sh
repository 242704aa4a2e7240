# r2o: source-level ncu pages of the bandwidth-class kernels (merge, TAA, search, sort, firefly, SVGF step 4)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_post_gpu.py tests/test_rpt_gpu.py tests/test_zz_bench_resolution_gpu.py tests/test_renderer_gpu.py -x -q -m gpu 2>&1 | tail -n 3
for scene in cornell tunnel; do timeout 300 python tools/bench_scenes.py $scene 6 2>&1 | tail -n 1 > gpurun_out/r2o_$scene.json; python -c "import json;d=json.load(open('gpurun_out/r2o_$scene.json'));print('$scene',d['ms_per_frame'],' '.join('%s=%.3f'%(a,b) for a,b in d['kernels_ms_per_frame'].items()))"; done
prof() {   # name regex skip
  ZR_DENOISE=2 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o /tmp/prof_$1 \
      python tools/bench_scenes.py cornell 2 > gpurun_out/r2o_ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/r2o_$1_details.txt 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv --print-source cuda,sass 2>&1 | gzip > gpurun_out/r2o_$1_source.csv.gz
}
prof k_spatial_merge k_spatial_merge 3
prof k_taa k_taa 3
prof k_spatial_search k_spatial_search 3
prof k_sort k_sort 3
prof k_firefly k_firefly 3
prof k_svgf_atrous_step4 k_svgf_atrous 7
prof k_temporal_merge k_temporal_merge 3
du -sh gpurun_out
