# r2h: SVGF parity (maps in global memory), RPT parity, SVGF timing, ncu launch list + full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_svgf_gpu.py -x -q -m gpu 2>&1 | grep -E "Error|error|differs|passed|failed" | head -n 20 > gpurun_out/r2h_svgf.log; cat gpurun_out/r2h_svgf.log
timeout 900 python -m pytest tests/test_rpt_gpu.py tests/test_zz_bench_resolution_gpu.py -x -q -m gpu 2>&1 | tail -n 5 > gpurun_out/r2h_rpt.log; cat gpurun_out/r2h_rpt.log
for r in 1 2; do
  ZR_DENOISE=$r timeout 200 python tools/bench_scenes.py cornell 8 2>&1 | tail -n 1 > gpurun_out/r2h_denoise$r.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2h_denoise$r.json")); k=d["kernels_ms_per_frame"]; print("denoise radius $r", d["ms_per_frame"], " ".join("%s=%.3f"%(a,b) for a,b in k.items()))
except Exception as e: print("denoise $r FAILED", e, open("gpurun_out/r2h_denoise$r.json").read()[-800:])
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/r2h_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_ncu_bench.log 2>&1
tail -n 3 gpurun_out/r2h_launches.csv
prof() {   # name regex skip
  ZR_DENOISE=2 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o /tmp/prof_$1 \
      python tools/bench_scenes.py cornell 2 > gpurun_out/r2h_ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/r2h_$1_details.txt 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/r2h_$1_raw.csv 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv 2>&1 | gzip > gpurun_out/r2h_$1_source.csv.gz
}
prof k_spatial_merge k_spatial_merge 3
prof k_shift_spatial_case1 "k_shift.*Li1ELb0ELb0" 3
prof k_temporal_merge k_temporal_merge 3
prof k_svgf_atrous_step1 k_svgf_atrous 5
prof k_svgf_atrous_step4 k_svgf_atrous 7
prof k_pathtrace k_pathtrace 3
du -sh gpurun_out
