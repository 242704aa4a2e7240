# r2s: whole GPU suite on the final kernels, bench line, ncu launch list + full captures of the judged kernels
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 6 > gpurun_out/r2s_tests.log; cat gpurun_out/r2s_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2s_bench.err | tail -n 1 > gpurun_out/r2s_bench.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2s_bench.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"]); print(d["roofline"]); print(d["c1_alias_table"]); print([(k["kernel"],k["ms_per_frame"],k["frac"]) for k in d["kernels"]])
except Exception as e: print("bench FAILED", e, open("gpurun_out/r2s_bench.err").read()[-1500:])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file gpurun_out/r2s_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_ncu_bench.log 2>&1
prof() {   # name regex skip
  ZR_DENOISE=2 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o /tmp/prof_$1 \
      python tools/bench_scenes.py cornell 2 > gpurun_out/r2s_ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/r2s_$1_details.txt 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/r2s_$1_raw.csv 2>&1
}
prof k_spatial_merge k_spatial_merge 3
prof k_shift_spatial_case1 k_shift 30
prof k_shift_temporal_case1 k_shift 24
prof k_temporal_merge k_temporal_merge 3
prof k_spatial_classify k_spatial_classify 3
prof k_svgf_atrous_step1 k_svgf_atrous 5
prof k_svgf_atrous_step4 k_svgf_atrous 7
prof k_svgf_temporal k_svgf_temporal 2
prof k_taa k_taa 3
prof k_spatial_search k_spatial_search 3
prof k_firefly k_firefly 3
prof k_pathtrace k_pathtrace 3
prof k_di_temporal k_di_temporal 3
du -sh gpurun_out
