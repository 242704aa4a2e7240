# r2w: final check of the shipped library -- smoke, the whole GPU suite, the bench line, the launch list and ncu captures of the three judged streaming kernels
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 2 > gpurun_out/r2w_tests.log; cat gpurun_out/r2w_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2w_bench.err | tail -n 1 > gpurun_out/r2w_bench.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2w_bench.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["with_svgf"]["value"]); print(d["roofline"]["frac"], d["roofline"]["traffic"]); print([(k["kernel"],k["ms_per_frame"],k["frac"]) for k in d["kernels"]])
except Exception as e: print("bench FAILED", e, open("gpurun_out/r2w_bench.err").read()[-1500:])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file gpurun_out/r2w_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_ncu_bench.log 2>&1
prof() {   # name regex skip
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o /tmp/prof_$1 \
      python tools/bench_scenes.py cornell 2 > gpurun_out/r2w_ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/r2w_$1_details.txt 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/r2w_$1_raw.csv 2>&1
}
prof k_spatial_merge k_spatial_merge 3
prof k_temporal_merge k_temporal_merge 3
prof k_taa k_taa 3
