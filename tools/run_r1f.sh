# GPU call after the degenerate-ray early-out: whole GPU suite, scene benches again, headline bench, ncu launch list of the
# bench, one ncu --set full capture of k_pathtrace on the 10^6-triangle tunnel (what bounds traversal on a real-size scene).
mkdir -p gpurun_out
echo "== GPU suite"; timeout 420 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 8 | tee gpurun_out/r1f_tests.log
echo "== scene benches"
timeout 200 python tools/bench_scenes.py atrium 10 2>&1 | tail -n 1 | tee gpurun_out/r1f_scene_atrium.json
timeout 240 python tools/bench_scenes.py tunnel 6 2>&1 | tail -n 1 | tee gpurun_out/r1f_scene_tunnel.json
echo "== headline bench"; timeout 300 python bench.py 2>&1 | tail -n 1 > gpurun_out/r1f_bench.json; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r1f_bench.json')); print('ours', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'], 'cpu', d['cpu_baseline']); print(' '.join('%s=%.3f'%(k['kernel'],k['ms_per_frame']) for k in d['kernels']))
except Exception as e: print('bench parse failed', e); print(open('gpurun_out/r1f_bench.json').read()[-2000:])
PY
echo "== ncu launch list"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/r1f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_ncu_bench.log 2>&1
echo "== ncu k_pathtrace on the tunnel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pathtrace -s 3 -c 1 -o /tmp/prof_pt python tools/bench_scenes.py tunnel 1 > gpurun_out/r1f_ncu_pt.log 2>&1
ncu -i /tmp/prof_pt.ncu-rep --page details > gpurun_out/r1f_tunnel_k_pathtrace_details.txt 2>&1
ncu -i /tmp/prof_pt.ncu-rep --page raw --csv 2>/dev/null | python -c "
import sys,csv
rows=list(csv.reader(sys.stdin))
if len(rows)>2:
    h=rows[0]; v=rows[-1]
    keep=[(a,b) for a,b in zip(h,v) if any(k in a for k in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','smsp__inst_executed.sum','sm__warps_active.avg.pct','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__t_sector_hit_rate','lts__t_sector_hit_rate','local_load','local_store','lsu_mem_local'))]
    print('\n'.join('%s = %s'%kv for kv in keep))
" > gpurun_out/r1f_tunnel_k_pathtrace_raw_selected.txt 2>&1
du -sh gpurun_out
