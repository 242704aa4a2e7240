python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 900 python bench.py 2>&1 | tail -n 1 > gpurun_out/bench_full.json; python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print('ours', d['value'], d['ms_per_step'], 'e2e', d['e2e'], 'launches', d['gpu_launches'], 'clocks', d['clocks'], 'roofline', d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic'], 'cpu', d['cpu_baseline'])"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -n 1 | cut -c1-400
bash tools/run_multi.sh 2
