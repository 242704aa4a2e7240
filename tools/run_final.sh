mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -n 6
timeout 900 python bench.py 2>&1 | tail -n 1 > gpurun_out/bench_full.json; python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print('ours', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['clocks'], 'roofline', d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'], 'cpu', d['cpu_baseline']['value']); print(' '.join('%s=%.2f'%(k['kernel'],k['ms_per_frame']) for k in d['kernels']))"
timeout 300 python tools/time_gi.py 2>&1 | tail -n 1 | tee gpurun_out/gi_timing.json
