timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
for fl in "--plain-order" ""; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $fl 2>&1 | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=1 $fl', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], ' '.join('%s=%.2f'%(k['kernel'],k['ms_per_frame']) for k in d['kernels'][:6]))"
done
bash tools/run_multi.sh 2 --plain-order
bash tools/run_multi.sh 2
