# usage: bash tools/run_multi.sh N [extra bench flags]
N=${1:-2}; shift
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=%d'%d['n_gpus'], d['value'], 'Mpaths/s', d['ms_per_step'], 'ms  e2e', d['e2e']['value'], d['config']['strips'], 'rank kernel ms', d['config']['kernel_ms_per_frame_by_rank'], 'streams', d['config']['streams'])
print('   ', ' '.join('%s=%.2f'%(k['kernel'],k['ms_per_frame']) for k in d['kernels']))
"
