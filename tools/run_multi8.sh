# 8-GPU box: the headline bench at N = 8 and BASELINE config 5 as named (tunnel 4K, ReSTIR PT 5 bounces, 8 GPUs)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 30 --warmup 5 2>gpurun_out/r2u_n8.err | tail -n 1 > gpurun_out/r2u_bench_n8.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2u_bench_n8.json")); print("bench N=8", d["value"], "Mpaths/s", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], d["config"].get("strips"), d["config"].get("kernel_ms_per_frame_by_rank"))
except Exception as e: print("N=8 FAILED", e, open("gpurun_out/r2u_n8.err").read()[-2000:])
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29539 tools/bench_scenes.py tunnel 6 2>gpurun_out/r2u_tunnel_n8.err | tail -n 1 > gpurun_out/r2u_tunnel_n8.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2u_tunnel_n8.json")); print("tunnel N=8", d["ms_per_frame"], "ms", d["mpaths_per_s"], "Mpaths/s", d.get("strips"), d.get("kernel_ms_per_frame_by_rank"))
except Exception as e: print("tunnel N=8 FAILED", e, open("gpurun_out/r2u_tunnel_n8.err").read()[-2000:])
PY
