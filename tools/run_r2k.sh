# r2k (2 GPUs): sharded parity (python + native), N=2 bench; SVGF parity after the shared-memory repack (GPU 0)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_svgf_gpu.py -x -q -m gpu 2>&1 | grep -E "Error|error|differs|passed|failed" | head -n 8
timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu 2>&1 | tail -n 15 > gpurun_out/r2k_sharded.log; cat gpurun_out/r2k_sharded.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2k_n1.err | tail -n 1 > gpurun_out/r2k_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/r2k_n2.err | tail -n 1 > gpurun_out/r2k_n2.json
python - <<PY
import json
for n in (1,2):
    try:
        d=json.load(open("gpurun_out/r2k_n%d.json"%n)); print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("strips"), d["config"].get("kernel_ms_per_frame_by_rank"))
    except Exception as e: print(n, "FAILED", e, open("gpurun_out/r2k_n%d.err"%n).read()[-1500:])
PY
