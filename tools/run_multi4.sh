# 4-GPU box: sharded parity (native PT, native GI, python transport; 2 and 4 ranks), the headline bench at N = 1, 2, 4, and the C4 / C5 stand-ins on 4 GPUs
mkdir -p gpurun_out
nvidia-smi -L | head -n 8
timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu 2>&1 | tail -n 6 > gpurun_out/r2q_sharded_tests.log; cat gpurun_out/r2q_sharded_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2q_n1.err | tail -n 1 > gpurun_out/r2q_bench_n1.json
for n in 2 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 2>gpurun_out/r2q_n$n.err | tail -n 1 > gpurun_out/r2q_bench_n$n.json
done
python - <<PY
import json
for n in (1,2,4):
    try:
        d=json.load(open("gpurun_out/r2q_bench_n%d.json"%n)); print("bench N=%d"%n, d["value"], "Mpaths/s", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], d["config"].get("strips"), d["config"].get("kernel_ms_per_frame_by_rank"))
    except Exception as e: print(n, "FAILED", e, open("gpurun_out/r2q_n%d.err"%n).read()[-1500:])
PY
for scene in atrium tunnel; do
  timeout 400 python tools/bench_scenes.py $scene 6 2>gpurun_out/r2q_${scene}_n1.err | tail -n 1 > gpurun_out/r2q_${scene}_n1.json
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 tools/bench_scenes.py $scene 6 2>gpurun_out/r2q_${scene}_n4.err | tail -n 1 > gpurun_out/r2q_${scene}_n4.json
  python - <<PY
import json
for n in (1,4):
    try:
        d=json.load(open("gpurun_out/r2q_${scene}_n%d.json"%n)); print("$scene N=%d"%n, d["ms_per_frame"], "ms", d["mpaths_per_s"], "Mpaths/s", d.get("strips"), d.get("kernel_ms_per_frame_by_rank"), " ".join("%s=%.2f"%(a,b) for a,b in list(d["kernels_ms_per_frame"].items())[:6]))
    except Exception as e: print("$scene", n, "FAILED", e, open("gpurun_out/r2q_${scene}_n%d.err"%n).read()[-1500:])
PY
done
