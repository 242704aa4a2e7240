# 4-GPU box: shift class launches on three streams -- parity (1 GPU + sharded), A/B at N = 1 and N = 4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rpt_gpu.py tests/test_zz_bench_resolution_gpu.py tests/test_zz_procedural_gpu.py tests/test_sharded_1gpu.py tests/test_renderer_gpu.py -x -q -m gpu 2>&1 | tail -n 4
timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu 2>&1 | tail -n 4
for one in 0 1; do
  for scene in cornell tunnel; do
    if [ $one = 1 ]; then export ZETARAY_B200_SHIFT_ONE_STREAM=1; else unset ZETARAY_B200_SHIFT_ONE_STREAM; fi
    timeout 300 python tools/bench_scenes.py $scene 8 2>&1 | tail -n 1 > gpurun_out/r2t_one${one}_$scene.json
    python -c "import json;d=json.load(open('gpurun_out/r2t_one${one}_$scene.json'));print('one_stream=$one $scene',d['ms_per_frame'],' '.join('%s=%.3f'%(a,b) for a,b in list(d['kernels_ms_per_frame'].items())[:8]))"
  done
done
run() {  # tag n flags...
  tag=$1; n=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 30 --warmup 5 "$@" 2>gpurun_out/r2t_$tag.err | tail -n 1 > gpurun_out/r2t_$tag.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2t_$tag.json")); print("$tag", d["value"], "Mpaths/s", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], d["config"].get("strips"), d["config"].get("kernel_ms_per_frame_by_rank"))
    print("    ", " ".join("%s=%.3f"%(k["kernel"],k["ms_per_frame"]) for k in d["kernels"]))
except Exception as e: print("$tag FAILED", e, open("gpurun_out/r2t_$tag.err").read()[-1200:])
PY
}
unset ZETARAY_B200_SHIFT_ONE_STREAM
run n4_three 4
run n2_three 2
export ZETARAY_B200_SHIFT_ONE_STREAM=1
run n4_one 4
unset ZETARAY_B200_SHIFT_ONE_STREAM
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 tools/bench_scenes.py tunnel 6 2>gpurun_out/r2t_tunnel_n4.err | tail -n 1 > gpurun_out/r2t_tunnel_n4.json
python -c "import json;d=json.load(open('gpurun_out/r2t_tunnel_n4.json'));print('tunnel N=4',d['ms_per_frame'],d['mpaths_per_s'])"
