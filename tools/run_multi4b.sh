# 4-GPU box: A/B of the block order (plain vs most-expensive-tile-first) and of one vs two streams on the strip-sharded Cornell frame
mkdir -p gpurun_out
run() {  # tag n flags...
  tag=$1; n=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 30 --warmup 5 "$@" 2>gpurun_out/r2r_$tag.err | tail -n 1 > gpurun_out/r2r_$tag.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2r_$tag.json")); print("$tag", d["value"], "Mpaths/s", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], d["config"].get("strips"), d["config"].get("kernel_ms_per_frame_by_rank"), d["config"].get("block_order"), "streams", d["config"].get("streams"))
    print("    ", " ".join("%s=%.3f"%(k["kernel"],k["ms_per_frame"]) for k in d["kernels"]))
except Exception as e: print("$tag FAILED", e, open("gpurun_out/r2r_$tag.err").read()[-1200:])
PY
}
run n4_plain 4
run n4_lpt 4 --schedule-by-cost
run n4_single 4 --single-stream
run n4_single_lpt 4 --single-stream --schedule-by-cost
run n2_lpt 2 --schedule-by-cost
run n4_plain_again 4
