mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_svgf_gpu.py -x -q -m gpu -k "temporal_stage" 2>&1 | grep -v "^$" | head -n 80 > gpurun_out/r2e_sanitizer.log; head -c 6000 gpurun_out/r2e_sanitizer.log
