mkdir -p gpurun_out
ZR_SVGF_DEBUG=1 timeout 600 python -m pytest tests/test_svgf_gpu.py -x -q -m gpu -k "frames and 1" 2>&1 | grep -E "Error|error|differs|passed|failed" | head -n 20 > gpurun_out/r2i_dbg.log; cat gpurun_out/r2i_dbg.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_svgf_gpu.py -x -q -m gpu -k "frames and 1" 2>&1 | grep -v "^$" | head -n 60 > gpurun_out/r2i_sanitizer.log; head -c 4000 gpurun_out/r2i_sanitizer.log
