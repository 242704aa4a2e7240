mkdir -p gpurun_out
ZR_SVGF_DEBUG=1 timeout 600 python -m pytest tests/test_svgf_gpu.py -x -q -m gpu -k "frames and 2" 2>&1 | grep -E "Error|error|differs|passed|failed" | head -n 20 > gpurun_out/r2g_dbg.log; cat gpurun_out/r2g_dbg.log
ZR_SVGF_DEBUG=1 timeout 600 python -m pytest tests/test_svgf_gpu.py -x -q -m gpu -k "temporal_stage" 2>&1 | grep -E "Error|error|differs|passed|failed" | head -n 20 > gpurun_out/r2g_dbg2.log; cat gpurun_out/r2g_dbg2.log
