# r2d: SVGF parity, bench-resolution parity, whole GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_svgf_gpu.py -x -q -m gpu 2>&1 | tail -n 12 > gpurun_out/r2d_svgf.log; cat gpurun_out/r2d_svgf.log
timeout 900 python -m pytest tests/test_zz_bench_resolution_gpu.py -x -q -m gpu 2>&1 | tail -n 12 > gpurun_out/r2d_benchres.log; cat gpurun_out/r2d_benchres.log
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_svgf_gpu.py --deselect tests/test_zz_bench_resolution_gpu.py 2>&1 | tail -n 8 > gpurun_out/r2d_all.log; cat gpurun_out/r2d_all.log
