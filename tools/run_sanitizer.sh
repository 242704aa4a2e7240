# compute-sanitizer evidence: memcheck of the smoke frames and of the SVGF / ReSTIR PT parity tests (TMA tiles, swizzled reads, queues), racecheck of the smoke frames
mkdir -p gpurun_out
timeout 60 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^$" | tail -n 6 > gpurun_out/r2x_memcheck_smoke.log; tail -n 3 gpurun_out/r2x_memcheck_smoke.log
timeout 110 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_svgf_gpu.py tests/test_rpt_gpu.py -x -q -m gpu 2>&1 | grep -v "^$" | tail -n 12 > gpurun_out/r2x_memcheck_tests.log; tail -n 5 gpurun_out/r2x_memcheck_tests.log
timeout 60 compute-sanitizer --tool racecheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^$" | tail -n 12 > gpurun_out/r2x_racecheck_smoke.log; tail -n 5 gpurun_out/r2x_racecheck_smoke.log
