mkdir -p gpurun_out
timeout 230 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^$" | tail -n 25 > gpurun_out/r2x_memcheck_smoke.log; tail -n 12 gpurun_out/r2x_memcheck_smoke.log
