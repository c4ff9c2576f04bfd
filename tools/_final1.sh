export BENCH_WATCHDOG_S=400
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_tests.log 2>&1; tail -4 gpurun_out/r2_final_tests.log
timeout 500 python bench.py > gpurun_out/r2_final_bench1.json 2> gpurun_out/r2_final_bench1.err; tail -c 1500 gpurun_out/r2_final_bench1.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
