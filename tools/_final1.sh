export BENCH_WATCHDOG_S=400
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_tests.log 2>&1; tail -2 gpurun_out/r2_final_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/r2_final_bench1.json 2> gpurun_out/r2_final_bench1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['cg_solve']['ms_per_iteration'], d['cfg2_100cube']['ms_per_step'], d['gpu_launches'])
PY
