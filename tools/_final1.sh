export BENCH_WATCHDOG_S=400
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_tests.log 2>&1; tail -3 gpurun_out/r2_final_tests.log
timeout 500 python bench.py > gpurun_out/r2_final_bench1.json 2> gpurun_out/r2_final_bench1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'], d['cg_solve']['ms_per_iteration'], d['cfg2_100cube']['ms_per_step'], d['clocks'])
PY
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --size 100 --cfg2-size 0 > /dev/null 2>&1; wc -l gpurun_out/r02_launches.csv
