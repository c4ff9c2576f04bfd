cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export BENCH_WATCHDOG_S=400
timeout 500 python bench.py --steps 5 > gpurun_out/r2_bench_1gpu_200.json 2> gpurun_out/r2_bench_1gpu_200.err
grep bench gpurun_out/r2_bench_1gpu_200.err | tail -8
