cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export BENCH_WATCHDOG_S=500
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --size 100 --cfg2-size 0 --no-cpu-baseline > gpurun_out/r2_launch_bench.json 2> gpurun_out/r2_launch_bench.err
# full sets of the two assembly kernels and the SpMV at cfg 2
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"element_dmma_kernel|gather_csr_kernel|spmv_block_fused" -s 6 -c 3 -o gpurun_out/r02_assembly_full python bench.py --steps 1 --warmup 3 --size 100 --cfg2-size 0 --no-cpu-baseline --no-solve > /dev/null 2> gpurun_out/r2_full.err
ls -la gpurun_out/r02_assembly_full.ncu-rep
