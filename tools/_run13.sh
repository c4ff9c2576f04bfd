cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export BENCH_WATCHDOG_S=500
for S in 100 200; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_traffic_$S.csv python bench.py --steps 2 --warmup 3 --no-solve --cfg2-size 0 --no-cpu-baseline --size $S > /dev/null 2> gpurun_out/r2_traffic_$S.err
done
FEM_ASSEMBLY=ring timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_traffic_100_ring.csv python bench.py --steps 2 --warmup 3 --no-solve --cfg2-size 0 --no-cpu-baseline --size 100 > /dev/null 2> gpurun_out/r2_traffic_100_ring.err
python tools/ncu_traffic.py gpurun_out/r2_traffic_100.csv elasticity 100 staged gpurun_out/r2_traffic_200.csv elasticity 200 staged gpurun_out/r2_traffic_100_ring.csv elasticity 100 ring
cp profiles/r02_traffic.json gpurun_out/
timeout 500 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; grep bench gpurun_out/r2_bench_default.err | tail -9
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 | head -c 600
