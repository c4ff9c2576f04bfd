cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export BENCH_WATCHDOG_S=300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_8gpu_200.json 2> gpurun_out/r2_bench_8gpu_200.err
grep bench gpurun_out/r2_bench_8gpu_200.err | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29658 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_4gpu_200.json 2> gpurun_out/r2_bench_4gpu_200.err
grep bench gpurun_out/r2_bench_4gpu_200.err | tail -4
