cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2_multi_test.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tests/multi_gpu_worker.py 2>&1 | grep -E "MULTI_OK|Error|error|assert" | head -5 | tee gpurun_out/r2_multi_worker.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --size 120 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu_120.json 2> gpurun_out/r2_bench_2gpu_120.err
tail -5 gpurun_out/r2_bench_2gpu_120.err; cat gpurun_out/r2_bench_2gpu_120.json | head -c 3000
timeout 900 python bench.py --size 120 --cfg2-size 0 --steps 5 --no-cpu-baseline > gpurun_out/r2_bench_1gpu_120.json 2> gpurun_out/r2_bench_1gpu_120.err
tail -4 gpurun_out/r2_bench_1gpu_120.err
