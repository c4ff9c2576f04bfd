cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29657 tools/nccl_micro.py 120 2>&1 | grep dist_cg_per | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    for k, v in d.items():
        if k.startswith('real'): print(k, v)"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu_200.json 2> gpurun_out/r2_bench_2gpu_200.err
grep bench gpurun_out/r2_bench_2gpu_200.err | tail -6
timeout 1500 python bench.py --steps 5 > gpurun_out/r2_bench_1gpu_200.json 2> gpurun_out/r2_bench_1gpu_200.err
grep bench gpurun_out/r2_bench_1gpu_200.err | tail -8
nvidia-smi --query-gpu=memory.used --format=csv | head -3
