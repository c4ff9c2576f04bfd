cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export BENCH_WATCHDOG_S=400
N=$1
run() { timeout 450 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $N --steps 5 --warmup 3 "${@:4}" > gpurun_out/$3.json 2> gpurun_out/$3.err; grep -E "bench.*(assembly|cg:|newton|adjoint)" gpurun_out/$3.err | tail -4; }
run x 29661 r2_bench_${N}gpu_simp_cfg5 --workload simp --adjoint --no-solve
if [ "$N" = "8" ]; then run x 29662 r2_bench_${N}gpu_neohookean_cfg3 --workload neohookean --newton --no-solve; fi
