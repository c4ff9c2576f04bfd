set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ring or fused_assembly_matches or renumbered" 2>&1 | tail -15 > gpurun_out/r2_t1.log
cat gpurun_out/r2_t1.log
timeout 600 python tools/ab_assembly.py 100 staged,ring 2>&1 | tail -5 | tee gpurun_out/r2_ab1.log
for s in 0 64 256; do FEM_RING_SLACK=$s timeout 300 python tools/ab_assembly.py 100 ring 2>&1 | tail -1 | tee -a gpurun_out/r2_ab1.log; done
for rb in 33554432 100663296; do FEM_RING_RING_BYTES=$rb FEM_RING_IN_FLIGHT=200 timeout 300 python tools/ab_assembly.py 100 ring 2>&1 | tail -1 | tee -a gpurun_out/r2_ab1.log; done
AB_STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed_pipe_fp64.sum --clock-control none -k regex:staged_assembly -s 3 -c 2 --csv --log-file gpurun_out/r2_ncu_staged1.csv python tools/ab_assembly.py 100 ring > gpurun_out/r2_ncu1.log 2>&1
tail -3 gpurun_out/r2_ncu_staged1.csv
