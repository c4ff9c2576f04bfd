cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ring or fused_assembly_matches or renumbered" 2>&1 | tail -3 | tee gpurun_out/r2_t3.log
export FEM_RING_IN_FLIGHT=200
for cfg in "128 83886080 2560" "256 100663296 2048" "512 117440512 2048" "768 150994944 2048" "1024 201326592 2048"; do
  set -- $cfg
  FEM_RING_SLACK=$1 FEM_RING_RING_BYTES=$2 FEM_RING_MARGIN=$3 timeout 300 python tools/ab_assembly.py 100 ring 2>&1 | tail -1 | sed "s/^/slack=$1 ring=$2 /" | tee -a gpurun_out/r2_ab3.log
done
FEM_RING_SLACK=512 FEM_RING_RING_BYTES=117440512 FEM_RING_MARGIN=2048 AB_STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:staged_assembly -s 3 -c 1 --csv --log-file gpurun_out/r2_ncu_staged3.csv python tools/ab_assembly.py 100 ring > gpurun_out/r2_ncu3.log 2>&1
tail -4 gpurun_out/r2_ncu_staged3.csv | cut -d, -f5,13-
