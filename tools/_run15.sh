cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ring or fused_assembly_matches or renumbered" 2>&1 | tail -6 | tee gpurun_out/r2_t15.log
for w in 16 19 12; do FEM_WARPS=$w timeout 300 python tools/ab_assembly.py 100 warp 2>&1 | tail -1 | sed "s/^/W=$w /" | tee -a gpurun_out/r2_ab_warp.log; done
timeout 300 python tools/ab_assembly.py 100 staged,warp 2>&1 | tail -2 | tee -a gpurun_out/r2_ab_warp.log
