cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export FEM_RING_IN_FLIGHT=200 FEM_RING_SLACK=512 FEM_RING_RING_BYTES=117440512 FEM_RING_MARGIN=2048
AB_STEPS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:staged_assembly -s 3 -c 1 -o gpurun_out/r2_ring_full2 python tools/ab_assembly.py 100 ring > gpurun_out/r2_ncu4.log 2>&1
tail -2 gpurun_out/r2_ncu4.log
