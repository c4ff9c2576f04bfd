set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/ab_assembly.py 100 ring 2>&1 | tail -1 | tee gpurun_out/r2_ab2.log
AB_STEPS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:staged_assembly -s 3 -c 1 -o gpurun_out/r2_ring_full python tools/ab_assembly.py 100 ring > gpurun_out/r2_ncu2.log 2>&1
tail -2 gpurun_out/r2_ncu2.log
