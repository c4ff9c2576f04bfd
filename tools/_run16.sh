cd $GRAFT_REPO_ROOT
AB_STEPS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:staged_warp -s 3 -c 1 -o gpurun_out/r2_warp_full python tools/ab_assembly.py 100 warp > gpurun_out/r2_ncu16.log 2>&1
tail -2 gpurun_out/r2_ncu16.log
