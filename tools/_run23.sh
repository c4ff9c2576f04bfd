cd $GRAFT_REPO_ROOT
AB_STEPS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:hex27_kernel -s 3 -c 1 -o gpurun_out/r2_hex27_full3 python tools/ab_assembly.py 60 staged hex27 > gpurun_out/r2_ncu23.log 2>&1
