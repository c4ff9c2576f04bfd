cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hex27 or HEX27" 2>&1 | tail -2
timeout 300 python tools/ab_assembly.py 60 staged hex27 2>&1 | tail -1
AB_STEPS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:hex27_kernel -s 3 -c 1 -o gpurun_out/r2_hex27_full2 python tools/ab_assembly.py 60 staged hex27 > gpurun_out/r2_ncu22.log 2>&1
