cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "native_plan or tile_major or ring_assembly or fused_assembly_matches or c_abi or renumbered" 2>&1 | tail -2
timeout 300 python tools/ab_assembly.py 100 staged 2>&1 | tail -1
AB_STEPS=2 timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__throughput.avg.pct_of_peak_sustained_active --clock-control none -k regex:gather_csr -s 3 -c 1 --csv python tools/ab_assembly.py 100 staged 2>&1 | grep gather_csr | awk -F'","' '{print $(NF-2), $NF}'
