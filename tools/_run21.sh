cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hex27 or HEX27" 2>&1 | tail -2
timeout 300 python tools/ab_assembly.py 60 staged hex27 2>&1 | tail -1
