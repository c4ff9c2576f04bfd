cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ring_assembly or renumbered" 2>&1 | tail -2
for w in 16 19; do FEM_WARPS=$w timeout 300 python tools/ab_assembly.py 100 warp 2>&1 | tail -1 | sed "s/^/W=$w /"; done
