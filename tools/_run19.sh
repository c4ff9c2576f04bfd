cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q -m gpu -k "tile_major or neohookean or hyper or golden or newton" 2>&1 | tail -2
for pth in blocks tiles; do FEM_ELEMENT_PATH=$pth timeout 300 python tools/ab_assembly.py 100 staged neohookean 2>&1 | tail -1 | sed "s/^/neohookean $pth /"; done
