cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ring_assembly or renumbered or fused_assembly_matches" 2>&1 | tail -2
export FEM_RING_IN_FLIGHT=200 FEM_RING_SLACK=512 FEM_RING_RING_BYTES=117440512 FEM_RING_MARGIN=2048
timeout 300 python tools/ab_assembly.py 100 ring 2>&1 | tail -1
