cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_gputests_b.log
for w in elasticity neohookean; do
 for pth in blocks tiles; do
  FEM_ELEMENT_PATH=$pth timeout 300 python tools/ab_assembly.py 100 staged $w 2>&1 | tail -1 | sed "s/^/$w $pth /" | tee -a gpurun_out/r2_ab_tiles.log
 done
done
