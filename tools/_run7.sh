cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29657 tools/nccl_micro.py 120 2>&1 | grep dist_cg_per | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    for k, v in d.items(): print(k, v)"
