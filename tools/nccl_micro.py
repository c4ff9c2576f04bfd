"""Micro-timing of the library's NCCL pieces against torch.distributed on the same ranks (run under torchrun)."""
import os, sys, time, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_problem
from jax_fem_b200 import _lib
from jax_fem_b200.distributed import NcclComm, TorchDistComm, Halo

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
size = int(sys.argv[1]) if len(sys.argv) > 1 else 100
comm = NcclComm()
prob, sp = build_problem(size, world, comm)
tcomm = TorchDistComm()
thalo = Halo(sp.part, tcomm, 3, prob.device)
n_local = prob.num_total_dofs_all_vars
x = torch.randn(n_local, dtype=torch.float64, device='cuda')
s4 = torch.ones(4, dtype=torch.float64, device='cuda')

def timeit(fn, reps=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return {"gpu_us": 1e3 * e0.elapsed_time(e1) / reps, "wall_us": 1e6 * (time.perf_counter() - w0) / reps}

out = {"native_halo": timeit(lambda: sp.halo.update(x)), "torch_halo": timeit(lambda: thalo.update(x)),
       "native_allreduce4": timeit(lambda: comm.allreduce(s4)), "torch_allreduce4": timeit(lambda: tcomm.allreduce(s4))}
lib = _lib.load()
ws = torch.zeros(lib.fem_krylov_workspace(n_local), dtype=torch.float64, device='cuda')
y = torch.empty_like(x)
ip, ix, dat = prob.newton_update([torch.zeros(n_local // 3, 3, dtype=torch.float64, device='cuda')]) and None, None, None
import jax_fem_b200 as jf
A = jf.get_A(prob)
indptr, indices, data = A.getValuesCSR()
P = _lib.ptr
n = sp.n_owned
out["spmv_only"] = timeit(lambda: _lib.check(lib.fem_dcg_spmv_dot(n, n_local, P(indptr), P(indices), P(data), 3, P(prob.plan.brow_ptr), P(prob.plan.bcol), P(x), P(y), 1, P(ws), _lib.stream_ptr())))
diag = A.diagonal()
b = torch.randn(n_local, dtype=torch.float64, device='cuda')
info = (_lib.ctypes.c_double * 4)()
def cg200():
    xx = torch.zeros(n_local, dtype=torch.float64, device='cuda')
    _lib.check(lib.fem_dist_pcg(sp.halo.handle, n, n_local, P(indptr), P(indices), P(data), 3, P(prob.plan.brow_ptr), P(prob.plan.bcol),
                                P(diag), P(b), P(xx), 0.0, 0.0, 200, 200, P(ws), info, _lib.stream_ptr()))
r = timeit(cg200, reps=3)
out["dist_cg_per_iteration_us"] = {k: v / 200 for k, v in r.items()}
out["debug_mask"] = os.environ.get("FEM_DIST_DEBUG", "0")
from jax_fem_b200.distributed import distributed_cg
dofs = torch.zeros(n_local, dtype=torch.float64, device='cuda')
r0 = jf.apply_bc_vec(prob.newton_update([dofs.reshape(-1, 3)])[0].reshape(-1), dofs, prob)
A0 = jf.get_A(prob)
for ce in (25, 200):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    xs, inf = distributed_cg(A0, -r0, torch.zeros_like(dofs), sp.part, sp.halo, comm, 3, check_every=ce)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[f"real_cg_native_check{ce}"] = {"iterations": inf["iterations"], "us_per_iteration": 1e6 * dt / inf["iterations"], "err": inf["err"]}
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
xs, inf = distributed_cg(A0, -r0, torch.zeros_like(dofs), sp.part, thalo, tcomm, 3)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
from jax_fem_b200.distributed import distributed_bicgstab
for _ in range(2):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    xs, inf2 = distributed_bicgstab(A0, -r0, torch.zeros_like(dofs), sp.part, sp.halo, comm, 3)
    torch.cuda.synchronize()
    dt2 = time.perf_counter() - t0
out["real_bicgstab_native"] = {"iterations": inf2["iterations"], "us_per_iteration": 1e6 * dt2 / inf2["iterations"], "err": inf2["err"]}
out["real_cg_python_loop_torch_comm"] = {"iterations": inf["iterations"], "us_per_iteration": 1e6 * dt / inf["iterations"], "err": inf["err"]}
if rank == 0:
    print(json.dumps(out))
comm.close(); dist.destroy_process_group()
