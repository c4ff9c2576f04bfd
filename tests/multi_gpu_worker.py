"""One rank of the real-NCCL correctness run (launched by tests/test_gpu_multi.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_worker.py

Every rank owns a slab of the mesh on its own GPU; halo exchange, all-reduces and the whole Krylov loops run in the library
over its own NCCL communicator (csrc/dist.cu).  Compared against the single-domain result computed on the same GPU:
linear solve (distributed Jacobi-CG), Newton solve (Neo-Hookean, per-iteration re-assembly), SIMP forward + implicit adjoint
(distributed Jacobi-BiCGSTAB on A^T) -- all to 1e-8 relative, iteration counts within a few of the single-domain ones."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import jax_fem_b200 as jf
    import gpu_problems as gp
    from jax_fem_b200 import laws
    from jax_fem_b200.distributed import NcclComm, ShardedProblem
    from jax_fem_b200.solver import implicit_vjp

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = NcclComm()
    report = {"world": world}
    TOL = 1e-8

    def relmax(a, b):
        return float(np.abs(a - b).max() / np.abs(b).max())

    # ---- 1. linear elasticity, distributed Jacobi-CG ----------------------------------------------------------------------
    class Elasticity(jf.Problem):
        def get_tensor_map(self):
            return laws.LinearElasticity(70e3, 0.3)

        def get_surface_maps(self):
            return [lambda u, x: np.array([0., 0., 100.])]

    m = jf.box_mesh(6 * world + 3, 6, 5, 3.0, 1.0, 0.8)
    rng = np.random.default_rng(0)
    pts, cells = m.points, m.cells_dict['hexahedron']
    left = lambda p: np.isclose(p[0], 0., atol=1e-5)
    right = lambda p: np.isclose(p[0], 3., atol=1e-5)
    pts = pts + 0.02 * rng.uniform(-1, 1, pts.shape) * (~np.isclose(pts[:, :1], 0.) & ~np.isclose(pts[:, :1], 3.))
    kw = dict(dirichlet_bc_info=[[left] * 3, [0, 1, 2], [lambda p: 0.] * 3], location_fns=[right])
    single = Elasticity(jf.Mesh(pts, cells), vec=3, dim=3, **kw)
    ref = jf.solver(single, {'jax_solver': {'method': 'cg'}})[0].cpu().numpy()
    sp = ShardedProblem(Elasticity, pts, cells, comm, vec=3, dim=3, **kw)
    sol = sp.solve_linear().cpu().numpy()
    e1 = relmax(sol, ref[sp.part.l2g])
    assert e1 <= TOL, f"sharded CG solve differs: {e1}"
    report["cg"] = {"err": e1, "iterations": sp.last_info["iterations"], "true_residual": sp.last_info["err"]}
    sol_b = sp.solve_linear(method='bicgstab').cpu().numpy()
    e1b = relmax(sol_b, ref[sp.part.l2g])
    assert e1b <= TOL, f"sharded BiCGSTAB solve differs: {e1b}"
    report["bicgstab"] = {"err": e1b, "iterations": sp.last_info["iterations"]}

    # ---- 1b. the same solves with the peer-memory halo exchange (stores into the neighbours' mailboxes instead of ncclSend /
    #          ncclRecv): the arithmetic is untouched, so iterates and solutions must be bit-identical --------------------------
    os.environ["FEM_HALO_P2P"] = "1"
    sp2 = ShardedProblem(Elasticity, pts, cells, comm, vec=3, dim=3, **kw)
    os.environ.pop("FEM_HALO_P2P")
    assert sp2.halo.p2p, "peer-memory exchange did not come up on this node"
    sol2 = sp2.solve_linear().cpu().numpy()
    it2 = sp2.last_info["iterations"]
    assert np.array_equal(sol2, sol) and it2 == report["cg"]["iterations"], "peer-memory CG differs from the NCCL one"
    sol2b = sp2.solve_linear(method='bicgstab').cpu().numpy()
    assert np.array_equal(sol2b, sol_b), "peer-memory BiCGSTAB differs from the NCCL one"
    report["p2p"] = {"cg_iterations": it2, "bit_identical": True}
    del sp2

    # ---- 2. Neo-Hookean Newton with per-iteration re-assembly ---------------------------------------------------------------
    class Hyper(jf.Problem):
        def get_tensor_map(self):
            return laws.NeoHookean(10.0, 0.3)

        def get_surface_maps(self):
            return [lambda u, x: np.array([0., 1e-3, 0.])]

    m = jf.box_mesh(4 * world, 4, 4, 1., 1., 1.)
    pts, cells = m.points, m.cells_dict['hexahedron']
    l1 = lambda p: np.isclose(p[0], 0., atol=1e-5)
    r1 = lambda p: np.isclose(p[0], 1., atol=1e-5)
    ymax = lambda p: np.isclose(p[1], 1., atol=1e-5)
    kw = dict(dirichlet_bc_info=[[l1] * 3 + [r1] * 3, [0, 1, 2] * 2, [lambda p: 0.] * 3 + [lambda p: 0.02] + [lambda p: 0.] * 2],
              location_fns=[ymax])
    single = Hyper(jf.Mesh(pts, cells), vec=3, dim=3, **kw)
    ref = jf.solver(single, {'jax_solver': {'method': 'cg'}})[0].cpu().numpy()
    sp = ShardedProblem(Hyper, pts, cells, comm, vec=3, dim=3, **kw)
    sol = sp.solve().cpu().numpy()
    e2 = relmax(sol, ref[sp.part.l2g])
    assert e2 <= TOL, f"sharded Newton solve differs: {e2}"
    assert sp.last_info['newton_iterations'] == single.last_newton_info['iterations']
    report["newton"] = {"err": e2, "newton_iterations": sp.last_info['newton_iterations'], "cg_iterations": sp.last_info['cg_iterations']}

    # ---- 3. SIMP: forward + implicit adjoint (A^T by distributed BiCGSTAB) + per-cell gradient -----------------------------
    m = jf.box_mesh(8 * world, 2, 4, 2.0, 0.5, 1.0)
    pts, cells = m.points, m.cells_dict['hexahedron']
    load = lambda p: np.isclose(p[0], 2.0, atol=1e-5)
    kw = dict(dirichlet_bc_info=[[l1] * 3, [0, 1, 2], [lambda p: 0.] * 3], location_fns=[load])
    theta = np.repeat((0.5 + 0.1 * np.random.default_rng(0).uniform(-1, 1, len(cells)))[:, None], 8, axis=1)
    single = gp.SIMPElasticity(jf.Mesh(pts, cells), vec=3, dim=3, **kw)
    single.internal_vars = [torch.from_numpy(theta).cuda()]
    ref_sol = jf.solver(single, {'jax_solver': {}})[0]
    ref_grad = implicit_vjp(single, [ref_sol], None, [-single._f_ext], {}).cpu().numpy()
    sp = ShardedProblem(gp.SIMPElasticity, pts, cells, comm, vec=3, dim=3, **kw)
    sp.problem.internal_vars = [torch.from_numpy(theta[sp.part.local_cells]).cuda()]
    sol = sp.solve_linear(method='bicgstab')
    f_ext = sp.problem._f_ext
    grad = sp.adjoint_gradient(sol, torch.zeros_like(sol) if f_ext is None else -f_ext).cpu().numpy()
    e3s = relmax(sol.cpu().numpy(), ref_sol.cpu().numpy()[sp.part.l2g])
    e3g = float(np.abs(grad - ref_grad[sp.part.local_cells]).max() / np.abs(ref_grad).max())
    assert e3s <= TOL and e3g <= TOL, f"sharded SIMP adjoint differs: {e3s}, {e3g}"
    report["simp_adjoint"] = {"sol_err": e3s, "grad_err": e3g, "adjoint_iterations": sp.last_info["iterations"],
                              "adjoint_true_residual": sp.last_info["err"]}

    errs = torch.tensor([e1, e1b, e2, e3s, e3g], dtype=torch.float64, device="cuda")
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    if rank == 0:
        report["max_err_over_ranks"] = errs.tolist()
        print("MULTI_OK " + json.dumps(report), flush=True)
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
