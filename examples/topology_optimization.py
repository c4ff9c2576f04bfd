"""SIMP compliance minimisation of a 3-D cantilever: the reference's topology-optimisation workflow
(docs/source/learn/topology_optimization/example.ipynb; 3-D form of applications/outdated/top_opt) on the B200 path.

    python examples/topology_optimization.py [--nx 80 --ny 20 --nz 40 --iters 30 --out /tmp/top_opt]

Forward solve, implicit adjoint and per-element gradient run in libfem_b200 (ad_wrapper); the MMA update and the
sensitivity filter (jax_fem_b200.mma.optimize) work on CUDA tensors; the design is written as .vtu every 10 iterations."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jax_fem_b200 as jf                      # noqa: E402
from jax_fem_b200 import laws, mma             # noqa: E402


class Elasticity(jf.Problem):
    def get_tensor_map(self):
        return laws.SIMP(70e3, 70.0, 0.3, 3.0)             # Emax, Emin = 1e-3 Emax, nu, penal

    def get_surface_maps(self):
        return [lambda u, x: np.array([0., 0., 100.])]     # residual convention: minus the traction (0, 0, -100)

    def set_params(self, params):                          # (num_cells,) densities -> theta at every quadrature point
        self.internal_vars = [params[:, None].expand(-1, self.fes[0].num_quads)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=80)
    ap.add_argument("--ny", type=int, default=20)
    ap.add_argument("--nz", type=int, default=40)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--vf", type=float, default=0.4)
    ap.add_argument("--out", default="/tmp/top_opt")
    args = ap.parse_args()

    Lx, Ly, Lz = 2.0, 0.5, 1.0
    m = jf.box_mesh(args.nx, args.ny, args.nz, Lx, Ly, Lz)
    left = lambda p: np.isclose(p[0], 0., atol=1e-5)
    # load band: lower 10 % of the right face (every vertex of a face must satisfy the predicate: needs nz >= 10)
    load = lambda p: np.isclose(p[0], Lx, atol=1e-5) & (p[2] <= 0.1 * Lz + 1e-5)
    problem = Elasticity(jf.Mesh(m.points, m.cells_dict['hexahedron']), vec=3, dim=3,
                         dirichlet_bc_info=[[left] * 3, [0, 1, 2], [lambda p: 0.] * 3], location_fns=[load])
    fe = problem.fes[0]
    fwd_pred = jf.ad_wrapper(problem, {'jax_solver': {'method': 'cg'}}, {'jax_solver': {}})
    f_ext = problem._f_ext
    n = fe.num_cells
    state = {"iter": 0}

    def objective(rho):
        params = rho[:, 0].clone().requires_grad_(True)
        sol = fwd_pred(params)[0]
        J = -(f_ext * sol).sum()                           # compliance = int t.u ds
        J.backward()
        state["iter"] += 1
        print(f"iter {state['iter']:3d}  compliance {float(J.detach()):.6e}  volume {float(rho.mean()):.4f}", flush=True)
        if state["iter"] % 10 == 0 or state["iter"] == args.iters:
            jf.save_sol(fe, sol.detach(), os.path.join(args.out, f"sol_{state['iter']:03d}.vtu"),
                        cell_infos=[('theta', rho[:, 0])])
        return J.detach(), params.grad.reshape(-1, 1)

    def constraint(rho, it):
        vc = torch.stack([rho.mean() / args.vf - 1.0])
        return vc, torch.full((1, n, 1), 1.0 / (n * args.vf), dtype=torch.float64, device=rho.device)

    rho0 = torch.full((n, 1), args.vf, dtype=torch.float64, device=problem.device)
    rho = mma.optimize(fe, rho0, {'movelimit': 0.1, 'maxIters': args.iters}, objective, constraint, 1)
    print("final volume fraction", float(rho.mean()))


if __name__ == "__main__":
    main()
