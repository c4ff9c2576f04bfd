"""Transient heat conduction with backward Euler, surface convection and a moving heat source: the thermal half of the
reference's applications/thermal_mechanical/example.py (mass map rho Cp (T - T_old) / dt :40-43, convection on the top
surface :45-56) on the B200 path.

    python examples/transient_heat.py [--n 60 --steps 20 --out /tmp/heat]

Per step: one assembly (Poisson element kernel + fem_mass_term for the heat capacity + fem_face_residual / fem_face_tangent
for the convective surface) and one Jacobi-CG solve, all in libfem_b200; T_old lives at the quadrature points and is updated
in place between the steps (laws.LinearMass reads its fields at every assembly)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jax_fem_b200 as jf                      # noqa: E402
from jax_fem_b200 import laws                  # noqa: E402

RHO, CP, K, H_CONV, T0 = 8440., 588., 15., 100., 300.        # the reference's material constants (SI units)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=60, help="cells along x and y (n // 4 along z)")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--dt", type=float, default=2e-2)
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    Lx, Ly, Lz = 2e-2, 2e-2, 5e-3
    nz = max(args.n // 4, 2)
    m = jf.box_mesh(args.n, args.n, nz, Lx, Ly, Lz)
    cells = m.cells_dict['hexahedron']
    top = lambda p: np.isclose(p[2], Lz, atol=1e-8)
    bottom = lambda p: np.isclose(p[2], 0., atol=1e-8)
    mass = laws.LinearMass(RHO * CP / args.dt, None)

    class Heat(jf.Problem):
        def get_tensor_map(self):
            return laws.Poisson(K)

        def get_mass_map(self):
            return mass

        def get_surface_maps(self):
            return [laws.RobinPower(H_CONV, power=1.0, u_ref=T0)]          # q_conv = h (T - T0) leaves through the top

    problem = Heat(jf.Mesh(m.points, cells), vec=1, dim=3, dirichlet_bc_info=[[bottom], [0], [lambda p: T0]],
                   location_fns=[top])
    fe = problem.fes[0]
    dev = problem.device
    N = torch.from_numpy(fe.shape_vals).to(dev)                            # (Q, nodes per cell)
    cells_d = torch.from_numpy(cells).to(dev)
    xq = torch.from_numpy(fe.get_physical_quad_points()).to(dev)           # (C, Q, 3)
    T = torch.full((fe.num_total_nodes, 1), T0, dtype=torch.float64, device=dev)
    coef = RHO * CP / args.dt
    print(f"{fe.num_total_nodes} nodes, {fe.num_cells} cells, dt = {args.dt}")
    for step in range(args.steps):
        t0 = time.perf_counter()
        T_q = torch.einsum('cn,qn->cq', T[cells_d, 0], N)                  # T_old at the quadrature points
        # volumetric heat source moving along x (a Gaussian spot under the surface), entered as part of the constant term
        cx = 0.25 * Lx + 0.5 * Lx * step / max(args.steps - 1, 1)
        src = 5e10 * torch.exp(-((xq[..., 0] - cx) ** 2 + (xq[..., 1] - 0.5 * Ly) ** 2 + (xq[..., 2] - Lz) ** 2) / (1e-3 ** 2))
        mass.const = -(coef * T_q + src)                                   # residual: rho Cp (T - T_old)/dt - s
        T = jf.solver(problem, {'jax_solver': {'method': 'cg'}})[0]
        torch.cuda.synchronize()
        print(f"step {step + 1:3d}  T max {float(T.max()):9.3f} K  T mean {float(T.mean()):8.3f} K  "
              f"{1e3 * (time.perf_counter() - t0):7.1f} ms", flush=True)
        if args.out:
            os.makedirs(args.out, exist_ok=True)
            jf.save_sol(fe, T.cpu().numpy(), os.path.join(args.out, f"T_{step + 1:04d}.vtu"))
    assert torch.isfinite(T).all() and float(T.max()) > T0
    return T


if __name__ == "__main__":
    main()
