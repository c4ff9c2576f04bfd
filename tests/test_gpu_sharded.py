"""Sharded (multi-rank) path on ONE GPU: ranks run as threads of this process (ThreadComm), each with its own local
Problem, halo plan and share of the distributed Jacobi-CG.  The result must equal the single-domain solve and the oracle."""
import threading

import numpy as np
import pytest
import torch

from oracle import fem, laws as olaws

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_linear_elasticity_matches_single_domain_and_oracle(world):
    import jax_fem_b200 as jf
    from jax_fem_b200 import laws
    from jax_fem_b200.distributed import ShardedProblem, ThreadComm

    class Elasticity(jf.Problem):
        def get_tensor_map(self):
            return laws.LinearElasticity(70e3, 0.3)

        def get_surface_maps(self):
            return [lambda u, x: np.array([0., 0., 100.])]

    m = jf.box_mesh(9, 4, 3, 3.0, 1.0, 0.8)
    pts, cells = m.points, m.cells_dict['hexahedron']
    left = lambda p: np.isclose(p[0], 0., atol=1e-5)
    right = lambda p: np.isclose(p[0], 3., atol=1e-5)
    kw = dict(dirichlet_bc_info=[[left] * 3, [0, 1, 2], [lambda p: 0.] * 3], location_fns=[right])

    single = Elasticity(jf.Mesh(pts, cells), vec=3, dim=3, **kw)
    ref = jf.solver(single, {'jax_solver': {'method': 'cg'}})[0].cpu().numpy()

    out, errs = {}, []

    def run(comm):
        try:
            torch.cuda.set_device(0)
            sp = ShardedProblem(Elasticity, pts, cells, comm, vec=3, dim=3, **kw)
            sol = sp.solve_linear()
            out[comm.rank] = (sp.part, sol.cpu().numpy(), sp.last_info)
        except Exception as e:                      # pragma: no cover
            errs.append(e)
            try:
                comm.sh.barrier.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=run, args=(c,)) for c in ThreadComm.group(world)]
    [t.start() for t in threads]
    [t.join(300) for t in threads]
    assert not errs, errs
    glob = np.full_like(ref, np.nan)
    for r in range(world):
        part, sol, info = out[r]
        glob[part.owned] = sol[:part.n_owned]
        assert np.abs(sol[part.n_owned:] - ref[part.ghosts]).max() <= 1e-8 * np.abs(ref).max()   # ghosts up to date
        assert 0 < info['iterations'] < 2000
    assert np.abs(glob - ref).max() <= 1e-8 * np.abs(ref).max()

    opb = fem.Problem(fem.Mesh(pts, cells), 3, 3, law=olaws.LinearElastic(70e3, 0.3), location_fns=[right],
                      surface_maps=[lambda u, x: np.array([0., 0., 100.]) + 0. * u], **{'dirichlet_bc_info': kw['dirichlet_bc_info']})
    osol = fem.solver(opb, method='cg')
    assert np.abs(glob - osol).max() <= 1e-8 * np.abs(osol).max()


def test_sharded_newton_neohookean_matches_single_domain():
    """BASELINE.json configs[2] in miniature: Neo-Hookean cube (hyperelastic3d_common.py:15-42, 83-103), Newton with
    per-iteration tangent re-assembly, cells sharded across 2 ranks."""
    import jax_fem_b200 as jf
    from jax_fem_b200 import laws
    from jax_fem_b200.distributed import ShardedProblem, ThreadComm

    class Hyper(jf.Problem):
        def get_tensor_map(self):
            return laws.NeoHookean(10.0, 0.3)

        def get_surface_maps(self):
            return [lambda u, x: np.array([0., 1e-3, 0.])]

    m = jf.box_mesh(6, 4, 4, 1., 1., 1.)
    pts, cells = m.points, m.cells_dict['hexahedron']
    left = lambda p: np.isclose(p[0], 0., atol=1e-5)
    right = lambda p: np.isclose(p[0], 1., atol=1e-5)
    ymax = lambda p: np.isclose(p[1], 1., atol=1e-5)
    kw = dict(dirichlet_bc_info=[[left] * 3 + [right] * 3, [0, 1, 2] * 2,
                                 [lambda p: 0.] * 3 + [lambda p: 0.02] + [lambda p: 0.] * 2], location_fns=[ymax])
    single = Hyper(jf.Mesh(pts, cells), vec=3, dim=3, **kw)
    ref = jf.solver(single, {'jax_solver': {'method': 'cg'}})[0].cpu().numpy()
    out, errs = {}, []

    def run(comm):
        try:
            torch.cuda.set_device(0)
            sp = ShardedProblem(Hyper, pts, cells, comm, vec=3, dim=3, **kw)
            out[comm.rank] = (sp.part, sp.solve().cpu().numpy(), sp.last_info)
        except Exception as e:                      # pragma: no cover
            errs.append(e)
            comm.sh.barrier.abort()

    threads = [threading.Thread(target=run, args=(c,)) for c in ThreadComm.group(2)]
    [t.start() for t in threads]
    [t.join(300) for t in threads]
    assert not errs, errs
    for r in range(2):
        part, sol, info = out[r]
        assert info['newton_iterations'] == single.last_newton_info['iterations']
        assert np.abs(sol - ref[part.l2g]).max() <= 1e-8 * np.abs(ref).max()


def test_sharded_simp_adjoint_matches_single_domain():
    """BASELINE.json configs[4] in miniature: SIMP cantilever, forward solve + implicit adjoint + per-element density
    gradient with the cells sharded across 2 ranks (A^T solved by the distributed Jacobi-BiCGSTAB)."""
    import jax_fem_b200 as jf
    import gpu_problems as gp
    from jax_fem_b200.distributed import ShardedProblem, ThreadComm
    from jax_fem_b200.solver import implicit_vjp

    m = jf.box_mesh(8, 2, 4, 2.0, 0.5, 1.0)
    pts, cells = m.points, m.cells_dict['hexahedron']
    left = lambda p: np.isclose(p[0], 0., atol=1e-5)
    load = lambda p: np.isclose(p[0], 2.0, atol=1e-5)
    kw = dict(dirichlet_bc_info=[[left] * 3, [0, 1, 2], [lambda p: 0.] * 3], location_fns=[load])
    theta = np.repeat((0.5 + 0.1 * np.random.default_rng(0).uniform(-1, 1, len(cells)))[:, None], 8, axis=1)

    single = gp.SIMPElasticity(jf.Mesh(pts, cells), vec=3, dim=3, **kw)
    single.internal_vars = [torch.from_numpy(theta).cuda()]
    ref_sol = jf.solver(single, {'jax_solver': {}})[0]
    ref_grad = implicit_vjp(single, [ref_sol], None, [-single._f_ext], {}).cpu().numpy()
    ref_sol = ref_sol.cpu().numpy()
    out, errs = {}, []

    def run(comm):
        try:
            torch.cuda.set_device(0)
            sp = ShardedProblem(gp.SIMPElasticity, pts, cells, comm, vec=3, dim=3, **kw)
            sp.problem.internal_vars = [torch.from_numpy(theta[sp.part.local_cells]).cuda()]
            sol = sp.solve_linear(method='bicgstab')        # the reference's default Krylov method, sharded
            f_ext = sp.problem._f_ext                       # None on a rank whose slab has no load face
            grad = sp.adjoint_gradient(sol, torch.zeros_like(sol) if f_ext is None else -f_ext)
            out[comm.rank] = (sp.part, sol.cpu().numpy(), grad.cpu().numpy(), sp.last_info)
        except Exception as e:                      # pragma: no cover
            errs.append(e)
            comm.sh.barrier.abort()

    threads = [threading.Thread(target=run, args=(c,)) for c in ThreadComm.group(2)]
    [t.start() for t in threads]
    [t.join(300) for t in threads]
    assert not errs, errs
    seen = np.zeros(len(cells), dtype=bool)
    for r in range(2):
        part, sol, grad, info = out[r]
        assert np.abs(sol - ref_sol[part.l2g]).max() <= 1e-8 * np.abs(ref_sol).max()
        assert 0 < info['iterations'] < 2000
        assert np.abs(grad - ref_grad[part.local_cells]).max() <= 1e-8 * np.abs(ref_grad).max()
        seen[part.local_cells] = True
    assert seen.all()


def test_sharded_solve_on_randomly_renumbered_mesh():
    """Ownership by node-id ranges on a mesh whose numbering has no spatial order: every rank's slab is scattered and
    nearly all of its neighbours are ghosts; the sharded solve must still equal the single-domain one."""
    import jax_fem_b200 as jf
    import gpu_problems as gp
    from jax_fem_b200.distributed import ShardedProblem, ThreadComm
    m = jf.box_mesh(7, 4, 3, 2.0, 1.0, 0.8)
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(m.points))
    pts = np.empty_like(m.points)
    pts[perm] = m.points
    cells = perm[m.cells_dict['hexahedron']][rng.permutation(len(m.cells_dict['hexahedron']))]
    left = lambda p: np.isclose(p[0], 0., atol=1e-5)
    kw = dict(dirichlet_bc_info=[[left] * 3, [0, 1, 2], [lambda p: 0., lambda p: 0.01, lambda p: 0.]])
    single = gp.PlainElasticity(jf.Mesh(pts, cells), vec=3, dim=3, **kw)
    ref = jf.solver(single, {'jax_solver': {'method': 'cg'}})[0].cpu().numpy()
    out, errs = {}, []

    def run(comm):
        try:
            torch.cuda.set_device(0)
            sp = ShardedProblem(gp.PlainElasticity, pts, cells, comm, vec=3, dim=3, **kw)
            out[comm.rank] = (sp.part, sp.solve_linear().cpu().numpy())
        except Exception as e:                      # pragma: no cover
            errs.append(e)
            comm.sh.barrier.abort()

    threads = [threading.Thread(target=run, args=(c,)) for c in ThreadComm.group(3)]
    [t.start() for t in threads]
    [t.join(300) for t in threads]
    assert not errs, errs
    for r in range(3):
        part, sol = out[r]
        assert np.abs(sol - ref[part.l2g]).max() <= 1e-8 * np.abs(ref).max()
