"""GPU parity tests: the CUDA path, called through the package's public API (which goes through the
C ABI of include/fem_b200.h), against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star): CSR pattern / index arrays / BC masks bit-exact; residual and
Jacobian values <= 1e-12 relative max-norm; solutions and adjoint gradients <= 1e-8 relative.
"""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import fem, laws as olaws

pytestmark = pytest.mark.gpu

VAL_TOL = 1e-12
SOL_TOL = 1e-8


def relmax(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def perturbed_box(N, seed=0, amp=0.2):
    import jax_fem_b200 as jf
    m = jf.box_mesh(N, N + 1, N + 2, 1.0, 1.2, 0.9)
    pts = m.points.copy()
    rng = np.random.default_rng(seed)
    pts += amp / (N + 2) * rng.uniform(-0.5, 0.5, pts.shape)        # non-affine cells, still valid
    return pts, m.cells_dict['hexahedron']


def host(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("law_name", ["poisson", "elastic", "neohookean", "neohookean_rho", "simp"])
def test_element_residual_and_jacobian_match_oracle(law_name):
    import jax_fem_b200 as jf
    import gpu_problems as gp
    pts, cells = perturbed_box(5)
    rng = np.random.default_rng(1)
    vec = 1 if law_name == "poisson" else 3
    iv = None
    if law_name == "poisson":
        prob, olaw = gp.LinearPoisson(jf.Mesh(pts, cells), vec=1, dim=3), olaws.Poisson(1.0)
    elif law_name == "elastic":
        prob, olaw = gp.PlainElasticity(jf.Mesh(pts, cells), vec=3, dim=3), olaws.LinearElastic(70e3, 0.3)
    elif law_name == "neohookean":
        prob, olaw = gp.HyperElasticity(jf.Mesh(pts, cells), vec=3, dim=3), olaws.NeoHookean(1e3, 0.3)
    elif law_name == "neohookean_rho":
        prob = gp.NeoHookeanInverse(jf.Mesh(pts, cells), vec=3, dim=3, location_fns=[lambda p: np.isclose(p[1], 1.2, atol=1e-5)])
        olaw = olaws.NeoHookean(10.0, 0.3, clamp_J=True)
        iv = 0.5 + rng.uniform(0, 1, (len(cells), 8))
    else:
        prob = gp.SIMPElasticity(jf.Mesh(pts, cells), vec=3, dim=3, location_fns=[lambda p: np.isclose(p[0], 1.0, atol=1e-5)])
        olaw = olaws.SIMP(70e3, 70.0, 0.3, 3.0)
        iv = 0.2 + 0.7 * rng.uniform(0, 1, (len(cells), 8))
    sol = (0.004 if law_name.startswith('neohookean') else 0.05) * rng.standard_normal((len(pts), vec))
    if iv is not None:
        prob.internal_vars = [torch.from_numpy(iv).cuda()]
    opb = fem.Problem(fem.Mesh(pts, cells), vec, 3, law=olaw, internal_vars=() if iv is None else [iv])
    prob.newton_update([torch.from_numpy(sol).cuda()])
    Ke, Re = host(prob.element_tangents()), host(prob._Re)
    assert relmax(Ke, opb.cell_jacobians(sol)) <= VAL_TOL
    law_only = fem.Problem(fem.Mesh(pts, cells), vec, 3, law=olaw, internal_vars=() if iv is None else [iv])
    assert relmax(Re.reshape(len(cells), 8, vec), law_only.cell_residuals(sol)) <= VAL_TOL
    # reference attribute V (problem.py:453) = cell blocks then zero face blocks
    assert prob.V.numel() == Ke.size + sum(len(b) for b in prob.boundary_inds_list) * (8 * vec) ** 2


def _cube_problems():
    import jax_fem_b200 as jf
    import gpu_problems as gp
    g = cases.load_golden("linear_elasticity_cube")
    prob = gp.LinearElasticityCube(jf.Mesh(g["points"], g["cells"]), vec=3, dim=3,
                                   dirichlet_bc_info=cases.CUBE_BC, location_fns=[cases.right])
    opb = fem.Problem(fem.Mesh(g["points"], g["cells"]), 3, 3, dirichlet_bc_info=cases.CUBE_BC,
                      location_fns=[cases.right], law=olaws.LinearElastic(70e3, 0.3),
                      mass_map=cases.cube_mass, surface_maps=[cases.cube_traction])
    return g, prob, opb


def test_pattern_bc_masks_bit_exact_and_values():
    import jax_fem_b200 as jf
    g, prob, opb = _cube_problems()
    rng = np.random.default_rng(3)
    sol = 1e-3 * rng.standard_normal(g["sol"].shape)
    res = prob.newton_update([torch.from_numpy(sol).cuda()])[0]
    A = jf.get_A(prob)
    ores = opb.newton_update(sol)
    oA = fem.get_A(opb)
    indptr, indices, data = [host(t) for t in A.getValuesCSR()]
    assert indptr.dtype == np.int32 and indices.dtype == np.int32
    assert np.array_equal(indptr, oA.indptr) and np.array_equal(indices, oA.indices)       # bit-exact pattern
    rows, vals, flag = [host(t) for t in prob.bc_data()]
    orows = np.unique(np.concatenate(opb.bc_rows()))
    assert np.array_equal(rows, orows)                                                    # bit-exact BC mask
    assert np.array_equal(np.flatnonzero(flag), orows)
    assert relmax(data, oA.data) <= VAL_TOL
    assert relmax(host(res), ores) <= VAL_TOL
    # residual with Dirichlet rows (apply_bc_vec) and the BC-exact initial guess
    dofs = torch.from_numpy(sol.reshape(-1)).cuda()
    rb = jf.apply_bc_vec(res.reshape(-1), dofs, prob)
    assert relmax(host(rb), fem.apply_bc_vec(ores.reshape(-1), sol.reshape(-1), opb)) <= VAL_TOL
    # COO attributes of the reference API
    I, J = opb.coo_pattern()
    assert np.array_equal(prob.I, I) and np.array_equal(prob.J, J)
    # run-to-run determinism: no atomics anywhere on the path
    prob.newton_update([torch.from_numpy(sol).cuda()])
    assert torch.equal(jf.get_A(prob).data, A.data)


def test_spmv_diag_transpose_match_scipy():
    import jax_fem_b200 as jf
    g, prob, opb = _cube_problems()
    sol = 1e-3 * np.random.default_rng(4).standard_normal(g["sol"].shape)
    prob.newton_update([torch.from_numpy(sol).cuda()])
    A = jf.get_A(prob)
    S = A.to_scipy()
    x = np.random.default_rng(0).standard_normal(S.shape[0])
    y = host(A @ torch.from_numpy(x).cuda())
    assert relmax(y, S @ x) <= 1e-13
    assert np.array_equal(host(A.diagonal()), S.diagonal())
    AT = A.transpose()
    ST = S.T.tocsr()
    ST.sort_indices()
    assert np.array_equal(host(AT.getValuesCSR()[1]), ST.indices)
    assert np.array_equal(host(AT.data), ST.data)                 # pure permutation: bit-exact


@pytest.mark.parametrize("method", ["bicgstab", "cg"])
def test_krylov_matches_oracle_iterates(method):
    import jax_fem_b200 as jf
    from jax_fem_b200.solver import jax_solve, newton_step
    g, prob, opb = _cube_problems()
    sol = np.zeros(g["sol"].shape)
    dofs = torch.zeros(sol.size, dtype=torch.float64, device='cuda')
    res = jf.apply_bc_vec(prob.newton_update([dofs.reshape(-1, 3)])[0].reshape(-1), dofs, prob)
    A = jf.get_A(prob)
    ores = fem.apply_bc_vec(opb.newton_update(sol).reshape(-1), sol.reshape(-1), opb)
    oA = fem.get_A(opb)
    ox0 = fem.assign_bc(np.zeros(sol.size), opb) - fem.copy_bc(sol.reshape(-1), opb)
    ox, ok = fem.jax_solve(oA, -ores, ox0, True, method, return_iters=True)
    x0 = torch.from_numpy(ox0).cuda()
    x, info = jax_solve(A, -res, x0, True, method=method, return_info=True)
    assert abs(info['iterations'] - ok) <= max(3, ok // 20)       # same recurrences; rounding may shift the exit by a few
    assert relmax(host(x), ox) <= SOL_TOL
    assert info['err'] < 1e-6


@pytest.mark.parametrize("case", ["linear_poisson", "linear_elasticity_cube", "hyperelasticity", "linear_elasticity_cylinder"])
@pytest.mark.parametrize("method", ["bicgstab", "cg"])
def test_reference_goldens_through_solver(case, method):
    """The reference's own four benchmark tests, run through solver(problem) on the GPU."""
    import jax_fem_b200 as jf
    import gpu_problems as gp
    g = cases.load_golden(case)
    mesh = jf.Mesh(g["points"], g["cells"])
    omesh = fem.Mesh(g["points"], g["cells"])
    if case == "linear_poisson":
        prob = gp.LinearPoisson(mesh, vec=1, dim=3, dirichlet_bc_info=cases.POISSON_BC)
        opb = fem.Problem(omesh, 1, 3, dirichlet_bc_info=cases.POISSON_BC, law=olaws.Poisson())
        decimal = 5
    elif case == "linear_elasticity_cube":
        prob = gp.LinearElasticityCube(mesh, vec=3, dim=3, dirichlet_bc_info=cases.CUBE_BC, location_fns=[cases.right])
        opb = fem.Problem(omesh, 3, 3, dirichlet_bc_info=cases.CUBE_BC, location_fns=[cases.right],
                          law=olaws.LinearElastic(70e3, 0.3), mass_map=cases.cube_mass, surface_maps=[cases.cube_traction])
        decimal = 5
    elif case == "hyperelasticity":
        prob = gp.HyperElasticity(mesh, vec=3, dim=3, dirichlet_bc_info=cases.HYPER_BC)
        opb = fem.Problem(omesh, 3, 3, dirichlet_bc_info=cases.HYPER_BC, law=olaws.NeoHookean(1e3, 0.3))
        decimal = 5
    else:
        prob = gp.LinearElasticityCylinder(mesh, vec=3, dim=3, dirichlet_bc_info=cases.CYL_BC, location_fns=[cases.top])
        opb = fem.Problem(omesh, 3, 3, dirichlet_bc_info=cases.CYL_BC, location_fns=[cases.top],
                          law=olaws.LinearElastic(70e3, 0.3), mass_map=cases.cyl_mass, surface_maps=[cases.cyl_traction])
        decimal = 3
    sol = host(jf.solver(prob, {'jax_solver': {'method': method}})[0])
    gold = g["sol"].reshape(sol.shape)
    np.testing.assert_array_almost_equal(gold, sol, decimal=decimal)          # the reference's assertion
    osol = fem.solver(opb, method=method)
    assert relmax(sol, osol.reshape(sol.shape)) <= SOL_TOL                     # parity with the oracle
    if case == "hyperelasticity":
        assert prob.last_newton_info['iterations'] == 4


def test_simp_adjoint_gradient_matches_oracle_and_fd():
    import jax_fem_b200 as jf
    import gpu_problems as gp
    m = jf.box_mesh(8, 2, 4, 2.0, 0.5, 1.0)
    pts, cells = m.points, m.cells_dict['hexahedron']
    left = lambda p: np.isclose(p[0], 0., atol=1e-5)
    load = lambda p: np.isclose(p[0], 2.0, atol=1e-5)
    bc = [[left] * 3, [0, 1, 2], [lambda p: 0.] * 3]
    prob = gp.SIMPElasticity(jf.Mesh(pts, cells), vec=3, dim=3, dirichlet_bc_info=bc, location_fns=[load])
    rng = np.random.default_rng(0)
    rho = 0.5 + 0.1 * rng.uniform(-1, 1, len(cells))
    fwd = jf.ad_wrapper(prob)
    params = torch.from_numpy(rho).cuda().requires_grad_(True)
    sol = fwd(params)[0]
    f_ext = prob._f_ext                      # residual convention: res = internal + f_ext, f_ext = -int t.v
    J = -(f_ext * sol).sum()                 # compliance = int t.u ds
    J.backward()
    grad = host(params.grad)

    otr = lambda u, x: -np.array([0., 0., -100.]) + 0. * u
    opb = fem.Problem(fem.Mesh(pts, cells), 3, 3, dirichlet_bc_info=bc, location_fns=[load],
                      law=olaws.SIMP(70e3, 70.0, 0.3, 3.0), surface_maps=[otr],
                      internal_vars=[np.repeat(rho[:, None], 8, axis=1)])
    osol = fem.solver(opb)
    assert relmax(host(sol), osol) <= SOL_TOL
    of = np.zeros_like(osol)
    np.add.at(of, opb.cells[opb.boundary_inds_list[0][:, 0]].reshape(-1), opb.face_residuals(osol, 0).reshape(-1, 3))
    ograd = fem.implicit_vjp(opb, osol, -of).sum(axis=1)
    assert relmax(grad, ograd) <= SOL_TOL
    assert abs(float(J) - float(-(of * osol).sum())) <= SOL_TOL * abs(float(J))
    # finite-difference check a la docs/source/learn/compute_gradients/example.ipynb cells 24-29
    k = int(np.argmax(np.abs(grad)))
    h = 1e-4
    vals = []
    for s in (+1, -1):
        r2 = rho.copy()
        r2[k] += s * h
        s2 = fwd(torch.from_numpy(r2).cuda())[0]
        vals.append(float(-(f_ext * s2).sum()))
    fd = (vals[0] - vals[1]) / (2 * h)
    assert abs(fd - grad[k]) <= 1e-5 * abs(grad[k])


@pytest.mark.parametrize("mesh_kind", ["affine", "curved"])
def test_hex27_simp_adjoint_gradient_matches_oracle_and_fd(mesh_kind):
    """ad_wrapper on the registered HEX27 + SIMP combination (fem_hex27_adjoint_param_grad): compliance gradient with respect to
    per-cell densities against the oracle's implicit adjoint and a central finite difference; on a box (forward solve through
    the affine-cell pass) and on curved cells (general kernel).  27-point rule to keep the oracle quick."""
    import jax_fem_b200 as jf
    from jax_fem_b200 import laws
    m = jf.box_mesh_hex27(3, 1, 2, 1.5, 0.5, 1.0)
    pts, cells = m.points.copy(), m.cells_dict['hexahedron27']
    rng = np.random.default_rng(3)
    left = lambda p: p[0] < 1e-5
    load = lambda p: p[0] > 1.5 - 1e-5
    if mesh_kind == "curved":
        inner = (pts[:, 0] > 1e-5) & (pts[:, 0] < 1.5 - 1e-5)
        pts[inner] += 0.01 * rng.uniform(-1, 1, (int(inner.sum()), 3))
    bc = [[left] * 3, [0, 1, 2], [lambda p: 0.] * 3]
    nq = 27

    class Simp27(jf.Problem):
        def get_tensor_map(self):
            return laws.SIMP(70e3, 70.0, 0.3, 3.0)

        def get_surface_maps(self):
            return [lambda u, x: np.array([0., 0., 100.])]

        def set_params(self, params):
            self.internal_vars = [params[:, None].expand(-1, nq)]

    prob = Simp27(jf.Mesh(pts, cells), vec=3, dim=3, ele_type='HEX27', quadrature_order=4, dirichlet_bc_info=bc, location_fns=[load])
    assert prob.fes[0].num_quads == nq
    rho = 0.5 + 0.2 * rng.uniform(-1, 1, len(cells))
    fwd = jf.ad_wrapper(prob)
    params = torch.from_numpy(rho).cuda().requires_grad_(True)
    sol = fwd(params)[0]
    assert len(prob.hex27_general_cells()) == (0 if mesh_kind == "affine" else len(cells))
    f_ext = prob._f_ext
    J = -(f_ext * sol).sum()
    J.backward()
    grad = host(params.grad)
    otr = lambda u, x: np.array([0., 0., 100.]) + 0. * u
    opb = fem.Problem(fem.Mesh(pts, cells), 3, 3, ele_type='HEX27', quadrature_order=4, dirichlet_bc_info=bc, location_fns=[load],
                      law=olaws.SIMP(70e3, 70.0, 0.3, 3.0), surface_maps=[otr], internal_vars=[np.repeat(rho[:, None], nq, axis=1)])
    osol = fem.solver(opb)
    assert relmax(host(sol), osol) <= SOL_TOL
    of = np.zeros_like(osol)
    np.add.at(of, opb.cells[opb.boundary_inds_list[0][:, 0]].reshape(-1), opb.face_residuals(osol, 0).reshape(-1, 3))
    ograd = fem.implicit_vjp(opb, osol, -of).sum(axis=1)
    assert relmax(grad, ograd) <= SOL_TOL
    k = int(np.argmax(np.abs(grad)))
    h = 1e-4
    vals = []
    for sgn in (+1, -1):
        r2 = rho.copy()
        r2[k] += sgn * h
        vals.append(float(-(f_ext * fwd(torch.from_numpy(r2).cuda())[0]).sum()))
    assert abs((vals[0] - vals[1]) / (2 * h) - grad[k]) <= 1e-5 * abs(grad[k])


def test_quad4_poisson_config1_plumbing():
    """BASELINE.json configs[0]: Poisson on QUAD4 rectangle_mesh 32x32 (Quickstart.md:15-51), bicgstab."""
    import jax_fem_b200 as jf
    from jax_fem_b200 import laws

    class Poisson(jf.Problem):
        def get_tensor_map(self):
            return laws.Poisson(1.0)

        def get_mass_map(self):
            return lambda u, x: np.array([-10. * np.exp(-((x[0] - 0.5) ** 2 + (x[1] - 0.5) ** 2) / 0.02)])

    m = jf.rectangle_mesh(32, 32, 1., 1.)
    fns = [lambda p: np.isclose(p[0], 0., atol=1e-5), lambda p: np.isclose(p[0], 1., atol=1e-5),
           lambda p: np.isclose(p[1], 0., atol=1e-5), lambda p: np.isclose(p[1], 1., atol=1e-5)]
    bc = [fns, [0] * 4, [lambda p: 0.] * 4]
    prob = Poisson(jf.Mesh(m.points, m.cells_dict['quad']), vec=1, dim=2, ele_type='QUAD4', dirichlet_bc_info=bc)
    sol = host(jf.solver(prob)[0])
    om = fem.rectangle_mesh(32, 32, 1., 1.)
    src = lambda u, x: (-10. * np.exp(-((x[..., 0] - 0.5) ** 2 + (x[..., 1] - 0.5) ** 2) / 0.02))[..., None]
    opb = fem.Problem(om, 1, 2, ele_type='QUAD4', dirichlet_bc_info=bc, law=olaws.Poisson(), mass_map=src)
    osol = fem.solver(opb)
    assert prob.plan.nnz == 9409 and sol.shape == (1089, 1)
    assert relmax(sol, osol) <= SOL_TOL
    assert sol.max() > 0.01


def test_unregistered_paths_raise():
    import jax_fem_b200 as jf
    from jax_fem_b200 import laws

    class Custom(jf.Problem):
        def get_tensor_map(self):
            return lambda u_grad: u_grad * 2.0

    class Universal(jf.Problem):
        def get_tensor_map(self):
            return laws.Poisson()

        def get_universal_kernel(self):
            return None

    m = jf.box_mesh(2, 2, 2, 1, 1, 1)
    mesh = jf.Mesh(m.points, m.cells_dict['hexahedron'])
    with pytest.raises(laws.UnregisteredLawError):
        Custom(mesh, vec=1, dim=3)
    with pytest.raises(NotImplementedError):
        Universal(mesh, vec=1, dim=3)
    with pytest.raises(NotImplementedError):
        class P2(jf.Problem):
            def get_tensor_map(self):
                return laws.NeoHookean(1., 0.3)
        P2(mesh, vec=1, dim=3)
    ok = type("Ok", (jf.Problem,), {"get_tensor_map": lambda self: laws.Poisson()})(mesh, vec=1, dim=3)
    with pytest.raises(NotImplementedError):
        jf.solver(ok, {'petsc_solver': {}})
    with pytest.raises(NotImplementedError):
        jf.solver(ok, {'arc_length': {}})


# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("order", [None, 4])
def test_hex27_dmma_element_assembly_and_solve(order):
    """BASELINE.json configs[3]: second-order HEX27 linear elasticity (81x81 element tangents on FP64 DMMA tiles);
    default quadrature degree 10 (216 points, basis.py:64) and the 27-point variant."""
    import jax_fem_b200 as jf
    import gpu_problems as gp
    m = jf.box_mesh_hex27(3, 2, 2, 1.5, 1.0, 0.8)
    pts, cells = m.points.copy(), m.cells_dict['hexahedron27']
    rng = np.random.default_rng(5)
    pts += 0.02 * rng.uniform(-1, 1, pts.shape)                      # curved, non-affine cells
    left = lambda p: p[0] < 0.03
    right = lambda p: p[0] > 1.47
    bc = [[left] * 3, [0, 1, 2], [lambda p: 0., lambda p: 0.01, lambda p: 0.]]

    class Elast27(jf.Problem):
        def get_tensor_map(self):
            return jf.laws.LinearElasticity(70e3, 0.3)

        def get_surface_maps(self):
            return [lambda u, x: np.array([0., 0., -50.])]

    prob = Elast27(jf.Mesh(pts, cells), vec=3, dim=3, ele_type='HEX27', quadrature_order=order,
                   dirichlet_bc_info=bc, location_fns=[right])
    opb = fem.Problem(fem.Mesh(pts, cells), 3, 3, ele_type='HEX27', quadrature_order=order, dirichlet_bc_info=bc,
                      location_fns=[right], law=olaws.LinearElastic(70e3, 0.3),
                      surface_maps=[lambda u, x: np.array([0., 0., -50.]) + 0. * u])
    assert prob.fes[0].num_quads == (216 if order is None else 27)
    sol = 1e-3 * rng.standard_normal((len(pts), 3))
    res = prob.newton_update([torch.from_numpy(sol).cuda()])[0]
    ores = opb.newton_update(sol)
    assert relmax(host(prob.element_tangents()), opb.cell_jacobians(sol)) <= VAL_TOL
    assert relmax(host(res), ores) <= VAL_TOL
    A = jf.get_A(prob)
    oA = fem.get_A(opb)
    indptr, indices, data = [host(t) for t in A.getValuesCSR()]
    assert np.array_equal(indptr, oA.indptr) and np.array_equal(indices, oA.indices)
    assert relmax(data, oA.data) <= VAL_TOL
    usol = host(jf.solver(prob, {'jax_solver': {'method': 'cg'}})[0])
    osol = fem.solver(opb, method='cg')
    assert relmax(usol, osol) <= SOL_TOL


@pytest.mark.parametrize("mesh_kind", ["sheared", "mixed"])
@pytest.mark.parametrize("law_kind", ["elastic", "simp_cell", "simp_graded"])
def test_hex27_affine_cell_pass_matches_oracle(mesh_kind, law_kind):
    """The HEX27 affine-cell pass (K_e = E detJ J^-T Ghat J^-1, R_e = K_e u_e) against the oracle's quadrature: a sheared box
    (every cell affine), a mesh with curved cells on one side only (both kernels in one call), cell-constant SIMP density (affine
    pass) and a density that varies inside the cells (listed for the general kernel); the list says which kernel ran."""
    import jax_fem_b200 as jf
    from jax_fem_b200 import laws
    m = jf.box_mesh_hex27(4, 2, 2, 2.0, 1.0, 0.8)
    pts, cells = m.points.copy(), m.cells_dict['hexahedron27']
    rng = np.random.default_rng(11)
    pts = pts @ np.array([[1.0, 0.2, 0.1], [0.05, 0.9, 0.3], [0.0, 0.1, 1.2]]).T + 0.3
    curved = np.zeros(len(cells), dtype=bool)
    if mesh_kind == "mixed":
        move = pts[:, 0] > np.median(pts[:, 0])
        pts[move] += 0.01 * rng.uniform(-1, 1, (int(move.sum()), 3))
        curved = move[cells].any(axis=1)
        assert curved.any() and not curved.all()
    nq = 216
    iv = None
    if law_kind != "elastic":
        iv = np.repeat(rng.uniform(0.2, 1.0, (len(cells), 1)), nq, axis=1)
        if law_kind == "simp_graded":
            iv[::2] += 0.05 * rng.uniform(0, 1, (len(iv[::2]), nq))
    law = laws.LinearElasticity(70e3, 0.3) if law_kind == "elastic" else laws.SIMP(70e3, 70.0, 0.3, 3.0)
    olaw = olaws.LinearElastic(70e3, 0.3) if law_kind == "elastic" else olaws.SIMP(70e3, 70.0, 0.3, 3.0)
    P = type("P27", (jf.Problem,), {"get_tensor_map": lambda self: law})
    kw = {} if iv is None else {"internal_vars": [iv]}
    prob = P(jf.Mesh(pts, cells), vec=3, dim=3, ele_type='HEX27')
    if iv is not None:
        prob.internal_vars = [torch.from_numpy(iv).cuda()]
    opb = fem.Problem(fem.Mesh(pts, cells), 3, 3, ele_type='HEX27', law=olaw, **kw)
    sol = 1e-3 * rng.standard_normal((len(pts), 3))
    res = prob.newton_update([torch.from_numpy(sol).cuda()])[0]
    ores = opb.newton_update(sol)
    flags = np.zeros(len(cells), dtype=bool)
    flags[prob.hex27_general_cells()] = True
    expect = curved.copy()
    if law_kind == "simp_graded":
        expect[::2] = True
    assert np.array_equal(flags, expect)
    assert relmax(host(prob.element_tangents()), opb.cell_jacobians(sol)) <= VAL_TOL
    assert relmax(host(res), ores) <= VAL_TOL
    data = host(jf.get_A(prob).getValuesCSR()[2])
    assert relmax(data, fem.get_A(opb).data) <= VAL_TOL
    # residual-only call (no tangent) goes through the same pass
    res2 = prob.compute_residual([torch.from_numpy(sol).cuda()])[0]
    assert relmax(host(res2), ores) <= VAL_TOL


def test_cuda_core_element_path_for_elasticity_in_subprocess():
    """The HEX8 elasticity tangent has two kernels (FP64 tensor-core tiles by default, CUDA cores with
    FEM_ELEMENT_PATH=dfma); the choice is read once per process, so the second one is checked in a child process."""
    import os
    import subprocess
    import sys
    code = r"""
import numpy as np, torch, sys
sys.path.insert(0, 'tests')
import jax_fem_b200 as jf, gpu_problems as gp
from oracle import fem, laws as olaws
m = jf.box_mesh(4, 3, 3, 1.0, 0.8, 0.9)
pts = m.points + 0.02 * np.random.default_rng(0).uniform(-1, 1, m.points.shape)
cells = m.cells_dict['hexahedron']
sol = 0.01 * np.random.default_rng(1).standard_normal(pts.shape)
prob = gp.PlainElasticity(jf.Mesh(pts, cells), vec=3, dim=3)
prob.newton_update([torch.from_numpy(sol).cuda()])
opb = fem.Problem(fem.Mesh(pts, cells), 3, 3, law=olaws.LinearElastic(70e3, 0.3))
K, Ko = prob.element_tangents().cpu().numpy(), opb.cell_jacobians(sol)
print('RELMAX', np.abs(K - Ko).max() / np.abs(Ko).max())
"""
    env = dict(os.environ, FEM_ELEMENT_PATH="dfma")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=200)
    assert out.returncode == 0, out.stderr[-2000:]
    assert float(out.stdout.split("RELMAX")[1]) <= VAL_TOL


# ---- fused owner-computes assembly (csrc/fused.cu) ------------------------------------------------------
def _assemble_both_ways(prob, sol, monkeypatch):
    import jax_fem_b200 as jf
    out = {}
    for mode in ("fused", "staged", "ring"):
        monkeypatch.setenv("FEM_ASSEMBLY", mode)
        assert prob.assembly_mode() == mode and prob.fused_assembly_enabled() == (mode == "fused")
        res = prob.newton_update([torch.from_numpy(sol).cuda()])[0]
        A = jf.get_A(prob)
        prob.check_assembly_status()
        out[mode] = (host(res), host(A.data))
    return out


@pytest.mark.parametrize("case,config", [("perturbed_box", 0), ("perturbed_box", 1), ("perturbed_box", 2), ("perturbed_box", 3),
                                         ("perturbed_box", 4), ("cylinder", 0), ("cylinder", 1), ("cylinder", 4), ("simp", 1), ("simp", 4)])
def test_fused_assembly_matches_oracle_and_staged_path(case, config, monkeypatch):
    """One-kernel assembly (element evaluation + CSR rows + Dirichlet rows + nodal residual) against the oracle's
    get_A / compute_residual and against the two-kernel path, on non-affine and unstructured meshes."""
    import jax_fem_b200 as jf
    import gpu_problems as gp
    rng = np.random.default_rng(11)
    iv = None
    monkeypatch.setenv("FEM_FUSED_CONFIG", str(config))
    if case == "cylinder":
        g = cases.load_golden("linear_elasticity_cylinder")
        pts, cells = g["points"], g["cells"]
        prob = gp.LinearElasticityCylinder(jf.Mesh(pts, cells), vec=3, dim=3, dirichlet_bc_info=cases.CYL_BC,
                                           location_fns=[cases.top])
        opb = fem.Problem(fem.Mesh(pts, cells), 3, 3, dirichlet_bc_info=cases.CYL_BC, location_fns=[cases.top],
                          law=olaws.LinearElastic(70e3, 0.3), mass_map=cases.cyl_mass, surface_maps=[cases.cyl_traction])
    else:
        pts, cells = perturbed_box(9, seed=5)
        bc = [[lambda p: np.isclose(p[0], 0., atol=0.03)] * 2 + [lambda p: np.isclose(p[2], 0.9, atol=0.03)], [0, 2, 1],
              [lambda p: 0.01, lambda p: 0., lambda p: -0.02]]
        if case == "simp":
            iv = 0.2 + 0.7 * rng.uniform(0, 1, (len(cells), 8))
            prob = gp.SIMPElasticity(jf.Mesh(pts, cells), vec=3, dim=3, dirichlet_bc_info=bc,
                                     location_fns=[lambda p: np.isclose(p[0], 1.0, atol=0.03)])
            prob.internal_vars = [torch.from_numpy(iv).cuda()]
            opb = fem.Problem(fem.Mesh(pts, cells), 3, 3, dirichlet_bc_info=bc, law=olaws.SIMP(70e3, 70.0, 0.3, 3.0),
                              internal_vars=[iv])
        else:
            prob = gp.PlainElasticity(jf.Mesh(pts, cells), vec=3, dim=3, dirichlet_bc_info=bc)
            opb = fem.Problem(fem.Mesh(pts, cells), 3, 3, dirichlet_bc_info=bc, law=olaws.LinearElastic(70e3, 0.3))
    sol = 0.01 * rng.standard_normal((len(pts), 3))
    out = _assemble_both_ways(prob, sol, monkeypatch)
    law_only = fem.Problem(fem.Mesh(pts, cells), 3, 3, law=opb.law, internal_vars=() if iv is None else [iv])
    ores = np.zeros((len(pts), 3))
    np.add.at(ores, cells.reshape(-1), law_only.cell_residuals(sol).reshape(-1, 3))
    f_ext = host(prob._f_ext) if prob._f_ext is not None else 0.0
    opb.newton_update(sol)
    oA = fem.get_A(opb)
    for mode in ("fused", "staged", "ring"):
        res, data = out[mode]
        assert relmax(data, oA.data) <= VAL_TOL, mode
        assert relmax(res, ores + f_ext) <= VAL_TOL, mode
    assert relmax(out["fused"][1], out["staged"][1]) <= 1e-13      # same blocks, different (fixed) summation order
    assert relmax(out["ring"][1], out["staged"][1]) <= 1e-13       # isotropic map applied after the sum instead of before
    # Dirichlet rows are exact unit rows in both
    rows = host(prob.bc_data()[0])
    indptr, indices = host(prob.plan.indptr), host(prob.plan.indices)
    for r in rows[:: max(1, len(rows) // 50)]:
        seg = out["fused"][1][indptr[r]:indptr[r + 1]]
        assert np.array_equal(seg, (indices[indptr[r]:indptr[r + 1]] == r).astype(float))
    # bit-reproducible, and reference attributes stay available after a fused assembly
    monkeypatch.setenv("FEM_ASSEMBLY", "fused")
    prob.newton_update([torch.from_numpy(sol).cuda()])
    assert np.array_equal(host(jf.get_A(prob).data), out["fused"][1])
    assert relmax(host(prob.element_tangents()), opb.cell_jacobians(sol)) <= VAL_TOL


def test_fused_assembly_rejects_unregistered_combination():
    from jax_fem_b200 import _lib
    lib = _lib.load()
    z = torch.zeros(64, dtype=torch.float64, device='cuda')
    zi = torch.zeros(64, dtype=torch.int32, device='cuda')
    P = _lib.ptr
    code = lib.fem_assemble_fused(0, 3, 2, _lib.host_doubles([1., .3]), P(z), P(z), None, P(z), 1, *([P(zi)] * 14),
                                  P(zi), None, P(z), P(z), 1, None)
    assert code == -1 and b"fused assembly is registered" in lib.fem_last_error()


@pytest.mark.parametrize("ring_kb,tile,slack,margin", [(400, 64, 2, 8), (150, 10 ** 9, 0, 0), (4000, 200, 16, 64)])
def test_ring_assembly_recycles_rows_correctly(ring_kb, tile, slack, margin, monkeypatch):
    """The one-kernel staged assembly with a ring far smaller than the element tangents (rows recycled many times, strip
    edges spilled) on a 24 x 12 x 10 non-affine box: CSR values and residual equal the two-kernel path bit for bit in the
    pattern and to 1e-13 in the values, every run gives the same bits, and no wait times out."""
    import jax_fem_b200 as jf
    import gpu_problems as gp
    m = jf.box_mesh(24, 12, 10, 2.4, 1.2, 1.0)
    rng = np.random.default_rng(3)
    pts = m.points + 0.015 * rng.uniform(-1, 1, m.points.shape)
    cells = m.cells_dict['hexahedron']
    bc = [[lambda p: p[0] < 0.03] * 3, [0, 1, 2], [lambda p: 0., lambda p: 0.01, lambda p: -0.01]]
    sol = torch.from_numpy(0.01 * rng.standard_normal((len(pts), 3))).cuda()
    for k, v in (("RING_BYTES", ring_kb * 1024), ("TILE_CELLS", tile), ("SLACK", slack), ("MARGIN", margin), ("IN_FLIGHT", 0)):
        monkeypatch.setenv("FEM_RING_" + k, str(v))
    monkeypatch.setenv("FEM_ASSEMBLY", "staged")
    ref = gp.PlainElasticity(jf.Mesh(pts, cells), vec=3, dim=3, dirichlet_bc_info=bc)
    res0 = ref.newton_update([sol])[0]
    data0 = jf.get_A(ref).data
    monkeypatch.setenv("FEM_ASSEMBLY", "ring")
    prob = gp.PlainElasticity(jf.Mesh(pts, cells), vec=3, dim=3, dirichlet_bc_info=bc)
    sp = prob.stage_plan
    if ring_kb == 400:
        assert sp.ring_rows > 0 and int(sp.prev_g.max()) >= 0 and 0 < sp.spill_fraction < 1
    res = prob.newton_update([sol])[0]
    data = jf.get_A(prob).data.clone()
    prob.check_assembly_status()
    assert relmax(host(data), host(data0)) <= 1e-13 and relmax(host(res), host(res0)) <= 1e-13
    for _ in range(5):
        prob.newton_update([sol])
        assert torch.equal(jf.get_A(prob).data, data)
    prob.check_assembly_status()


@pytest.mark.parametrize("case", ["box", "renumbered", "quad4", "hex27", "cylinder"])
def test_native_plan_equals_torch_plan(case):
    """fem_plan_create (csrc/plan.cu: Thrust on raw device buffers) must produce exactly the tables of the torch construction
    (jax_fem_b200/plan.py, the one the CPU tests pin against the oracle's COO pattern): pattern, node-block graph, corner
    order, gather schedule, transpose map and, for a Dirichlet mask, the emeta rows -- bit for bit."""
    import jax_fem_b200 as jf
    from jax_fem_b200.plan import build_plan, build_plan_native
    rng = np.random.default_rng(4)
    vec = 3
    if case == "quad4":
        m = jf.rectangle_mesh(13, 9, 1.3, 0.9)
        cells, n_nodes, vec = m.cells_dict['quad'], len(m.points), 2
    elif case == "hex27":
        m = jf.box_mesh_hex27(4, 3, 3, 1., 1., 1.)
        cells, n_nodes = m.cells_dict['hexahedron27'], len(m.points)
    elif case == "cylinder":
        g = cases.load_golden("linear_elasticity_cylinder")
        cells, n_nodes = g["cells"], len(g["points"])
    else:
        m = jf.box_mesh(11, 7, 5, 1., 1., 1.)
        cells, n_nodes = m.cells_dict['hexahedron'], len(m.points)
        if case == "renumbered":
            perm = rng.permutation(n_nodes)
            cells = perm[cells][rng.permutation(len(cells))]
    ct = torch.from_numpy(np.ascontiguousarray(cells).astype(np.int32)).cuda()
    a, b = build_plan(ct, n_nodes, vec), build_plan_native(ct, n_nodes, vec)
    for name in ("brow_ptr", "bcol", "src_ptr", "src", "nc_ptr", "nc", "corner_pos", "indptr", "indices", "gdesc", "m_sb", "m_se",
                 "m_ent", "m_add", "edst", "erow"):
        assert torch.equal(getattr(a, name).to(torch.int32), getattr(b, name)), name
    assert torch.equal(a.tperm, b.tperm)
    flag = torch.from_numpy((rng.uniform(size=n_nodes * vec) < 0.2).astype(np.uint8)).cuda()
    assert torch.equal(a.entry_meta(flag), b.entry_meta(flag)) and torch.equal(a.entry_meta(None), b.entry_meta(None))
    assert (a.nnz, a.nnzb, a.n_gather_blocks) == (b.nnz, b.nnzb, b.n_gather_blocks)


def test_c_abi_pipeline_without_torch():
    """plan -> element kernel -> gathers -> Dirichlet rows -> Jacobi-CG through ctypes on cudaMalloc'd buffers, in an
    interpreter that never imports torch (tests/no_torch_pipeline.py), against the oracle."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "no_torch_pipeline.py")], cwd=root, capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    assert "C_ABI_PIPELINE_OK" in out.stdout


@pytest.mark.parametrize("law", ["elastic", "simp", "neohookean"])
def test_tile_major_staging_equals_block_staging(law, monkeypatch):
    """The default HEX8 hot path stages tile-major rows written straight from the tensor-core fragments and applies the
    isotropic map after the sum (fem_element_tiles + fem_gather_csr_tiles); FEM_ELEMENT_PATH=blocks stages K blocks in the
    reference's V layout.  Same CSR values (1e-13: the map is applied to the sum instead of to every block), same residual,
    and problem.V / element_tangents() keep the reference layout either way."""
    import jax_fem_b200 as jf
    import gpu_problems as gp
    pts, cells = perturbed_box(8, seed=3)
    rng = np.random.default_rng(17)
    bc = [[lambda p: p[0] < 0.05] * 3, [0, 1, 2], [lambda p: 0., lambda p: 0.01, lambda p: -0.01]]
    sol = torch.from_numpy(0.01 * rng.standard_normal((len(pts), 3))).cuda()
    out = {}
    for path in ("blocks", "tiles"):
        monkeypatch.setenv("FEM_ELEMENT_PATH", path)
        if law == "neohookean":
            prob = gp.NeoHookeanInverse(jf.Mesh(pts, cells), vec=3, dim=3, dirichlet_bc_info=bc, location_fns=[lambda p: p[1] > 0.95])
            prob.internal_vars = [torch.from_numpy(0.5 + rng.uniform(0, 1, (len(cells), 8))).cuda()] if path == "blocks" else out["iv"]
            out["iv"] = prob.internal_vars
        elif law == "simp":
            prob = gp.SIMPElasticity(jf.Mesh(pts, cells), vec=3, dim=3, dirichlet_bc_info=bc, location_fns=[lambda p: p[0] > 0.95])
            prob.internal_vars = [torch.from_numpy(0.2 + 0.7 * rng.uniform(0, 1, (len(cells), 8))).cuda()] if path == "blocks" else out["iv"]
            out["iv"] = prob.internal_vars
        else:
            prob = gp.PlainElasticity(jf.Mesh(pts, cells), vec=3, dim=3, dirichlet_bc_info=bc)
        assert prob.tiles_enabled() == (path == "tiles")
        res = prob.newton_update([sol])[0]
        A = jf.get_A(prob)
        out[path] = (host(res), host(A.data), host(prob.element_tangents()))
    assert relmax(out["tiles"][1], out["blocks"][1]) <= 1e-13
    assert relmax(out["tiles"][0], out["blocks"][0]) <= 1e-13
    assert np.array_equal(out["tiles"][2], out["blocks"][2])


@pytest.mark.parametrize("case", ["robin_poisson", "spring_elasticity", "robin_quad4"])
@pytest.mark.parametrize("mode", ["staged", "ring", "fused"])
def test_solution_dependent_surface_maps_match_oracle(case, mode, monkeypatch):
    """registered u-dependent surface maps (laws.RobinPower; the reference's robin_bc map 5 u^2 and a
    spring foundation).  Face residual and face tangent are added by csrc/faces.cu to the nodal residual and to the assembled CSR
    values; residual, CSR values, problem.V (face blocks) and the Newton solution must equal the oracle's."""
    import jax_fem_b200 as jf
    import gpu_problems as gp
    rng = np.random.default_rng(8)
    if case == "robin_quad4":
        if mode != "staged":
            pytest.skip("one-kernel modes are registered for HEX8 elasticity")
        m = jf.rectangle_mesh(9, 7, 1., 1.)
        pts, cells, ele, dim, vec = m.points + 0.01 * rng.uniform(-1, 1, m.points.shape), m.cells_dict['quad'], 'QUAD4', 2, 1
        m_pts = m.points
    else:
        m_pts, cells = perturbed_box(6, seed=2)
        pts, ele, dim, vec = m_pts, 'HEX8', 3, (1 if case == "robin_poisson" else 3)
    lo = lambda p: p[0] < 0.02
    hi = lambda p: p[0] > 0.98
    side = lambda p: p[1] > 0.98
    monkeypatch.setenv("FEM_ASSEMBLY", mode)
    if vec == 1:
        bc = [[lo], [0], [lambda p: 0.4]]
        cls = type("RobinQ", (gp.RobinPoisson,), {"get_mass_map": lambda self: (lambda u, x: -np.array([3.0]))}) if dim == 2 else gp.RobinPoisson
        prob = cls(jf.Mesh(pts, cells), vec=1, dim=dim, ele_type=ele, dirichlet_bc_info=bc, location_fns=[hi, side])
        mass = (lambda u, x: -3.0 + 0. * u) if dim == 2 else \
            (lambda u, x: -10. * np.exp(-((x[..., 0] - .5) ** 2 + (x[..., 1] - .5) ** 2) / 0.02)[..., None] + 0. * u)
        opb = fem.Problem(fem.Mesh(pts, cells), 1, dim, ele_type=ele, dirichlet_bc_info=bc, location_fns=[hi, side], law=olaws.Poisson(1.0),
                          mass_map=mass, surface_maps=[lambda u, x: 5 * u ** 2, lambda u, x: 2.0 * (u - 0.3)],
                          surface_map_jacs=[lambda u, x: (10 * u)[..., None], lambda u, x: 2.0 * np.ones(u.shape + (1,))])
    else:
        bc = [[lo] * 3, [0, 1, 2], [lambda p: 0., lambda p: 0.01, lambda p: 0.]]
        k = np.array([3e3, 5e3, 7e3])
        prob = gp.SpringFoundation(jf.Mesh(pts, cells), vec=3, dim=3, dirichlet_bc_info=bc, location_fns=[hi, side])
        opb = fem.Problem(fem.Mesh(pts, cells), 3, 3, dirichlet_bc_info=bc, location_fns=[hi, side], law=olaws.LinearElastic(70e3, 0.3),
                          surface_maps=[lambda u, x: k * u, lambda u, x: np.array([0., 0., 100.]) + 0. * u],
                          surface_map_jacs=[lambda u, x: np.broadcast_to(np.diag(k), u.shape + (3,)), None])
    if prob.assembly_mode() != mode:
        pytest.skip(f"{mode} is not registered for this problem")
    sol = 0.3 + 0.1 * rng.standard_normal((len(pts), vec))
    res = prob.newton_update([torch.from_numpy(sol).cuda()])[0]
    A = jf.get_A(prob)
    ores = opb.newton_update(sol)
    oA = fem.get_A(opb)
    assert np.array_equal(host(A.getValuesCSR()[1]), oA.indices)
    assert relmax(host(A.data), oA.data) <= VAL_TOL and relmax(host(res), ores) <= VAL_TOL
    assert relmax(host(prob.V), opb.coo_values()) <= VAL_TOL                    # face blocks of the reference's V
    x = jf.solver(prob, {'jax_solver': {}})[0]
    assert relmax(host(x), fem.solver(opb)) <= SOL_TOL


def test_solution_dependent_surface_map_on_hex27():
    """The registered Robin law on the 9-node faces of HEX27 cells (spring foundation + dead load), curved cells: residual, CSR
    values and the Newton solution against the oracle.  (The face set must hold every cell node that lives on the face, not only
    the four vertices fe.face_inds lists.)"""
    import jax_fem_b200 as jf
    import gpu_problems as gp
    from jax_fem_b200 import laws
    m = jf.box_mesh_hex27(3, 2, 2, 1.5, 1.0, 0.8)
    pts, cells = m.points.copy(), m.cells_dict['hexahedron27']
    rng = np.random.default_rng(12)
    inner = (pts[:, 0] > 0.03) & (pts[:, 0] < 1.47) & (pts[:, 1] < 0.97)
    pts[inner] += 0.01 * rng.uniform(-1, 1, (int(inner.sum()), 3))
    lo = lambda p: p[0] < 0.02
    hi = lambda p: p[0] > 1.48
    side = lambda p: p[1] > 0.98
    bc = [[lo] * 3, [0, 1, 2], [lambda p: 0., lambda p: 0.01, lambda p: 0.]]
    k = np.array([3e3, 5e3, 7e3])
    prob = gp.SpringFoundation(jf.Mesh(pts, cells), vec=3, dim=3, ele_type='HEX27', quadrature_order=4, dirichlet_bc_info=bc,
                               location_fns=[hi, side])
    opb = fem.Problem(fem.Mesh(pts, cells), 3, 3, ele_type='HEX27', quadrature_order=4, dirichlet_bc_info=bc, location_fns=[hi, side],
                      law=olaws.LinearElastic(70e3, 0.3),
                      surface_maps=[lambda u, x: k * u, lambda u, x: np.array([0., 0., 100.]) + 0. * u],
                      surface_map_jacs=[lambda u, x: np.broadcast_to(np.diag(k), u.shape + (3,)), None])
    sol = 0.01 * rng.standard_normal((len(pts), 3))
    res = prob.newton_update([torch.from_numpy(sol).cuda()])[0]
    A = jf.get_A(prob)
    ores = opb.newton_update(sol)
    oA = fem.get_A(opb)
    assert np.array_equal(host(A.getValuesCSR()[1]), oA.indices)
    assert relmax(host(A.data), oA.data) <= VAL_TOL and relmax(host(res), ores) <= VAL_TOL
    x = jf.solver(prob, {'jax_solver': {}})[0]
    assert relmax(host(x), fem.solver(opb)) <= SOL_TOL


@pytest.mark.parametrize("case", ["heat_step_hex8", "heat_step_quad4", "phase_field_hex8", "foundation_hex8", "foundation_quad4",
                                  "foundation_hex27"])
def test_solution_dependent_mass_maps_match_oracle(case):
    """registered u-dependent mass maps (laws.LinearMass, csrc/mass.cu): the backward-Euler heat capacity
    rho Cp (T - T_old) / dt of applications/thermal_mechanical (constant coefficient, per-point T_old), the phase-field driving
    term (G_c / l + 2 H) d - 2 H (per-point coefficient and constant), and an elastic foundation k u with a body force on
    vec 3 / vec 2 elasticity.  Residual, CSR values, problem.V and the solution of the step must equal the oracle's; the fields
    are then changed in place (next time step) and the same objects must follow."""
    import jax_fem_b200 as jf
    from jax_fem_b200 import laws
    rng = np.random.default_rng(21)
    quad = case.endswith("quad4")
    extra = {}
    if quad:
        m = jf.rectangle_mesh(8, 6, 1., 1.)
        pts, cells, ele, dim = m.points + 0.01 * rng.uniform(-1, 1, m.points.shape), m.cells_dict['quad'], 'QUAD4', 2
    elif case.endswith("hex27"):
        m = jf.box_mesh_hex27(2, 2, 1, 1., 1., 0.5)
        pts, cells, ele, dim = m.points.copy(), m.cells_dict['hexahedron27'], 'HEX27', 3
        inner = (pts[:, 0] > 0.02) & (pts[:, 0] < 0.98)
        pts[inner] += 0.01 * rng.uniform(-1, 1, (int(inner.sum()), 3))
        extra = {"quadrature_order": 4}
    else:
        pts, cells = perturbed_box(5, seed=4)
        ele, dim = 'HEX8', 3
    nq = 4 if quad else (27 if ele == 'HEX27' else 8)
    C = len(cells)
    lo = lambda p: p[0] < 0.02
    hi = lambda p: p[0] > 0.98
    if case.startswith("foundation"):
        vec = dim
        kf, body = 2.5e3, np.array([0., -40., 15.])[:vec]
        law, olaw = laws.LinearElasticity(70e3, 0.3), olaws.LinearElastic(70e3, 0.3)
        mass = laws.LinearMass(kf, body)
        fields = lambda: (kf * np.ones((C, nq)), np.broadcast_to(body, (C, nq, vec)))
        bc = [[lo] * vec, list(range(vec)), [lambda p: 0.] * vec]
    else:
        vec = 1
        law, olaw = laws.Poisson(2.0), olaws.Poisson(2.0)
        bc = [[lo], [0], [lambda p: 0.4]]
        if case.startswith("heat"):
            coef = 7.0 / 0.05                                                  # rho Cp / dt
            T_old = 0.5 + 0.2 * rng.standard_normal((C, nq))
            mass = laws.LinearMass(coef, -coef * T_old)
            fields = lambda: (coef * np.ones((C, nq)), np.asarray(mass.const)[..., None])
        else:
            H = rng.uniform(0., 3., (C, nq))
            mass = laws.LinearMass(torch.from_numpy(1.5 + 2 * H).cuda(), torch.from_numpy(-2 * H).cuda())
            fields = lambda: (host(mass.coef), host(mass.const)[..., None])
    P = type("MassProblem", (jf.Problem,), {"get_tensor_map": lambda self: law, "get_mass_map": lambda self: mass,
                                            "get_surface_maps": lambda self: [lambda u, x: np.full(vec, 3.0)]})
    prob = P(jf.Mesh(pts, cells), vec=vec, dim=dim, ele_type=ele, dirichlet_bc_info=bc, location_fns=[hi], **extra)
    opb = fem.Problem(fem.Mesh(pts, cells), vec, dim, ele_type=ele, dirichlet_bc_info=bc, location_fns=[hi], law=olaw, **extra,
                      mass_map=lambda u, x: fields()[0][..., None] * u + fields()[1],
                      mass_map_jac=lambda u, x: fields()[0][..., None, None] * np.eye(vec),
                      surface_maps=[lambda u, x: 3.0 + 0. * u])
    assert prob.assembly_mode() == 'staged' and not prob.tiles_enabled()
    sol = 0.3 + 0.1 * rng.standard_normal((len(pts), vec))
    for step in range(2):
        res = prob.newton_update([torch.from_numpy(sol).cuda()])[0]
        A = jf.get_A(prob)
        ores = opb.newton_update(sol)
        oA = fem.get_A(opb)
        assert np.array_equal(host(A.getValuesCSR()[1]), oA.indices)
        assert relmax(host(A.data), oA.data) <= VAL_TOL and relmax(host(res), ores) <= VAL_TOL
        assert relmax(host(prob.V), opb.coo_values()) <= VAL_TOL
        assert relmax(host(prob.compute_residual([torch.from_numpy(sol).cuda()])[0]), ores) <= VAL_TOL    # residual-only call
        x = host(jf.solver(prob, {'jax_solver': {}})[0])
        assert relmax(x, fem.solver(opb)) <= SOL_TOL
        if case.startswith("heat"):                                            # next time step: T_old <- T at the points
            T_q = np.einsum('cn,qn->cq', x[cells, 0], prob.fes[0].shape_vals)
            mass.const = -coef * T_q
        elif case.startswith("phase"):
            mass.coef.mul_(1.1)
        else:
            break


def test_c_abi_empty_inputs_and_argument_errors():
    """Edge cases straight at the C ABI: empty meshes / boundary sets are no-ops that return FEM_OK, missing pointers and
    unregistered combinations return FEM_EINVAL with a message in fem_last_error (never a crash, never a silent fallback)."""
    from jax_fem_b200 import _lib
    lib, P = _lib.load(), _lib.ptr
    dev = 'cuda'
    d = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=dev)
    i32 = lambda *shape: torch.zeros(shape, dtype=torch.int32, device=dev)
    pts, cells, sol, ref, Re = d(8, 3), i32(1, 8), d(8, 3), d(8 * 8 * 3 + 8), d(1, 24)
    params = _lib.host_doubles([70e3, 0.3])
    ok = lambda rc: rc == 0
    einval = lambda rc, text: rc == -1 and text in lib.fem_last_error().decode()
    # zero cells / zero boundary nodes / zero blocks: nothing is launched
    assert ok(lib.fem_element_residual_jacobian(0, 3, 1, params, P(pts), P(cells), 0, P(sol), None, P(ref), None, None, P(Re), None))
    assert ok(lib.fem_element_tiles(1, params, P(pts), P(cells), 0, P(sol), None, P(ref), None, P(d(72)), P(Re),
                                    (_lib.ctypes.c_double * 3)(), None))
    assert ok(lib.fem_hex27_residual_jacobian(1, params, P(pts), P(cells), 0, P(sol), None, P(ref), None, 27, None, None, None,
                                              None, P(Re), None))
    assert ok(lib.fem_mass_term(0, 1, P(pts), P(cells), 0, P(sol), P(ref), P(d(64)), 8, 1.0, None, _lib.host_doubles([0.]), None,
                                None, None, P(Re), None))
    assert ok(lib.fem_gather_csr(3, 8, 0, P(i32(4)), P(i32(4)), P(i32(4)), P(d(72)), P(d(9)), None))
    assert ok(lib.fem_apply_bc_vec(0, None, None, 1.0, P(d(3)), P(d(3)), None))
    # missing pointers
    assert einval(lib.fem_element_residual_jacobian(0, 3, 1, params, None, P(cells), 1, P(sol), None, P(ref), None, None, P(Re), None),
                  "null pointer")
    assert einval(lib.fem_mass_term(0, 1, P(pts), P(cells), 1, P(sol), P(ref), None, 8, 1.0, None, _lib.host_doubles([0.]), None,
                                    None, None, P(Re), None), "null pointer")
    # unregistered combinations
    assert einval(lib.fem_element_residual_jacobian(0, 2, 1, params, P(pts), P(cells), 1, P(sol), None, P(ref), None, None, P(Re), None), "")
    assert einval(lib.fem_mass_term(2, 1, P(pts), P(cells), 1, P(sol), P(ref), P(d(64)), 8, 1.0, None, _lib.host_doubles([0.]),
                                    None, None, None, P(Re), None), "unregistered")
    assert einval(lib.fem_mass_term(0, 1, P(pts), P(cells), 1, P(sol), P(ref), P(d(64)), 64, 1.0, None, _lib.host_doubles([0.]), None,
                                    None, None, P(Re), None), "quadrature")
    assert einval(lib.fem_hex27_residual_jacobian(2, params, P(pts), P(cells), 1, P(sol), None, P(ref), None, 27, None, None, None,
                                                  None, P(Re), None), "HEX27")
    # the affine pass needs both its tables and its workspace
    assert einval(lib.fem_hex27_residual_jacobian(1, params, P(pts), P(cells), 1, P(sol), None, P(ref), None, 27, P(d(8)), None, None,
                                                  None, P(Re), None), "affine")
    # SIMP without its density field
    assert einval(lib.fem_hex27_residual_jacobian(3, params, P(pts), P(cells), 1, P(sol), None, P(ref), None, 27, None, None, None,
                                                  None, P(Re), None), "density")
    torch.cuda.synchronize()


def test_csr_diagonal_beyond_2_30_nonzeros():
    """The Jacobi preconditioner of the 200^3 mesh (nnz = 1.95e9): row offsets above 2^30 must not overflow the binary search
    for the diagonal entry (lo + hi in int32 did: the first 200^3 solve on one GPU hung).  Synthetic banded matrix with
    nnz = 1.09e9 whose value at (i, j) is j, so diag[i] must be i."""
    from jax_fem_b200 import _lib
    n, w = 39_000_000, 28
    dev = 'cuda'
    start = (torch.arange(n, device=dev, dtype=torch.int64) - 14).clamp_(0, n - w)
    indptr = (torch.arange(n + 1, device=dev, dtype=torch.int64) * w).to(torch.int32)
    assert int(indptr[-1]) > 2 ** 30
    indices = torch.empty(n * w, dtype=torch.int32, device=dev)
    for a in range(0, n, 1 << 22):
        b = min(n, a + (1 << 22))
        indices[a * w:b * w] = (start[a:b, None] + torch.arange(w, device=dev)[None, :]).reshape(-1).to(torch.int32)
    data = indices.to(torch.float64)
    diag = torch.empty(n, dtype=torch.float64, device=dev)
    P = _lib.ptr
    _lib.check(_lib.load().fem_csr_diagonal(n, P(indptr), P(indices), P(data), P(diag), None))
    torch.cuda.synchronize()
    assert torch.equal(diag, torch.arange(n, device=dev, dtype=torch.float64))


# ---- full-size (BASELINE.json configs[1]) checks through size-independent properties -------------------------------
def test_full_size_cfg2_properties(monkeypatch):
    """HEX8 100^3 linear elasticity (3.09 M DOF, nnz 245 M) -- too large for the oracle, so the assembled operator is
    checked through properties that hold at any size: pattern formula, rigid-body null space (translations and
    infinitesimal rotations), symmetry (via the transpose kernel), linearity res(u) = A u + res(0), bit-reproducibility,
    fused == two-kernel assembly, unit Dirichlet rows, and a Jacobi-CG solve whose true residual meets the tolerance."""
    import jax_fem_b200 as jf
    import gpu_problems as gp
    from jax_fem_b200.solver import jax_solve
    N = 100
    m = jf.box_mesh(N, N, N, 1., 1., 1.)
    pts, cells = m.points, m.cells_dict['hexahedron']
    free = gp.PlainElasticity(jf.Mesh(pts, cells), vec=3, dim=3)                     # no Dirichlet rows
    n = free.num_total_dofs_all_vars
    assert free.plan.nnz == 9 * (3 * (N + 1) - 2) ** 3 and n == 3 * (N + 1) ** 3
    rng = np.random.default_rng(7)
    u = torch.from_numpy(1e-3 * rng.standard_normal((len(pts), 3))).cuda()
    monkeypatch.setenv("FEM_ASSEMBLY", "staged")
    res_u = free.newton_update([u])[0]
    A = jf.get_A(free)
    data0 = A.data.clone()
    free.newton_update([u])
    assert torch.equal(jf.get_A(free).data, data0)                                   # bit-reproducible at full size
    scale = float(A.data.abs().max())
    X = torch.from_numpy(pts).cuda()
    rigid = [torch.tensor(t, dtype=torch.float64, device='cuda').expand(len(pts), 3).contiguous() for t in
             ([1., 0, 0], [0, 1., 0], [0, 0, 1.])]
    rigid += [torch.stack([-X[:, 1], X[:, 0], 0 * X[:, 0]], 1), torch.stack([0 * X[:, 0], -X[:, 2], X[:, 1]], 1),
              torch.stack([X[:, 2], 0 * X[:, 0], -X[:, 0]], 1)]
    for v in rigid:
        y = A @ v.reshape(-1).contiguous()
        assert float(y.abs().max()) <= 1e-12 * scale * 27 * float(v.abs().max())
    assert float((A.transpose().data - A.data).abs().max()) <= 1e-12 * scale           # K_ba = K_ab^T
    zero = torch.zeros_like(u)
    res_0 = free.compute_residual([zero])[0]
    lin = (res_u - res_0).reshape(-1) - A @ u.reshape(-1).contiguous()
    assert float(lin.abs().max()) <= 1e-12 * scale * float(u.abs().max()) * 27
    monkeypatch.setenv("FEM_ASSEMBLY", "fused")
    res_f = free.newton_update([u])[0]
    assert float((jf.get_A(free).data - data0).abs().max()) <= 1e-13 * scale
    assert float((res_f - res_u).abs().max()) <= 1e-12 * float(res_u.abs().max())
    # the default: one persistent kernel, element tangents staged in the L2-resident ring (csrc/staged.cu)
    monkeypatch.setenv("FEM_ASSEMBLY", "ring")
    res_r = free.newton_update([u])[0]
    data_r = jf.get_A(free).data.clone()
    free.check_assembly_status()
    assert free.stage_plan.ring_rows > 0 and free.stage_plan.spill_fraction < 0.2
    assert free.stage_plan.staging_bytes < 0.2 * 8 * 72 * 8 * len(cells)            # no full-size element-tangent buffer
    assert float((data_r - data0).abs().max()) <= 1e-13 * scale
    assert float((res_r - res_u).abs().max()) <= 1e-12 * float(res_u.abs().max())
    for _ in range(3):                                                              # bit-reproducible under any schedule
        free.newton_update([u])
        assert torch.equal(jf.get_A(free).data, data_r)
    free.check_assembly_status()
    del free, A, data0, data_r
    torch.cuda.empty_cache()

    # cfg 2 proper: u = 0 on x = 0, traction on x = 1, Jacobi-CG to 1e-10 (default assembly mode)
    monkeypatch.delenv("FEM_ASSEMBLY")
    left = lambda p: np.isclose(p[0], 0., atol=1e-5)
    right = lambda p: np.isclose(p[0], 1., atol=1e-5)
    cls = type("Cfg2", (jf.Problem,), {"get_tensor_map": lambda self: jf.laws.LinearElasticity(70e3, 0.3),
                                       "get_surface_maps": lambda self: [lambda u, x: np.array([0., 0., 100.])]})
    pb = cls(jf.Mesh(pts, cells), vec=3, dim=3, dirichlet_bc_info=[[left] * 3, [0, 1, 2], [lambda p: 0.] * 3],
             location_fns=[right])
    dofs = torch.zeros(n, dtype=torch.float64, device='cuda')
    b = -jf.apply_bc_vec(pb.newton_update([dofs.reshape(-1, 3)])[0].reshape(-1), dofs, pb)
    A = jf.get_A(pb)
    rows = pb.bc_data()[0].long()
    indptr, indices, data = A.getValuesCSR()
    r0 = int(rows[len(rows) // 2])
    seg = slice(int(indptr[r0]), int(indptr[r0 + 1]))
    assert torch.equal(data[seg], (indices[seg] == r0).double())                      # unit Dirichlet row, pattern kept
    assert len(rows) == 3 * (N + 1) ** 2
    x, info = jax_solve(A, b, torch.zeros_like(b), True, method='cg', return_info=True)
    assert 0 < info['iterations'] < 3000
    true_res = float((A @ x - b).norm())
    assert true_res <= 1e-9 * max(float(b.norm()), 1.0) and info['err'] < 1e-6
    assert float(x[rows].abs().max()) == 0.0


def test_full_size_neohookean_tangent_is_the_derivative_of_the_residual():
    """cfg 3's law at 100^3 cells (one GPU's share of the 8-GPU run): the assembled tangent (FP64 DMMA kernel) must be the
    directional derivative of the assembled residual, A(u) v = d/de res(u + e v) (central difference, O(e^2)), and must be
    symmetric (hyperelastic).  Size-independent, no oracle needed."""
    import jax_fem_b200 as jf
    import gpu_problems as gp
    N = 100
    m = jf.box_mesh(N, N, N, 1., 1., 1.)
    pts, cells = m.points, m.cells_dict['hexahedron']
    prob = gp.HyperElasticity(jf.Mesh(pts, cells), vec=3, dim=3)
    rng = np.random.default_rng(3)
    u = torch.from_numpy(2e-4 * rng.standard_normal((len(pts), 3))).cuda()            # |grad u| ~ 0.03: det F > 0 everywhere
    v = torch.from_numpy(rng.standard_normal((len(pts), 3))).cuda()
    prob.newton_update([u])
    A = jf.get_A(prob)
    Av = A @ v.reshape(-1).contiguous()
    eps = 2e-7                                          # truncation ~ (eps |grad v|)^2 ~ 4e-9, round-off ~ 1e-16 |res| / eps
    fd = (prob.compute_residual([u + eps * v])[0] - prob.compute_residual([u - eps * v])[0]).reshape(-1) / (2 * eps)
    err, ref = float((Av - fd).abs().max()), float(Av.abs().max())
    assert err <= 1e-7 * ref, (err, ref)
    assert float((A.transpose().data - A.data).abs().max()) <= 1e-11 * float(A.data.abs().max())


@pytest.mark.parametrize("law_name", ["elastic_plane_strain", "elastic_plane_stress", "simp_default_plane_stress"])
def test_quad4_elasticity_and_plane_stress_simp_match_oracle(law_name):
    """2-D elasticity on QUAD4: plane strain, plane stress, and SIMP with the reference's own 2-D choice (plane stress,
    docs/source/learn/topology_optimization/example.ipynb cell 9), element values and assembled operator vs the oracle."""
    import jax_fem_b200 as jf
    from jax_fem_b200 import laws
    m = jf.rectangle_mesh(7, 5, 1.4, 1.0)
    pts = m.points.copy()
    rng = np.random.default_rng(2)
    pts += 0.03 * rng.uniform(-0.5, 0.5, pts.shape)
    cells = m.cells_dict['quad']
    iv = None
    if law_name == "simp_default_plane_stress":
        law, olaw = laws.SIMP(70e3, 70.0, 0.3, 3.0), olaws.SIMP(70e3, 70.0, 0.3, 3.0, plane_stress=True)
        iv = 0.2 + 0.7 * rng.uniform(0, 1, (len(cells), 4))
    else:
        ps = law_name.endswith("stress")
        law, olaw = laws.LinearElasticity(70e3, 0.3, plane_stress=ps), olaws.LinearElastic(70e3, 0.3, plane_stress=ps)
    left = lambda p: np.isclose(p[0], 0., atol=0.05)
    bc = [[left, left], [0, 1], [lambda p: 0., lambda p: 0.01]]
    cls = type("P2D", (jf.Problem,), {"get_tensor_map": lambda self: law})
    prob = cls(jf.Mesh(pts, cells), vec=2, dim=2, ele_type='QUAD4', dirichlet_bc_info=bc)
    assert law.plane_stress == (law_name != "elastic_plane_strain")
    if iv is not None:
        prob.internal_vars = [torch.from_numpy(iv).cuda()]
    opb = fem.Problem(fem.Mesh(pts, cells), 2, 2, ele_type='QUAD4', dirichlet_bc_info=bc, law=olaw,
                      internal_vars=() if iv is None else [iv])
    sol = 0.01 * rng.standard_normal((len(pts), 2))
    res = prob.newton_update([torch.from_numpy(sol).cuda()])[0]
    A = jf.get_A(prob)
    assert relmax(host(prob.element_tangents()), opb.cell_jacobians(sol)) <= VAL_TOL
    ores = opb.newton_update(sol)
    oA = fem.get_A(opb)
    assert np.array_equal(host(A.getValuesCSR()[1]), oA.indices) and relmax(host(A.data), oA.data) <= VAL_TOL
    assert relmax(host(res), ores) <= VAL_TOL
    with pytest.raises(laws.UnregisteredLawError):
        laws.resolve(laws.LinearElasticity(1., .3, plane_stress=True), 'HEX8', 3)


@pytest.mark.parametrize("mode", ["staged", "fused", "ring"])
def test_assembly_on_randomly_renumbered_mesh(mode, monkeypatch):
    """Generality of the plans: a non-affine box whose nodes and cells are randomly renumbered (no tensor-grid structure in
    the numbering, scattered CSR rows, bin-based patches) must assemble to the oracle's operator in both assembly modes."""
    import jax_fem_b200 as jf
    import gpu_problems as gp
    pts, cells = perturbed_box(12, seed=9)
    rng = np.random.default_rng(21)
    perm = rng.permutation(len(pts))                 # new id of old node i
    pts2 = np.empty_like(pts)
    pts2[perm] = pts
    cells2 = perm[cells][rng.permutation(len(cells))]
    bc = [[lambda p: p[0] < 0.05] * 3, [0, 1, 2], [lambda p: 0., lambda p: 0.01, lambda p: -0.01]]
    monkeypatch.setenv("FEM_ASSEMBLY", mode)
    prob = gp.PlainElasticity(jf.Mesh(pts2, cells2), vec=3, dim=3, dirichlet_bc_info=bc)
    opb = fem.Problem(fem.Mesh(pts2, cells2), 3, 3, dirichlet_bc_info=bc, law=olaws.LinearElastic(70e3, 0.3))
    sol = 0.01 * rng.standard_normal((len(pts2), 3))
    res = prob.newton_update([torch.from_numpy(sol).cuda()])[0]
    A = jf.get_A(prob)
    ores = opb.newton_update(sol)
    oA = fem.get_A(opb)
    indptr, indices, data = [host(t) for t in A.getValuesCSR()]
    assert np.array_equal(indptr, oA.indptr) and np.array_equal(indices, oA.indices)
    assert relmax(data, oA.data) <= VAL_TOL and relmax(host(res), ores) <= VAL_TOL
    x = jf.solver(prob, {'jax_solver': {'method': 'cg'}})[0]
    assert relmax(host(x), fem.solver(opb, method='cg')) <= SOL_TOL


def test_topology_optimisation_loop_matches_oracle():
    """BASELINE.json configs[4] end to end in miniature: SIMP cantilever, compliance objective through ad_wrapper (forward
    solve + implicit adjoint + per-element gradient on the GPU), volume constraint, sensitivity filter and three MMA
    iterations (jax_fem_b200/mma.py::optimize) against the oracle's NumPy chain (oracle/fem.py + oracle/mma.py)."""
    import jax_fem_b200 as jf
    import gpu_problems as gp
    from jax_fem_b200 import mma
    from oracle import mma as omma
    m = jf.box_mesh(8, 2, 4, 2.0, 0.5, 1.0)
    pts, cells = m.points, m.cells_dict['hexahedron']
    left = lambda p: np.isclose(p[0], 0., atol=1e-5)
    load = lambda p: np.isclose(p[0], 2.0, atol=1e-5)
    bc = [[left] * 3, [0, 1, 2], [lambda p: 0.] * 3]
    prob = gp.SIMPElasticity(jf.Mesh(pts, cells), vec=3, dim=3, dirichlet_bc_info=bc, location_fns=[load])
    fwd = jf.ad_wrapper(prob)
    n, vf = len(cells), 0.5
    f_ext = prob._f_ext
    history = []

    def objective(rho):
        params = rho[:, 0].clone().requires_grad_(True)
        sol = fwd(params)[0]
        J = -(f_ext * sol).sum()
        J.backward()
        history.append(float(J.detach()))
        return J.detach(), params.grad.reshape(-1, 1)

    def constraint(rho, it):
        return torch.stack([rho.mean() / vf - 1.0]), torch.full((1, n, 1), 1.0 / (n * vf), dtype=torch.float64, device='cuda')

    rho0 = torch.full((n, 1), vf, dtype=torch.float64, device='cuda')
    rho = mma.optimize(prob.fes[0], rho0, {'movelimit': 0.1, 'maxIters': 3}, objective, constraint, 1)

    otr = lambda u, x: -np.array([0., 0., -100.]) + 0. * u
    opb = fem.Problem(fem.Mesh(pts, cells), 3, 3, dirichlet_bc_info=bc, location_fns=[load],
                      law=olaws.SIMP(70e3, 70.0, 0.3, 3.0), surface_maps=[otr], internal_vars=[np.full((n, 8), vf)])
    ohist = []

    def oobjective(r):
        opb.internal_vars = [np.repeat(r, 8, axis=1)]
        osol = fem.solver(opb)
        of = np.zeros_like(osol)
        np.add.at(of, opb.cells[opb.boundary_inds_list[0][:, 0]].reshape(-1), opb.face_residuals(osol, 0).reshape(-1, 3))
        ohist.append(float(-(of * osol).sum()))
        return ohist[-1], fem.implicit_vjp(opb, osol, -of).sum(axis=1)[:, None]

    ocon = lambda r, it: (np.array([r.mean() / vf - 1.0]), np.full((1, n, 1), 1.0 / (n * vf)))
    fe = fem.FiniteElement(fem.Mesh(pts, cells), 3, 3, 'HEX8')
    oH = omma.kd_filter(pts, cells, fe.get_shape_grads()[1], 3)
    orho = omma.optimize(oH, np.full((n, 1), vf), {'movelimit': 0.1, 'maxIters': 3}, oobjective, ocon, 1)
    assert len(history) == 3 and relmax(history, ohist) <= SOL_TOL
    assert history[2] < history[0]                                   # compliance goes down
    assert np.abs(host(rho) - orho).max() <= 1e-6                    # the sub-problem is solved to 1e-7 on both sides


def test_newton_line_search_matches_oracle():
    """solver_options['line_search_flag'] (step halving, jax_fem/solver.py:416-462): same Newton path as the oracle --
    iteration count and converged state -- on a Neo-Hookean block stretched by 60 % in one load step."""
    import jax_fem_b200 as jf
    import gpu_problems as gp
    m = jf.box_mesh(4, 4, 4, 1., 1., 1.)
    pts, cells = m.points, m.cells_dict['hexahedron']
    left = lambda p: np.isclose(p[0], 0., atol=1e-5)
    right = lambda p: np.isclose(p[0], 1., atol=1e-5)
    bc = [[left] * 3 + [right] * 3, [0, 1, 2] * 2, [lambda p: 0.] * 3 + [lambda p: 0.6] + [lambda p: 0.3] * 2]
    prob = gp.HyperElasticity(jf.Mesh(pts, cells), vec=3, dim=3, dirichlet_bc_info=bc)
    sol = jf.solver(prob, {'line_search_flag': True, 'jax_solver': {}})[0]
    opb = fem.Problem(fem.Mesh(pts, cells), 3, 3, dirichlet_bc_info=bc, law=olaws.NeoHookean(1e3, 0.3))
    log = []
    osol = fem.solver(opb, log=log, line_search_flag=True)
    assert prob.last_newton_info['iterations'] == len(log) - 1 > 5
    assert relmax(host(sol), osol) <= SOL_TOL
