"""Analytic anchors for what no reference test pins: tangent moduli vs finite
differences, rigid-body null space, partition of unity, CSR pattern formula, PETSc-style COO->CSR."""
import numpy as np
import pytest
from oracle import basis, fem, laws


@pytest.mark.parametrize("law,iv", [
    (laws.Poisson(2.5), ()),
    (laws.LinearElastic(70e3, 0.3), ()),
    (laws.NeoHookean(10.0, 0.3), ()),
    (laws.NeoHookean(10.0, 0.3, clamp_J=True), (np.array([0.7, 1.3]),)),
    (laws.SIMP(70e3, 70.0, 0.3), (np.array([0.4, 0.9]),)),
])
def test_tangent_matches_finite_difference(law, iv):
    rng = np.random.default_rng(0)
    ug = 0.1 * rng.standard_normal((2, 3, 3))
    A = law.tangent(ug, *iv)
    A_fd = laws.fd_tangent(law, ug, *iv, h=1e-6)
    assert np.abs(A - A_fd).max() <= 1e-7 * max(1.0, np.abs(A).max())


def test_dstress_dparam_fd():
    rng = np.random.default_rng(1)
    ug = 0.05 * rng.standard_normal((4, 3, 3))
    th = np.array([0.3, 0.5, 0.7, 0.9])
    for law in (laws.SIMP(70e3, 70.0, 0.3), laws.NeoHookean(10.0, 0.3, True)):
        fd = (law.stress(ug, th + 1e-6) - law.stress(ug, th - 1e-6)) / 2e-6
        assert np.abs(law.dstress_dparam(ug, th) - fd).max() <= 1e-6 * np.abs(fd).max()


@pytest.mark.parametrize("ele", ["HEX8", "QUAD4", "HEX27"])
def test_partition_of_unity_and_weights(ele):
    vals, grads, w = basis.get_shape_vals_and_grads(ele)
    assert np.allclose(vals.sum(1), 1.0, atol=1e-13)
    assert np.allclose(grads.sum(1), 0.0, atol=1e-12)
    assert abs(w.sum() - 1.0) < 1e-14
    fv, fg, fw, fn, fi = basis.get_face_shape_vals_and_grads(ele)
    assert np.allclose(fv.sum(2), 1.0, atol=1e-13) and np.allclose(fw.sum(1), 1.0)
    assert {"HEX8": (8, 8), "QUAD4": (4, 4), "HEX27": (216, 27)}[ele] == vals.shape


def test_hex27_nodes_are_vtk_triquadratic():
    """After re_order the nodal points must be the VTK_TRIQUADRATIC_HEXAHEDRON lattice."""
    _, _, degree, re = basis.get_elements("HEX27")
    nodes = basis._lagrange_nodes(3, 2)[re]
    vals, _ = basis.tabulate(3, 2, nodes)
    assert np.allclose(vals[:, re], np.eye(27), atol=1e-13)          # nodal basis
    vtk_corners = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    assert np.allclose(nodes[:8], vtk_corners)
    edges = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
    for k, (a, b) in enumerate(edges):
        assert np.allclose(nodes[8 + k], 0.5 * (vtk_corners[a] + vtk_corners[b]))
    faces = [(0, 4, 7, 3), (1, 2, 6, 5), (0, 1, 5, 4), (3, 2, 6, 7), (0, 1, 2, 3), (4, 5, 6, 7)]
    for k, f in enumerate(faces):
        assert np.allclose(nodes[20 + k], vtk_corners[list(f)].mean(0))
    assert np.allclose(nodes[26], 0.5)


def test_rigid_body_null_space_and_symmetry():
    mesh = fem.box_mesh(3, 2, 2, 1.0, 0.7, 0.9)
    rng = np.random.default_rng(0)
    mesh.points = mesh.points + 0.03 * rng.standard_normal(mesh.points.shape)      # non-affine cells
    pb = fem.Problem(mesh, 3, 3, law=laws.LinearElastic(70e3, 0.3))
    pb.newton_update(np.zeros((len(mesh.points), 3)))
    A = fem.get_A(pb)
    assert abs(A - A.T).max() < 1e-9 * abs(A).max()
    x = mesh.points
    modes = [np.tile(e, (len(x), 1)) for e in np.eye(3)] + \
            [np.cross(e, x) for e in np.eye(3)]
    for m in modes:
        assert np.abs(A @ m.reshape(-1)).max() < 1e-9 * abs(A).max()


def test_pattern_formula_and_coo_semantics():
    N = 5
    mesh = fem.box_mesh(N, N, N, 1, 1, 1)
    pb = fem.Problem(mesh, 3, 3, dirichlet_bc_info=[[lambda p: np.isclose(p[0], 0.)], [1], [lambda p: 0.5]],
                     law=laws.LinearElastic(1.0, 0.3))
    pb.newton_update(np.zeros((len(mesh.points), 3)))
    A = fem.get_A(pb)
    assert A.nnz == 9 * (3 * (N + 1) - 2) ** 3
    indptr, indices = fem.csr_pattern_from_cells(pb.cells, 3, pb.num_total_dofs_all_vars)
    assert np.array_equal(indptr, A.indptr) and np.array_equal(indices, A.indices)
    rows = pb.bc_rows()[0]
    for r in rows[:7]:
        row = A.getrow(r)
        assert row.nnz == indptr[r + 1] - indptr[r]                   # pattern kept
        assert row[0, r] == 1.0 and abs(row).sum() == 1.0             # zeroRows: unit diagonal


def test_krylov_restatements_agree_with_direct_solve():
    import scipy.sparse.linalg as spla
    mesh = fem.box_mesh(4, 4, 4, 1, 1, 1)
    bc = [[lambda p: np.isclose(p[0], 0.)] * 3, [0, 1, 2], [lambda p: 0.] * 3]
    pb = fem.Problem(mesh, 3, 3, dirichlet_bc_info=bc, law=laws.LinearElastic(70e3, 0.3))
    sol = np.zeros((len(mesh.points), 3))
    res = fem.apply_bc_vec(pb.newton_update(sol).reshape(-1), sol.reshape(-1), pb)
    A = fem.get_A(pb)
    b = np.random.default_rng(0).standard_normal(A.shape[0])
    b[np.concatenate(pb.bc_rows())] = 0.0
    x_ref = spla.spsolve(A.tocsc(), b)
    for method in ("bicgstab", "cg"):
        x, k = fem.jax_solve(A, b, np.zeros_like(b), True, method, return_iters=True)
        assert 0 < k < 500
        assert np.abs(x - x_ref).max() <= 1e-8 * np.abs(x_ref).max()


def test_plane_stress_simp_is_the_topology_optimisation_notebook_law():
    """docs/source/learn/topology_optimization/example.ipynb cell 9 (restated): plane-stress Hooke's law with the SIMP
    modulus.  The oracle expresses it as the isotropic form with lambda* = E nu / ((1+nu)(1-nu))."""
    from oracle import laws
    rng = np.random.default_rng(0)
    ug = rng.standard_normal((5, 4, 2, 2)) * 0.01
    theta = rng.uniform(0.1, 1.0, (5, 4))
    Emax, nu, penal = 70e3, 0.3, 3.0
    Emin = 1e-3 * Emax
    law = laws.SIMP(Emax, Emin, nu, penal, plane_stress=True)
    E = Emin + (Emax - Emin) * theta ** penal
    eps = 0.5 * (ug + np.swapaxes(ug, -1, -2))
    s11 = E / (1 + nu) / (1 - nu) * (eps[..., 0, 0] + nu * eps[..., 1, 1])
    s22 = E / (1 + nu) / (1 - nu) * (nu * eps[..., 0, 0] + eps[..., 1, 1])
    s12 = E / (1 + nu) * eps[..., 0, 1]
    ref = np.stack([np.stack([s11, s12], -1), np.stack([s12, s22], -1)], -2)
    assert np.abs(law.stress(ug, theta) - ref).max() <= 1e-13 * np.abs(ref).max()
    # tangent and parameter derivative are consistent with the stress (finite differences)
    assert np.abs(law.tangent(ug, theta) - laws.fd_tangent(law, ug, theta)).max() <= 1e-5 * np.abs(law.tangent(ug, theta)).max()
    h = 1e-6
    fd = (law.stress(ug, theta + h) - law.stress(ug, theta - h)) / (2 * h)
    assert np.abs(law.dstress_dparam(ug, theta) - fd).max() <= 1e-6 * np.abs(fd).max()
    assert np.abs(laws.SIMP(Emax, Emin, nu, penal).stress(ug, theta) - ref).max() > 1e-3 * np.abs(ref).max()   # plane strain differs


def test_newton_with_step_halving_line_search_reaches_the_same_equilibrium():
    """solver.py:424-462 restated in the oracle: with a 60 % stretch imposed in one step the halving rule is active (more,
    shorter Newton steps), and the converged state is the one plain Newton finds."""
    from oracle import fem, laws
    m = fem.box_mesh(4, 4, 4, 1., 1., 1.)
    left = lambda p: np.isclose(p[0], 0., atol=1e-5)
    right = lambda p: np.isclose(p[0], 1., atol=1e-5)
    bc = [[left] * 3 + [right] * 3, [0, 1, 2] * 2, [lambda p: 0.] * 3 + [lambda p: 0.6] + [lambda p: 0.3] * 2]
    sols, its = [], []
    for flag in (False, True):
        pb = fem.Problem(m, 3, 3, dirichlet_bc_info=bc, law=laws.NeoHookean(1e3, 0.3))
        log = []
        sols.append(fem.solver(pb, log=log, line_search_flag=flag))
        its.append(len(log) - 1)
    assert its[1] > its[0] and np.abs(sols[0] - sols[1]).max() <= 1e-8 * np.abs(sols[0]).max()


def test_oracle_face_tangent_is_the_derivative_of_the_face_residual():
    """u-dependent surface maps (Robin terms, problem.py:238-259 + 289-325): the oracle's analytic face tangent must be the
    central difference of its own residual, for the reference's nonlinear Robin map 5 u^2 (applications/robin_bc) and for a
    vector spring foundation."""
    import numpy as np
    from oracle import fem, laws
    m = fem.box_mesh(3, 2, 2, 1.5, 1.0, 1.0)
    rng = np.random.default_rng(0)
    pts = m.points + 0.03 * rng.uniform(-1, 1, m.points.shape)
    right = lambda p: np.isclose(p[0], 1.5, atol=0.05)
    top = lambda p: np.isclose(p[2], 1.0, atol=0.05)
    for vec, law, maps, jacs in (
            (1, laws.Poisson(1.0), [lambda u, x: 5 * u ** 2, lambda u, x: 2.0 * (u - 0.3)],
             [lambda u, x: (10 * u)[..., None], lambda u, x: 2.0 * np.ones(u.shape + (1,))]),
            (3, laws.LinearElastic(70e3, 0.3), [lambda u, x: np.array([3., 5., 7.]) * u, lambda u, x: 0. * u + np.array([0., 0., 1.])],
             [lambda u, x: np.broadcast_to(np.diag([3., 5., 7.]), u.shape + (3,)), None])):
        pb = fem.Problem(fem.Mesh(pts, m.cells), vec, 3, location_fns=[right, top], law=law, surface_maps=maps, surface_map_jacs=jacs)
        sol = 0.2 + 0.1 * rng.standard_normal((len(pts), vec))
        pb.newton_update(sol)
        A = fem.get_A(pb).toarray()
        n = A.shape[0]
        h = 1e-6
        for j in rng.integers(0, n, 12):
            e = np.zeros(n)
            e[j] = h
            col = (pb.compute_residual(sol + e.reshape(sol.shape)) - pb.compute_residual(sol - e.reshape(sol.shape))).reshape(-1) / (2 * h)
            assert np.abs(col - A[:, j]).max() <= 1e-6 * max(np.abs(A[:, j]).max(), 1.0)


def test_oracle_mass_map_jacobian_matches_finite_differences():
    """oracle.fem.Problem.cell_jacobians with a u-dependent mass map (the checker of csrc/mass.cu) against central
    differences of cell_residuals."""
    import jax_fem_b200 as jf
    from oracle import fem, laws as olaws
    rng = np.random.default_rng(3)
    m = jf.box_mesh(2, 2, 1, 1., 1., 0.5)
    pts = m.points + 0.03 * rng.uniform(-1, 1, m.points.shape)
    cells = m.cells_dict['hexahedron']
    a = rng.uniform(1, 3, (len(cells), 8))
    b = rng.standard_normal((len(cells), 8, 1))
    pb = fem.Problem(fem.Mesh(pts, cells), 1, 3, law=olaws.Poisson(2.0), mass_map=lambda u, x: a[..., None] * u + b,
                     mass_map_jac=lambda u, x: a[..., None, None] * np.ones((1, 1)))
    sol = rng.standard_normal((len(pts), 1))
    K = pb.cell_jacobians(sol)
    h = 1e-6
    for c in (0, 3):
        for j in range(8):
            e = np.zeros_like(sol)
            e[cells[c, j]] = h
            col = (pb.cell_residuals(sol + e)[c] - pb.cell_residuals(sol - e)[c]).reshape(-1) / (2 * h)
            assert np.abs(col - K[c][:, j]).max() <= 1e-7 * np.abs(K[c]).max()
