"""Pin the CPU oracle against every golden the reference's own tests hold for the hot path
(tests/benchmarks/{linear_poisson,linear_elasticity_cube,hyperelasticity,linear_elasticity_cylinder}).
The reference asserts decimal=5 (cylinder: decimal=3, area decimal=4) on float32-rounded output;
the oracle is held to much tighter bounds where the discretisations coincide."""
import numpy as np
from oracle import fem, laws
import cases


def test_linear_poisson_golden():
    g = cases.load_golden("linear_poisson")
    pb = fem.Problem(fem.Mesh(g["points"], g["cells"]), 1, 3, dirichlet_bc_info=cases.POISSON_BC, law=laws.Poisson())
    assert abs(pb.JxW.sum() - 1.0) < 1e-13
    sol = fem.solver(pb)
    np.testing.assert_array_almost_equal(g["sol"], sol[:, 0], decimal=5)      # the reference's assertion
    assert np.abs(sol[:, 0] - g["sol"]).max() < 1e-8


def test_linear_elasticity_cube_golden():
    g = cases.load_golden("linear_elasticity_cube")
    pb = fem.Problem(fem.Mesh(g["points"], g["cells"]), 3, 3, dirichlet_bc_info=cases.CUBE_BC,
                     location_fns=[cases.right], law=laws.LinearElastic(70e3, 0.3),
                     mass_map=cases.cube_mass, surface_maps=[cases.cube_traction])
    assert len(pb.boundary_inds_list[0]) == 100
    assert abs(pb.face_data[0][1].sum() - 1.0) < 1e-13        # area of the x=1 face (facet 3)
    sol = fem.solver(pb)
    np.testing.assert_array_almost_equal(g["sol"], sol, decimal=5)
    assert np.abs(sol - g["sol"]).max() < 1e-8


def test_hyperelasticity_golden():
    g = cases.load_golden("hyperelasticity")
    pb = fem.Problem(fem.Mesh(g["points"], g["cells"]), 3, 3, dirichlet_bc_info=cases.HYPER_BC,
                     law=laws.NeoHookean(1e3, 0.3))
    log = []
    sol = fem.solver(pb, log=log)
    np.testing.assert_array_almost_equal(g["sol"], sol, decimal=5)
    assert np.abs(sol - g["sol"]).max() < 1e-11
    assert len(log) == 5 and log[-1] < 1e-6                   # quadratic Newton convergence: tangent is exact
    # compute_traction, test_hyper_elasticity.py:36-70 (facet 5, z = H)
    b = pb.fe.get_boundary_conditions_inds([cases.top])[0]
    fg, nanson = pb.fe.get_face_shape_grads(b)
    ug = np.einsum('fnv,fqnd->fqvd', sol[pb.cells[b[:, 0]]], fg)
    tz = np.einsum('fqv,fq->v', pb.law.stress(ug)[..., 2], nanson)[2]
    np.testing.assert_almost_equal(float(g["traction"]), tz, decimal=5)


def test_linear_elasticity_cylinder_golden():
    g = cases.load_golden("linear_elasticity_cylinder")
    pb = fem.Problem(fem.Mesh(g["points"], g["cells"]), 3, 3, dirichlet_bc_info=cases.CYL_BC,
                     location_fns=[cases.top], law=laws.LinearElastic(70e3, 0.3),
                     mass_map=cases.cyl_mass, surface_maps=[cases.cyl_traction])
    sol = fem.solver(pb)
    np.testing.assert_array_almost_equal(g["sol"], sol, decimal=3)            # the reference's tolerance
    np.testing.assert_almost_equal(float(g["surface_area"]), pb.face_data[0][1].sum(), decimal=4)
    assert abs(pb.face_data[0][1].sum() - float(g["surface_area"])) < 1e-11
