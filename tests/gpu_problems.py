"""Problem subclasses for the GPU parity tests: the reference's own test problems
(tests/benchmarks/*/test_*.py) written against jax_fem_b200 with registered laws."""
import numpy as np
from jax_fem_b200 import Problem, laws


class LinearPoisson(Problem):                 # test_linear_poisson.py:15-17
    def get_tensor_map(self):
        return laws.Poisson(1.0)


class LinearElasticityCube(Problem):          # test_linear_elasticity_cube.py:15-36
    def get_tensor_map(self):
        return laws.LinearElasticity(70e3, 0.3)

    def get_mass_map(self):
        return lambda u, x: -np.array([0., 10., 10.])

    def get_surface_maps(self):
        return [lambda u, x: -np.array([10., 0., 0.])]


class LinearElasticityCylinder(Problem):      # test_linear_elasticity_cylinder.py:15-37
    def get_tensor_map(self):
        return laws.LinearElasticity(70e3, 0.3)

    def get_mass_map(self):
        return lambda u, x: -np.array([1e3 * x[0], 2e3 * x[1], 3e3 * x[2]])

    def get_surface_maps(self):
        return [lambda u, x: -np.array([1e3 * x[0] ** 2 + 1e3 * x[1] ** 2, 0. * x[0], 0. * x[0]])]


class HyperElasticity(Problem):               # test_hyper_elasticity.py:15-34
    def get_tensor_map(self):
        return laws.NeoHookean(1e3, 0.3)


class PlainElasticity(Problem):
    def custom_init(self, E=70e3, nu=0.3):
        self.E, self.nu = E, nu

    def get_tensor_map(self):
        return laws.LinearElasticity(self.E, self.nu)


class SIMPElasticity(Problem):                # topology_optimization example.ipynb cell 9 (3-D isotropic form)
    def custom_init(self, traction=(0., 0., -100.)):
        self.traction = np.array(traction)

    def get_tensor_map(self):
        return laws.SIMP(70e3, 70.0, 0.3, 3.0)

    def get_surface_maps(self):
        t = self.traction
        return [lambda u, x: -t]

    def set_params(self, params):
        # params: (num_cells,) densities -> theta at every quadrature point (np.repeat in the reference)
        self.internal_vars = [params[:, None].expand(-1, self.fes[0].num_quads)]


class NeoHookeanInverse(Problem):             # hyperelastic3d_common.py:45-80
    def get_tensor_map(self):
        return laws.NeoHookean(10.0, 0.3, clamp_J=True)

    def get_surface_maps(self):
        return [lambda u, x: np.array([0., 1e-3, 0.])]

    def set_params(self, rho):
        self.internal_vars = [rho]


class RobinPoisson(Problem):                   # applications/robin_bc/example.py:59-67: surface_map(u, x) = 5 u^2 on two faces
    def get_tensor_map(self):
        return laws.Poisson(1.0)

    def get_mass_map(self):
        return lambda u, x: -np.array([10. * np.exp(-((x[0] - .5) ** 2 + (x[1] - .5) ** 2) / 0.02)])

    def get_surface_maps(self):
        return [laws.RobinPower(5.0, power=2.0), laws.RobinPower(2.0, power=1.0, u_ref=0.3)]


class SpringFoundation(Problem):               # elastic foundation k u on one face + a dead load on another
    def get_tensor_map(self):
        return laws.LinearElasticity(70e3, 0.3)

    def get_surface_maps(self):
        return [laws.RobinPower([3e3, 5e3, 7e3]), lambda u, x: np.array([0., 0., 100.])]
