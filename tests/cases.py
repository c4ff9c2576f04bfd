"""The reference's own benchmark problems (tests/benchmarks/*/test_*.py), stated once for the
oracle (numpy callables) and reused by the GPU parity tests."""
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def isclose(a, b):
    return np.isclose(a, b, atol=1e-5)


def left(p):
    return isclose(p[0], 0.)


def right(p):
    return isclose(p[0], 1.)


def bottom(p):
    return isclose(p[2], 0.)


def top(p):
    return isclose(p[2], 10.)


def zero(p):
    return 0.


def one(p):
    return 1.


# (dirichlet_bc_info, location_fns) exactly as the reference tests build them
POISSON_BC = [[left, right], [0, 0], [zero, one]]                                  # test_linear_poisson.py:38-53
CUBE_BC = [[left] * 3, [0, 1, 2], [one] * 3]                                       # test_linear_elasticity_cube.py:56-68
HYPER_BC = [[bottom] * 3 + [top] * 3, [0, 1, 2] * 2, [zero] * 5 + [one]]            # test_hyper_elasticity.py:102-112
CYL_BC = [[bottom] * 3, [0, 1, 2], [zero] * 3]                                     # test_linear_elasticity_cylinder.py:106-116


def cube_mass(u, x):            # test_linear_elasticity_cube.py:27-31
    return -np.array([0., 10., 10.]) + 0. * u


def cube_traction(u, x):        # test_linear_elasticity_cube.py:33-36
    return -np.array([10., 0., 0.]) + 0. * u


def cyl_mass(u, x):             # test_linear_elasticity_cylinder.py:27-31
    return -np.stack([1e3 * x[..., 0], 2e3 * x[..., 1], 3e3 * x[..., 2]], -1)


def cyl_traction(u, x):         # test_linear_elasticity_cylinder.py:33-37
    z = 0. * x[..., 0]
    return -np.stack([1e3 * x[..., 0] ** 2 + 1e3 * x[..., 1] ** 2, z, z], -1)
