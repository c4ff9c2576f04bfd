"""Registry of constitutive laws that have a hand-written sm_100a element kernel.

In the reference ``Problem.get_tensor_map()`` returns an arbitrary Python function that JAX traces
and differentiates (jax_fem/problem.py:189-214, 262-266).  A CUDA kernel cannot be generated from
an arbitrary callable, so on this path ``get_tensor_map()`` must return one of the objects below;
anything else raises ``UnregisteredLawError`` -- it never falls back to another implementation.

Each law cites the reference ``get_tensor_map`` body it reproduces (formulas in SURVEY.md 8a).
"""
from . import _lib


class UnregisteredLawError(NotImplementedError):
    pass


class Law:
    law_id = -1
    n_internal_vars = 0       # per-quadrature-point parameter arrays the kernel accepts
    requires_internal_var = False

    def params(self):
        raise NotImplementedError

    def __call__(self, *a, **k):
        raise UnregisteredLawError("registered laws are evaluated by CUDA kernels, not called from Python")


class Poisson(Law):
    """f(grad u) = k grad u.  tests/benchmarks/linear_poisson/test_linear_poisson.py:15-17 (k = 1),
    applications/thermal_mechanical/example.py:35-38 (heat conduction).  An optional per-quad
    internal variable scales k."""
    law_id = 0
    n_internal_vars = 1

    def __init__(self, k=1.0):
        self.k = float(k)

    def params(self):
        return [self.k]


Heat = Poisson


class LinearElasticity(Law):
    """sigma = lambda tr(eps) I + 2 mu eps.
    tests/benchmarks/linear_elasticity_cube/test_linear_elasticity_cube.py:17-25."""
    law_id = 1

    def __init__(self, E, nu, plane_stress=False):
        """plane_stress (2-D elements only): lambda* = E nu / ((1+nu)(1-nu)) instead of the plane-strain / 3-D lambda."""
        self.E, self.nu, self.plane_stress = float(E), float(nu), bool(plane_stress)

    def params(self):
        return [self.E, self.nu, 1.0 if self.plane_stress else 0.0]


class NeoHookean(Law):
    """psi = mu/2 (J^-2/3 I1 - 3) + kappa/2 (J-1)^2, P = d psi/dF.
    tests/benchmarks/hyperelasticity/test_hyper_elasticity.py:16-34;
    applications/scalability/hyperelastic3d_common.py:18-35; with ``clamp_J`` and a per-quad
    modulus scale rho: hyperelastic3d_common.py:52-70."""
    law_id = 2
    n_internal_vars = 1

    def __init__(self, E, nu, clamp_J=False):
        self.E, self.nu, self.clamp_J = float(E), float(nu), bool(clamp_J)

    def params(self):
        return [self.E, self.nu, 1.0 if self.clamp_J else 0.0]


class SIMP(Law):
    """E(theta) = Emin + (Emax - Emin) theta^penal, then isotropic linear elasticity.
    docs/source/learn/topology_optimization/example.ipynb cell 9;
    applications/outdated/top_opt/fem_model.py:68-80 (3-D)."""
    law_id = 3
    n_internal_vars = 1
    requires_internal_var = True

    def __init__(self, Emax, Emin, nu, penal=3.0, plane_stress=None):
        """plane_stress: None = the reference's own choice for the element (its only 2-D SIMP law, the topology
        optimisation notebook, is plane stress; its 3-D one is the full isotropic law): True on QUAD4, False on 3-D."""
        self.Emax, self.Emin, self.nu, self.penal = float(Emax), float(Emin), float(nu), float(penal)
        self.plane_stress = plane_stress

    def params(self):
        return [self.Emax, self.Emin, self.nu, self.penal, 1.0 if self.plane_stress else 0.0]


class SurfaceLaw:
    """A solution-dependent surface map with hand-written face kernels (csrc/faces.cu).  The reference accepts any callable
    ``surface_map(u, x)`` and differentiates it (problem.py:238-259, 289-325); here a u-dependent map must be one of the
    registered objects below -- plain callables stay supported as long as they do not depend on u."""

    def law_host(self, vec):
        raise NotImplementedError


class RobinPower(SurfaceLaw):
    """val_i(u) = coef_i (u_i - u_ref_i)^power on a boundary set: convective / Robin boundary condition or elastic foundation
    (power 1), the nonlinear Robin map ``5 * u**2`` of applications/robin_bc/example.py:59-67 (coef 5, power 2)."""

    def __init__(self, coef, power=1.0, u_ref=0.0):
        import numpy as np
        self.coef = np.atleast_1d(np.asarray(coef, dtype=np.float64))
        self.u_ref = np.atleast_1d(np.asarray(u_ref, dtype=np.float64))
        self.power = float(power)

    def law_host(self, vec):
        import numpy as np
        coef = np.broadcast_to(self.coef, (vec,)) if self.coef.size in (1, vec) else None
        uref = np.broadcast_to(self.u_ref, (vec,)) if self.u_ref.size in (1, vec) else None
        if coef is None or uref is None:
            raise ValueError(f"RobinPower: coef / u_ref must be scalars or have {vec} components")
        out = np.zeros(7)
        out[:vec], out[3:3 + vec], out[6] = coef, uref, self.power
        return out

    def value(self, u):
        """NumPy evaluation (host-side post-processing only)."""
        return self.coef * (u - self.u_ref) ** self.power


class MassLaw:
    """A solution-dependent mass map with a hand-written kernel (csrc/mass.cu).  The reference accepts any callable
    ``mass_map(u, x, *internal_vars)`` and differentiates it (problem.py:216-236, 262-266); here a u-dependent map must be one
    of the registered objects below -- plain callables stay supported as long as they do not depend on u."""


class LinearMass(MassLaw):
    """m_i(u) = coef * u_i + const_i at every quadrature point.  ``coef``: a number or a (num_cells, num_quads) field;
    ``const``: None, a number / vec numbers, or a (num_cells, num_quads[, vec]) field.  Both are read at every assembly, so a
    time stepper updates them in place between solves: backward-Euler heat capacity ``rho Cp (T - T_old) / dt`` is
    ``coef = rho Cp / dt, const = -coef * T_old(q)`` (applications/thermal_mechanical/example.py:40-43), the phase-field
    driving term ``(G_c / l + 2 H) d - 2 H`` is ``coef = G_c / l + 2 H(q), const = -2 H(q)``
    (applications/phase_field_fracture/example.py:54-57)."""

    def __init__(self, coef, const=None):
        self.coef = coef
        self.const = const

    def fields(self, num_cells, num_quads, vec, device):
        """-> (coef scalar, coef field or None, const host vector (3,), const field (C, Q, vec) or None)."""
        import numpy as np
        import torch

        def as_field(v, shape):
            t = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v, dtype=np.float64))
            t = t.detach().to(device=device, dtype=torch.float64)
            if t.ndim == 2 and len(shape) == 3:                  # one field for every component
                t = t.unsqueeze(-1).expand(*t.shape, vec)
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"LinearMass: field of shape {tuple(t.shape)}, expected {tuple(shape)}")
            return t.contiguous()

        ndim = lambda v: v.ndim if hasattr(v, 'ndim') else np.ndim(v)
        size = lambda v: v.numel() if isinstance(v, torch.Tensor) else np.size(v)
        coef, coef_f = 0.0, None
        if ndim(self.coef) == 0:
            coef = float(self.coef)
        else:
            coef_f = as_field(self.coef, (num_cells, num_quads))
        cst, cst_f = np.zeros(3), None
        if self.const is None:
            pass
        elif ndim(self.const) <= 1 and size(self.const) in (1, vec):
            host = self.const.detach().cpu().numpy() if isinstance(self.const, torch.Tensor) else self.const
            cst[:vec] = np.broadcast_to(np.asarray(host, dtype=np.float64).reshape(-1), (vec,))
        else:
            cst_f = as_field(self.const, (num_cells, num_quads, vec))
        return coef, coef_f, cst, cst_f


REGISTERED = {
    ('HEX8', 1, Poisson), ('HEX8', 3, LinearElasticity), ('HEX8', 3, NeoHookean), ('HEX8', 3, SIMP),
    ('QUAD4', 1, Poisson), ('QUAD4', 2, LinearElasticity), ('QUAD4', 2, SIMP),
    ('HEX27', 3, LinearElasticity), ('HEX27', 3, SIMP),
}


def resolve(tensor_map, ele_type, vec):
    if not isinstance(tensor_map, Law):
        raise UnregisteredLawError(
            f"get_tensor_map() returned {type(tensor_map).__name__}; the B200 hot path only runs registered "
            f"laws (jax_fem_b200.laws.Poisson/Heat, LinearElasticity, NeoHookean, SIMP). "
            f"Unregistered tensor maps raise; they do not fall back.")
    if (ele_type, vec, type(tensor_map)) not in REGISTERED:
        raise UnregisteredLawError(
            f"no sm_100a kernel is registered for (ele_type={ele_type}, vec={vec}, law={type(tensor_map).__name__})")
    if isinstance(tensor_map, (LinearElasticity, SIMP)):
        two_d = ele_type == 'QUAD4'
        if tensor_map.plane_stress is None:
            tensor_map.plane_stress = two_d
        if tensor_map.plane_stress and not two_d:
            raise UnregisteredLawError("plane_stress is a 2-D assumption; it is not registered for 3-D elements")
    return tensor_map
