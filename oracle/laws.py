"""CPU ORACLE (test infrastructure only) -- the registered constitutive laws.

Each law restates a ``get_tensor_map`` of the reference and supplies the
closed-form tangent moduli A[i,J,k,L] = d sigma_iJ / d (grad u)_kL that the
reference obtains by forward-mode AD (jax_fem/problem.py:262-266).

  Poisson / heat : tests/benchmarks/linear_poisson/test_linear_poisson.py:15-17,
                   applications/thermal_mechanical/example.py:35-38
  LinearElastic  : tests/benchmarks/linear_elasticity_cube/test_linear_elasticity_cube.py:17-25
  NeoHookean     : tests/benchmarks/hyperelasticity/test_hyper_elasticity.py:16-34,
                   applications/scalability/hyperelastic3d_common.py:18-35 (classic),
                   :52-70 (per-quad rho scaling of E, J = max(J, 1e-14))
  SIMP           : docs/source/learn/topology_optimization/example.ipynb cell 9 (E(theta)),
                   applications/outdated/top_opt/fem_model.py:68-80 (3-D isotropic stress)

All arrays carry leading batch axes (..., vec, dim); internal variables (...,).
"""
import numpy as np


def _eye(d):
    return np.eye(d)


class Poisson:
    """sigma = k * grad u (k = 1: identity tensor map)."""

    def __init__(self, k=1.0):
        self.k = float(k)

    def stress(self, ug, kq=None):
        k = self.k if kq is None else self.k * kq[..., None, None]
        return k * ug

    def tangent(self, ug, kq=None):
        v, d = ug.shape[-2:]
        A = np.einsum('ik,jl->ijkl', _eye(v), _eye(d))
        A = np.broadcast_to(A, ug.shape[:-2] + A.shape)
        k = self.k if kq is None else self.k * kq[..., None, None, None, None]
        return k * A

    def dstress_dparam(self, ug, kq):
        return self.k * ug


def _iso(ug, lam, mu):
    d = ug.shape[-1]
    eps = 0.5 * (ug + np.swapaxes(ug, -1, -2))
    tr = np.trace(eps, axis1=-2, axis2=-1)
    return lam * tr[..., None, None] * _eye(d) + 2.0 * mu * eps


def _iso_tangent(d):
    I = _eye(d)
    t_lam = np.einsum('ij,kl->ijkl', I, I)
    t_mu = np.einsum('ik,jl->ijkl', I, I) + np.einsum('il,jk->ijkl', I, I)
    return t_lam, t_mu


class LinearElastic:
    """plane_stress=True (2-D only) is the law of docs/source/learn/topology_optimization/example.ipynb cell 9:
    sig11 = E/((1+nu)(1-nu)) (eps11 + nu eps22), sig12 = E/(1+nu) eps12, i.e. the isotropic form with
    lambda* = E nu / ((1+nu)(1-nu))."""

    def __init__(self, E, nu, plane_stress=False):
        self.E, self.nu, self.plane_stress = float(E), float(nu), bool(plane_stress)

    def lame(self, E):
        nu = self.nu
        lam = E * nu / ((1 + nu) * (1 - nu)) if self.plane_stress else E * nu / ((1 + nu) * (1 - 2 * nu))
        return lam, E / (2. * (1. + nu))

    def stress(self, ug):
        lam, mu = self.lame(self.E)
        return _iso(ug, lam, mu)

    def tangent(self, ug):
        lam, mu = self.lame(self.E)
        t_lam, t_mu = _iso_tangent(ug.shape[-1])
        A = lam * t_lam + mu * t_mu
        return np.broadcast_to(A, ug.shape[:-2] + A.shape)


class SIMP(LinearElastic):
    """E(theta) = Emin + (Emax - Emin) theta^p, then isotropic linear elasticity."""

    def __init__(self, Emax, Emin, nu, p=3.0, plane_stress=False):
        self.Emax, self.Emin, self.nu, self.p = float(Emax), float(Emin), float(nu), float(p)
        self.plane_stress = bool(plane_stress)

    def E_of(self, theta):
        return self.Emin + (self.Emax - self.Emin) * theta ** self.p

    def stress(self, ug, theta):
        lam, mu = self.lame(self.E_of(theta))
        return _iso(ug, lam[..., None, None], mu[..., None, None])

    def tangent(self, ug, theta):
        lam, mu = self.lame(self.E_of(theta))
        t_lam, t_mu = _iso_tangent(ug.shape[-1])
        return lam[..., None, None, None, None] * t_lam + mu[..., None, None, None, None] * t_mu

    def dstress_dparam(self, ug, theta):
        dE = (self.Emax - self.Emin) * self.p * theta ** (self.p - 1.0)
        lam, mu = self.lame(dE)
        return _iso(ug, lam[..., None, None], mu[..., None, None])


class NeoHookean:
    """psi(F) = mu/2 (J^{-2/3} I1 - 3) + kappa/2 (J - 1)^2, F = grad u + I, P = d psi / d F.

    With the optional per-quad scaling rho of E (hyperelastic3d_common.py:52-70) the reference also
    clamps J = max(J, 1e-14); the clamp is inactive for any admissible deformation and is mirrored
    only in the value of J used by the formulas.
    """

    def __init__(self, E, nu, clamp_J=False):
        self.E, self.nu, self.clamp_J = float(E), float(nu), clamp_J

    def _consts(self, rho):
        E = self.E if rho is None else self.E * rho
        return E / (2. * (1. + self.nu)), E / (3. * (1. - 2. * self.nu))

    def _kin(self, ug):
        F = ug + _eye(3)
        J = np.linalg.det(F)
        if self.clamp_J:
            J = np.maximum(J, 1e-14)
        Finv = np.linalg.inv(F)
        FinvT = np.swapaxes(Finv, -1, -2)
        I1 = np.einsum('...ij,...ij->...', F, F)
        return F, J, Finv, FinvT, I1

    def stress(self, ug, rho=None):
        mu, kappa = self._consts(rho)
        F, J, _, H, I1 = self._kin(ug)
        c = lambda a: np.asarray(a)[..., None, None]
        return c(mu * J ** (-2. / 3.)) * (F - c(I1 / 3.) * H) + c(kappa * (J - 1.) * J) * H

    def tangent(self, ug, rho=None):
        mu, kappa = self._consts(rho)
        F, J, Finv, H, I1 = self._kin(ug)
        c = lambda a: np.asarray(a)[..., None, None, None, None]
        I = _eye(3)
        dd = np.einsum('ik,jl->ijkl', I, I)
        FH = np.einsum('...ij,...kl->...ijkl', F, H)
        HF = np.einsum('...ij,...kl->...ijkl', H, F)
        HH = np.einsum('...ij,...kl->...ijkl', H, H)
        X = np.einsum('...li,...jk->...ijkl', Finv, Finv)          # F^{-1}_{Li} F^{-1}_{Jk}
        m = mu * J ** (-2. / 3.)
        return (c(m) * (dd - (2. / 3.) * FH - (2. / 3.) * HF + c(2. / 9. * I1) * HH + c(I1 / 3.) * X)
                + c(kappa * (2. * J - 1.) * J) * HH - c(kappa * (J - 1.) * J) * X)

    def dstress_dparam(self, ug, rho):
        return self.stress(ug, np.ones_like(rho))     # P is linear in E


def fd_tangent(law, ug, *iv, h=1e-6):
    """Central finite difference of stress wrt grad u (test helper)."""
    v, d = ug.shape[-2:]
    A = np.zeros(ug.shape[:-2] + (v, d, v, d))
    for k in range(v):
        for l in range(d):
            e = np.zeros_like(ug)
            e[..., k, l] = h
            A[..., :, :, k, l] = (law.stress(ug + e, *iv) - law.stress(ug - e, *iv)) / (2 * h)
    return A
