"""Topology-optimisation driver around the hot path: k-d-tree filters and the MMA optimiser, with the reference's
interface (jax_fem/mma.py: compute_filter_kd_tree :27, applySensitivityFilter :59, applyDensityFilter :64, class MMA :68,
subsolv :207, optimize :415) -- SURVEY.md 8(f) row 1, the caller of ad_wrapper's (J, dJ/dtheta).

In the reference this module is host NumPy code around the jitted FE solve (only the filter matrix lives in a BCOO); here
the design vector, the filter and every O(n) update of the MMA step are torch tensors on the device the densities live
on (the B200 next to the FE solve, or the CPU in the host tests), so the optimisation loop never copies the n design
variables to the host; only the (m+1) x (m+1) Newton systems and the scalars that steer the interior-point iteration
are read back.  The filter matrix is built once on the host (SciPy k-d tree, as in the reference) and applied as a CSR
mat-vec.  The arithmetic is elementwise and memory-trivial next to the FE solve; it uses torch operators, not hand-written
kernels.

Method: K. Svanberg's MMA (IJNME 24, 1987; constants and sub-problem solver of his 2007 note), as the reference
implements it.  Vectors are 1-D internally; (n, 1) / (m, 1) arrays of the reference's interface are accepted and
returned.
"""
import numpy as np
import torch

density_filtering = False        # module switches of the reference (mma.py:24-25)
sensitivity_filtering = True

_EPSIMIN, _RAA0, _ALBEFA, _ASYINIT, _ASYINCR, _ASYDECR = 1e-7, 1e-5, 0.1, 0.5, 1.2, 0.7


def _t(a, like=None, device=None):
    if isinstance(a, torch.Tensor):
        return a.to(dtype=torch.float64)
    dev = like.device if like is not None else device
    return torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev)


# ---- filters -----------------------------------------------------------------------------------------------------------
def compute_filter_kd_tree(fe, device=None, num_nbs=20):
    """H_ij = max(rmin - |c_i - c_j|, 0) over the 20 nearest cell centroids, rmin = 1.5 (mean cell volume)^(1/dim), and
    its row sums (mma.py:27-57).  ``fe.flex_inds`` (cells open to optimisation) defaults to all cells.
    Returns (H as a torch sparse CSR tensor, Hs)."""
    import scipy.spatial
    points, cells = np.asarray(fe.points), np.asarray(fe.cells)
    cent = np.mean(np.take(points, cells, axis=0), axis=1)
    flex = np.asarray(getattr(fe, 'flex_inds', np.arange(len(cells))))
    fc = np.take(cent, flex, axis=0)
    JxW = fe.get_JxW() if hasattr(fe, 'get_JxW') else fe.JxW
    rmin = 1.5 * (float(np.sum(JxW)) / fe.num_cells) ** (1.0 / fe.dim)
    k = min(num_nbs, len(fc))
    dd, ii = scipy.spatial.KDTree(fc).query(fc, k)
    dd, ii = dd.reshape(len(fc), k), ii.reshape(len(fc), k)
    vals = np.where(rmin - dd > 0.0, rmin - dd, 0.0)
    order = np.argsort(ii, axis=1, kind='stable')                       # CSR wants ascending columns inside a row
    ii, vals = np.take_along_axis(ii, order, 1), np.take_along_axis(vals, order, 1)
    crow = torch.arange(0, (len(fc) + 1) * k, k, dtype=torch.int64)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)                    # torch: "sparse CSR support is in beta state"
        H = torch.sparse_csr_tensor(crow, torch.from_numpy(ii.reshape(-1).astype(np.int64)),
                                    torch.from_numpy(vals.reshape(-1)), size=(len(fc), len(fc)), dtype=torch.float64)
    Hs = torch.from_numpy(vals.sum(axis=1))
    if device is not None:
        H, Hs = H.to(device), Hs.to(device)
    return H, Hs


def applySensitivityFilter(ft, rho, dJ, dvc):
    """rho, dJ: (n, 1); dvc: (m, n, 1) -> filtered (dJ, dvc) (mma.py:59-62)."""
    H, Hs = ft['H'], ft['Hs']
    rho, dJ, dvc = _t(rho, Hs), _t(dJ, Hs), _t(dvc, Hs)
    w = rho / torch.clamp(rho, min=1e-3) / Hs[:, None]
    dJ_f = H @ (w * dJ)
    dvc_f = torch.stack([H @ (w * g) for g in dvc]) if dvc.shape[0] else dvc
    return dJ_f, dvc_f


def applyDensityFilter(ft, rho):
    return (ft['H'] @ _t(rho, ft['Hs'])) / ft['Hs'][:, None]


# ---- sub-problem solver --------------------------------------------------------------------------------------------------
def _residual(x, y, z, lam, xsi, eta, mu, zet, s, epsi, low, upp, alfa, beta, p0, q0, P, Q, a0, a, b, c, d):
    ux, xl = upp - x, x - low
    plam, qlam = p0 + P.T @ lam, q0 + Q.T @ lam
    gvec = P @ (1.0 / ux) + Q @ (1.0 / xl)
    return torch.cat([plam / ux ** 2 - qlam / xl ** 2 - xsi + eta,
                      c + d * y - mu - lam,
                      (a0 - zet - a @ lam).reshape(1),
                      gvec - a * z - y + s - b,
                      xsi * (x - alfa) - epsi, eta * (beta - x) - epsi,
                      mu * y - epsi, (zet * z - epsi).reshape(1), lam * s - epsi])


def subsolv(m, n, epsimin, low, upp, alfa, beta, p0, q0, P, Q, a0, a, b, c, d):
    """Primal-dual interior-point Newton method for the MMA sub-problem (mma.py:207-413): barrier parameter 1 -> epsimin in
    decades; Newton steps on the perturbed KKT system reduced to the (m+1) x (m+1) Schur complement in (lambda, z) when
    m < n, to the (n+1) x (n+1) one in (x, z) otherwise; step length = the largest one keeping all positive quantities
    positive (factor 1.01), halved until the residual norm decreases.
    Returns x, y, z, lam, xsi, eta, mu, zet, s."""
    dev, f64 = low.device, torch.float64
    one = lambda k: torch.ones(k, dtype=f64, device=dev)
    x = 0.5 * (alfa + beta)
    y, lam, s = one(m), one(m), one(m)
    z = zet = torch.ones((), dtype=f64, device=dev)
    xsi = torch.clamp(1.0 / (x - alfa), min=1.0)
    eta = torch.clamp(1.0 / (beta - x), min=1.0)
    mu = torch.maximum(one(m), 0.5 * c)
    args = (low, upp, alfa, beta, p0, q0, P, Q, a0, a, b, c, d)
    epsi = 1.0
    while epsi > epsimin:
        res = _residual(x, y, z, lam, xsi, eta, mu, zet, s, epsi, *args)
        resnorm, resmax = float(res.norm()), float(res.abs().max())
        it = 0
        while resmax > 0.9 * epsi and it < 200:
            it += 1
            ux, xl = upp - x, x - low
            plam, qlam = p0 + P.T @ lam, q0 + Q.T @ lam
            gvec = P @ (1.0 / ux) + Q @ (1.0 / xl)
            GG = P / ux ** 2 - Q / xl ** 2
            delx = plam / ux ** 2 - qlam / xl ** 2 - epsi / (x - alfa) + epsi / (beta - x)
            dely = c + d * y - lam - epsi / y
            delz = a0 - a @ lam - epsi / z
            dellam = gvec - a * z - y - b + epsi / lam
            diagx = 2.0 * (plam / ux ** 3 + qlam / xl ** 3) + xsi / (x - alfa) + eta / (beta - x)
            diagy = d + mu / y
            diaglamyi = s / lam + 1.0 / diagy
            if m < n:
                blam = dellam + dely / diagy - GG @ (delx / diagx)
                AA = torch.zeros((m + 1, m + 1), dtype=f64, device=dev)
                AA[:m, :m] = torch.diag(diaglamyi) + (GG / diagx) @ GG.T
                AA[:m, m] = a
                AA[m, :m] = a
                AA[m, m] = -zet / z
                sol = torch.linalg.solve(AA, torch.cat([blam, delz.reshape(1)]))
                dlam, dz = sol[:m], sol[m]
                dx = -delx / diagx - (GG.T @ dlam) / diagx
            else:
                dellamyi = dellam + dely / diagy
                AA = torch.zeros((n + 1, n + 1), dtype=f64, device=dev)
                AA[:n, :n] = torch.diag(diagx) + (GG.T / diaglamyi) @ GG
                axz = -GG.T @ (a / diaglamyi)
                AA[:n, n] = axz
                AA[n, :n] = axz
                AA[n, n] = zet / z + a @ (a / diaglamyi)
                bb = -torch.cat([delx + GG.T @ (dellamyi / diaglamyi), (delz - a @ (dellamyi / diaglamyi)).reshape(1)])
                sol = torch.linalg.solve(AA, bb)
                dx, dz = sol[:n], sol[n]
                dlam = (GG @ dx) / diaglamyi - dz * (a / diaglamyi) + dellamyi / diaglamyi
            dy = -dely / diagy + dlam / diagy
            dxsi = -xsi + epsi / (x - alfa) - xsi * dx / (x - alfa)
            deta = -eta + epsi / (beta - x) + eta * dx / (beta - x)
            dmu = -mu + epsi / y - mu * dy / y
            dzet = -zet + epsi / z - zet * dz / z
            ds = -s + epsi / lam - s * dlam / lam
            cur = torch.cat([y, z.reshape(1), lam, xsi, eta, mu, zet.reshape(1), s])
            step = torch.cat([dy, dz.reshape(1), dlam, dxsi, deta, dmu, dzet.reshape(1), ds])
            stminv = max(float((-1.01 * step / cur).max()), float((-1.01 * dx / (x - alfa)).max()),
                         float((1.01 * dx / (beta - x)).max()), 1.0)
            steg = 1.0 / stminv
            old = (x, y, z, lam, xsi, eta, mu, zet, s)
            newton = (dx, dy, dz, dlam, dxsi, deta, dmu, dzet, ds)
            newnorm, itto = 2.0 * resnorm, 0
            while newnorm > resnorm and itto < 50:
                itto += 1
                x, y, z, lam, xsi, eta, mu, zet, s = (o + steg * dv for o, dv in zip(old, newton))
                res = _residual(x, y, z, lam, xsi, eta, mu, zet, s, epsi, *args)
                newnorm = float(res.norm())
                steg *= 0.5
            resnorm, resmax = newnorm, float(res.abs().max())
        epsi *= 0.1
    return x, y, z, lam, xsi, eta, mu, zet, s


# ---- the optimiser object (same setters / getters as the reference) -------------------------------------------------------
class MMA:
    def __init__(self):
        self.epoch = 0

    def resetMMACounter(self):
        self.epoch = 0

    def registerMMAIter(self, xval, xold1, xold2):
        self.epoch += 1
        self.xval, self.xold1, self.xold2 = xval, xold1, xold2

    def setNumConstraints(self, numConstraints):
        self.numConstraints = numConstraints

    def setNumDesignVariables(self, numDesVar):
        self.numDesignVariables = numDesVar

    def setMinandMaxBoundsForDesignVariables(self, xmin, xmax):
        self.xmin, self.xmax = xmin, xmax

    def setObjectiveWithGradient(self, obj, objGrad):
        self.objective, self.objectiveGradient = obj, objGrad

    def setConstraintWithGradient(self, cons, consGrad):
        self.constraint, self.consGrad = cons, consGrad

    def setScalingParams(self, zconst, zscale, ylinscale, yquadscale):
        self.zconst, self.zscale, self.ylinscale, self.yquadscale = zconst, zscale, ylinscale, yquadscale

    def setMoveLimit(self, movelim):
        self.moveLimit = movelim

    def setLowerAndUpperAsymptotes(self, low, upp):
        self.lowAsymp, self.upAsymp = low, upp

    def getOptimalValues(self):
        return self.xmma, self.ymma, self.zmma

    def getLagrangeMultipliers(self):
        return self.lam, self.xsi, self.eta, self.mu, self.zet

    def getSlackValue(self):
        return self.slack

    def getAsymptoteValues(self):
        return self.lowAsymp, self.upAsymp

    def mmasub(self, xval):
        """One MMA update (mma.py:114-205): moving asymptotes, move limits, the separable convex approximations p/(U-x) +
        q/(x-L) of objective and constraints, then subsolv."""
        m, n = self.numConstraints, self.numDesignVariables
        xv = _t(xval).reshape(-1)
        v = lambda a: _t(a, xv).reshape(-1)
        xmin, xmax, xold1, xold2 = v(self.xmin), v(self.xmax), v(self.xold1), v(self.xold2)
        df0dx, fval = v(self.objectiveGradient), v(self.constraint)
        dfdx = _t(self.consGrad, xv).reshape(m, n)
        low, upp = v(self.lowAsymp), v(self.upAsymp)
        a0 = float(np.asarray(self.zconst if not isinstance(self.zconst, torch.Tensor) else self.zconst.cpu()).reshape(-1)[0])
        a, c, d = v(self.zscale), v(self.ylinscale), v(self.yquadscale)
        move, span = self.moveLimit, xmax - xmin
        if self.epoch <= 2:
            low, upp = xv - _ASYINIT * span, xv + _ASYINIT * span
        else:
            osc = (xv - xold1) * (xold1 - xold2)
            factor = torch.where(osc > 0, torch.full_like(osc, _ASYINCR), torch.ones_like(osc))      # float64 throughout
            factor = torch.where(osc < 0, torch.full_like(osc, _ASYDECR), factor)
            low = xv - factor * (xold1 - low)
            upp = xv + factor * (upp - xold1)
            low = torch.minimum(torch.maximum(low, xv - 10 * span), xv - 0.01 * span)
            upp = torch.maximum(torch.minimum(upp, xv + 10 * span), xv + 0.01 * span)
        alfa = torch.maximum(torch.maximum(low + _ALBEFA * (xv - low), xv - move * span), xmin)
        beta = torch.minimum(torch.minimum(upp - _ALBEFA * (upp - xv), xv + move * span), xmax)
        inv_span = 1.0 / torch.clamp(span, min=1e-5)
        ux2, xl2 = (upp - xv) ** 2, (xv - low) ** 2
        p0, q0 = torch.clamp(df0dx, min=0), torch.clamp(-df0dx, min=0)
        pq0 = 0.001 * (p0 + q0) + _RAA0 * inv_span
        p0, q0 = (p0 + pq0) * ux2, (q0 + pq0) * xl2
        P, Q = torch.clamp(dfdx, min=0), torch.clamp(-dfdx, min=0)
        PQ = 0.001 * (P + Q) + _RAA0 * inv_span[None, :]
        P, Q = (P + PQ) * ux2[None, :], (Q + PQ) * xl2[None, :]
        b = P @ (1.0 / (upp - xv)) + Q @ (1.0 / (xv - low)) - fval
        x, y, z, lam, xsi, eta, mu, zet, s = subsolv(m, n, _EPSIMIN, low, upp, alfa, beta, p0, q0, P, Q, a0, a, b, c, d)
        col = lambda t: t.reshape(-1, 1)
        self.xmma, self.ymma, self.zmma = col(x), col(y), z.reshape(1, 1)
        self.lam, self.xsi, self.eta, self.mu, self.zet = col(lam), col(xsi), col(eta), col(mu), zet.reshape(1, 1)
        self.slack = col(s)
        self.lowAsymp, self.upAsymp = col(low), col(upp)


def optimize(fe, rho_ini, optimizationParams, objectiveHandle, consHandle, numConstraints):
    """Topology optimisation loop of the reference (mma.py:415-528), same signature and conventions:
    rho_ini (num_rho_vars, 1); ``J, dJ = objectiveHandle(rho_physical)``; ``vc, dvc = consHandle(rho_physical, iter)`` with
    vc (m,) and dvc (m, ...); optimizationParams = {'movelimit': .., 'maxIters': ..}.  Returns the optimised rho
    (a torch tensor on rho_ini's device)."""
    from . import logger
    rho = _t(rho_ini)
    dev = rho.device
    H, Hs = compute_filter_kd_tree(fe, device=dev)
    ft = {'H': H, 'Hs': Hs}
    m, n = numConstraints, rho.numel()
    mma = MMA()
    mma.setNumConstraints(m)
    mma.setNumDesignVariables(n)
    mma.setMinandMaxBoundsForDesignVariables(torch.zeros(n, 1, dtype=torch.float64, device=dev),
                                             torch.ones(n, 1, dtype=torch.float64, device=dev))
    xval = rho.reshape(-1, 1).clone()
    xold1, xold2 = xval.clone(), xval.clone()
    mma.registerMMAIter(xval, xold1, xold2)
    mma.setLowerAndUpperAsymptotes(torch.ones(n, 1, dtype=torch.float64, device=dev),
                                   torch.ones(n, 1, dtype=torch.float64, device=dev))
    mma.setScalingParams(1.0, torch.zeros(m, 1, dtype=torch.float64, device=dev),
                         10000 * torch.ones(m, 1, dtype=torch.float64, device=dev),
                         torch.zeros(m, 1, dtype=torch.float64, device=dev))
    mma.setMoveLimit(optimizationParams['movelimit'])
    loop = 0
    while loop < optimizationParams['maxIters']:
        loop += 1
        rho_physical = applyDensityFilter(ft, rho) if density_filtering else rho
        J, dJ = objectiveHandle(rho_physical)
        vc, dvc = consHandle(rho_physical, loop)
        dJ, dvc, vc = _t(dJ, rho), _t(dvc, rho), _t(vc, rho)
        if sensitivity_filtering:
            dJ, dvc = applySensitivityFilter(ft, rho, dJ.reshape(n, 1), dvc.reshape(m, n, 1))
        mma.setObjectiveWithGradient(float(J), dJ.reshape(-1, 1))
        mma.setConstraintWithGradient(vc.reshape(-1, 1), dvc.reshape(m, -1))
        mma.mmasub(xval)
        xmma, _, _ = mma.getOptimalValues()
        xold2, xold1, xval = xold1.clone(), xval.clone(), xmma.clone()
        mma.registerMMAIter(xval, xold1, xold2)
        rho = xval.reshape(rho.shape)
        logger.info("MMA iter %d; J %.5f; constraint %s", loop, float(J), vc.reshape(-1).tolist())
    return rho
