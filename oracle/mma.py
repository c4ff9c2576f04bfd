"""CPU restatement (NumPy/SciPy, float64) of the reference's topology-optimisation driver: the k-d-tree
sensitivity / density filter and the MMA optimiser (jax_fem/mma.py) -- SURVEY.md 8(f) row 1, the caller of the
hot path's (J, dJ/dtheta).  TEST INFRASTRUCTURE ONLY: nothing in jax_fem_b200/ may import this module.

Follows, function by function:
  kd_filter            <- compute_filter_kd_tree        (mma.py:27-57)
  sensitivity_filter   <- applySensitivityFilter         (mma.py:59-62)
  density_filter       <- applyDensityFilter             (mma.py:64-65)
  mma_step             <- MMA.mmasub                     (mma.py:114-205): asymptotes, move limits, p/q approximations
  subsolv              <- subsolv                        (mma.py:207-413): primal-dual interior-point Newton method
  optimize             <- optimize                       (mma.py:415-528)

Algorithm: K. Svanberg, "The method of moving asymptotes -- a new method for structural optimization", IJNME 24 (1987),
in the form of his 2007 note "MMA and GCMMA -- two methods for nonlinear optimization" (constants asyinit 0.5, asyincr 1.2,
asydecr 0.7, albefa 0.1, raa0 1e-5, epsimin 1e-7), which is what the reference's code implements.  Vectors are 1-D here
((n,), (m,)); the reference carries (n,1) / (m,1) columns.

Parity unpinned: the reference holds no golden vector or test for this module and cannot be imported here (jax); the
restatement is anchored by analytic checks in tests/test_oracle_mma.py (KKT conditions of the sub-problem, a convex
problem with a known optimum, filter weights against brute-force distances).
"""
import numpy as np
import scipy.sparse
import scipy.spatial


# ---- filters ---------------------------------------------------------------------------------------------------------
def kd_filter(points, cells, JxW, dim, flex_inds=None, num_nbs=20):
    """H (CSR, flex x flex) with H_ij = max(rmin - |c_i - c_j|, 0) over the num_nbs nearest centroids of cell i,
    rmin = 1.5 * (mean cell volume)^(1/dim); Hs = row sums.  mma.py:27-57."""
    cent = np.mean(np.take(points, cells, axis=0), axis=1)
    flex = np.arange(len(cells)) if flex_inds is None else np.asarray(flex_inds)
    fc = cent[flex]
    rmin = 1.5 * (np.sum(JxW) / len(cells)) ** (1.0 / dim)
    k = min(num_nbs, len(fc))
    dd, ii = scipy.spatial.KDTree(fc).query(fc, k)
    dd, ii = dd.reshape(len(fc), k), ii.reshape(len(fc), k)
    vals = np.where(rmin - dd > 0.0, rmin - dd, 0.0)
    rows = np.repeat(np.arange(len(fc)), k)
    H = scipy.sparse.csr_matrix((vals.reshape(-1), (rows, ii.reshape(-1))), shape=(len(fc), len(fc)))
    return H, np.asarray(H.sum(axis=1)).reshape(-1)


def sensitivity_filter(H, Hs, rho, dJ, dvc):
    """rho, dJ: (n, 1); dvc: (m, n, 1).  mma.py:59-62 (the scaling sits inside the sum, as the reference writes it)."""
    w = rho / np.maximum(1e-3, rho) / Hs[:, None]
    dJ_f = H @ (w * dJ)
    dvc_f = np.stack([H @ (w * g) for g in dvc])
    return dJ_f, dvc_f


def density_filter(H, Hs, rho):
    return (H @ rho) / Hs[:, None]


# ---- MMA sub-problem ---------------------------------------------------------------------------------------------------
EPSIMIN, RAA0, ALBEFA, ASYINIT, ASYINCR, ASYDECR = 1e-7, 1e-5, 0.1, 0.5, 1.2, 0.7


def _residual(x, y, z, lam, xsi, eta, mu, zet, s, epsi, low, upp, alfa, beta, p0, q0, P, Q, a0, a, b, c, d):
    ux, xl = upp - x, x - low
    plam, qlam = p0 + P.T @ lam, q0 + Q.T @ lam
    gvec = P @ (1.0 / ux) + Q @ (1.0 / xl)
    return np.concatenate([
        plam / ux ** 2 - qlam / xl ** 2 - xsi + eta,          # d/dx
        c + d * y - mu - lam,                                 # d/dy
        [a0 - zet - a @ lam],                                 # d/dz
        gvec - a * z - y + s - b,                             # constraints
        xsi * (x - alfa) - epsi, eta * (beta - x) - epsi,     # complementarity
        mu * y - epsi, [zet * z - epsi], lam * s - epsi])


def subsolv(m, n, epsimin, low, upp, alfa, beta, p0, q0, P, Q, a0, a, b, c, d):
    """Primal-dual Newton method for the MMA sub-problem (mma.py:207-413).  Returns x, y, z, lam, xsi, eta, mu, zet, s."""
    x = 0.5 * (alfa + beta)
    y, lam, s = np.ones(m), np.ones(m), np.ones(m)
    z = zet = 1.0
    xsi = np.maximum(1.0 / (x - alfa), 1.0)
    eta = np.maximum(1.0 / (beta - x), 1.0)
    mu = np.maximum(1.0, 0.5 * c)
    epsi = 1.0
    args = (low, upp, alfa, beta, p0, q0, P, Q, a0, a, b, c, d)
    while epsi > epsimin:
        res = _residual(x, y, z, lam, xsi, eta, mu, zet, s, epsi, *args)
        resnorm, resmax = np.sqrt(res @ res), np.abs(res).max()
        it = 0
        while resmax > 0.9 * epsi and it < 200:
            it += 1
            ux, xl = upp - x, x - low
            plam, qlam = p0 + P.T @ lam, q0 + Q.T @ lam
            gvec = P @ (1.0 / ux) + Q @ (1.0 / xl)
            GG = P / ux ** 2 - Q / xl ** 2                                         # (m, n)
            delx = plam / ux ** 2 - qlam / xl ** 2 - epsi / (x - alfa) + epsi / (beta - x)
            dely = c + d * y - lam - epsi / y
            delz = a0 - a @ lam - epsi / z
            dellam = gvec - a * z - y - b + epsi / lam
            diagx = 2.0 * (plam / ux ** 3 + qlam / xl ** 3) + xsi / (x - alfa) + eta / (beta - x)
            diagy = d + mu / y
            diaglamyi = s / lam + 1.0 / diagy
            if m < n:                                                               # Schur complement on (lam, z)
                blam = dellam + dely / diagy - GG @ (delx / diagx)
                Alam = np.diag(diaglamyi) + (GG / diagx) @ GG.T
                AA = np.block([[Alam, a[:, None]], [a[None, :], np.array([[-zet / z]])]])
                sol = np.linalg.solve(AA, np.concatenate([blam, [delz]]))
                dlam, dz = sol[:m], sol[m]
                dx = -delx / diagx - (GG.T @ dlam) / diagx
            else:                                                                   # Schur complement on (x, z)
                dellamyi = dellam + dely / diagy
                Axx = np.diag(diagx) + (GG.T / diaglamyi) @ GG
                azz = zet / z + a @ (a / diaglamyi)
                axz = -GG.T @ (a / diaglamyi)
                AA = np.block([[Axx, axz[:, None]], [axz[None, :], np.array([[azz]])]])
                bb = -np.concatenate([delx + GG.T @ (dellamyi / diaglamyi), [delz - a @ (dellamyi / diaglamyi)]])
                sol = np.linalg.solve(AA, bb)
                dx, dz = sol[:n], sol[n]
                dlam = (GG @ dx) / diaglamyi - dz * (a / diaglamyi) + dellamyi / diaglamyi
            dy = -dely / diagy + dlam / diagy
            dxsi = -xsi + epsi / (x - alfa) - xsi * dx / (x - alfa)
            deta = -eta + epsi / (beta - x) + eta * dx / (beta - x)
            dmu = -mu + epsi / y - mu * dy / y
            dzet = -zet + epsi / z - zet * dz / z
            ds = -s + epsi / lam - s * dlam / lam
            # largest step that keeps every positive quantity positive (factor 1.01), at most 1
            cur = np.concatenate([y, [z], lam, xsi, eta, mu, [zet], s])
            step = np.concatenate([dy, [dz], dlam, dxsi, deta, dmu, [dzet], ds])
            stminv = max((-1.01 * step / cur).max(), (-1.01 * dx / (x - alfa)).max(), (1.01 * dx / (beta - x)).max(), 1.0)
            steg = 1.0 / stminv
            old = (x, y, z, lam, xsi, eta, mu, zet, s)
            newnorm, itto = 2.0 * resnorm, 0
            while newnorm > resnorm and itto < 50:                                  # halve until the residual drops
                itto += 1
                x, y, z, lam, xsi, eta, mu, zet, s = (o + steg * dv for o, dv in
                                                      zip(old, (dx, dy, dz, dlam, dxsi, deta, dmu, dzet, ds)))
                res = _residual(x, y, z, lam, xsi, eta, mu, zet, s, epsi, *args)
                newnorm = np.sqrt(res @ res)
                steg *= 0.5
            resnorm, resmax = newnorm, np.abs(res).max()
        epsi *= 0.1
    return x, y, z, lam, xsi, eta, mu, zet, s


class MMAState:
    """What the reference keeps on its MMA object between iterations (mma.py:71-112)."""

    def __init__(self, x0, xmin, xmax, m, move, a0=1.0, a=None, c=None, d=None):
        n = len(x0)
        self.xval, self.xold1, self.xold2 = x0.copy(), x0.copy(), x0.copy()
        self.xmin, self.xmax = xmin, xmax
        self.low, self.upp = np.ones(n), np.ones(n)
        self.a0 = a0
        self.a = np.zeros(m) if a is None else a
        self.c = 10000.0 * np.ones(m) if c is None else c
        self.d = np.zeros(m) if d is None else d
        self.move, self.epoch, self.m, self.n = move, 1, m, n


def mma_step(st, f0val, df0dx, fval, dfdx):
    """One MMA update (MMA.mmasub, mma.py:114-205): returns the new design and the sub-problem's dual variables;
    st.low / st.upp are updated, the caller shifts xold2 <- xold1 <- xval <- xnew and bumps st.epoch."""
    xval, xmin, xmax, low, upp = st.xval, st.xmin, st.xmax, st.low, st.upp
    span = xmax - xmin
    if st.epoch <= 2:
        low, upp = xval - ASYINIT * span, xval + ASYINIT * span
    else:
        osc = (xval - st.xold1) * (st.xold1 - st.xold2)
        factor = np.where(osc > 0, ASYINCR, np.where(osc < 0, ASYDECR, 1.0))
        low = xval - factor * (st.xold1 - low)
        upp = xval + factor * (upp - st.xold1)
        low = np.minimum(np.maximum(low, xval - 10 * span), xval - 0.01 * span)
        upp = np.maximum(np.minimum(upp, xval + 10 * span), xval + 0.01 * span)
    alfa = np.maximum(np.maximum(low + ALBEFA * (xval - low), xval - st.move * span), xmin)
    beta = np.minimum(np.minimum(upp - ALBEFA * (upp - xval), xval + st.move * span), xmax)
    inv_span = 1.0 / np.maximum(span, 1e-5)
    ux2, xl2 = (upp - xval) ** 2, (xval - low) ** 2
    p0, q0 = np.maximum(df0dx, 0), np.maximum(-df0dx, 0)
    pq0 = 0.001 * (p0 + q0) + RAA0 * inv_span
    p0, q0 = (p0 + pq0) * ux2, (q0 + pq0) * xl2
    P, Q = np.maximum(dfdx, 0), np.maximum(-dfdx, 0)
    PQ = 0.001 * (P + Q) + RAA0 * inv_span[None, :]
    P, Q = (P + PQ) * ux2[None, :], (Q + PQ) * xl2[None, :]
    b = P @ (1.0 / (upp - xval)) + Q @ (1.0 / (xval - low)) - fval
    out = subsolv(st.m, st.n, EPSIMIN, low, upp, alfa, beta, p0, q0, P, Q, st.a0, st.a, b, st.c, st.d)
    st.low, st.upp = low, upp
    return out


def optimize(filter_HHs, rho_ini, params, objective, constraint, num_constraints, density_filtering=False,
             sensitivity_filtering=True, log=None):
    """mma.py:415-528.  rho_ini (n, 1); objective(rho) -> (J, dJ (n,1)); constraint(rho, iter) -> (vc (m,), dvc (m,n,1))."""
    H, Hs = filter_HHs
    rho = rho_ini
    n, m = rho.size, num_constraints
    st = MMAState(rho.reshape(-1).copy(), np.zeros(n), np.ones(n), m, params['movelimit'])
    for loop in range(1, params['maxIters'] + 1):
        rho_phys = density_filter(H, Hs, rho) if density_filtering else rho
        J, dJ = objective(rho_phys)
        vc, dvc = constraint(rho_phys, loop)
        if sensitivity_filtering:
            dJ, dvc = sensitivity_filter(H, Hs, rho, dJ, dvc)
        xnew = mma_step(st, float(J), np.asarray(dJ).reshape(-1), np.asarray(vc).reshape(-1),
                        np.asarray(dvc).reshape(m, -1))[0]
        st.xold2, st.xold1, st.xval = st.xold1, st.xval, xnew.copy()
        st.epoch += 1
        rho = xnew.reshape(rho.shape)
        if log is not None:
            log.append((float(J), np.asarray(vc).reshape(-1).copy()))
    return rho
