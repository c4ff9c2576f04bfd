"""Analytic anchors of the oracle's MMA optimiser and k-d-tree filters (oracle/mma.py).  The reference
holds no test or golden vector for jax_fem/mma.py and cannot be imported here, so the restatement is checked against
the mathematics: optimality conditions of the sub-problem, a convex problem with a known optimum, brute-force filter
weights."""
import numpy as np

from oracle import fem, mma


def _random_subproblem(m, n, seed):
    rng = np.random.default_rng(seed)
    xval = rng.uniform(0.2, 0.8, n)
    xmin, xmax = np.zeros(n), np.ones(n)
    st = mma.MMAState(xval, xmin, xmax, m, move=0.2)
    return st, rng.standard_normal(n), rng.standard_normal(m) * 0.1, rng.standard_normal((m, n))


def test_subproblem_solution_satisfies_the_kkt_conditions():
    for m, n, seed in [(1, 40, 0), (3, 25, 1), (6, 4, 2)]:          # m < n and m >= n branches
        st, df0, f, df = _random_subproblem(m, n, seed)
        x, y, z, lam, xsi, eta, mu, zet, s = mma.mma_step(st, 1.0, df0, f, df)
        assert np.all(x >= st.xmin - 1e-12) and np.all(x <= st.xmax + 1e-12)
        assert np.all(np.abs(x - st.xval) <= 0.2 + 1e-9)                          # move limit
        assert min(y.min(), z, lam.min(), xsi.min(), eta.min(), mu.min(), zet, s.min()) > 0
        # rebuild the sub-problem data exactly as mma_step does and evaluate the residual at the final barrier value
        low, upp = st.low, st.upp
        span = st.xmax - st.xmin
        alfa = np.maximum(np.maximum(low + 0.1 * (st.xval - low), st.xval - 0.2 * span), st.xmin)
        beta = np.minimum(np.minimum(upp - 0.1 * (upp - st.xval), st.xval + 0.2 * span), st.xmax)
        ux2, xl2 = (upp - st.xval) ** 2, (st.xval - low) ** 2
        p0, q0 = np.maximum(df0, 0), np.maximum(-df0, 0)
        pq0 = 0.001 * (p0 + q0) + 1e-5 / span
        p0, q0 = (p0 + pq0) * ux2, (q0 + pq0) * xl2
        P, Q = np.maximum(df, 0), np.maximum(-df, 0)
        PQ = 0.001 * (P + Q) + 1e-5 / span[None, :]
        P, Q = (P + PQ) * ux2[None], (Q + PQ) * xl2[None]
        b = P @ (1 / (upp - st.xval)) + Q @ (1 / (st.xval - low)) - f
        res = mma._residual(x, y, z, lam, xsi, eta, mu, zet, s, 1e-7, low, upp, alfa, beta, p0, q0, P, Q, st.a0, st.a, b,
                            st.c, st.d)
        assert np.abs(res).max() <= 0.9e-7 * 1.0001                                # the method's own stopping rule
        # the approximating functions interpolate value and gradient at xval: g_i(xval) = f_i, so b + f = g(xval) - ... >= 0 form
        g_at_xval = P @ (1 / (upp - st.xval)) + Q @ (1 / (st.xval - low)) - b
        assert np.abs(g_at_xval - f).max() < 1e-12


def test_mma_converges_to_the_known_optimum_of_a_convex_problem():
    """min sum (x - t)^2  s.t.  mean(x) <= v, 0 <= x <= 1: the optimum is the clipped shift x = clip(t - nu, 0, 1)."""
    rng = np.random.default_rng(3)
    n, v = 60, 0.4
    t = rng.uniform(0.1, 0.95, n)
    lo, hi = -1.0, 1.0
    for _ in range(200):                                       # bisection on the multiplier
        nu = 0.5 * (lo + hi)
        lo, hi = (nu, hi) if np.clip(t - nu, 0, 1).mean() > v else (lo, nu)
    x_star = np.clip(t - nu, 0, 1)
    obj = lambda rho: (float(((rho[:, 0] - t) ** 2).sum()), 2 * (rho - t[:, None]))
    con = lambda rho, it: (np.array([rho.mean() / v - 1.0]), (np.ones((1, n, 1)) / (n * v)))
    H = mma.scipy.sparse.identity(n, format='csr')
    log = []
    rho = mma.optimize((H, np.ones(n)), np.full((n, 1), v), {'movelimit': 0.2, 'maxIters': 40}, obj, con, 1,
                       sensitivity_filtering=False, log=log)
    assert np.abs(rho[:, 0] - x_star).max() < 2e-4
    assert log[-1][1][0] < 1e-6 and log[-1][0] < log[0][0]


def test_kd_filter_weights_match_brute_force_distances():
    m = fem.box_mesh(5, 4, 3, 1.0, 0.8, 0.6)
    fe = fem.FiniteElement(m, 3, 3, 'HEX8')
    JxW = fe.get_shape_grads()[1]
    H, Hs = mma.kd_filter(m.points, m.cells, JxW, 3)
    cent = m.points[m.cells].mean(axis=1)
    rmin = 1.5 * (JxW.sum() / len(m.cells)) ** (1 / 3)
    D = np.linalg.norm(cent[:, None] - cent[None], axis=-1)
    ref = np.maximum(rmin - D, 0.0)                             # at most 19 positive weights per row < 20 neighbours
    assert np.abs(H.toarray() - ref).max() < 1e-13
    assert np.abs(Hs - ref.sum(1)).max() < 1e-13
    rho = np.random.default_rng(0).uniform(0.2, 1.0, (len(m.cells), 1))
    dJ = np.random.default_rng(1).standard_normal((len(m.cells), 1))
    dJf, dvcf = mma.sensitivity_filter(H, Hs, rho, dJ, dJ[None])
    assert np.abs(dJf - ref @ (dJ / ref.sum(1)[:, None])).max() < 1e-12 and np.array_equal(dvcf[0], dJf)
    assert np.abs(mma.density_filter(H, Hs, np.ones((len(m.cells), 1))) - 1.0).max() < 1e-14
