"""CPU ORACLE (test infrastructure only; never imported by the product path).

NumPy/SciPy float64 restatement of the reference hot path
    per-cell residual/tangent -> COO/CSR assembly -> Dirichlet row elimination
    -> Jacobi-preconditioned Krylov solve -> implicit adjoint
following, function by function,
    jax_fem/fe.py:112-141,143-183,185-218,220-258,271-322
    jax_fem/problem.py:82-107,130-182,189-259,426-460
    jax_fem/solver.py:63-92,290-363,390-421,469-553,1285-1356,1362-1418
    jax_fem/generate_mesh.py:120-189

Third-party arithmetic that is NOT under /root/reference and is restated from
its published algorithm (call sites in parentheses):
  * jax.scipy.sparse.linalg.bicgstab / cg (JAX, unpinned in the reference;
    solver.py:78-84)  -> bicgstab(), cg() below;
  * petsc4py==3.25.1 setPreallocationCOO/setValuesCOO/zeroRows/getValuesCSR/
    transpose (solver.py:476-478,525-528,65,1407-1408) -> coo_to_csr(), zero_rows();
  * forward-mode AD of the element kernel (problem.py:262-266) -> closed-form
    tangent moduli in oracle/laws.py, FD-checked in tests/test_oracle_laws.py.

Parity status: PINNED for final solutions by the reference's four FEniCSx
goldens (tests/golden/*.npz, made by tests/golden/make_fixtures.py; checked in
tests/test_oracle_golden.py at the reference's own tolerances and much
tighter).  Element values, CSR pattern, BC masks, Krylov iterates and adjoint
gradients are not pinned by any reference test; for
those this oracle is the authority, anchored by the goldens plus analytic
checks (rigid-body null space, FD of tangent and of the adjoint gradient).
"""
import numpy as np
import scipy.sparse as sp

from . import basis as _basis

try:
    from threadpoolctl import threadpool_limits as _tpl

    def _blas_single_thread():
        return _tpl(limits=1, user_api='blas')
except Exception:            # pragma: no cover
    import contextlib

    def _blas_single_thread():
        return contextlib.nullcontext()


# ----------------------------------------------------------------------------
# generate_mesh.py:120-189
class Mesh:
    def __init__(self, points, cells, ele_type=None):
        self.points = np.asarray(points, dtype=np.float64)
        self.cells = np.asarray(cells)
        self.ele_type = ele_type


def rectangle_mesh(Nx, Ny, domain_x, domain_y):
    """generate_mesh.py:120-148 (QUAD4)."""
    x = np.linspace(0, domain_x, Nx + 1)
    y = np.linspace(0, domain_y, Ny + 1)
    xv, yv = np.meshgrid(x, y, indexing='ij')
    points = np.stack((xv, yv), axis=2).reshape(-1, 2)
    ids = np.arange(len(points)).reshape(Nx + 1, Ny + 1)
    cells = np.stack((ids[:-1, :-1], ids[1:, :-1], ids[1:, 1:], ids[:-1, 1:]), axis=2).reshape(-1, 4)
    return Mesh(points, cells, 'QUAD4')


def box_mesh(Nx, Ny, Nz, domain_x, domain_y, domain_z):
    """generate_mesh.py:151-189 (HEX8)."""
    x = np.linspace(0, domain_x, Nx + 1)
    y = np.linspace(0, domain_y, Ny + 1)
    z = np.linspace(0, domain_z, Nz + 1)
    xv, yv, zv = np.meshgrid(x, y, z, indexing='ij')
    points = np.stack((xv, yv, zv), axis=3).reshape(-1, 3)
    ids = np.arange(len(points)).reshape(Nx + 1, Ny + 1, Nz + 1)
    cells = np.stack((ids[:-1, :-1, :-1], ids[1:, :-1, :-1], ids[1:, 1:, :-1], ids[:-1, 1:, :-1],
                      ids[:-1, :-1, 1:], ids[1:, :-1, 1:], ids[1:, 1:, 1:], ids[:-1, 1:, 1:]),
                     axis=3).reshape(-1, 8)
    return Mesh(points, cells, 'HEX8')


# ----------------------------------------------------------------------------
def _det_inv(J):
    """Closed-form det / inverse of (...,d,d), d in {2,3} (np.linalg at fe.py:134-135)."""
    return np.linalg.det(J), np.linalg.inv(J)


class FiniteElement:
    """fe.py:77-110."""

    def __init__(self, mesh, vec, dim, ele_type, quadrature_order=None, dirichlet_bc_info=None):
        self.mesh, self.vec, self.dim, self.ele_type = mesh, vec, dim, ele_type
        self.points = mesh.points
        self.cells = np.asarray(mesh.cells)
        self.num_cells = len(self.cells)
        self.num_total_nodes = len(self.points)
        self.num_total_dofs = self.num_total_nodes * vec
        self.shape_vals, self.shape_grads_ref, self.quad_weights = \
            _basis.get_shape_vals_and_grads(ele_type, quadrature_order)
        (self.face_shape_vals, self.face_shape_grads_ref, self.face_quad_weights,
         self.face_normals, self.face_inds) = _basis.get_face_shape_vals_and_grads(ele_type, quadrature_order)
        self.num_quads, self.num_nodes = self.shape_vals.shape
        self.num_faces = self.face_shape_vals.shape[0]
        self.num_face_quads = self.face_quad_weights.shape[1]
        self.node_inds_list, self.vec_inds_list, self.vals_list = \
            self.Dirichlet_boundary_conditions(dirichlet_bc_info)

    def get_shape_grads(self):
        """fe.py:112-141."""
        coos = self.points[self.cells]                                            # (C,N,dim)
        J = np.einsum('cnd,qne->cqde', coos, self.shape_grads_ref)                # fe.py:132
        det, inv = _det_inv(J)
        grads = np.einsum('qne,cqed->cqnd', self.shape_grads_ref, inv)            # fe.py:138-139
        return grads, det * self.quad_weights[None, :]                           # fe.py:140

    def get_face_shape_grads(self, boundary_inds):
        """fe.py:143-183."""
        coos = self.points[self.cells][boundary_inds[:, 0]]                       # (F,N,dim)
        gref = self.face_shape_grads_ref[boundary_inds[:, 1]]                     # (F,FQ,N,dim)
        normals = self.face_normals[boundary_inds[:, 1]]                          # (F,dim)
        J = np.einsum('fnd,fqne->fqde', coos, gref)                               # fe.py:170
        det, inv = _det_inv(J)
        grads = np.einsum('fqne,fqed->fqnd', gref, inv)                           # fe.py:176
        nanson = np.linalg.norm(np.einsum('fe,fqed->fqd', normals, inv), axis=-1)  # fe.py:180
        return grads, nanson * det * self.face_quad_weights[boundary_inds[:, 1]]  # fe.py:181-182

    def get_physical_quad_points(self):
        """fe.py:185-197."""
        return np.einsum('qn,cnd->cqd', self.shape_vals, self.points[self.cells])

    def get_physical_surface_quad_points(self, boundary_inds):
        """fe.py:199-218."""
        coos = self.points[self.cells][boundary_inds[:, 0]]
        return np.einsum('fqn,fnd->fqd', self.face_shape_vals[boundary_inds[:, 1]], coos)

    def Dirichlet_boundary_conditions(self, info):
        """fe.py:220-258.  location_fn(point[, ind]) -> bool; value_fn(point) -> float."""
        node_inds_list, vec_inds_list, vals_list = [], [], []
        if info is not None:
            location_fns, vecs, value_fns = info
            assert len(location_fns) == len(value_fns) == len(vecs)
            for i in range(len(location_fns)):
                flags = _eval_location(location_fns[i], self.points, np.arange(self.num_total_nodes))
                node_inds = np.argwhere(flags).reshape(-1)
                vals = np.array([float(value_fns[i](p)) for p in self.points[node_inds]], dtype=np.float64)
                node_inds_list.append(node_inds)
                vec_inds_list.append(np.ones_like(node_inds, dtype=np.int32) * vecs[i])
                vals_list.append(vals)
        return node_inds_list, vec_inds_list, vals_list

    def get_boundary_conditions_inds(self, location_fns):
        """fe.py:271-322: a face is selected iff ALL its vertices satisfy the predicate."""
        cell_face_inds = self.cells[:, self.face_inds]                            # (C,F,V)
        out = []
        if location_fns is not None:
            for fn in location_fns:
                flags = _eval_location(fn, self.points, np.arange(self.num_total_nodes))
                out.append(np.argwhere(np.all(flags[cell_face_inds], axis=-1)))   # (S,2)
        return out


def _eval_location(fn, points, inds):
    nargs = fn.__code__.co_argcount
    try:    # NumPy-style predicates evaluate on the transposed array in one call (point[d] -> all d-coordinates)
        flags = np.asarray(fn(points.T) if nargs == 1 else fn(points.T, inds))
        if flags.shape == (len(points),) and flags.dtype == np.bool_:
            return flags
    except Exception:
        pass
    if nargs == 1:
        return np.array([bool(fn(p)) for p in points])
    if nargs == 2:
        return np.array([bool(fn(p, i)) for p, i in zip(points, inds)])
    raise ValueError(f"Wrong number of arguments for location_fn: must be 1 or 2, get {nargs}")


# ----------------------------------------------------------------------------
class Problem:
    """problem.py:46-128 for ONE variable (the hot path's scope).

    law          : oracle.laws.* object (tensor map + closed-form tangent)
    mass_map     : vectorised f(u (...,vec), x (...,dim)) -> (...,vec) or None; if it depends on u, mass_map_jac gives d f / d u
    surface_maps : list of such functions, one per location_fn
    internal_vars: list of (C,Q,...) arrays forwarded to the law
    """

    def __init__(self, mesh, vec, dim, ele_type='HEX8', quadrature_order=None, dirichlet_bc_info=None,
                 location_fns=None, law=None, mass_map=None, surface_maps=None, internal_vars=(), surface_map_jacs=None,
                 mass_map_jac=None):
        self.fe = FiniteElement(mesh, vec, dim, ele_type, quadrature_order, dirichlet_bc_info)
        self.fes = [self.fe]
        fe = self.fe
        self.vec, self.dim, self.law = vec, dim, law
        self.mass_map, self.surface_maps = mass_map, surface_maps or []
        self.mass_map_jac = mass_map_jac      # vectorised d mass_map / d u -> (..., vec, vec); None: the map does not depend on u
        # d(surface_map)/du (..., vec, vec) per boundary set, or None for a u-independent map (the reference gets it from jacfwd
        # of the surface kernel, problem.py:299-301)
        self.surface_map_jacs = surface_map_jacs or [None] * len(self.surface_maps)
        self.internal_vars = list(internal_vars)
        self.num_cells = fe.num_cells
        self.cells = fe.cells
        self.boundary_inds_list = fe.get_boundary_conditions_inds(location_fns)
        assert len(self.boundary_inds_list) == len(self.surface_maps), "Missing definitions for surface integral"
        self.num_total_dofs_all_vars = fe.num_total_dofs
        # problem.py:86-107 -- the COO pattern, cell blocks then face blocks
        self.inds = (vec * fe.cells[:, :, None] + np.arange(vec)[None, None, :]).reshape(self.num_cells, -1)
        self.ndof = self.inds.shape[1]
        self.shape_grads, self.JxW = fe.get_shape_grads()
        self.physical_quad_points = fe.get_physical_quad_points()
        self.face_data = []
        for b in self.boundary_inds_list:
            _, nanson = fe.get_face_shape_grads(b)
            self.face_data.append((fe.face_shape_vals[b[:, 1]], nanson, fe.get_physical_surface_quad_points(b)))

    # problem.py:95-107 (materialised only on request: C*ndof^2 integers)
    def coo_pattern(self):
        def blocks(inds):
            n = inds.shape[1]
            return (np.repeat(inds[:, :, None], n, axis=2).reshape(-1),
                    np.repeat(inds[:, None, :], n, axis=1).reshape(-1))
        I, J = blocks(self.inds)
        for b in self.boundary_inds_list:
            If, Jf = blocks(self.inds[b[:, 0]])
            I, J = np.hstack((I, If)), np.hstack((J, Jf))
        return I, J

    def _u_grads(self, sol, sl):
        return np.einsum('cnv,cqnd->cqvd', sol[self.cells[sl]], self.shape_grads[sl])   # problem.py:204-205

    def _iv(self, sl):
        return [v[sl] for v in self.internal_vars]

    def cell_residuals(self, sol, sl=slice(None)):
        """laplace + mass kernels, problem.py:189-236 -> (C,N,vec)."""
        fe = self.fe
        out = np.zeros((len(self.cells[sl]), fe.num_nodes, self.vec))
        if self.law is not None:
            sig = self.law.stress(self._u_grads(sol, sl), *self._iv(sl))               # problem.py:208
            out += np.einsum('cqvd,cqnd,cq->cnv', sig, self.shape_grads[sl], self.JxW[sl])  # problem.py:210
        if self.mass_map is not None:
            u = np.einsum('cnv,qn->cqv', sol[self.cells[sl]], fe.shape_vals)            # problem.py:229
            val = self.mass_map(u, self.physical_quad_points[sl])
            out += np.einsum('cqv,qn,cq->cnv', np.broadcast_to(val, u.shape), fe.shape_vals, self.JxW[sl])
        return out

    def cell_jacobians(self, sol, sl=slice(None)):
        """value_and_jacfwd of the cell kernel, problem.py:262-266 -> (C, ndof, ndof), row = test dof."""
        A = self.law.tangent(self._u_grads(sol, sl), *self._iv(sl))                     # (C,Q,v,d,v,d)
        g = self.shape_grads[sl]
        C, Q, N, d = g.shape
        v = self.vec
        # B[(i,d),(n,i')] = delta_ii' g[n,d]  =>  K_e = sum_q JxW B^T A B  (batched GEMMs instead of a 6-index einsum)
        B = np.einsum('cqnd,ij->cqidnj', g, np.eye(v)).reshape(C, Q, v * d, N * v)
        with _blas_single_thread():      # tiny GEMMs: BLAS thread hand-offs dominate on many-core hosts
            AB = np.matmul(A.reshape(C, Q, v * d, v * d) * self.JxW[sl][:, :, None, None], B)
            K = np.matmul(B.transpose(0, 1, 3, 2), AB).sum(axis=1)
        K = K.reshape(K.shape[0], self.ndof, self.ndof)
        if self.mass_map_jac is not None:                                               # problem.py:216-236 differentiated
            u = np.einsum('cnv,qn->cqv', sol[self.cells[sl]], self.fe.shape_vals)
            dm = np.broadcast_to(self.mass_map_jac(u, self.physical_quad_points[sl]), u.shape + (v,))
            M = np.einsum('cqik,qa,qb,cq->caibk', dm, self.fe.shape_vals, self.fe.shape_vals, self.JxW[sl])
            K = K + M.reshape(K.shape)
        return K

    def face_residuals(self, sol, k):
        """surface kernel, problem.py:238-259 -> (S,N,vec)."""
        b = self.boundary_inds_list[k]
        vals, nanson, x = self.face_data[k]
        u = np.einsum('fnv,fqn->fqv', sol[self.cells[b[:, 0]]], vals)
        t = np.broadcast_to(self.surface_maps[k](u, x), u.shape)
        return np.einsum('fqv,fqn,fq->fnv', t, vals, nanson)

    def face_jacobians(self, sol, k):
        """d(face residual)/d(cell dofs) of boundary set k -> (S, ndof, ndof), row = test dof (problem.py:289-325)."""
        b = self.boundary_inds_list[k]
        vals, nanson, x = self.face_data[k]
        if self.surface_map_jacs[k] is None:
            return np.zeros((len(b), self.ndof, self.ndof))
        u = np.einsum('fnv,fqn->fqv', sol[self.cells[b[:, 0]]], vals)
        dt = np.broadcast_to(self.surface_map_jacs[k](u, x), u.shape + (self.vec,))          # (S,FQ,vec,vec)
        K = np.einsum('fqik,fqa,fqb,fq->faibk', dt, vals, vals, nanson)
        return K.reshape(len(b), self.ndof, self.ndof)

    def compute_residual(self, sol):
        """problem.py:426-445 -> (nodes, vec)."""
        res = np.zeros((self.fe.num_total_nodes, self.vec))
        np.add.at(res, self.cells.reshape(-1), self.cell_residuals(sol).reshape(-1, self.vec))
        for k, b in enumerate(self.boundary_inds_list):
            np.add.at(res, self.cells[b[:, 0]].reshape(-1), self.face_residuals(sol, k).reshape(-1, self.vec))
        return res

    def newton_update(self, sol):
        """problem.py:447-460: residual + COO values V (cells, then zero face blocks)."""
        self.V_cells = self.cell_jacobians(sol)
        self.V_faces = [self.face_jacobians(sol, k) for k in range(len(self.boundary_inds_list))]
        return self.compute_residual(sol)

    def coo_values(self):
        V = self.V_cells.reshape(-1)
        faces = getattr(self, 'V_faces', None)
        for k, b in enumerate(self.boundary_inds_list):       # u-independent loads: exact zeros (problem.py:456-458)
            V = np.hstack((V, faces[k].reshape(-1) if faces is not None else np.zeros(len(b) * self.ndof ** 2)))
        return V

    # solver.py:511-516
    def bc_rows(self):
        fe = self.fe
        return [np.asarray(fe.node_inds_list[i] * fe.vec + fe.vec_inds_list[i]) for i in range(len(fe.node_inds_list))]


# ----------------------------------------------------------------------------
# PETSc semantics (solver.py:469-553)
def coo_to_csr(I, J, V, n):
    """setPreallocationCOO + setValuesCOO: duplicates summed, explicit zeros kept, columns ascending."""
    A = sp.coo_matrix((V, (I, J)), shape=(n, n)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def csr_pattern_from_cells(cells, vec, n):
    """Pattern only, without materialising C*ndof^2 COO entries (same result as coo_to_csr's pattern)."""
    C, N = cells.shape
    a = np.repeat(cells[:, :, None], N, axis=2).reshape(-1)
    b = np.repeat(cells[:, None, :], N, axis=1).reshape(-1)
    G = sp.coo_matrix((np.ones(len(a), dtype=np.int8), (a, b)), shape=(n // vec, n // vec)).tocsr()
    G.sum_duplicates()
    G.sort_indices()
    B = sp.kron(G, np.ones((vec, vec), dtype=np.int8), format='csr')
    B.sort_indices()
    return B.indptr.astype(np.int32), B.indices.astype(np.int32)


def zero_rows(A, rows_list):
    """Mat.zeroRows with KEEP_NONZERO_PATTERN (solver.py:477,527-528): row <- 0, diagonal <- 1."""
    A = A.copy()
    if len(rows_list) == 0:
        return A
    mask = np.zeros(A.shape[0], dtype=bool)
    mask[np.concatenate(rows_list)] = True
    rowid = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    sel = mask[rowid]
    A.data[sel] = (A.indices[sel] == rowid[sel]).astype(np.float64)
    return A


def get_A(problem):
    """solver.py:540-553 (no P_mat)."""
    I, J = problem.coo_pattern()
    A = coo_to_csr(I, J, problem.coo_values(), problem.num_total_dofs_all_vars)
    return zero_rows(A, problem.bc_rows())


# solver.py:290-363
def apply_bc_vec(res_vec, dofs, problem, scale=1.):
    fe = problem.fe
    res = res_vec.reshape(-1, fe.vec).copy()
    sol = dofs.reshape(-1, fe.vec)
    for i in range(len(fe.node_inds_list)):
        n, v = fe.node_inds_list[i], fe.vec_inds_list[i]
        res[n, v] = sol[n, v]
        res[n, v] = res[n, v] - fe.vals_list[i] * scale
    return res.reshape(-1)


def assign_bc(dofs, problem):
    fe = problem.fe
    sol = dofs.reshape(-1, fe.vec).copy()
    for i in range(len(fe.node_inds_list)):
        sol[fe.node_inds_list[i], fe.vec_inds_list[i]] = fe.vals_list[i]
    return sol.reshape(-1)


def copy_bc(dofs, problem):
    fe = problem.fe
    sol = dofs.reshape(-1, fe.vec)
    new = np.zeros_like(sol)
    for i in range(len(fe.node_inds_list)):
        n, v = fe.node_inds_list[i], fe.vec_inds_list[i]
        new[n, v] = sol[n, v]
    return new.reshape(-1)


# ----------------------------------------------------------------------------
# jax.scipy.sparse.linalg (third party; restated)
def bicgstab(A, b, x0=None, M=None, tol=1e-10, atol=1e-10, maxiter=10000):
    """Returns (x, iterations).  Same recurrences, stopping rule and early exit as JAX's _bicgstab_solve."""
    M = M or (lambda v: v)
    x = np.zeros_like(b) if x0 is None else x0.copy()
    atol2 = max(tol ** 2 * float(b @ b), atol ** 2)
    r = b - A @ x
    rhat = r.copy()
    rho = alpha = omega = 1.0
    p = r.copy()
    q = r.copy()
    k = 0
    while float(r @ r) > atol2 and 0 <= k < maxiter:
        rho_ = float(rhat @ r)
        beta = rho_ / rho * alpha / omega
        p = r + beta * (p - omega * q)
        phat = M(p)
        q = A @ phat
        alpha = rho_ / float(rhat @ q)
        s = r - alpha * q
        if float(s @ s) < atol2:
            x = x + alpha * phat
            r = s
            k += 1
            rho = rho_
            continue
        shat = M(s)
        t = A @ shat
        omega = float(t @ s) / float(t @ t)
        x = x + alpha * phat + omega * shat
        r = s - omega * t
        rho = rho_
        k = -11 if (omega == 0 or alpha == 0) else k + 1
        if rho_ == 0:
            k = -10
    return x, k


def cg(A, b, x0=None, M=None, tol=1e-10, atol=1e-10, maxiter=10000):
    """Returns (x, iterations).  JAX's _cg_solve with a preconditioner (true residual norm test)."""
    M = M or (lambda v: v)
    x = np.zeros_like(b) if x0 is None else x0.copy()
    atol2 = max(tol ** 2 * float(b @ b), atol ** 2)
    r = b - A @ x
    z = M(r)
    p = z.copy()
    gamma = float(r @ z)
    k = 0
    while float(r @ r) > atol2 and k < maxiter:
        Ap = A @ p
        alpha = gamma / float(p @ Ap)
        x = x + alpha * p
        r = r - alpha * Ap
        z = M(r)
        gamma_ = float(r @ z)
        p = z + (gamma_ / gamma) * p
        gamma = gamma_
        k += 1
    return x, k


def jax_solve(A, b, x0, precond=True, method='bicgstab', return_iters=False):
    """solver.py:63-92 (method='cg' is the north-star's Jacobi-CG on the same operator)."""
    jacobi = A.diagonal()
    M = (lambda v: v / jacobi) if precond else None
    fn = bicgstab if method == 'bicgstab' else cg
    x, k = fn(A, b, x0=x0, M=M, tol=1e-10, atol=1e-10, maxiter=10000)
    err = np.linalg.norm(A @ x - b)
    assert err < 0.1, f"linear solver failed to converge with err = {err}"
    return (x, k) if return_iters else x


# ----------------------------------------------------------------------------
def line_search(problem, dofs, inc):
    """Step halving of solver.py:424-462: alpha = 1, then up to three halvings, each kept only while the norm of the
    residual (with Dirichlet rows) keeps decreasing."""
    def res_norm(alpha):
        d = dofs + alpha * inc
        return np.linalg.norm(apply_bc_vec(problem.compute_residual(d.reshape(-1, problem.vec)).reshape(-1), d, problem))
    alpha = 1.0
    norm = res_norm(alpha)
    for _ in range(3):
        alpha *= 0.5
        half = res_norm(alpha)
        if half > norm:
            alpha *= 2.0
            break
        norm = half
    return dofs + alpha * inc


def solver(problem, tol=1e-6, rel_tol=1e-8, method='bicgstab', initial_guess=None, log=None, line_search_flag=False):
    """Newton loop, solver.py:1285-1356 + newton_step :390-421.  Returns (nodes, vec) solution."""
    n = problem.num_total_dofs_all_vars
    dofs = np.zeros(n) if initial_guess is None else np.asarray(initial_guess, dtype=np.float64).reshape(-1).copy()

    def helper(dofs):
        res = problem.newton_update(dofs.reshape(-1, problem.vec)).reshape(-1)
        return apply_bc_vec(res, dofs, problem), get_A(problem)

    res_vec, A = helper(dofs)
    res_val = res0 = np.linalg.norm(res_vec)
    hist = [res_val]
    while res_val / res0 > rel_tol and res_val > tol:
        x0 = assign_bc(np.zeros(n), problem) - copy_bc(dofs, problem)     # solver.py:402-409
        inc = jax_solve(A, -res_vec, x0, True, method)
        dofs = line_search(problem, dofs, inc) if line_search_flag else dofs + inc            # solver.py:416-419
        res_vec, A = helper(dofs)
        res_val = np.linalg.norm(res_vec)
        hist.append(res_val)
    assert np.isfinite(res_val) and np.all(np.isfinite(dofs))
    if log is not None:
        log.extend(hist)
    return dofs.reshape(-1, problem.vec)


def implicit_vjp(problem, sol, v, method='bicgstab'):
    """solver.py:1362-1418 for per-quadrature-point parameters internal_vars[0] of shape (C,Q).

    Returns dL/dtheta (C,Q) given v = dL/dsol (nodes, vec):  -lambda^T dc/dtheta  with
    A^T lambda = v (x0=None) and c = apply_bc(compute_residual)  (BC rows of c do not depend on theta).
    """
    problem.newton_update(sol)
    A = get_A(problem)
    lam = jax_solve(A.T.tocsr(), v.reshape(-1), None, True, method).reshape(-1, problem.vec).copy()
    fe = problem.fe
    for i in range(len(fe.node_inds_list)):
        lam[fe.node_inds_list[i], fe.vec_inds_list[i]] = 0.0
    ds = problem.law.dstress_dparam(problem._u_grads(sol, slice(None)), *problem.internal_vars)  # (C,Q,v,d)
    return -np.einsum('cnv,cqvd,cqnd,cq->cq', lam[problem.cells], ds, problem.shape_grads, problem.JxW)
