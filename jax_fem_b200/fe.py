"""Per-variable finite-element data (host side).

Mirror of jax_fem/fe.py::FiniteElement: reference tables, Dirichlet index sets (:220-258),
boundary-face selection (:271-322) and the geometric helpers (:112-218) that user post-processing
calls.  The hot path never materialises (C,Q,N,dim) shape gradients -- the CUDA element kernels
recompute them per cell -- so ``get_shape_grads`` & co. are provided for post-processing only and
run on the host in NumPy.
"""
import time
from dataclasses import dataclass
from typing import Any, Optional

import numpy as np

from . import logger
from .basis import get_face_shape_vals_and_grads, get_shape_vals_and_grads
from .generate_mesh import Mesh

def _arg_count(fn):
    """Number of positional parameters of a user predicate (functions, functools.partial, callable objects)."""
    code = getattr(fn, '__code__', None)
    if code is not None:
        return code.co_argcount
    import inspect
    params = inspect.signature(fn).parameters.values()
    return sum(p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) for p in params)


def _pointwise_equal(call_one, batch, n, same):
    """True when the batched result equals the per-point calls at EVERY point (the reference vmaps per point,
    fe.py:252; a predicate that mixes per-point values with reductions over the batch must not slip through)."""
    return all(same(call_one(i), batch[i]) for i in range(n))


def evaluate_location_fn(fn, points, inds=None):
    """Boolean flag per point for a reference-style predicate ``fn(point[, ind])``.

    The reference vmaps the predicate over points (fe.py:252, 312-318).  Here the predicate is first called once on
    the transposed array (point[d] becomes the vector of d-coordinates, which is what NumPy-style predicates such as
    ``np.isclose(point[0], 0., atol=1e-5)`` need).  The batched result is only used if it agrees with the per-point
    call at every point where either says True and on a deterministic spread of the others; a predicate that cannot be
    batched is evaluated point by point.
    """
    n = len(points)
    inds = np.arange(n) if inds is None else inds
    nargs = _arg_count(fn)
    if nargs not in (1, 2):
        raise ValueError(f"Wrong number of arguments for location_fn: must be 1 or 2, get {nargs}")
    call = (lambda p, i: fn(p)) if nargs == 1 else fn
    try:
        flags = np.asarray(call(points.T, inds))
        if flags.shape == (n,) and flags.dtype == np.bool_:
            # Dirichlet / load sets are small against the mesh: check every selected point, every point of a shuffled copy
            # of the batch (a reduction over the batch changes with the batch), and a spread of unselected points
            perm = np.random.default_rng(12345).permutation(n)
            shuffled = np.asarray(call(points[perm].T, inds[perm]))
            if shuffled.shape == (n,) and np.array_equal(shuffled, flags[perm]):
                check = np.union1d(np.flatnonzero(flags)[:4096], np.linspace(0, n - 1, min(n, 256)).astype(np.int64))
                if all(bool(call(points[i], inds[i])) == bool(flags[i]) for i in check):
                    return flags
    except Exception:
        pass
    return np.array([bool(call(points[i], inds[i])) for i in range(n)], dtype=bool)


def evaluate_point_fn(fn, points, out_shape=()):
    """Values of ``fn(point)`` for every point -> (n, *out_shape), batched like evaluate_location_fn (same safeguards:
    the batched values must not change when the batch is shuffled and must equal per-point calls on a spread of points)."""
    n = len(points)
    if n == 0:
        return np.zeros((0,) + tuple(out_shape))

    def batched(pts):
        val = np.asarray(fn(pts.T), dtype=np.float64)
        if val.shape == tuple(out_shape):
            return np.broadcast_to(val, (len(pts),) + tuple(out_shape)).copy()
        if val.shape == tuple(out_shape) + (len(pts),):
            return np.moveaxis(val, -1, 0).copy()
        raise ValueError

    try:
        full = batched(points)
        perm = np.random.default_rng(12345).permutation(n)
        if np.array_equal(batched(points[perm]), full[perm]):
            check = np.linspace(0, n - 1, min(n, 256)).astype(np.int64)
            if all(np.allclose(np.asarray(fn(points[i]), dtype=np.float64), full[i], rtol=1e-14, atol=0) for i in check):
                return full
    except Exception:
        pass
    return np.array([np.asarray(fn(p), dtype=np.float64) for p in points]).reshape((n,) + tuple(out_shape))


@dataclass
class FiniteElement:
    mesh: Mesh
    vec: int
    dim: int
    ele_type: str
    quadrature_rule: Any = None
    quadrature_order: Optional[int] = None
    dirichlet_bc_info: Optional[list] = None

    def __post_init__(self):
        self.points = self.mesh.points
        self.cells = self.mesh.cells
        self.num_cells = len(self.cells)
        self.num_total_nodes = len(self.mesh.points)
        self.num_total_dofs = self.num_total_nodes * self.vec
        start = time.time()
        logger.info("Computing shape function values, gradients, etc.")
        self.shape_vals, self.shape_grads_ref, self.quad_weights = get_shape_vals_and_grads(
            self.ele_type, quadrature_rule=self.quadrature_rule, quadrature_order=self.quadrature_order)
        (self.face_shape_vals, self.face_shape_grads_ref, self.face_quad_weights, self.face_normals,
         self.face_inds) = get_face_shape_vals_and_grads(
            self.ele_type, quadrature_rule=self.quadrature_rule, quadrature_order=self.quadrature_order)
        self.num_quads = self.shape_vals.shape[0]
        self.num_nodes = self.shape_vals.shape[1]
        self.num_faces = self.face_shape_vals.shape[0]
        self.num_face_quads = self.face_quad_weights.shape[1]
        assert self.cells.shape[1] == self.num_nodes, \
            f"{self.ele_type} needs {self.num_nodes} nodes per cell, mesh has {self.cells.shape[1]}"
        assert self.points.shape[1] == self.dim
        self.node_inds_list, self.vec_inds_list, self.vals_list = \
            self.Dirichlet_boundary_conditions(self.dirichlet_bc_info)
        logger.info(f"Solving a problem with {len(self.cells)} cells, {self.num_total_nodes}x{self.vec} = "
                    f"{self.num_total_dofs} dofs.")
        logger.info(f"Element type is {self.ele_type}, using {self.num_quads} quad points per element.")
        logger.info(f"Pre-computations took {time.time() - start:.3f} [s]")

    # ---- geometry helpers (post-processing; host NumPy) ------------------------------------------
    def _cell_coos(self, points, sel=None):
        points = self.points if points is None else np.asarray(points)
        cells = self.cells if sel is None else self.cells[sel]
        return points[cells]

    def get_shape_grads(self, points=None):
        """-> shape_grads_physical (C,Q,N,dim), JxW (C,Q)."""
        coos = self._cell_coos(points)
        J = np.einsum('cnd,qne->cqde', coos, self.shape_grads_ref)
        inv = np.linalg.inv(J)
        return np.einsum('qne,cqed->cqnd', self.shape_grads_ref, inv), np.linalg.det(J) * self.quad_weights[None, :]

    def get_JxW(self, points=None, chunk=1 << 17):
        """JxW (C,Q) alone, in cell chunks (the load-vector assembly needs it for every cell but must not materialise
        the (C,Q,N,dim) gradients)."""
        out = np.empty((self.num_cells, self.num_quads))
        pts = self.points if points is None else np.asarray(points)
        for s in range(0, self.num_cells, chunk):
            J = np.einsum('cnd,qne->cqde', pts[self.cells[s:s + chunk]], self.shape_grads_ref)
            out[s:s + chunk] = np.linalg.det(J) * self.quad_weights[None, :]
        return out

    def get_face_shape_grads(self, boundary_inds, points=None):
        """-> face_shape_grads_physical (S,FQ,N,dim), nanson_scale (S,FQ)."""
        boundary_inds = np.asarray(boundary_inds)
        coos = self._cell_coos(points, boundary_inds[:, 0])
        gref = self.face_shape_grads_ref[boundary_inds[:, 1]]
        J = np.einsum('fnd,fqne->fqde', coos, gref)
        inv = np.linalg.inv(J)
        grads = np.einsum('fqne,fqed->fqnd', gref, inv)
        scale = np.linalg.norm(np.einsum('fe,fqed->fqd', self.face_normals[boundary_inds[:, 1]], inv), axis=-1)
        return grads, scale * np.linalg.det(J) * self.face_quad_weights[boundary_inds[:, 1]]

    def get_physical_quad_points(self, points=None):
        return np.einsum('qn,cnd->cqd', self.shape_vals, self._cell_coos(points))

    def get_physical_surface_quad_points(self, boundary_inds, points=None):
        boundary_inds = np.asarray(boundary_inds)
        return np.einsum('fqn,fnd->fqd', self.face_shape_vals[boundary_inds[:, 1]],
                         self._cell_coos(points, boundary_inds[:, 0]))

    # ---- Dirichlet data ----------------------------------------------------------------------------
    def Dirichlet_boundary_conditions(self, dirichlet_bc_info):
        """[location_fns, vecs, value_fns] -> node_inds_list, vec_inds_list, vals_list (ascending node ids)."""
        node_inds_list, vec_inds_list, vals_list = [], [], []
        if dirichlet_bc_info is not None:
            location_fns, vecs, value_fns = dirichlet_bc_info
            assert len(location_fns) == len(value_fns) and len(value_fns) == len(vecs)
            for i in range(len(location_fns)):
                flags = evaluate_location_fn(location_fns[i], self.points)
                node_inds = np.flatnonzero(flags)
                node_inds_list.append(node_inds)
                vec_inds_list.append(np.full_like(node_inds, vecs[i], dtype=np.int32))
                vals_list.append(evaluate_point_fn(value_fns[i], self.points[node_inds]).reshape(-1))
        return node_inds_list, vec_inds_list, vals_list

    def update_Dirichlet_boundary_conditions(self, dirichlet_bc_info):
        self.node_inds_list, self.vec_inds_list, self.vals_list = self.Dirichlet_boundary_conditions(dirichlet_bc_info)

    def get_boundary_conditions_inds(self, location_fns):
        """Faces whose vertices ALL satisfy the predicate -> list of (num_selected_faces, 2) [cell, local face]."""
        out = []
        if location_fns is not None:
            cell_face_nodes = self.cells[:, self.face_inds]                     # (C,F,V)
            for fn in location_fns:
                flags = evaluate_location_fn(fn, self.points)
                out.append(np.argwhere(np.all(flags[cell_face_nodes], axis=-1)))
        return out

    # ---- interpolation helpers -----------------------------------------------------------------------
    def convert_from_dof_to_quad(self, sol):
        return np.einsum('cnv,qn->cqv', _host(sol)[self.cells], self.shape_vals)

    def convert_from_dof_to_face_quad(self, sol, boundary_inds):
        boundary_inds = np.asarray(boundary_inds)
        return np.einsum('fnv,fqn->fqv', _host(sol)[self.cells[boundary_inds[:, 0]]],
                         self.face_shape_vals[boundary_inds[:, 1]])

    def sol_to_grad(self, sol):
        grads, _ = self.get_shape_grads()
        return np.einsum('cnv,cqnd->cqvd', _host(sol)[self.cells], grads)


def _host(x):
    return x.detach().cpu().numpy() if hasattr(x, 'detach') else np.asarray(x)
