"""Mesh container and the structured generators whose node/cell numbering parity depends on.

Mirrors jax_fem/generate_mesh.py: Mesh (:13-55), get_meshio_cell_type (:83-117), rectangle_mesh
(:120-148), box_mesh (:151-189).  The gmsh-based generators (:192-361) are out of scope.  The
generators return a small meshio-like object (``.points``, ``.cells_dict``) so that reference user
code ``Mesh(m.points, m.cells_dict[cell_type])`` runs unchanged.
"""
import numpy as np

from . import basis as _basis

_MESHIO_NAME = {'HEX8': 'hexahedron', 'HEX27': 'hexahedron27', 'QUAD4': 'quad'}


class Mesh:
    """points (num_total_nodes, dim) float64, cells (num_cells, num_nodes) int32."""

    def __init__(self, points, cells, ele_type=None):
        self.points = np.ascontiguousarray(np.asarray(points, dtype=np.float64))
        self.cells = np.ascontiguousarray(np.asarray(cells, dtype=np.int32))
        self.ele_type = ele_type

    def count_selected_faces(self, location_fn):
        """generate_mesh.py:29-55."""
        from .fe import evaluate_location_fn
        *_, face_inds = _basis.get_face_shape_vals_and_grads(self.ele_type)
        flags = evaluate_location_fn(location_fn, self.points)
        return int(np.all(flags[self.cells[:, face_inds]], axis=-1).sum())


class StructuredMesh:
    """Stand-in for the meshio.Mesh the reference generators return."""

    def __init__(self, points, cell_type, cells):
        self.points = points
        self.cells_dict = {cell_type: cells}


def get_meshio_cell_type(ele_type):
    if ele_type not in _MESHIO_NAME:
        raise NotImplementedError(f"element type {ele_type!r} is not registered on the B200 hot path")
    return _MESHIO_NAME[ele_type]


def _lattice_points(counts, lengths):
    axes = [np.linspace(0, L, n + 1) for n, L in zip(counts, lengths)]
    grids = np.meshgrid(*axes, indexing='ij')
    pts = np.stack(grids, axis=len(counts)).reshape(-1, len(counts))
    ids = np.arange(len(pts), dtype=np.int64).reshape([n + 1 for n in counts])
    return pts, ids


def rectangle_mesh(Nx, Ny, domain_x, domain_y):
    """QUAD4 mesh, node id = i*(Ny+1)+j, cell corners counter-clockwise from (i,j)."""
    pts, ids = _lattice_points((Nx, Ny), (domain_x, domain_y))
    lo, hi = slice(None, -1), slice(1, None)
    corners = [(lo, lo), (hi, lo), (hi, hi), (lo, hi)]
    cells = np.stack([ids[c] for c in corners], axis=2).reshape(-1, 4).astype(np.int32)
    return StructuredMesh(pts, 'quad', cells)


def box_mesh(Nx, Ny, Nz, domain_x, domain_y, domain_z):
    """HEX8 mesh, node id = (i*(Ny+1)+j)*(Nz+1)+k, VTK corner order (bottom face ccw, then top)."""
    pts, ids = _lattice_points((Nx, Ny, Nz), (domain_x, domain_y, domain_z))
    lo, hi = slice(None, -1), slice(1, None)
    corners = [(lo, lo, lo), (hi, lo, lo), (hi, hi, lo), (lo, hi, lo),
               (lo, lo, hi), (hi, lo, hi), (hi, hi, hi), (lo, hi, hi)]
    cells = np.stack([ids[c] for c in corners], axis=3).reshape(-1, 8).astype(np.int32)
    return StructuredMesh(pts, 'hexahedron', cells)


def box_mesh_hex27(Nx, Ny, Nz, domain_x, domain_y, domain_z):
    """Second-order box mesh in VTK_TRIQUADRATIC_HEXAHEDRON order (the reference can only obtain
    HEX27 meshes through gmsh, generate_mesh.py:192-262; this structured generator is what
    BASELINE.json config 4 is run on).  Nodes live on the (2Nx+1)(2Ny+1)(2Nz+1) lattice."""
    pts, ids = _lattice_points((2 * Nx, 2 * Ny, 2 * Nz), (domain_x, domain_y, domain_z))
    lattice = _basis.get_elements('HEX27')[3]
    base = ids[:-1:2, :-1:2, :-1:2]
    sy, sx = 2 * Nz + 1, (2 * Ny + 1) * (2 * Nz + 1)
    cells = np.stack([base + a * sx + b * sy + c for (a, b, c) in lattice], axis=3).reshape(-1, 27).astype(np.int32)
    return StructuredMesh(pts, 'hexahedron27', cells)
