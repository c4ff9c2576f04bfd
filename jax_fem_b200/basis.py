"""Reference-cell tables for the tensor-product elements on the hot path (host side, NumPy).

Mirror of the interface of jax_fem/basis.py (get_elements :19-114, get_shape_vals_and_grads
:141-175, get_face_shape_vals_and_grads :178-250).  The reference obtains these numbers from
fenics-basix; here they are closed-form tensor products of 1-D Lagrange polynomials on the
lattice {0, 1/p, ..., 1} evaluated at Gauss-Legendre points, emitted directly in meshio/VTK node
order (so no ``re_order`` gather is needed afterwards).  tests/test_host_logic.py checks them
entry by entry against the oracle's basix restatement.
"""
import numpy as np

# lattice coordinates (in units of 1/degree) of every node, in meshio / VTK order
_HEX8 = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
_QUAD4 = [(0, 0), (1, 0), (1, 1), (0, 1)]


def _hex27_lattice():
    c = [(2 * i, 2 * j, 2 * k) for (i, j, k) in _HEX8]
    mid = lambda a, b: tuple((x + y) // 2 for x, y in zip(c[a], c[b]))
    edges = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
    faces = [(0, 7), (1, 6), (0, 5), (3, 6), (0, 2), (4, 6)]       # diagonals of x-,x+,y-,y+,z-,z+ faces
    return c + [mid(a, b) for a, b in edges] + [mid(a, b) for a, b in faces] + [(1, 1, 1)]


_LATTICE = {'HEX8': (3, 1, _HEX8), 'QUAD4': (2, 1, _QUAD4), 'HEX27': (3, 2, _hex27_lattice())}
_DEFAULT_QUAD_DEGREE = {'HEX8': 2, 'QUAD4': 2, 'HEX27': 10}     # basis.py:56,90,64

# facets as meshio-local vertex ids, in basix facet order, with outward reference normals
_FACETS = {
    3: ([(0, 1, 3, 2), (0, 1, 4, 5), (0, 3, 4, 7), (1, 2, 5, 6), (3, 2, 7, 6), (4, 5, 7, 6)],
        [(0, 0, -1), (0, -1, 0), (-1, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]),
    2: ([(0, 1), (0, 3), (1, 2), (3, 2)], [(0, -1), (-1, 0), (1, 0), (0, 1)]),
}


def get_elements(ele_type):
    """(dim, default quadrature degree, Lagrange degree, node lattice).  Unknown types are not registered."""
    if ele_type not in _LATTICE:
        raise NotImplementedError(f"element type {ele_type!r} is not registered on the B200 hot path "
                                  f"(registered: {sorted(_LATTICE)})")
    dim, degree, lattice = _LATTICE[ele_type]
    return dim, _DEFAULT_QUAD_DEGREE[ele_type], degree, np.array(lattice)


def gauss_points(dim, degree):
    """Gauss-Legendre rule of basix.make_quadrature(cell, degree): m = (degree+2)//2 points per axis,
    first axis slowest, on [0,1]^dim."""
    m = (degree + 2) // 2
    x, w = np.polynomial.legendre.leggauss(m)
    x, w = 0.5 * (x + 1.0), 0.5 * w
    grids = np.meshgrid(*([x] * dim), indexing='ij')
    wgrids = np.meshgrid(*([w] * dim), indexing='ij')
    pts = np.stack([g.reshape(-1) for g in grids], axis=1)
    wts = np.prod(np.stack([g.reshape(-1) for g in wgrids], axis=1), axis=1)
    return pts, wts


def _lagrange_1d(degree, x):
    """(vals, ders), each (len(x), degree+1)."""
    x = np.asarray(x, dtype=np.float64)
    if degree == 1:
        return np.stack([1 - x, x], 1), np.stack([-np.ones_like(x), np.ones_like(x)], 1)
    if degree == 2:
        return (np.stack([(2 * x - 1) * (x - 1), 4 * x * (1 - x), x * (2 * x - 1)], 1),
                np.stack([4 * x - 3, 4 - 8 * x, 4 * x - 1], 1))
    raise NotImplementedError(degree)


def tabulate(ele_type, pts):
    """Shape values (P,N) and reference gradients (P,N,dim) at points (P,dim), meshio node order."""
    dim, _, degree, lattice = get_elements(ele_type)
    one_d = [_lagrange_1d(degree, pts[:, d]) for d in range(dim)]
    vals = np.ones((len(pts), len(lattice)))
    grads = np.ones((len(pts), len(lattice), dim))
    for d in range(dim):
        v, dv = one_d[d][0][:, lattice[:, d]], one_d[d][1][:, lattice[:, d]]
        vals *= v
        for e in range(dim):
            grads[:, :, e] *= dv if e == d else v
    return vals, grads


def get_shape_vals_and_grads(ele_type, quadrature_rule=None, quadrature_order=None):
    """-> shape_values (Q,N), shape_grads_ref (Q,N,dim), weights (Q,)   [basis.py:141-175]"""
    if quadrature_rule is not None:
        raise NotImplementedError("only the default (Gauss-Jacobi) quadrature rule is registered")
    dim, q_default, _, _ = get_elements(ele_type)
    pts, w = gauss_points(dim, q_default if quadrature_order is None else quadrature_order)
    vals, grads = tabulate(ele_type, pts)
    return vals, grads, w


def get_face_shape_vals_and_grads(ele_type, quadrature_rule=None, quadrature_order=None):
    """-> face_shape_vals (F,FQ,N), face_shape_grads_ref (F,FQ,N,dim), face_weights (F,FQ),
    face_normals (F,dim), face_inds (F,V)   [basis.py:178-250]"""
    if quadrature_rule is not None:
        raise NotImplementedError("only the default (Gauss-Jacobi) quadrature rule is registered")
    dim, q_default, degree, lattice = get_elements(ele_type)
    fp, fw = gauss_points(dim - 1, q_default if quadrature_order is None else quadrature_order)
    facets, normals = _FACETS[dim]
    corners = lattice[:2 ** dim] / float(degree)                    # meshio vertex coordinates
    vals, grads, weights = [], [], []
    for f in facets:
        v = corners[list(f)]
        if dim == 3:                                                # affine map of the facet's (s,t) square
            pts = v[0] + fp[:, :1] * (v[1] - v[0]) + fp[:, 1:2] * (v[2] - v[0])
            size = np.linalg.norm(np.cross(v[1] - v[0], v[2] - v[0]))
        else:
            pts = v[0] + fp[:, :1] * (v[1] - v[0])
            size = np.linalg.norm(v[1] - v[0])
        a, b = tabulate(ele_type, pts)
        vals.append(a)
        grads.append(b)
        weights.append(fw * size)
    return (np.stack(vals), np.stack(grads), np.stack(weights),
            np.array(normals, dtype=np.float64), np.array(facets))
