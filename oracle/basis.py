"""CPU ORACLE (test infrastructure only) -- reference-cell tabulation.

Restates jax_fem/basis.py:19-114 (get_elements), :141-175
(get_shape_vals_and_grads) and :178-250 (get_face_shape_vals_and_grads).

The reference delegates the arithmetic to the third-party package
fenics-basix==0.10.0 (environment.yml), which is NOT under /root/reference and
not installed here.  What is restated is basix's published conventions for the
calls made at basis.py:168,170-171,210,213-217,231,240,244-245:

* reference cells are [0,1]^d, vertices numbered x-fastest
  ((0,0,0),(1,0,0),(0,1,0),(1,1,0),(0,0,1),...);
* ``make_quadrature(cell, degree)`` (default rule) on interval/quadrilateral/
  hexahedron is Gauss-Jacobi == Gauss-Legendre with m=(degree+2)//2 points per
  direction, tensor order first-axis-slowest, points ascending on [0,1];
* Lagrange P1/P2 dofs: vertices, then edges, then faces, then interior, on the
  equispaced lattice; edges of the hexahedron (0,1),(0,2),(0,4),(1,3),(1,5),
  (2,3),(2,6),(3,7),(4,5),(4,6),(5,7),(6,7); facets (0,1,2,3),(0,1,4,5),
  (0,2,4,6),(1,3,5,7),(2,3,6,7),(4,5,6,7) with outward normals;
  quadrilateral facets (0,1),(0,2),(1,3),(2,3).

Pinned by: tests/test_oracle_golden.py (the four FEniCSx goldens of
tests/benchmarks/* are reproduced through these tables, including facets 3
and 5 via the traction / surface-area goldens).
"""
import itertools
import numpy as np

_HEX_EDGES = [(0, 1), (0, 2), (0, 4), (1, 3), (1, 5), (2, 3), (2, 6), (3, 7), (4, 5), (4, 6), (5, 7), (6, 7)]
_HEX_FACETS = [(0, 1, 2, 3), (0, 1, 4, 5), (0, 2, 4, 6), (1, 3, 5, 7), (2, 3, 6, 7), (4, 5, 6, 7)]
_HEX_NORMALS = np.array([[0, 0, -1], [0, -1, 0], [-1, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
_QUAD_FACETS = [(0, 1), (0, 2), (1, 3), (2, 3)]
_QUAD_NORMALS = np.array([[0, -1], [-1, 0], [1, 0], [0, 1]], dtype=np.float64)


def cell_vertices(dim):
    """basix.geometry(cell): vertices of [0,1]^dim, x fastest."""
    v = []
    for idx in itertools.product((0.0, 1.0), repeat=dim):
        v.append(idx[::-1])  # last axis of product varies fastest -> make x fastest
    return np.array(v, dtype=np.float64)


def get_elements(ele_type):
    """basis.py:19-114 restricted to the tensor-product cells on the hot path.

    Returns (cell_dim, default quadrature degree, Lagrange degree, re_order).
    """
    if ele_type == 'HEX8':
        return 3, 2, 1, [0, 1, 3, 2, 4, 5, 7, 6]                      # basis.py:52-57
    if ele_type == 'HEX27':
        return 3, 10, 2, [0, 1, 3, 2, 4, 5, 7, 6, 8, 11, 13, 9, 16, 18, 19,
                          17, 10, 12, 15, 14, 22, 23, 21, 24, 20, 25, 26]  # basis.py:58-65
    if ele_type == 'QUAD4':
        return 2, 2, 1, [0, 1, 3, 2]                                   # basis.py:86-91
    raise NotImplementedError(ele_type)


def gauss_legendre_01(m):
    x, w = np.polynomial.legendre.leggauss(m)
    return 0.5 * (x + 1.0), 0.5 * w


def make_quadrature(dim, degree):
    """basix.make_quadrature(cell, degree) for interval / quadrilateral / hexahedron."""
    m = (degree + 2) // 2
    x, w = gauss_legendre_01(m)
    if dim == 1:
        return x[:, None].copy(), w.copy()
    pts, wts = [], []
    for idx in itertools.product(range(m), repeat=dim):   # first axis slowest
        pts.append([x[i] for i in idx])
        wts.append(np.prod([w[i] for i in idx]))
    return np.array(pts), np.array(wts)


def _lagrange_nodes(dim, degree):
    """Node coordinates of the basix Lagrange element in basix dof order."""
    verts = cell_vertices(dim)
    if degree == 1:
        return verts
    assert degree == 2
    nodes = [v for v in verts]
    if dim == 2:
        for (a, b) in _QUAD_FACETS:
            nodes.append(0.5 * (verts[a] + verts[b]))
        nodes.append(verts.mean(axis=0))
    else:
        for (a, b) in _HEX_EDGES:
            nodes.append(0.5 * (verts[a] + verts[b]))
        for f in _HEX_FACETS:
            nodes.append(verts[list(f)].mean(axis=0))
        nodes.append(verts.mean(axis=0))
    return np.array(nodes)


def _lagrange_1d(degree, k, x):
    """Value and derivative at x of the 1-D Lagrange polynomial attached to lattice node k/degree."""
    nodes = np.linspace(0.0, 1.0, degree + 1)
    val = np.ones_like(x)
    for j in range(degree + 1):
        if j != k:
            val = val * (x - nodes[j]) / (nodes[k] - nodes[j])
    der = np.zeros_like(x)
    for i in range(degree + 1):
        if i == k:
            continue
        term = np.ones_like(x) / (nodes[k] - nodes[i])
        for j in range(degree + 1):
            if j != k and j != i:
                term = term * (x - nodes[j]) / (nodes[k] - nodes[j])
        der = der + term
    return val, der


def tabulate(dim, degree, pts):
    """element.tabulate(1, pts) of the Lagrange element, basix dof order.

    Returns vals (P, N) and grads (P, N, dim).
    """
    nodes = _lagrange_nodes(dim, degree)
    P, N = len(pts), len(nodes)
    vals = np.ones((P, N))
    grads = np.ones((P, N, dim))
    for n in range(N):
        one_d = []
        for d in range(dim):
            k = int(round(nodes[n, d] * degree))
            one_d.append(_lagrange_1d(degree, k, pts[:, d]))
        for d in range(dim):
            vals[:, n] *= one_d[d][0]
            for e in range(dim):
                grads[:, n, e] *= one_d[d][1] if d == e else one_d[d][0]
    return vals, grads


def get_shape_vals_and_grads(ele_type, quadrature_order=None):
    """basis.py:141-175 -> shape_values (Q,N), shape_grads_ref (Q,N,dim), weights (Q,)."""
    dim, q_default, degree, re_order = get_elements(ele_type)
    if quadrature_order is None:
        quadrature_order = q_default
    pts, w = make_quadrature(dim, quadrature_order)
    vals, grads = tabulate(dim, degree, pts)
    return vals[:, re_order], grads[:, re_order, :], w


def get_face_shape_vals_and_grads(ele_type, quadrature_order=None):
    """basis.py:178-250 -> face_shape_vals (F,FQ,N), face_shape_grads_ref (F,FQ,N,dim),
    face_weights (F,FQ), face_normals (F,dim), face_inds (F, n_face_vertices)."""
    dim, q_default, degree, re_order = get_elements(ele_type)
    if quadrature_order is None:
        quadrature_order = q_default
    pts, w = make_quadrature(dim - 1, quadrature_order)
    map_vals, _ = tabulate(dim - 1, 1, pts)                 # basis.py:212-214
    verts = cell_vertices(dim)
    facets = _HEX_FACETS if dim == 3 else _QUAD_FACETS
    normals = _HEX_NORMALS if dim == 3 else _QUAD_NORMALS
    face_pts, face_w, face_inds = [], [], []
    re = np.array(re_order)
    for f in facets:
        fv = verts[list(f)]
        face_pts.append(map_vals @ fv)                      # basis.py:224-229
        if dim == 2:                                        # basis.py:231-236
            size = np.linalg.norm(fv[1] - fv[0])
        else:
            size = np.linalg.norm(np.cross(fv[1] - fv[0], fv[2] - fv[0]))
        face_w.append(w * size)
        face_inds.append([int(np.argwhere(re == i)[0, 0]) for i in f])   # basis.py:117-124,242
    face_pts = np.stack(face_pts)
    F, FQ, _ = face_pts.shape
    vals, grads = tabulate(dim, degree, face_pts.reshape(-1, dim))
    vals = vals[:, re_order].reshape(F, FQ, -1)
    grads = grads[:, re_order, :].reshape(F, FQ, -1, dim)
    return vals, grads, np.stack(face_w), normals.copy(), np.array(face_inds)
