"""Mesh ingestion without meshio / gmsh: the on-disk formats either side of the hot path.

The reference generates meshes with gmsh, writes them as MSH 2.2 ("Mesh.MshFileVersion", 2.2) and reads them back with
``meshio.read`` (jax_fem/generate_mesh.py:238-260, 353-361); users also bring Abaqus .inp files and .vtu results
(jax_fem/utils.py:60-80 reads .vtu through meshio).  Neither package exists in this image, so the three ASCII formats are
parsed here with the standard library.  ``read_mesh`` returns the same meshio-like object as the generators of
generate_mesh.py (``.points``, ``.cells_dict`` keyed by meshio cell names, ``.point_data``), so reference code such as
``Mesh(m.points, m.cells_dict['hexahedron'])`` runs unchanged.

First-order cells (hexahedron / quad / tetra / triangle: their node order is the same in Gmsh, Abaqus, VTK and meshio) and
the 27-node hexahedron of the path's HEX27 element: VTK type 29 is already in the meshio / VTK order the kernels use, Gmsh
type 12 numbers its edge and face nodes differently and is permuted (the permutation is derived from the two documented
edge / face tables below, not typed in).  Anything else raises -- a silently permuted higher-order cell would corrupt every
element matrix.
"""
import os
import xml.etree.ElementTree as ET

import numpy as np


class MeshFile:
    def __init__(self, points, cells_dict, point_data=None):
        self.points = points
        self.cells_dict = cells_dict
        self.point_data = point_data or {}


# nodes per cell of the first-order types, by format-specific tag
_GMSH = {1: ('line', 2), 2: ('triangle', 3), 3: ('quad', 4), 4: ('tetra', 4), 5: ('hexahedron', 8), 15: ('vertex', 1)}
_VTK = {3: ('line', 2), 5: ('triangle', 3), 9: ('quad', 4), 10: ('tetra', 4), 12: ('hexahedron', 8), 72: ('hexahedron', 8)}
_ABAQUS = {'C3D8': 'hexahedron', 'C3D8R': 'hexahedron', 'C3D4': 'tetra', 'CPS4': 'quad', 'CPE4': 'quad', 'CPS4R': 'quad',
           'CPE4R': 'quad', 'S4': 'quad', 'S4R': 'quad', 'CPS3': 'triangle', 'CPE3': 'triangle', 'S3': 'triangle'}
_NODES = {'hexahedron': 8, 'tetra': 4, 'quad': 4, 'triangle': 3, 'line': 2, 'vertex': 1, 'hexahedron27': 27}
_GMSH[12] = ('hexahedron27', 27)
_VTK[29] = ('hexahedron27', 27)


def _hex27_gmsh_to_vtk():
    """perm such that vtk_cell = gmsh_cell[perm].  Both formats number the 8 vertices alike, then the 12 edge midpoints, the 6
    face centres and the cell centre; only the order of the edges and faces differs (Gmsh reference manual, "Node ordering";
    VTK_TRIQUADRATIC_HEXAHEDRON)."""
    corner = [(0, 0, 0), (2, 0, 0), (2, 2, 0), (0, 2, 0), (0, 0, 2), (2, 0, 2), (2, 2, 2), (0, 2, 2)]
    centre = lambda vs: tuple(sum(corner[v][d] for v in vs) // len(vs) for d in range(3))
    gmsh_edges = [(0, 1), (0, 3), (0, 4), (1, 2), (1, 5), (2, 3), (2, 6), (3, 7), (4, 5), (4, 7), (5, 6), (6, 7)]
    gmsh_faces = [(0, 3, 2, 1), (0, 1, 5, 4), (0, 4, 7, 3), (1, 2, 6, 5), (2, 3, 7, 6), (4, 5, 6, 7)]
    vtk_edges = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
    vtk_faces = [(0, 4, 7, 3), (1, 2, 6, 5), (0, 1, 5, 4), (2, 3, 7, 6), (0, 1, 2, 3), (4, 5, 6, 7)]
    where = lambda edges, faces: corner + [centre(e) for e in edges] + [centre(f) for f in faces] + [(1, 1, 1)]
    gmsh = where(gmsh_edges, gmsh_faces)
    return np.array([gmsh.index(p) for p in where(vtk_edges, vtk_faces)])


_HEX27_GMSH_TO_VTK = _hex27_gmsh_to_vtk()


def _stack(groups):
    return {name: np.asarray(rows, dtype=np.int64) for name, rows in groups.items() if rows}


def read_msh(path):
    """Gmsh MSH 2.2 ASCII (what the reference's gmsh generators write)."""
    lines = open(path).read().split('\n')
    points, ids, groups = None, None, {}
    i = 0
    while i < len(lines):
        tag = lines[i].strip()
        if tag == '$MeshFormat':
            version, file_type = lines[i + 1].split()[:2]
            if not version.startswith('2') or file_type != '0':
                raise NotImplementedError(f"{path}: only MSH 2.x ASCII is supported (got version {version}, type {file_type})")
            i += 3
        elif tag == '$Nodes':
            n = int(lines[i + 1])
            block = np.array([l.split() for l in lines[i + 2:i + 2 + n]], dtype=np.float64)
            ids, points = block[:, 0].astype(np.int64), block[:, 1:4]
            i += n + 3
        elif tag == '$Elements':
            n = int(lines[i + 1])
            for l in lines[i + 2:i + 2 + n]:
                f = l.split()
                etype, ntags = int(f[1]), int(f[2])
                if etype not in _GMSH:
                    raise NotImplementedError(f"{path}: Gmsh element type {etype} is not registered (first-order cells and the "
                                              "27-node hexahedron are); other higher-order node orders are not converted here")
                name, per = _GMSH[etype]
                groups.setdefault(name, []).append([int(v) for v in f[3 + ntags:3 + ntags + per]])
            i += n + 3
        else:
            i += 1
    if points is None:
        raise ValueError(f"{path}: no $Nodes section")
    lookup = np.full(int(ids.max()) + 1, -1, dtype=np.int64)
    lookup[ids] = np.arange(len(ids))
    cells = {k: lookup[v] for k, v in _stack(groups).items()}
    if 'hexahedron27' in cells:
        cells['hexahedron27'] = cells['hexahedron27'][:, _HEX27_GMSH_TO_VTK]
    return MeshFile(points, cells)


def read_inp(path):
    """Abaqus .inp: *NODE and *ELEMENT, TYPE=... blocks (everything else is skipped)."""
    ids, pts, groups, mode, etype = [], [], {}, None, None
    for raw in open(path):
        line = raw.strip()
        if not line or line.startswith('**'):
            continue
        if line.startswith('*'):
            key = line.split(',')[0].strip().upper()
            mode = None
            if key == '*NODE':
                mode = 'node'
            elif key == '*ELEMENT':
                opts = {k.strip().upper(): v.strip().upper() for k, v in (o.split('=') for o in line.split(',')[1:] if '=' in o)}
                etype = opts.get('TYPE')
                if etype not in _ABAQUS:
                    raise NotImplementedError(f"{path}: Abaqus element type {etype} is not a registered first-order cell")
                mode = 'element'
            continue
        f = [v for v in line.replace(',', ' ').split()]
        if mode == 'node':
            ids.append(int(f[0]))
            pts.append([float(v) for v in f[1:4]] + [0.0] * (4 - len(f)))
        elif mode == 'element':
            name = _ABAQUS[etype]
            groups.setdefault(name, []).append([int(v) for v in f[1:1 + _NODES[name]]])
    if not ids:
        raise ValueError(f"{path}: no *NODE block")
    ids = np.asarray(ids, dtype=np.int64)
    lookup = np.full(int(ids.max()) + 1, -1, dtype=np.int64)
    lookup[ids] = np.arange(len(ids))
    return MeshFile(np.asarray(pts, dtype=np.float64), {k: lookup[v] for k, v in _stack(groups).items()})


def read_vtu(path):
    """ASCII VTK XML UnstructuredGrid (what save_sol writes; the reference's FEniCSx goldens use the same layout)."""
    dtypes = {"Float64": np.float64, "Float32": np.float32, "Int32": np.int32, "Int64": np.int64, "Int8": np.int8, "UInt8": np.uint8}
    piece = ET.parse(path).getroot().find(".//Piece")
    n_pts, n_cells = int(piece.attrib["NumberOfPoints"]), int(piece.attrib["NumberOfCells"])

    def arr(da):
        if da.attrib.get("format", "ascii") != "ascii":
            raise NotImplementedError(f"{path}: only ASCII data arrays are supported")
        return np.array(da.text.split(), dtype=dtypes[da.attrib["type"]])

    points = arr(piece.find("Points/DataArray")).reshape(n_pts, 3)
    named = {da.attrib.get("Name"): arr(da) for da in piece.find("Cells").iter("DataArray")}
    conn, offs, types = named["connectivity"], named["offsets"], named["types"]
    starts = np.concatenate([[0], offs[:-1]])
    groups = {}
    for t in np.unique(types):
        if int(t) not in _VTK:
            raise NotImplementedError(f"{path}: VTK cell type {int(t)} is not registered (first-order cells and type 29)")
        name, per = _VTK[int(t)]
        sel = np.flatnonzero(types == t)
        if not np.all(offs[sel] - starts[sel] == per):
            raise NotImplementedError(f"{path}: VTK cell type {int(t)} with {int((offs[sel] - starts[sel])[0])} nodes (higher order)")
        groups[name] = conn[starts[sel][:, None] + np.arange(per)[None, :]].astype(np.int64)
    point_data = {}
    pd = piece.find("PointData")
    if pd is not None:
        for da in pd.iter("DataArray"):
            a, nc = arr(da), int(da.attrib.get("NumberOfComponents", "1"))
            point_data[da.attrib["Name"]] = a.reshape(n_pts, nc) if nc > 1 else a
    return MeshFile(points, groups, point_data)


_READERS = {'.msh': read_msh, '.inp': read_inp, '.vtu': read_vtu}


def read_mesh(path):
    """meshio.read for the three ASCII formats on the path's boundary: Gmsh MSH 2.2, Abaqus .inp, VTK .vtu."""
    ext = os.path.splitext(path)[1].lower()
    if ext not in _READERS:
        raise NotImplementedError(f"no reader for {ext!r} files (registered: {sorted(_READERS)})")
    return _READERS[ext](path)
