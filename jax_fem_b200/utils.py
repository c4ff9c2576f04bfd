"""Solution output: ``save_sol`` of the reference (jax_fem/utils.py:13-57) without meshio.

The reference writes a .vtu through ``meshio.Mesh(...).write``; here the same content -- points, cells, the point field
``sol`` and optional cell / point fields, all down-cast to float32 exactly as the reference does -- is written as an
ASCII VTK XML UnstructuredGrid with the standard library only.  Host-side I/O after the solve: not part of the hot path
(SURVEY.md 8f row 3)."""
import os

import numpy as np

# VTK cell type ids and the meshio names the reference uses (jax_fem/generate_mesh.py:20-50)
_VTK_TYPE = {'HEX8': 12, 'HEX27': 29, 'QUAD4': 9}


def _to_numpy(a):
    if hasattr(a, 'detach'):
        a = a.detach().cpu().numpy()
    return np.asarray(a)


def _data_array(name, a, vtk_type, ncomp=None):
    a = np.ascontiguousarray(a)
    comp = f' NumberOfComponents="{ncomp}"' if ncomp and ncomp > 1 else ''
    fmt = ('%.17g' if a.dtype == np.float64 else '%.9g') if a.dtype.kind == 'f' else '%d'     # round-trip exact
    body = '\n'.join(' '.join(fmt % v for v in row) for row in a.reshape(len(a), -1)) if a.size else ''
    return f'<DataArray type="{vtk_type}" Name="{name}"{comp} format="ascii">\n{body}\n</DataArray>\n'


def save_sol(fe, sol, sol_file, cell_infos=None, point_infos=None):
    """Save a vertex-based solution (num_total_nodes, vec) and optional fields to ``sol_file`` (.vtu).

    cell_infos: [(name, data (num_cells,)), ...]; point_infos: [(name, data (num_total_nodes, ...)), ...] -- same
    contract and the same assertions as the reference (utils.py:44-55)."""
    if fe.ele_type not in _VTK_TYPE:
        raise NotImplementedError(f"save_sol: element type {fe.ele_type} is not registered")
    sol = _to_numpy(sol)
    sol_dir = os.path.dirname(sol_file)
    if sol_dir:
        os.makedirs(sol_dir, exist_ok=True)
    points = np.asarray(fe.points, dtype=np.float64)
    if points.shape[1] == 2:                                      # VTK points are 3-D
        points = np.concatenate([points, np.zeros((len(points), 1))], axis=1)
    cells = np.asarray(fe.cells, dtype=np.int64)
    n_pts, (n_cells, per) = len(points), cells.shape
    point_data = [('sol', sol.astype(np.float32).reshape(n_pts, -1))]
    for name, data in (point_infos or []):
        data = _to_numpy(data)
        assert len(data) == len(sol), "point data wrong shape!"
        point_data.append((name, data.astype(np.float32).reshape(n_pts, -1)))
    cell_data = []
    for name, data in (cell_infos or []):
        data = _to_numpy(data)
        assert data.shape == (fe.num_cells,), f"cell data wrong shape, get {data.shape}, while num_cells = {fe.num_cells}"
        cell_data.append((name, data.astype(np.float32).reshape(n_cells, 1)))
    with open(sol_file, 'w') as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian">\n')
        f.write(f'<UnstructuredGrid>\n<Piece NumberOfPoints="{n_pts}" NumberOfCells="{n_cells}">\n')
        f.write('<Points>\n' + _data_array('Points', points, 'Float64', 3) + '</Points>\n')
        f.write('<Cells>\n' + _data_array('connectivity', cells, 'Int64'))
        f.write(_data_array('offsets', per * np.arange(1, n_cells + 1, dtype=np.int64), 'Int64'))
        f.write(_data_array('types', np.full(n_cells, _VTK_TYPE[fe.ele_type], dtype=np.int64), 'Int64') + '</Cells>\n')
        f.write('<PointData>\n' + ''.join(_data_array(n, a, 'Float32', a.shape[1]) for n, a in point_data) + '</PointData>\n')
        if cell_data:
            f.write('<CellData>\n' + ''.join(_data_array(n, a, 'Float32') for n, a in cell_data) + '</CellData>\n')
        f.write('</Piece>\n</UnstructuredGrid>\n</VTKFile>\n')
