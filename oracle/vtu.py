"""Minimal stdlib reader for the ASCII .vtu goldens (TEST INFRASTRUCTURE ONLY).

The reference reads them with meshio after rewriting the header version
(jax_fem/utils.py:60-80, tests/benchmarks/*/test_*.py:29-34); the files are
plain ASCII XML so xml.etree is enough.  VTK_LAGRANGE_HEXAHEDRON (type 72)
at first order has VTK hexahedron node order == meshio 'hexahedron' order.
"""
import xml.etree.ElementTree as ET
import numpy as np

_DT = {"Float64": np.float64, "Float32": np.float32, "Int32": np.int32, "Int64": np.int64,
       "Int8": np.int8, "UInt8": np.uint8}


def read_vtu(path):
    root = ET.parse(path).getroot()
    piece = root.find(".//Piece")
    n_pts = int(piece.attrib["NumberOfPoints"])
    n_cells = int(piece.attrib["NumberOfCells"])

    def arr(da):
        return np.array(da.text.split(), dtype=_DT[da.attrib["type"]])

    points = arr(piece.find("Points/DataArray")).reshape(n_pts, 3)
    conn = offs = None
    for da in piece.find("Cells").iter("DataArray"):
        if da.attrib.get("Name") == "connectivity":
            conn = arr(da)
        elif da.attrib.get("Name") == "offsets":
            offs = arr(da)
    per = int(offs[0])
    assert np.all(np.diff(offs) == per) and len(offs) == n_cells
    cells = conn.reshape(n_cells, per)
    point_data = {}
    pd = piece.find("PointData")
    if pd is not None:
        for da in pd.iter("DataArray"):
            a = arr(da)
            nc = int(da.attrib.get("NumberOfComponents", "1"))
            point_data[da.attrib["Name"]] = a.reshape(n_pts, nc) if nc > 1 else a
    return points, cells, point_data
