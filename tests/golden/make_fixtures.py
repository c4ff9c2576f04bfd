"""Generate the golden fixtures under tests/golden/ from the reference's own test data.

Run ONCE in the build container (where /root/reference exists):

    python tests/golden/make_fixtures.py

Sources (public, ASCII VTU written by FEniCSx; see reference
tests/benchmarks/fenicsx_gold.py:27-66,123-183,265-358):
    tests/benchmarks/<case>/fenicsx/sol_p0_000000.vtu
    tests/benchmarks/hyperelasticity/fenicsx/traction.npy
    tests/benchmarks/linear_elasticity_cylinder/fenicsx/surface_area.npy

Each fixture is a compressed .npz holding the golden mesh (points, HEX8
connectivity in VTK order, usable as meshio 'hexahedron' directly) and the
golden nodal solution in float64, exactly as stored in the VTU.  Nothing at
test/bench time reads /root/reference.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.vtu import read_vtu  # noqa: E402

REF = os.environ.get("JAXFEM_REFERENCE", "/root/reference")
CASES = ["linear_poisson", "linear_elasticity_cube", "hyperelasticity", "linear_elasticity_cylinder"]


def main():
    for case in CASES:
        d = os.path.join(REF, "tests", "benchmarks", case, "fenicsx")
        points, cells, point_data = read_vtu(os.path.join(d, "sol_p0_000000.vtu"))
        out = dict(points=points, cells=cells.astype(np.int32), sol=point_data["sol"])
        if case == "hyperelasticity":
            out["traction"] = np.load(os.path.join(d, "traction.npy"))
        if case == "linear_elasticity_cylinder":
            out["surface_area"] = np.load(os.path.join(d, "surface_area.npy"))
        path = os.path.join(HERE, case + ".npz")
        np.savez_compressed(path, **out)
        print(case, {k: np.asarray(v).shape for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
