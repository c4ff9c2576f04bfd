"""B200-native drop-in for JAX-FEM's data-parallel hot path.

Public surface mirrors the reference (docs/source/more/api/api_problem.rst, api_solver.rst,
api_fe.rst): Problem, solver, ad_wrapper, FiniteElement, Mesh, box_mesh, rectangle_mesh.
All hot-path arithmetic runs in libfem_b200.so (hand-written sm_100a CUDA, float64); there is no
CPU fallback and unregistered constitutive laws raise.
"""
import os as _os
from .logger_setup import setup_logger as _setup_logger

logger = _setup_logger(level=int(_os.environ.get("JAX_FEM_B200_LOGLEVEL", "30")))

from . import laws                                              # noqa: E402
from .generate_mesh import Mesh, box_mesh, box_mesh_hex27, rectangle_mesh, get_meshio_cell_type  # noqa: E402
from .fe import FiniteElement                                   # noqa: E402
from .problem import Problem                                    # noqa: E402
from .solver import solver, ad_wrapper, get_A, apply_bc_vec, linear_solver   # noqa: E402
from .utils import save_sol                                     # noqa: E402
from .mesh_io import read_mesh                                  # noqa: E402

__all__ = ["Problem", "solver", "ad_wrapper", "get_A", "apply_bc_vec", "linear_solver", "FiniteElement", "Mesh",
           "box_mesh", "box_mesh_hex27", "rectangle_mesh", "get_meshio_cell_type", "laws", "logger", "save_sol", "read_mesh"]
