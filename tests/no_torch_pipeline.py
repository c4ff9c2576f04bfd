"""The C ABI on its own: plan -> element kernel -> residual gather -> CSR gather -> Jacobi-CG driven from cudaMalloc'd buffers
with ctypes only (no torch, no Python plan), checked against the NumPy oracle.  This is what a jax.ffi / C caller binds
(include/fem_b200.h); run by tests/test_gpu_parity.py::test_c_abi_pipeline_without_torch in a fresh interpreter."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fem, laws as olaws            # the checker (test infrastructure)

lib = ctypes.CDLL(os.path.join(ROOT, "jax_fem_b200", "lib", "libfem_b200.so"))
lib.fem_last_error.restype = ctypes.c_char_p
rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = ctypes.CDLL(name)
        break
    except OSError:
        pass
if rt is None:
    import glob
    rt = ctypes.CDLL(sorted(glob.glob("/usr/local/cuda*/lib64/libcudart.so*"))[0])
vp, i64, c_int, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double


def ck(code):
    assert code == 0, (code, lib.fem_last_error())


def dev(arr):
    arr = np.ascontiguousarray(arr)
    p = vp()
    assert rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(max(arr.nbytes, 16))) == 0
    assert rt.cudaMemcpy(p, arr.ctypes.data_as(vp), ctypes.c_size_t(arr.nbytes), c_int(1)) == 0
    return p


def dev_empty(nbytes):
    p = vp()
    assert rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(max(nbytes, 16))) == 0
    return p


def host(p, shape, dtype):
    out = np.empty(shape, dtype=dtype)
    assert rt.cudaDeviceSynchronize() == 0
    assert rt.cudaMemcpy(out.ctypes.data_as(vp), p, ctypes.c_size_t(out.nbytes), c_int(2)) == 0
    return out


# ---- the reference's linear-elasticity cube on a non-affine 7 x 5 x 4 box -----------------------------------------------------
om = fem.box_mesh(7, 5, 4, 1.4, 1.0, 0.8)
rng = np.random.default_rng(2)
pts = om.points + 0.02 * rng.uniform(-1, 1, om.points.shape)
cells = om.cells.astype(np.int32)
left = lambda p: p[0] < 0.03
bc = [[left] * 3, [0, 1, 2], [lambda p: 0., lambda p: 0.01, lambda p: -0.01]]
opb = fem.Problem(fem.Mesh(pts, cells), 3, 3, dirichlet_bc_info=bc, law=olaws.LinearElastic(70e3, 0.3))
sol = 0.01 * rng.standard_normal(pts.shape)
n_nodes, C, N, vec = len(pts), len(cells), 8, 3
n = n_nodes * vec

plan = vp()
ck(lib.fem_plan_create(dev(cells), i64(C), i64(n_nodes), c_int(N), c_int(vec), None, ctypes.byref(plan)))
sizes = (i64 * 8)()
ck(lib.fem_plan_sizes(plan, sizes))
nnzb, nnz, n_items, n_rows, n_src, n_dofs, row_block, n_row_blocks = list(sizes)
assert n_dofs == n and n_row_blocks == C * N


def table(which):
    p, cnt = vp(), i64()
    ck(lib.fem_plan_table(plan, c_int(which), ctypes.byref(p), ctypes.byref(cnt)))
    return p, cnt.value


BROW_PTR, BCOL, INDPTR, INDICES, CORNER_POS, NC_PTR, NC, GDESC, SRC = range(9)
opb.newton_update(sol)
oA_pattern = fem.get_A(opb)
indptr = host(table(INDPTR)[0], (n + 1,), np.int32)
indices = host(table(INDICES)[0], (nnz,), np.int32)
assert np.array_equal(indptr, oA_pattern.indptr) and np.array_equal(indices, oA_pattern.indices), "pattern must be bit-exact"

# Dirichlet mask -> emeta
flag = np.zeros(n, dtype=np.uint8)
rows = np.unique(np.concatenate(opb.bc_rows()))
flag[rows] = 1
emeta = dev_empty(16 * n_rows)
ck(lib.fem_plan_entry_meta(plan, dev(flag), emeta, None))

# element kernel -> residual gather -> CSR gather
ref = np.concatenate([opb.fe.shape_grads_ref.reshape(-1), opb.fe.quad_weights])
params = (dbl * 8)(70e3, 0.3, 0, 0, 0, 0, 0, 0)
Ke, Re = dev_empty(8 * n_row_blocks * row_block), dev_empty(8 * C * N * vec)
d_pts, d_cells, d_sol = dev(pts), dev(cells), dev(sol)
ck(lib.fem_element_residual_jacobian(c_int(0), c_int(vec), c_int(1), params, d_pts, d_cells, i64(C), d_sol, None, dev(ref),
                                     table(CORNER_POS)[0], Ke, Re, None))
res = dev_empty(8 * n)
ck(lib.fem_gather_residual(c_int(vec), c_int(N), i64(n_nodes), table(NC_PTR)[0], table(NC)[0], Re, None, res, None))
data = dev_empty(8 * nnz)
ck(lib.fem_gather_csr(c_int(vec), c_int(N), i64(n_items), table(GDESC)[0], emeta, table(SRC)[0], Ke, data, None))
vals = host(data, (nnz,), np.float64)
ores = opb.newton_update(sol)
oA = fem.get_A(opb)
assert np.abs(vals - oA.data).max() <= 1e-12 * np.abs(oA.data).max(), "CSR values differ from the oracle"
r = host(res, (n_nodes, vec), np.float64)

# apply_bc_vec + x0, then the whole Jacobi-CG solve in the library
bc_rows = np.asarray(rows, dtype=np.int32)
bc_vals = fem.assign_bc(np.zeros(n), opb)[bc_rows]          # merged with the reference's "later groups overwrite" rule
d_rows, d_vals = dev(bc_rows), dev(np.asarray(bc_vals, dtype=np.float64))
ck(lib.fem_apply_bc_vec(i64(len(bc_rows)), d_rows, d_vals, dbl(1.0), d_sol, res, None))
rb = host(res, (n,), np.float64)
assert np.abs(rb - fem.apply_bc_vec(ores.reshape(-1).copy(), sol.reshape(-1), opb)).max() <= 1e-12 * np.abs(rb).max()
b = dev(-rb)
x = dev_empty(8 * n)
ck(lib.fem_bc_initial_guess(i64(n), i64(len(bc_rows)), d_rows, d_vals, d_sol, x, None))
diag = dev_empty(8 * n)
ck(lib.fem_csr_diagonal(i64(n), table(INDPTR)[0], table(INDICES)[0], data, diag, None))
lib.fem_krylov_workspace.restype = i64
ws = dev_empty(8 * lib.fem_krylov_workspace(i64(n)))
info = (dbl * 4)()
ck(lib.fem_pcg(i64(n), table(INDPTR)[0], table(INDICES)[0], data, c_int(vec), table(BROW_PTR)[0], table(BCOL)[0], diag, b, x,
               dbl(1e-10), dbl(1e-10), c_int(10000), c_int(25), ws, info, None))
inc = host(x, (n,), np.float64)
import scipy.sparse.linalg as spla
want = spla.spsolve(oA.tocsc(), -rb)
assert np.abs(inc - want).max() <= 1e-8 * np.abs(want).max(), "CG solution differs from the direct solve of the oracle's system"
ck(lib.fem_plan_destroy(plan))
assert "torch" not in sys.modules, "this pipeline must not need torch"
print(f"C_ABI_PIPELINE_OK iterations={int(info[0])} err={info[2]:.2e} nnz={nnz}")
