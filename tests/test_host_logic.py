"""CPU tests of the host side: reference tables, meshes, plan (pattern bit-exact vs the oracle), Dirichlet /
face index sets, load vectors, law registry, and that the C-ABI library loads and exports every declared
symbol.  No compute call is made (there is no GPU here and the product has no CPU fallback)."""
import numpy as np
import pytest
import torch

import cases
from oracle import basis as obasis, fem, laws as olaws
import jax_fem_b200 as jf
from jax_fem_b200 import _lib, basis, laws
from jax_fem_b200.fe import FiniteElement, evaluate_location_fn, evaluate_point_fn
from jax_fem_b200.plan import build_plan


@pytest.mark.parametrize("ele", ["HEX8", "QUAD4", "HEX27"])
def test_reference_tables_match_oracle(ele):
    for a, b in zip(obasis.get_shape_vals_and_grads(ele), basis.get_shape_vals_and_grads(ele)):
        assert np.abs(a - b).max() < 1e-14
    for a, b in zip(obasis.get_face_shape_vals_and_grads(ele), basis.get_face_shape_vals_and_grads(ele)):
        assert np.abs(a - b).max() < 1e-14


def test_mesh_generators_match_oracle_numbering():
    o, p = fem.box_mesh(4, 3, 2, 1., 2., 3.), jf.box_mesh(4, 3, 2, 1., 2., 3.)
    assert np.array_equal(o.points, p.points) and np.array_equal(o.cells, p.cells_dict['hexahedron'])
    o, p = fem.rectangle_mesh(5, 3, 1., 2.), jf.rectangle_mesh(5, 3, 1., 2.)
    assert np.array_equal(o.points, p.points) and np.array_equal(o.cells, p.cells_dict['quad'])
    h = jf.box_mesh_hex27(2, 3, 2, 1., 1., 1.)
    cells, pts = h.cells_dict['hexahedron27'], h.points
    assert cells.shape == (12, 27) and len(pts) == 5 * 7 * 5 and len(np.unique(cells)) == len(pts)
    # every cell's nodes must sit on the VTK triquadratic lattice of its own corner box
    lattice = basis.get_elements('HEX27')[3] / 2.0
    for c in cells[[0, 5, 11]]:
        x = pts[c]
        lo, hi = x[0], x[6]
        assert np.allclose(x, lo + lattice * (hi - lo))


@pytest.mark.parametrize("N,vec", [(3, 3), (6, 3), (5, 1)])
def test_plan_pattern_bit_exact(N, vec):
    m = jf.box_mesh(N, N + 1, N, 1, 1, 1)
    cells = m.cells_dict['hexahedron']
    plan = build_plan(torch.from_numpy(cells), len(m.points), vec)
    indptr, indices = fem.csr_pattern_from_cells(cells, vec, vec * len(m.points))
    assert plan.indptr.dtype == torch.int32 and plan.indices.dtype == torch.int32
    assert np.array_equal(plan.indptr.numpy(), indptr) and np.array_equal(plan.indices.numpy(), indices)
    # source lists: every element block (corner position, b) appears exactly once and lands in the right entry
    src = plan.src.numpy().astype(np.int64)
    assert np.array_equal(np.sort(src), np.arange(cells.shape[0] * 64))
    cpos = plan.corner_pos.numpy()
    assert np.array_equal(np.sort(cpos), np.arange(cells.size))
    corner_of_pos = np.argsort(cpos)                       # position -> c*8 + a
    corner, b = corner_of_pos[src // 8], src % 8
    c, a = corner // 8, corner % 8
    counts = np.diff(plan.src_ptr.numpy())
    brow = np.repeat(np.arange(len(m.points)), np.diff(plan.brow_ptr.numpy()))
    assert np.array_equal(np.repeat(brow, counts), cells[c, a])
    assert np.array_equal(np.repeat(plan.bcol.numpy(), counts), cells[c, b])
    code = (c * 8 + a) * 8 + b                             # fixed summation order: ascending (c, a, b) inside an entry
    seg = np.repeat(np.arange(len(counts)), counts)
    assert np.all((np.diff(code) > 0) | (np.diff(seg) > 0))
    assert np.array_equal(cells.reshape(-1)[corner_of_pos], np.sort(cells.reshape(-1)))   # node-sorted corners
    # gather work split: CTA b owns the nodes whose first corner lies in [32 b, 32 (b+1))
    gd = plan.gdesc.numpy().reshape(-1, 4)
    ncp, sp, bp = plan.nc_ptr.numpy(), plan.src_ptr.numpy(), plan.brow_ptr.numpy()
    assert len(gd) == plan.n_gather_blocks + 1 and gd[0, 0] == 0 and gd[-1, 0] == cells.size
    assert gd[-1, 1] == plan.nnzb and gd[-1, 2] == plan.n_items
    node0 = np.searchsorted(ncp[:-1], 32 * np.arange(len(gd)))
    assert np.array_equal(gd[:, 0], ncp[node0]) and np.array_equal(gd[:, 1], bp[node0])
    assert np.array_equal(gd[:, 2], sp[bp[node0]])
    assert np.all(np.diff(gd[:, 0]) <= 64) and np.all(np.diff(gd[:, 0]) >= 0)
    ed, erow = plan.edst.numpy(), plan.erow.numpy()
    assert np.array_equal(erow, brow)
    assert np.array_equal(plan.indices.numpy()[ed], vec * plan.bcol.numpy())
    assert np.all((plan.indptr.numpy()[vec * erow] <= ed) & (ed < plan.indptr.numpy()[vec * erow + 1]))
    flag = torch.zeros(plan.n, dtype=torch.uint8)
    flag[vec * 5 + (vec - 1)] = 1
    info = plan.entry_info(flag).numpy()
    assert np.array_equal(info & 0xffff, vec * np.diff(plan.brow_ptr.numpy())[erow])
    assert np.array_equal((info >> 16) & 1, (plan.bcol.numpy() == erow).astype(np.int32))
    assert np.array_equal(info >> 17, np.where(erow == 5, 1 << (vec - 1), 0))
    # gather rows: every entry's source range is covered once, in order, by one row or by a (store, add) pair of halves
    em = plan.entry_meta(flag).numpy()
    assert em.shape[0] == gd[-1, 3] and gd[0, 3] == 0 and np.all(np.diff(gd[:, 3]) >= np.diff(gd[:, 1]))
    ment, madd = plan.m_ent.numpy(), plan.m_add.numpy()
    spn = plan.src_ptr.numpy()
    for e in (0, 5, len(spn) // 2, len(spn) - 2):
        rows = np.flatnonzero(ment == e)
        rows = rows[np.argsort(madd[rows])]
        assert em[rows[0], 0] == spn[e] and em[rows[-1], 1] == spn[e + 1] and list(madd[rows]) == list(range(len(rows)))
        assert len(rows) == (2 if spn[e + 1] - spn[e] > 4 else 1)
        if len(rows) == 2:
            assert em[rows[0], 1] == em[rows[1], 0] and (em[rows[1], 3] >> 20) & 1 and not (em[rows[0], 3] >> 20) & 1
        assert np.all(em[rows, 2] == ed[e]) and np.all((em[rows, 3] & 0xfffff) == info[e])
    cta_rows = np.searchsorted(gd[1:, 3], np.arange(len(em)), side='right')
    assert np.array_equal(cta_rows, np.searchsorted(gd[1:, 1], ment, side='right'))       # rows stay inside their CTA
    # transpose permutation is an involution mapping (n,m) -> (m,n)
    t = plan.tperm.numpy()
    assert np.array_equal(t[t], np.arange(len(t)))
    assert np.array_equal(plan.bcol.numpy()[t], brow)
    nc = plan.nc.numpy()
    assert np.array_equal(np.repeat(np.arange(len(m.points)), np.diff(plan.nc_ptr.numpy())), cells.reshape(-1)[nc])


def test_plan_on_unstructured_golden_mesh():
    g = cases.load_golden("hyperelasticity")
    plan = build_plan(torch.from_numpy(g["cells"]), len(g["points"]), 3)
    indptr, indices = fem.csr_pattern_from_cells(g["cells"], 3, 3 * len(g["points"]))
    assert np.array_equal(plan.indptr.numpy(), indptr) and np.array_equal(plan.indices.numpy(), indices)


def test_dirichlet_and_face_sets_match_oracle():
    g = cases.load_golden("linear_elasticity_cylinder")
    mesh = jf.Mesh(g["points"], g["cells"])
    fe = FiniteElement(mesh, 3, 3, 'HEX8', dirichlet_bc_info=cases.CYL_BC)
    ofe = fem.FiniteElement(fem.Mesh(g["points"], g["cells"]), 3, 3, 'HEX8', dirichlet_bc_info=cases.CYL_BC)
    for a, b in zip(fe.node_inds_list, ofe.node_inds_list):
        assert np.array_equal(a, b)
    for a, b in zip(fe.vals_list, ofe.vals_list):
        assert np.array_equal(a, b)
    b1 = fe.get_boundary_conditions_inds([cases.top])[0]
    b2 = ofe.get_boundary_conditions_inds([cases.top])[0]
    assert np.array_equal(b1, b2)
    g1, n1 = fe.get_face_shape_grads(b1)
    g2, n2 = ofe.get_face_shape_grads(b2)
    assert np.abs(g1 - g2).max() < 1e-13 and np.abs(n1 - n2).max() < 1e-14
    np.testing.assert_almost_equal(n1.sum(), float(g["surface_area"]), decimal=10)
    s1, j1 = fe.get_shape_grads()
    s2, j2 = ofe.get_shape_grads()
    assert np.abs(s1 - s2).max() < 1e-13 and np.abs(j1 - j2).max() < 1e-14


def test_vectorised_predicates_fall_back_to_pointwise():
    pts = np.random.default_rng(0).uniform(0, 1, (200, 3))
    vect = lambda p: np.isclose(p[0], pts[7, 0], atol=1e-5)
    scalar_only = lambda p: bool(p[0] > 0.5 and p[1] < 0.5)          # `and` breaks on arrays -> per-point loop
    with_index = lambda p, i: (p[2] > 0.3) & (i % 2 == 0)
    assert np.array_equal(evaluate_location_fn(vect, pts), np.isclose(pts[:, 0], pts[7, 0], atol=1e-5))
    assert np.array_equal(evaluate_location_fn(scalar_only, pts), (pts[:, 0] > 0.5) & (pts[:, 1] < 0.5))
    assert np.array_equal(evaluate_location_fn(with_index, pts), (pts[:, 2] > 0.3) & (np.arange(200) % 2 == 0))
    with pytest.raises(ValueError):
        evaluate_location_fn(lambda a, b, c: True, pts)
    assert np.array_equal(evaluate_point_fn(lambda p: 1.5, pts), np.full(200, 1.5))
    assert np.allclose(evaluate_point_fn(lambda p: np.array([p[0], 2 * p[1], 0.]), pts, (3,)),
                       np.stack([pts[:, 0], 2 * pts[:, 1], 0 * pts[:, 0]], 1))
    assert evaluate_point_fn(lambda p: 1.0, pts[:0]).shape == (0,)   # empty Dirichlet set


def test_law_registry_raises_on_unregistered():
    assert laws.resolve(laws.NeoHookean(1., 0.3), 'HEX8', 3).law_id == 2
    with pytest.raises(laws.UnregisteredLawError):
        laws.resolve(lambda u_grad: u_grad, 'HEX8', 3)
    with pytest.raises(laws.UnregisteredLawError):
        laws.resolve(laws.NeoHookean(1., 0.3), 'QUAD4', 2)
    with pytest.raises(NotImplementedError):
        basis.get_elements('TET4')


def test_library_exports_every_declared_symbol_and_has_no_cpu_fallback():
    lib = _lib.load()
    declared = _lib.declared_symbols()
    assert len(declared) >= 17
    for s in declared:
        assert hasattr(lib, s), s
    assert set(declared) == set(_lib._SIGNATURES)
    assert lib.fem_version() >= 100
    if not torch.cuda.is_available():
        assert lib.fem_device_count() == -3                                      # FEM_ENODEV
        assert lib.fem_spmv(0, None, None, None, 1, None, None, None, None, None) == -3         # refuses to compute without a GPU
        assert b"no CUDA device" in lib.fem_last_error()
        m = jf.box_mesh(2, 2, 2, 1, 1, 1)
        with pytest.raises(RuntimeError):
            type("P", (jf.Problem,), {"get_tensor_map": lambda self: laws.Poisson()})(
                jf.Mesh(m.points, m.cells_dict['hexahedron']), vec=1, dim=3)


def test_product_never_imports_the_oracle():
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "jax_fem_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def _patch_plan_case(points, cells, bc_info, law, config):
    from jax_fem_b200.patch_plan import CONFIGS, build_patch_plan
    from patch_emulator import emulate
    pb = fem.Problem(fem.Mesh(points, cells), 3, 3, dirichlet_bc_info=bc_info, law=law)
    nn = len(points)
    sol = np.random.default_rng(1).standard_normal((nn, 3)) * 1e-3
    Ke = pb.cell_jacobians(sol).reshape(len(cells), 8, 3, 8, 3)
    Re = pb.cell_residuals(sol).reshape(len(cells), 8, 3)
    pb.newton_update(sol)
    A = fem.get_A(pb)
    plan = build_plan(torch.from_numpy(cells), nn, 3)
    pp = build_patch_plan(torch.from_numpy(points), torch.from_numpy(cells), nn, 3, plan.brow_ptr, plan.bcol, config=config)
    cfg = CONFIGS[config]
    assert pp.ck_rnd.max() <= cfg.rmax and np.diff(pp.ck_cell.numpy()).max() <= cfg.chunk
    flag = np.zeros(3 * nn, dtype=np.uint8)
    for rows in pb.bc_rows():
        flag[rows] = 1
    f_ext = np.random.default_rng(2).standard_normal((nn, 3))
    data, res = emulate(pp, Ke, Re, flag, f_ext, plan.nnz)
    assert np.array_equal(A.indptr, plan.indptr.numpy()) and np.array_equal(A.indices, plan.indices.numpy())
    assert not np.isnan(data).any() and not np.isnan(res).any()
    # (first lane, owned mask) of every patch-cell agree with the lane list
    lm = pp.pc_lm.numpy()
    assert np.array_equal(lm[:, 0], np.concatenate([[0], np.cumsum([bin(m).count('1') for m in lm[:, 1]])[:-1]]))
    assert lm[:, 1].min() >= 1 and lm[:, 0][-1] + bin(lm[-1, 1]).count('1') == pp.n_lanes
    assert np.abs(data - A.data).max() <= 1e-13 * np.abs(A.data).max()
    ref = np.zeros((nn, 3))
    np.add.at(ref, cells.reshape(-1), Re.reshape(-1, 3))
    assert np.abs(res - (ref + f_ext)).max() <= 1e-13 * max(1.0, np.abs(ref).max())
    return pp


@pytest.mark.parametrize("config", [0, 1, 4])
def test_patch_plan_structured_box(config):
    """Fused owner-computes assembly tables: walking them in the kernel's order reproduces the oracle's CSR
    (Dirichlet rows included) and nodal residual on a box whose sides are not multiples of the patch edge."""
    m = fem.box_mesh(6, 9, 5, 1., 1.5, 1.)
    pp = _patch_plan_case(m.points, m.cells, cases.CUBE_BC, olaws.LinearElastic(70e3, 0.3), config)
    assert pp.n_patches == {0: 2, 1: 4, 4: 2}[config] * 3 * 2          # 7 x 10 x 6 nodes in bricks of 4 (2) x 4 x 4 layers
    hdr = pp.phdr.numpy()
    assert np.diff(hdr[:, 0]).max() == {0: 64, 1: 32, 4: 64}[config] and np.diff(hdr[:, 1]).max() <= {0: 216, 1: 144, 4: 216}[config]
    assert np.array_equal(np.sort(pp.pn_node.numpy()), np.arange(len(m.points)))


@pytest.mark.parametrize("config", [0, 1, 4])
def test_patch_plan_unstructured_golden_mesh(config):
    from jax_fem_b200.patch_plan import CONFIGS
    g = cases.load_golden("linear_elasticity_cylinder")
    pp = _patch_plan_case(g["points"], g["cells"], cases.CYL_BC, olaws.LinearElastic(70e3, 0.3), config)
    hdr, cfg = pp.phdr.numpy(), CONFIGS[config]
    assert np.diff(hdr[:, 0]).max() <= cfg.max_owned and np.diff(hdr[:, 1]).max() <= cfg.max_local
    assert hdr[:-1, 4].max() <= cfg.acc_doubles


def test_save_sol_round_trips_through_the_vtu_reader(tmp_path):
    """save_sol (jax_fem/utils.py:13-57): float32 point field `sol`, optional cell / point fields; the file is read back
    with the stdlib reader that parses the reference's own .vtu goldens."""
    from oracle.vtu import read_vtu
    m = jf.box_mesh(3, 2, 2, 1., 1., 1.)
    fe = FiniteElement(jf.Mesh(m.points, m.cells_dict['hexahedron']), 3, 3, 'HEX8')
    rng = np.random.default_rng(0)
    sol = rng.standard_normal((fe.num_total_nodes, 3))
    rho = rng.uniform(0, 1, fe.num_cells)
    T = rng.standard_normal(fe.num_total_nodes)
    path = str(tmp_path / "out" / "u.vtu")
    jf.save_sol(fe, torch.from_numpy(sol), path, cell_infos=[('rho', rho)], point_infos=[('T', T)])
    pts, cells, pd = read_vtu(path)
    assert np.array_equal(pts, fe.points) and np.array_equal(cells, fe.cells)
    assert np.array_equal(pd['sol'], sol.astype(np.float32)) and np.array_equal(pd['T'], T.astype(np.float32))
    text = open(path).read()
    assert 'Name="rho"' in text and text.count('<DataArray') == 7
    with pytest.raises(AssertionError):
        jf.save_sol(fe, sol, path, cell_infos=[('bad', rho[:-1])])
    q = jf.rectangle_mesh(2, 2, 1., 1.)
    fe2 = FiniteElement(jf.Mesh(q.points, q.cells_dict['quad']), 1, 2, 'QUAD4')
    jf.save_sol(fe2, np.zeros((fe2.num_total_nodes, 1)), str(tmp_path / "q.vtu"))
    assert read_vtu(str(tmp_path / "q.vtu"))[0].shape == (9, 3)


def test_header_is_plain_c_and_every_entry_point_cites_the_reference(tmp_path):
    """The drop-in boundary is a C ABI: include/fem_b200.h must compile as C (no CUDA / C++ / torch types in any
    signature) and document which reference seam each group of entry points replaces (file:line)."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "fem_b200.h")
    src = tmp_path / "abi.c"
    src.write_text('#include "fem_b200.h"\nint main(void) { int (*f)(void) = fem_version; return f == 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.dirname(hdr), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(hdr).read()
    assert not re.search(r"\b(cudaStream_t|torch|at::|std::)\b", re.sub(r"/\*.*?\*/", "", text, flags=re.S))
    for seam in ("jax_fem/problem.py:439-460", "jax_fem/solver.py:469-553", "jax_fem/solver.py:290-363",
                 "jax_fem/solver.py:63-92", "jax_fem/solver.py:1386-1394", "jax_fem/fe.py:112-141"):
        assert seam in text, seam


def test_mma_optimiser_and_filters_match_the_oracle():
    """jax_fem_b200/mma.py (torch, device-agnostic; here on the CPU) against the oracle's NumPy restatement of
    jax_fem/mma.py: filter matrix and filters, one sub-problem solve incl. dual variables, and the design after a few
    iterations of optimize() with the reference's calling convention."""
    from oracle import mma as omma
    from jax_fem_b200 import mma
    m = jf.box_mesh(5, 4, 3, 1.0, 0.8, 0.6)
    fe = FiniteElement(jf.Mesh(m.points, m.cells_dict['hexahedron']), 3, 3, 'HEX8')
    H, Hs = mma.compute_filter_kd_tree(fe)
    oH, oHs = omma.kd_filter(fe.points, fe.cells, fe.get_JxW(), 3)
    assert np.abs(H.to_dense().numpy() - oH.toarray()).max() < 1e-14 and np.abs(Hs.numpy() - oHs).max() < 1e-14
    n = fe.num_cells
    rng = np.random.default_rng(4)
    rho, dJ, dvc = rng.uniform(0.0005, 1.0, (n, 1)), rng.standard_normal((n, 1)), rng.standard_normal((2, n, 1))
    ft = {'H': H, 'Hs': Hs}
    a, b = mma.applySensitivityFilter(ft, rho, dJ, dvc)
    oa, ob = omma.sensitivity_filter(oH, oHs, rho, dJ, dvc)
    assert np.abs(a.numpy() - oa).max() < 1e-13 and np.abs(b.numpy() - ob).max() < 1e-13
    assert np.abs(mma.applyDensityFilter(ft, rho).numpy() - omma.density_filter(oH, oHs, rho)).max() < 1e-14

    # one sub-problem, both Schur branches (m < n and m >= n), third iteration so that the asymptote update is exercised
    for mm, nn, seed in [(2, 30, 0), (5, 4, 1)]:
        r = np.random.default_rng(seed)
        xval, xo1, xo2 = r.uniform(0.3, 0.7, nn), r.uniform(0.3, 0.7, nn), r.uniform(0.3, 0.7, nn)
        df0, f, df = r.standard_normal(nn), 0.1 * r.standard_normal(mm), r.standard_normal((mm, nn))
        st = omma.MMAState(xval, np.zeros(nn), np.ones(nn), mm, move=0.2)
        st.xold1, st.xold2, st.epoch = xo1, xo2, 3
        st.low, st.upp = xval - 0.4, xval + 0.45
        ref = omma.mma_step(st, 1.0, df0, f, df)
        opt = mma.MMA()
        opt.setNumConstraints(mm); opt.setNumDesignVariables(nn)
        opt.setMinandMaxBoundsForDesignVariables(np.zeros((nn, 1)), np.ones((nn, 1)))
        opt.registerMMAIter(xval[:, None], xo1[:, None], xo2[:, None])
        opt.epoch = 3
        opt.setLowerAndUpperAsymptotes((xval - 0.4)[:, None], (xval + 0.45)[:, None])
        opt.setScalingParams(1.0, np.zeros((mm, 1)), 10000 * np.ones((mm, 1)), np.zeros((mm, 1)))
        opt.setMoveLimit(0.2)
        opt.setObjectiveWithGradient(1.0, df0[:, None])
        opt.setConstraintWithGradient(f[:, None], df)
        opt.mmasub(xval[:, None])
        x, y, z = opt.getOptimalValues()
        lam = opt.getLagrangeMultipliers()[0]
        # both solvers stop when the KKT residual drops below 0.9e-7 (epsimin): the iterates agree to that accuracy
        assert np.abs(x.numpy()[:, 0] - ref[0]).max() < 1e-7 and np.abs(y.numpy()[:, 0] - ref[1]).max() < 1e-7
        assert abs(float(z) - ref[2]) < 1e-7 and np.abs(lam.numpy()[:, 0] - ref[3]).max() < 1e-6 * max(1.0, np.abs(ref[3]).max())
        low, upp = opt.getAsymptoteValues()
        assert np.array_equal(low.numpy()[:, 0], st.low) and np.array_equal(upp.numpy()[:, 0], st.upp)

    # the loop: compliance-like separable objective with a volume constraint, reference calling convention
    t = rng.uniform(0.1, 0.95, n)
    v = 0.4
    obj_np = lambda r_: (float(((r_[:, 0] - t) ** 2).sum()), 2 * (r_ - t[:, None]))
    con_np = lambda r_, it: (np.array([r_.mean() / v - 1.0]), np.ones((1, n, 1)) / (n * v))
    tt = torch.from_numpy(t)
    obj_t = lambda r_: (((r_[:, 0] - tt) ** 2).sum(), 2 * (r_ - tt[:, None]))
    con_t = lambda r_, it: (torch.stack([r_.mean() / v - 1.0]), torch.ones((1, n, 1), dtype=torch.float64) / (n * v))
    params = {'movelimit': 0.2, 'maxIters': 6}
    out = mma.optimize(fe, np.full((n, 1), v), params, obj_t, con_t, 1)
    oout = omma.optimize((oH, oHs), np.full((n, 1), v), params, obj_np, con_np, 1)
    assert out.shape == (n, 1) and np.abs(out.numpy() - oout).max() < 1e-6


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_plans_on_random_connectivity(seed):
    """Plans only see connectivity: on RANDOM cell -> node tables (no geometric structure at all, valence up to the
    limits) the CSR pattern must equal the oracle's and the fused-assembly tables, walked in the kernel's order with
    random element blocks, must reproduce the COO -> CSR sum of those blocks."""
    import scipy.sparse as sp
    from jax_fem_b200.patch_plan import CONFIGS, build_patch_plan
    from patch_emulator import emulate
    rng = np.random.default_rng(seed)
    nn, C, v = 60, 45, 3
    cells = np.stack([np.concatenate([[c % nn], rng.choice(np.delete(np.arange(nn), c % nn), 7, replace=False)])
                      for c in range(C)])                         # 8 DISTINCT nodes per cell; every low node id is used
    points = rng.uniform(0, 1, (nn, 3))
    plan = build_plan(torch.from_numpy(cells), nn, v)
    indptr, indices = fem.csr_pattern_from_cells(cells, v, v * nn)
    assert np.array_equal(plan.indptr.numpy(), indptr) and np.array_equal(plan.indices.numpy(), indices)
    Ke = rng.standard_normal((C, 8, v, 8, v))
    Re = rng.standard_normal((C, 8, v))
    dof = (v * cells[:, :, None] + np.arange(v)).reshape(C, -1)
    I = np.repeat(dof[:, :, None], 8 * v, axis=2).reshape(-1)
    J = np.repeat(dof[:, None, :], 8 * v, axis=1).reshape(-1)
    A = sp.coo_matrix((Ke.reshape(-1), (I, J)), shape=(v * nn, v * nn)).tocsr()
    A.sort_indices()
    assert np.array_equal(A.indices, indices)
    flag = np.zeros(v * nn, dtype=np.uint8)
    flag[rng.choice(v * nn, 11, replace=False)] = 1
    ref = A.data.copy()
    for r in np.flatnonzero(flag):
        seg = slice(indptr[r], indptr[r + 1])
        ref[seg] = (indices[seg] == r).astype(float)
    res_ref = np.zeros((nn, v))
    np.add.at(res_ref, cells.reshape(-1), Re.reshape(-1, v))
    for config in (0, 1, 4):
        pp = build_patch_plan(torch.from_numpy(points), torch.from_numpy(cells), nn, v, plan.brow_ptr, plan.bcol, config=config)
        cfg = CONFIGS[config]
        assert pp.ck_rnd.max() <= cfg.rmax and np.diff(pp.ck_cell.numpy()).max() <= cfg.chunk
        data, res = emulate(pp, Ke, Re, flag, None, plan.nnz)
        assert np.abs(data - ref).max() <= 1e-12 * np.abs(ref).max() and np.abs(res - res_ref).max() <= 1e-12
    if seed == 0:                                                  # a cell that repeats a node is rejected, not mis-assembled
        bad = cells.copy()
        bad[3, 5] = bad[3, 2]
        bplan = build_plan(torch.from_numpy(bad), nn, v)
        with pytest.raises(ValueError, match="same node twice"):
            build_patch_plan(torch.from_numpy(points), torch.from_numpy(bad), nn, v, bplan.brow_ptr, bplan.bcol, config=4)


def test_xla_ffi_shim_compiles_against_the_stub_headers_and_fails_loudly_without_jax():
    """csrc/fem_b200_xla.cc (the jax.ffi custom calls over the C ABI) must at least be well-formed C++: jax is not
    installable here, so it is compiled against the syntax stub of the FFI headers; building or registering for real must
    raise instead of silently doing nothing."""
    import re
    import pytest
    from jax_fem_b200 import xla_ffi
    assert xla_ffi.check_syntax()
    src = open(xla_ffi.SHIM_SRC).read()
    assert sorted(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+)", src)) == sorted(xla_ffi.TARGETS)
    if xla_ffi.include_dir() is None:
        with pytest.raises(RuntimeError):
            xla_ffi.build()
        with pytest.raises(RuntimeError):
            xla_ffi.register()


def test_mesh_files_round_trip_through_the_three_readers(tmp_path):
    """read_mesh (the meshio.read replacement: Gmsh MSH 2.2 as the reference's gmsh generators write it, Abaqus .inp, ASCII
    .vtu) must return the generator's points and connectivity; the .vtu golden of the reference is read identically by
    the product reader and by the oracle's reader; higher-order cells are refused, never silently permuted."""
    import os
    import numpy as np
    import pytest
    from jax_fem_b200.generate_mesh import box_mesh, rectangle_mesh
    from jax_fem_b200.mesh_io import read_mesh
    from oracle import vtu as ovtu
    m = box_mesh(3, 2, 2, 1.5, 1.0, 1.0)
    pts, cells = m.points, m.cells_dict['hexahedron']
    ids = 10 + 3 * np.arange(len(pts))                     # non-contiguous node numbers, as files in the wild have
    msh = tmp_path / "box.msh"
    with open(msh, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % len(pts))
        f.write("".join("%d %.17g %.17g %.17g\n" % (i, *p) for i, p in zip(ids, pts)))
        f.write("$EndNodes\n$Elements\n%d\n" % (len(cells) + 1))
        f.write("1 15 2 0 1 %d\n" % ids[0])                 # a point element, as gmsh emits for physical points
        f.write("".join("%d 5 2 0 1 %s\n" % (k + 2, " ".join(str(ids[n]) for n in c)) for k, c in enumerate(cells)))
        f.write("$EndElements\n")
    got = read_mesh(str(msh))
    assert np.array_equal(got.points, pts) and np.array_equal(got.cells_dict['hexahedron'], cells)
    inp = tmp_path / "box.inp"
    with open(inp, "w") as f:
        f.write("*HEADING\n** comment\n*NODE\n" + "".join("%d, %.17g, %.17g, %.17g\n" % (i, *p) for i, p in zip(ids, pts)))
        f.write("*ELEMENT, TYPE=C3D8, ELSET=ALL\n" + "".join("%d, %s\n" % (k + 1, ", ".join(str(ids[n]) for n in c)) for k, c in enumerate(cells)))
        f.write("*END STEP\n")
    got = read_mesh(str(inp))
    assert np.array_equal(got.points, pts) and np.array_equal(got.cells_dict['hexahedron'], cells)
    q = rectangle_mesh(4, 3, 1., 1.)
    with open(inp, "w") as f:
        f.write("*NODE\n" + "".join("%d, %.17g, %.17g\n" % (i + 1, *p) for i, p in enumerate(q.points)))
        f.write("*ELEMENT, TYPE=CPS4\n" + "".join("%d, %s\n" % (k + 1, ", ".join(str(n + 1) for n in c)) for k, c in enumerate(q.cells_dict['quad'])))
    got = read_mesh(str(inp))
    assert np.array_equal(got.points[:, :2], q.points) and np.array_equal(got.cells_dict['quad'], q.cells_dict['quad'])
    with open(inp, "w") as f:
        f.write("*NODE\n1, 0, 0, 0\n*ELEMENT, TYPE=C3D20\n1, 1\n")
    with pytest.raises(NotImplementedError):
        read_mesh(str(inp))
    golden = "/root/reference/tests/benchmarks/linear_elasticity_cube/fenicsx/sol_p0_000000.vtu"
    if os.path.exists(golden):                              # only in the build container; the GPU box has no /root/reference
        a, (p2, c2, pd2) = read_mesh(golden), ovtu.read_vtu(golden)
        assert np.array_equal(a.points, p2) and np.array_equal(a.cells_dict['hexahedron'], c2)
        assert all(np.array_equal(a.point_data[k], pd2[k]) for k in pd2)
    with pytest.raises(NotImplementedError):
        read_mesh("mesh.xdmf")


def test_hex27_affine_tables_reproduce_the_quadrature_on_affine_cells():
    """The tables Problem hands to the HEX27 affine-cell pass (dN at the reference nodes, reference Gram tables) and the
    formula the kernel evaluates with them, restated in numpy against the oracle's 216-point quadrature on a sheared box."""
    import jax_fem_b200 as jf
    from jax_fem_b200 import basis
    from oracle import fem, laws as olaws
    m = jf.box_mesh_hex27(2, 1, 1, 1.5, 1.0, 0.8)
    pts, cells = m.points.copy(), m.cells_dict['hexahedron27']
    pts = pts @ np.array([[1.0, 0.2, 0.1], [0.05, 0.9, 0.3], [0.0, 0.1, 1.2]]).T + 0.3
    E, nu = 70e3, 0.3
    opb = fem.Problem(fem.Mesh(pts, cells), 3, 3, ele_type='HEX27', law=olaws.LinearElastic(E, nu))
    sol = 1e-3 * np.random.default_rng(0).standard_normal((len(pts), 3))
    Kref = opb.cell_jacobians(sol)
    _, dN, w = basis.get_shape_vals_and_grads('HEX27')
    xi = basis.get_elements('HEX27')[3] / 2.0
    corner = [int(np.flatnonzero((xi == t).all(axis=1))[0]) for t in ([0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1])]
    gram = np.einsum('q,qae,qbf->efab', w, dN, dN)
    mu1, lam1 = 1 / (2 * (1 + nu)), nu / ((1 + nu) * (1 - 2 * nu))

    def affine_map(X):
        J = np.stack([X[corner[e + 1]] - X[corner[0]] for e in range(3)], axis=1)
        return J, np.abs(X - (X[corner[0]] + xi @ J.T)).max() / np.abs(J).max()

    for c in range(len(cells)):
        J, dev = affine_map(pts[cells[c]])
        assert dev <= 1e-13                                                      # the kernel's affinity test
        inv, det = np.linalg.inv(J), np.linalg.det(J)
        G = E * det * np.einsum('ei,efab,fk->abik', inv, gram, inv)
        K = lam1 * G + mu1 * G.transpose(0, 1, 3, 2) + mu1 * np.einsum('abii->ab', G)[:, :, None, None] * np.eye(3)
        K = K.transpose(0, 2, 1, 3).reshape(81, 81)
        assert np.abs(K - Kref[c].reshape(81, 81)).max() <= 1e-13 * np.abs(Kref[c]).max()
    # a curved cell fails the test
    pts2 = pts.copy()
    pts2[cells[0][20]] += 0.01
    assert affine_map(pts2[cells[0]])[1] > 1e-3


def _undefined_names(path):
    """Names a function reads or deletes that are bound nowhere it can see (its own scope, an enclosing function, the module,
    builtins).  A small stand-in for pyflakes, which this image does not have: bench.py runs unattended on the GPU box."""
    import ast
    import builtins
    tree = ast.parse(open(path).read(), path)
    problems = []

    def bound_in(node):
        names = set()
        args = getattr(node, 'args', None)
        if isinstance(args, ast.arguments):
            for a in args.posonlyargs + args.args + args.kwonlyargs + [args.vararg, args.kwarg]:
                if a is not None:
                    names.add(a.arg)

        def visit(n):
            for child in ast.iter_child_nodes(n):
                if isinstance(child, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                    names.add(child.name)
                    continue                      # own scope
                if isinstance(child, ast.Lambda):
                    continue
                if isinstance(child, ast.Name) and isinstance(child.ctx, ast.Store):
                    names.add(child.id)
                elif isinstance(child, (ast.Import, ast.ImportFrom)):
                    for al in child.names:
                        names.add((al.asname or al.name).split('.')[0])
                elif isinstance(child, ast.ExceptHandler) and child.name:
                    names.add(child.name)
                elif isinstance(child, (ast.Global, ast.Nonlocal)):
                    names.update(child.names)
                elif isinstance(child, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
                    for g in child.generators:
                        for t in ast.walk(g.target):
                            if isinstance(t, ast.Name):
                                names.add(t.id)
                elif isinstance(child, ast.NamedExpr):
                    names.add(child.target.id)
                visit(child)
        visit(node)
        return names

    def check(node, visible):
        scope = visible | bound_in(node)
        for child in ast.iter_child_nodes(node):
            walk(child, scope)

    def walk(n, scope):
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
            for d in getattr(n, 'decorator_list', []):
                walk(d, scope)
            check(n, scope)
            return
        if isinstance(n, ast.ClassDef):
            check(n, scope)
            return
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Load, ast.Del)) and n.id not in scope:
            problems.append(f"{path}:{n.lineno}: {n.id}")
        for child in ast.iter_child_nodes(n):
            walk(child, scope)

    check(tree, set(dir(builtins)) | {'__file__', '__name__', '__doc__'})
    return problems


def test_scripts_and_package_have_no_undefined_names():
    import glob
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join(root, f) for f in ("bench.py", "__graft_entry__.py")]
    files += sorted(glob.glob(os.path.join(root, "jax_fem_b200", "*.py"))) + sorted(glob.glob(os.path.join(root, "examples", "*.py")))
    files += sorted(glob.glob(os.path.join(root, "tests", "*.py"))) + sorted(glob.glob(os.path.join(root, "oracle", "*.py")))
    problems = [p for f in files for p in _undefined_names(f)]
    assert not problems, "\n".join(problems)


def test_hex27_meshes_through_the_readers(tmp_path):
    """The 27-node hexahedron of cfg 4 through the file formats: save_sol writes VTK type 29 in the kernels' own node order
    and read_mesh reads it back unchanged; a Gmsh MSH 2.2 file (element type 12) numbers edge and face nodes differently and
    must come back in the kernels' order -- checked geometrically: every node of every cell must sit at the reference
    position the element's lattice gives it."""
    import numpy as np
    import jax_fem_b200 as jf
    from jax_fem_b200 import basis, mesh_io
    from jax_fem_b200.mesh_io import read_mesh
    m = jf.box_mesh_hex27(2, 2, 1, 2.0, 1.0, 0.5)
    pts, cells = m.points, m.cells_dict['hexahedron27']
    fe = jf.FiniteElement(jf.Mesh(pts, cells), vec=3, dim=3, ele_type='HEX27', quadrature_order=4)
    jf.save_sol(fe, np.zeros((len(pts), 3)), str(tmp_path / "h27.vtu"))
    got = read_mesh(str(tmp_path / "h27.vtu"))
    assert np.array_equal(got.cells_dict['hexahedron27'], cells) and np.allclose(got.points, pts)
    # the derivation's VTK table is the element's lattice
    lattice = basis.get_elements('HEX27')[3]
    perm = mesh_io._HEX27_GMSH_TO_VTK
    assert sorted(perm.tolist()) == list(range(27)) and perm[:8].tolist() == list(range(8))
    # Gmsh file: gmsh_cell[perm] = vtk_cell  <=>  gmsh_cell = vtk_cell[inverse]
    inv = np.argsort(perm)
    with open(tmp_path / "h27.msh", "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % len(pts))
        f.write("".join("%d %.17g %.17g %.17g\n" % (i + 1, *p) for i, p in enumerate(pts)))
        f.write("$EndNodes\n$Elements\n%d\n" % len(cells))
        f.write("".join("%d 12 2 0 1 %s\n" % (k + 1, " ".join(str(n + 1) for n in c[inv])) for k, c in enumerate(cells)))
        f.write("$EndElements\n")
    got = read_mesh(str(tmp_path / "h27.msh"))
    c2 = got.cells_dict['hexahedron27']
    assert np.array_equal(c2, cells)
    # independent geometric check of the Gmsh convention itself: in a Gmsh cell, node 9 is the midpoint of vertices 0 and 3,
    # node 20 the centre of face (0, 3, 2, 1), node 26 the cell centre (Gmsh reference manual, hexahedron27)
    g = cells[0][inv]
    P = got.points
    assert np.allclose(P[g[9]], 0.5 * (P[g[0]] + P[g[3]])) and np.allclose(P[g[11]], 0.5 * (P[g[1]] + P[g[2]]))
    assert np.allclose(P[g[20]], 0.25 * (P[g[0]] + P[g[3]] + P[g[2]] + P[g[1]])) and np.allclose(P[g[26]], P[g[:8]].mean(axis=0))
    # and of the result: every node where the lattice says
    X = P[c2[0]]
    lo, ext = X[:8].min(axis=0), X[:8].max(axis=0) - X[:8].min(axis=0)
    assert np.allclose((X - lo) / ext, lattice / 2.0)


def test_linear_mass_law_field_shapes():
    """laws.LinearMass.fields: scalars stay scalars (passed by value to fem_mass_term), per-point fields are validated and
    broadcast to (cells, quads, vec); a field of the wrong shape is refused."""
    import numpy as np
    import pytest
    import torch
    from jax_fem_b200 import laws
    C, Q = 5, 8
    coef, coef_f, cst, cst_f = laws.LinearMass(2.5).fields(C, Q, 1, 'cpu')
    assert coef == 2.5 and coef_f is None and cst_f is None and np.array_equal(cst, np.zeros(3))
    coef, coef_f, cst, cst_f = laws.LinearMass(1.0, [0., -40., 15.]).fields(C, Q, 3, 'cpu')
    assert coef_f is None and cst_f is None and np.array_equal(cst, [0., -40., 15.])
    a, b = np.arange(C * Q, dtype=np.float64).reshape(C, Q), np.ones((C, Q))
    coef, coef_f, cst, cst_f = laws.LinearMass(a, torch.from_numpy(b)).fields(C, Q, 1, 'cpu')
    assert coef_f.shape == (C, Q) and coef_f.is_contiguous() and cst_f.shape == (C, Q, 1) and torch.equal(coef_f, torch.from_numpy(a))
    _, _, _, cst_f = laws.LinearMass(1.0, b).fields(C, Q, 3, 'cpu')              # one field for every component
    assert cst_f.shape == (C, Q, 3) and cst_f.is_contiguous() and bool((cst_f == 1).all())
    with pytest.raises(ValueError):
        laws.LinearMass(np.ones((C, Q + 1))).fields(C, Q, 1, 'cpu')
    with pytest.raises(ValueError):
        laws.LinearMass(1.0, np.ones((C, Q, 2))).fields(C, Q, 3, 'cpu')
