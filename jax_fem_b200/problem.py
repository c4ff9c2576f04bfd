"""Weak-form container with the reference's Problem API, executing on the B200.

Mirror of jax_fem/problem.py::Problem (ctor fields :46-54, custom_init :184, compute_residual :462,
newton_update :477, set_params :493, internal_vars :125).  What differs is where the work happens:

  reference                                             here
  ---------                                             ----
  get_tensor_map() -> Python fn, traced + jacfwd'ed     -> a registered law (jax_fem_b200.laws.*), run by the
  (problem.py:189-214, 262-266)                            hand-written element kernels of csrc/element.cu
  shape_grads/JxW/v_grads_JxW (C,Q,N,dim) in memory     recomputed per cell inside the kernel
  I, J (C*ndof^2 ints, :95-107)                         AssemblyPlan (node-block graph + slot maps); I/J/V
                                                        are materialised only if the caller reads them
  20 batches + host vstack of Jacobians (:359-396)      one launch, element matrices stay in HBM
  .at[].add scatter of the residual (:426-437)          deterministic per-node gather

Only one FE variable is supported (every configuration on the hot path is single-variable);
multi-variable problems, universal kernels and u-dependent mass/surface maps are not registered and
raise NotImplementedError instead of falling back.
"""
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

from . import _lib, laws, logger
from .fe import FiniteElement, evaluate_point_fn
from .generate_mesh import Mesh
from .patch_plan import DEFAULT_CONFIG, build_patch_plan
from .plan import build_plan, build_plan_native
from .stage_plan import StageConfig, build_stage_plan


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("jax_fem_b200 needs a CUDA device: the hot path is sm_100a CUDA with no CPU fallback")
    return torch.device('cuda', torch.cuda.current_device())


def _eval_load_map(fn, u, x, vec):
    """Evaluate a reference-style mass/surface map ``fn(u, x)`` at many points -> (P, vec).

    Registered loads are u-independent (true for every configuration on the hot path: body forces and
    tractions).  The map is evaluated at u = 0; a second evaluation at a perturbed u must agree,
    otherwise the map contributes to the tangent and is not registered."""
    def at(uvals):
        return evaluate_point_fn(lambda z: fn(z[:vec], z[vec:]), np.concatenate([uvals, x], axis=1), (vec,))
    f0 = at(np.zeros((len(x), vec)))
    f1 = at(np.full((len(x), vec), 0.37))
    if not np.array_equal(f0, f1):
        raise NotImplementedError("a mass / surface map given as a Python callable must not depend on u; solution-dependent "
                                  "surface maps are registered as jax_fem_b200.laws.RobinPower (csrc/faces.cu), "
                                  "solution-dependent mass maps are not registered; nothing falls back")
    return f0


@dataclass
class Problem:
    mesh: Mesh
    vec: int
    dim: int
    ele_type: str = 'HEX8'
    quadrature_rule: Any = None
    quadrature_order: int = None
    dirichlet_bc_info: list = None
    location_fns: list = None
    additional_info: tuple = ()

    def __post_init__(self):
        if isinstance(self.mesh, list):
            if len(self.mesh) != 1:
                raise NotImplementedError("multi-variable problems are outside the B200 hot path (SURVEY.md 8)")
            self.mesh, self.vec, self.ele_type = self.mesh[0], self.vec[0], self.ele_type[0]
            for name in ('quadrature_rule', 'quadrature_order'):
                v = getattr(self, name)
                if isinstance(v, list):
                    setattr(self, name, v[0])
            # the reference indexes dirichlet_bc_info[i] per variable (problem.py:56-75): [info] -> info
            d = self.dirichlet_bc_info
            if isinstance(d, list) and len(d) == 1 and (d[0] is None or (isinstance(d[0], (list, tuple)) and len(d[0]) == 3
                                                                           and isinstance(d[0][0], (list, tuple)))):
                self.dirichlet_bc_info = d[0]
        for hook in ('get_universal_kernel', 'get_universal_kernels_surface'):
            if hasattr(self, hook):
                raise NotImplementedError(f"{hook} is not registered on the B200 hot path; it does not fall back")
        self.device = _device()
        self.num_vars = 1
        fe = FiniteElement(mesh=self.mesh, vec=self.vec, dim=self.dim, ele_type=self.ele_type,
                           quadrature_rule=self.quadrature_rule, quadrature_order=self.quadrature_order,
                           dirichlet_bc_info=self.dirichlet_bc_info)
        self.fes = [fe]
        self.cells_list = [fe.cells]
        self.num_cells = fe.num_cells
        self.boundary_inds_list = fe.get_boundary_conditions_inds(self.location_fns)
        self.offset = [0]
        self.num_total_dofs_all_vars = fe.num_total_dofs
        self.cells_list_face_list = [[fe.cells[b[:, 0]]] for b in self.boundary_inds_list]

        dev = self.device
        self._points = torch.from_numpy(fe.points).to(dev)
        self._cells = torch.from_numpy(fe.cells).to(dev)
        self._ref = torch.from_numpy(np.concatenate([fe.shape_grads_ref.reshape(-1), fe.quad_weights])).to(dev)
        # the same dN with the point index fastest (HEX27 kernel: coalesced reads when thread = quadrature point)
        self._ref_t = torch.from_numpy(np.ascontiguousarray(fe.shape_grads_ref.reshape(fe.num_quads, -1).T)).to(dev)
        # HEX27: tables of the affine-cell pass (fem_b200.h: reference node coordinates, then the reference Gram tables
        # Ghat[e][f][a][b] = sum_q w_q dN_a^e dN_b^f); FEM_HEX27_AFFINE=0 sends every cell through the general kernel
        self._hex27_affine = self._hex27_list = None
        import os
        if self.ele_type == 'HEX27' and os.environ.get('FEM_HEX27_AFFINE', '1') != '0':
            from . import basis
            xi = basis.get_elements('HEX27')[3] / 2.0
            gram = np.einsum('q,qae,qbf->efab', fe.quad_weights, fe.shape_grads_ref, fe.shape_grads_ref)
            self._hex27_affine = torch.from_numpy(np.concatenate([xi.reshape(-1), gram.reshape(-1)])).to(dev)
            # workspace: [num_cells][12] doubles of per-cell records, then the int32 list (count, ids) of the non-affine cells
            self._hex27_list = torch.zeros(100 * fe.num_cells + 16, dtype=torch.uint8, device=dev)
        # the plan is built by the library (fem_plan_create, csrc/plan.cu); FEM_PLAN=torch selects the torch construction
        # of plan.py (the one the CPU tests exercise), both give identical tables
        builder = build_plan if os.environ.get('FEM_PLAN', 'native') == 'torch' else build_plan_native
        self.plan = builder(self._cells, fe.num_total_nodes, fe.vec)
        self._Ke = None
        self._Re = None
        self._patch_plan = None
        self._stage_plan = None
        self._ring = None              # (staging buffer, ctrl, pinned status, event) of the one-kernel staged assembly
        self._A_data = None            # CSR values left by the fused assembly of the last newton_update
        self._A_bc_key = None
        self._last_sol = None
        self._bc_cache = None
        self._law = None

        self.internal_vars = ()
        self.internal_vars_surfaces = [() for _ in range(len(self.boundary_inds_list))]
        self.custom_init(*self.additional_info)
        self.pre_jit_fns()

    # ---- hooks kept from the reference ------------------------------------------------------------
    def custom_init(self):
        """Child class should override if more things need to be done in initialization."""
        pass

    def set_params(self, params):
        raise NotImplementedError("Child class must implement this function!")

    def pre_jit_fns(self):
        """The reference jits/vmaps its Python kernels here (problem.py:261-356); this version resolves the
        registered law and precomputes the (solution-independent) load vector."""
        if hasattr(self, 'get_tensor_map'):
            self._law = laws.resolve(self.get_tensor_map(), self.ele_type, self.vec)
        else:
            raise NotImplementedError("a Problem without get_tensor_map has no registered kernel")
        num_surfaces = len(self.boundary_inds_list)
        if hasattr(self, 'get_surface_maps'):
            assert num_surfaces == len(self.get_surface_maps())
        else:
            assert num_surfaces == 0, "Missing definitions for surface integral"
        self._f_ext = self._assemble_loads()

    def _assemble_loads(self):
        """mass_kernel + surface_kernel of the reference (problem.py:216-259) for u-independent maps: a constant
        nodal vector, assembled once on the host (face sets are small; cell loads use bincount)."""
        fe = self.fes[0]
        f = np.zeros((fe.num_total_nodes, fe.vec))
        used = False
        self._mass_law = None
        mass_map = self.get_mass_map() if hasattr(self, 'get_mass_map') else None
        if isinstance(mass_map, laws.MassLaw):                                     # u-dependent: device kernel (csrc/mass.cu)
            self._mass_law = mass_map
            self._shape_vals = torch.from_numpy(np.ascontiguousarray(fe.shape_vals)).to(self.device)
        elif mass_map is not None:
            x = fe.get_physical_quad_points()                                      # (C,Q,dim)
            JxW = fe.get_JxW()
            val = _eval_load_map(mass_map, None, x.reshape(-1, self.dim), fe.vec).reshape(*x.shape[:2], fe.vec)
            contrib = np.einsum('cqv,qn,cq->cnv', val, fe.shape_vals, JxW)
            for i in range(fe.vec):
                f[:, i] += np.bincount(fe.cells.reshape(-1), weights=contrib[:, :, i].reshape(-1), minlength=fe.num_total_nodes)
            used = True
        self._face_sets = []
        if hasattr(self, 'get_surface_maps'):
            for k, b in enumerate(self.boundary_inds_list):
                if len(b) == 0:
                    continue
                if isinstance(self.get_surface_maps()[k], laws.SurfaceLaw):       # u-dependent: device face kernels
                    self._face_sets.append(self._build_face_set(k, b, self.get_surface_maps()[k]))
                    continue
                x = fe.get_physical_surface_quad_points(b)
                _, nanson = fe.get_face_shape_grads(b)
                val = _eval_load_map(self.get_surface_maps()[k], None, x.reshape(-1, self.dim), fe.vec)
                val = val.reshape(*x.shape[:2], fe.vec)
                contrib = np.einsum('fqv,fqn,fq->fnv', val, fe.face_shape_vals[b[:, 1]], nanson)
                nodes = fe.cells[b[:, 0]].reshape(-1)
                for i in range(fe.vec):
                    f[:, i] += np.bincount(nodes, weights=contrib[:, :, i].reshape(-1), minlength=fe.num_total_nodes)
                used = True
        return torch.from_numpy(f).to(self.device) if used else None

    def _build_face_set(self, k, b, law):
        """Device tables of one boundary set for fem_face_residual / fem_face_tangent: its faces, the (u-independent) Nanson
        scale x weight per face quadrature point, and for every boundary node its (face, local node) pairs."""
        fe, dev = self.fes[0], self.device
        _, nanson = fe.get_face_shape_grads(b)                                     # (F, FQ)
        # local nodes of the cell that live on each local face: those whose face shape values do not vanish (fe.face_inds only
        # lists the face's VERTICES, which is all of them for HEX8 / QUAD4 but 4 of the 9 for HEX27)
        support = [np.flatnonzero(np.abs(fe.face_shape_vals[l]).max(axis=0) > 1e-12) for l in range(len(fe.face_shape_vals))]
        assert len({len(sup) for sup in support}) == 1
        face_nodes = np.stack(support)[b[:, 1]]                                    # (F, V)
        glob = fe.cells[b[:, 0][:, None], face_nodes]                              # (F, V) global nodes
        fidx = np.repeat(np.arange(len(b)), face_nodes.shape[1])
        order = np.lexsort((fidx, glob.reshape(-1)))                               # by node, then ascending face
        nodes_sorted = glob.reshape(-1)[order]
        bnode, counts = np.unique(nodes_sorted, return_counts=True)
        i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
        f64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        return dict(k=k, law=law, n_bnodes=len(bnode), bnode=i32(bnode), bf_ptr=i32(np.concatenate([[0], np.cumsum(counts)])),
                    bf_face=i32(fidx[order]), bf_local=i32(face_nodes.reshape(-1)[order]), face_cell=i32(b[:, 0]),
                    face_lid=i32(b[:, 1]), nanson=f64(nanson), fvals=f64(fe.face_shape_vals), fq=fe.num_face_quads,
                    law_host=law.law_host(fe.vec))

    def _face_args(self, fs, sol):
        fe, P = self.fes[0], _lib.ptr
        law = (_lib.ctypes.c_double * 7)(*fs['law_host'])
        return (fe.vec, fe.num_nodes, fs['fq'], fs['n_bnodes'], P(fs['bnode']), P(fs['bf_ptr']), P(fs['bf_face']), P(fs['bf_local']),
                P(fs['face_cell']), P(fs['face_lid']), P(fs['nanson']), P(fs['fvals']), P(self._cells), P(sol), law)

    def _add_face_residual(self, sol, res):
        """+= surface kernel of the registered u-dependent surface maps (problem.py:238-259)."""
        for fs in self._face_sets:
            _lib.check(_lib.load().fem_face_residual(*self._face_args(fs, sol), _lib.ptr(res), _lib.stream_ptr()))
        return res

    def add_face_tangent(self, sol, data):
        """+= d(surface kernel)/du in the assembled CSR values (the face blocks of problem.V, problem.py:456-458)."""
        if self._face_sets:
            p = self.plan
            _, _, flag = self.bc_data()
            for fs in self._face_sets:
                _lib.check(_lib.load().fem_face_tangent(*self._face_args(fs, sol), _lib.ptr(p.brow_ptr), _lib.ptr(p.bcol),
                                                        _lib.ptr(flag), _lib.ptr(data), _lib.stream_ptr()))
        return data

    # ---- flat <-> list helpers (jax.flatten_util in the reference) -----------------------------------
    def unflatten_fn_sol_list(self, dofs):
        return [dofs.reshape(self.fes[0].num_total_nodes, self.fes[0].vec)]

    # ---- Dirichlet rows, merged with the reference's "later groups overwrite" rule --------------------
    def bc_data(self):
        """(rows int32, vals float64, flag uint8 per dof) on the device.  The row set (and everything derived from it) is
        cached on the identity of the index arrays, as _PetscTangentCache._refresh_bc_rows_if_needed does
        (solver.py:494-520); the VALUES are read from fe.vals_list on every call, as the reference's apply_bc_vec does
        (solver.py:297-301), so editing them in place takes effect."""
        fe = self.fes[0]
        key = tuple(id(a) for lst in (fe.node_inds_list, fe.vec_inds_list) for a in lst)
        if self._bc_cache is None or self._bc_cache[0] != key:
            flag = np.zeros(self.num_total_dofs_all_vars, dtype=np.uint8)
            for i in range(len(fe.node_inds_list)):
                flag[np.asarray(fe.node_inds_list[i]) * fe.vec + np.asarray(fe.vec_inds_list[i])] = 1
            rows = np.flatnonzero(flag).astype(np.int32)
            dev = self.device
            flag_dev = torch.from_numpy(flag).to(dev)
            # position of every group's entries in the merged, ascending row list
            pos = [np.searchsorted(rows, np.asarray(fe.node_inds_list[i]) * fe.vec + np.asarray(fe.vec_inds_list[i]))
                   for i in range(len(fe.node_inds_list))]
            self._bc_cache = [key, torch.from_numpy(rows).to(dev), None, flag_dev,
                              (list(fe.node_inds_list), list(fe.vec_inds_list)),   # keep the arrays alive: ids stay unique
                              self.plan.entry_meta(flag_dev), pos, None]
        c = self._bc_cache
        host_vals = np.zeros(c[1].numel())
        for i, pos in enumerate(c[6]):
            host_vals[pos] = np.asarray(fe.vals_list[i], dtype=np.float64)      # later groups overwrite earlier ones
        if c[7] is None or not np.array_equal(c[7], host_vals):
            c[7] = host_vals
            c[2] = torch.from_numpy(host_vals).to(self.device)
        return c[1], c[2], c[3]

    def entry_meta(self):
        """Per-entry (source range, CSR destination, row stride / diagonal / Dirichlet bits) of the CSR gather kernel."""
        self.bc_data()
        return self._bc_cache[5]

    # ---- the hot path -----------------------------------------------------------------------------------
    def hex27_general_cells(self):
        """Ids of the cells the last HEX27 element call left to the general DMMA kernel (not affine / graded density)."""
        if self._hex27_list is None:
            return np.arange(self.num_cells)
        left = self._hex27_list[96 * self.num_cells:].view(torch.int32).cpu().numpy()
        return np.sort(left[1:1 + left[0]])

    def _internal_var(self):
        law = self._law
        iv = list(self.internal_vars)
        if len(iv) > law.n_internal_vars:
            raise NotImplementedError(f"{type(law).__name__} takes {law.n_internal_vars} internal variable(s), got {len(iv)}")
        if not iv:
            if law.requires_internal_var:
                raise ValueError(f"{type(law).__name__} needs internal_vars = [theta (num_cells, num_quads)]")
            return None
        t = iv[0]
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t), dtype=torch.float64)
        t = t.detach().to(device=self.device, dtype=torch.float64).contiguous()
        assert t.shape == (self.num_cells, self.fes[0].num_quads), \
            f"internal variable must have shape (num_cells, num_quads) = {(self.num_cells, self.fes[0].num_quads)}"
        return t

    def _as_sol(self, sol_list):
        sol = sol_list[0] if isinstance(sol_list, (list, tuple)) else sol_list
        if not isinstance(sol, torch.Tensor):
            sol = torch.as_tensor(np.asarray(sol), dtype=torch.float64)
        return sol.detach().to(device=self.device, dtype=torch.float64).contiguous()

    # ---- assembly modes -------------------------------------------------------------------------------------
    def assembly_mode(self):
        """'staged' : element kernel -> HBM staging buffer -> gather kernel (every registered combination; the default);
           'ring'   : element evaluation + CSR gather as work items of ONE persistent kernel, the element tangents staged
                      in an L2-resident ring (csrc/staged.cu; HEX8 / vec 3 / isotropic elasticity).  HBM traffic drops from
                      12.1 to 3.4 GB per assembly at 100^3 (ncu), but the kernel is latency-bound and measures 3.4 ms
                      against 2.45 ms for 'staged' (profiles/r02_ring_assembly.md), so it is opt-in;
           'fused'  : owner-computes patches (csrc/fused.cu; measured slower).
        FEM_ASSEMBLY in the environment selects the mode; every mode is parity-tested."""
        import os
        eligible = (self.ele_type == 'HEX8' and self.fes[0].vec == 3 and self._mass_law is None
                    and self._law.law_id in (laws.LinearElasticity.law_id, laws.SIMP.law_id))
        mode = os.environ.get('FEM_ASSEMBLY', 'staged')
        if mode not in ('ring', 'staged', 'fused'):
            raise ValueError(f"FEM_ASSEMBLY={mode!r}: registered modes are 'ring', 'staged', 'fused'")
        return mode if eligible else 'staged'

    def fused_assembly_enabled(self):
        return self.assembly_mode() == 'fused'

    @property
    def stage_plan(self):
        if self._stage_plan is None:
            import os
            cfg = StageConfig()
            for name in ('ring_bytes', 'tile_cells', 'slack', 'margin', 'in_flight'):
                v = os.environ.get('FEM_RING_' + name.upper())
                if v is not None:
                    setattr(cfg, name, int(v))
            self._stage_plan = build_stage_plan(self.plan, self._cells, self._points, cfg)
        return self._stage_plan

    def _run_ring(self, sol):
        fe, p, sp = self.fes[0], self.plan, self.stage_plan
        lib, P = _lib.load(), _lib.ptr
        dev = self.device
        if self._ring is None:
            stage = torch.empty(sp.n_rows * 72, dtype=torch.float64, device=dev)
            ctrl = torch.zeros(lib.fem_staged_ctrl_ints(sp.n_e, sp.n_g), dtype=torch.int32, device=dev)
            self._ring = (stage, ctrl, torch.zeros(1, dtype=torch.int32).pin_memory(), torch.cuda.Event())
        stage, ctrl, status, event = self._ring
        self.check_assembly_status(block=False)
        if self._Re is None:
            self._Re = torch.empty((self.num_cells, fe.num_nodes * fe.vec), dtype=torch.float64, device=dev)
        iv = self._internal_var()
        emeta = self.entry_meta()
        data = torch.empty(p.nnz, dtype=torch.float64, device=dev)
        _lib.check(lib.fem_assemble_staged(
            self._law.law_id, _lib.host_doubles(self._law.params()), P(self._points), P(sol), P(iv), P(self._ref),
            self.num_cells, P(sp.cells_p), P(sp.corder), P(sp.dest_row), P(sp.prev_g), sp.n_g, P(sp.tdesc), P(sp.gdep),
            P(emeta), P(p.src), P(stage), P(ctrl), P(self._Re), P(data), _lib.stream_ptr()))
        status.copy_(ctrl[3:4], non_blocking=True)
        event.record()
        self._ring_pending = True
        res = torch.empty((fe.num_total_nodes, fe.vec), dtype=torch.float64, device=dev)
        _lib.check(lib.fem_gather_residual(fe.vec, fe.num_nodes, fe.num_total_nodes, P(p.nc_ptr), P(p.nc),
                                           P(self._Re), P(self._f_ext), P(res), _lib.stream_ptr()))
        self._A_data, self._A_bc_key = self.add_face_tangent(sol, data), self._bc_cache[0]
        return self._add_face_residual(sol, res)

    def check_assembly_status(self, block=True):
        """Raise if a wait of the last one-kernel assembly exceeded its limit (a schedule bug: the values are invalid).
        Non-blocking calls only look at a status copy that has already arrived."""
        if self._ring is None or not getattr(self, '_ring_pending', False):
            return
        _, _, status, event = self._ring
        if block:
            event.synchronize()
        elif not event.query():
            return
        self._ring_pending = False
        if int(status[0]) != 0:
            raise RuntimeError("staged assembly: a dependency wait timed out (schedule bug); results are invalid")

    @property
    def patch_plan(self):
        if self._patch_plan is None:
            import os
            self._patch_plan = build_patch_plan(self._points, self._cells, self.fes[0].num_total_nodes, self.fes[0].vec,
                                                self.plan.brow_ptr, self.plan.bcol,
                                                config=int(os.environ.get('FEM_FUSED_CONFIG', DEFAULT_CONFIG)))
        return self._patch_plan

    def _run_fused(self, sol):
        fe = self.fes[0]
        pp, p = self.patch_plan, self.plan
        _, _, flag = self.bc_data()
        iv = self._internal_var()
        data = torch.empty(p.nnz, dtype=torch.float64, device=self.device)
        res = torch.empty((fe.num_total_nodes, fe.vec), dtype=torch.float64, device=self.device)
        P = _lib.ptr
        _lib.check(_lib.load().fem_assemble_fused(
            _lib.ELE[self.ele_type], fe.vec, self._law.law_id, _lib.host_doubles(self._law.params()),
            P(self._points), P(sol), P(iv), P(self._ref), pp.n_patches, P(pp.phdr), P(pp.pn_node), P(pp.pn_out),
            P(pp.pn_acc), P(pp.pn_info), P(pp.lnodes), P(pp.pc_cell), P(pp.pc_ln), P(pp.pc_lm), P(pp.ck_cell), P(pp.ck_lane),
            P(pp.ck_rnd), P(pp.ln_desc), P(pp.ln_slot), P(flag), P(self._f_ext), P(data), P(res), pp.config,
            _lib.stream_ptr()))
        self._A_data, self._A_bc_key = self.add_face_tangent(sol, data), self._bc_cache[0]
        return self._add_face_residual(sol, res)

    def assembled_values(self):
        """CSR values of the last newton_update for get_A: the fused kernel's output when it ran with the current
        Dirichlet sets, otherwise None (get_A then gathers the staged element tangents)."""
        if self._A_data is None:
            return None
        self.bc_data()
        return self._A_data if self._A_bc_key == self._bc_cache[0] else None

    def tiles_enabled(self):
        """HEX8 / vec 3 / {linear elasticity, SIMP, Neo-Hookean} on the FP64 tensor cores: the element tangents are staged as
        tile-major rows written straight from the accumulator fragments (fem_element_tiles) and the isotropic map is applied by
        the CSR gather after the sum (fem_gather_csr_tiles).  FEM_ELEMENT_PATH=dfma / blocks selects the reference-layout
        kernels instead (A/B measurements; all paths are parity-tested)."""
        import os
        return (os.environ.get('FEM_ELEMENT_PATH', 'tiles') == 'tiles' and self.ele_type == 'HEX8' and self.fes[0].vec == 3
                and self._mass_law is None      # the mass kernel adds to row blocks in the reference layout
                and self._law.law_id in (laws.LinearElasticity.law_id, laws.SIMP.law_id, laws.NeoHookean.law_id))

    def _run_element_kernel(self, sol, jac, tiles=None):
        fe = self.fes[0]
        tiles = (jac and self.tiles_enabled()) if tiles is None else tiles
        if self._Re is None:
            self._Re = torch.empty((self.num_cells, fe.num_nodes * fe.vec), dtype=torch.float64, device=self.device)
        if jac and self._Ke is None:
            # one row block (N blocks of vec x vec, padded to an even number of doubles) per corner (cell, a),
            # stored in the plan's node-sorted corner order
            self._Ke = torch.empty((self.num_cells * fe.num_nodes, self.plan.row_block), dtype=torch.float64,
                                   device=self.device)
        iv = self._internal_var()
        lib = _lib.load()
        if self.ele_type == 'HEX27':
            _lib.check(lib.fem_hex27_residual_jacobian(
                self._law.law_id, _lib.host_doubles(self._law.params()), _lib.ptr(self._points), _lib.ptr(self._cells),
                self.num_cells, _lib.ptr(sol), _lib.ptr(iv), _lib.ptr(self._ref), _lib.ptr(self._ref_t), fe.num_quads,
                _lib.ptr(self._hex27_affine), _lib.ptr(self._hex27_list), _lib.ptr(self.plan.corner_pos), _lib.ptr(self._Ke) if jac else None, _lib.ptr(self._Re), _lib.stream_ptr()))
        elif tiles:
            post = (_lib.ctypes.c_double * 3)()
            _lib.check(lib.fem_element_tiles(
                self._law.law_id, _lib.host_doubles(self._law.params()), _lib.ptr(self._points), _lib.ptr(self._cells),
                self.num_cells, _lib.ptr(sol), _lib.ptr(iv), _lib.ptr(self._ref), _lib.ptr(self.plan.corner_pos),
                _lib.ptr(self._Ke), _lib.ptr(self._Re), post, _lib.stream_ptr()))
            self._Ke_post = post
        else:
            self._launch_element(lib, fe, sol, iv, jac)
        if jac:
            self._Ke_tiles = tiles
        if self._mass_law is not None:
            coef, coef_f, cst, cst_f = self._mass_law.fields(self.num_cells, fe.num_quads, fe.vec, self.device)
            _lib.check(lib.fem_mass_term(
                _lib.ELE[self.ele_type], fe.vec, _lib.ptr(self._points), _lib.ptr(self._cells), self.num_cells, _lib.ptr(sol),
                _lib.ptr(self._ref), _lib.ptr(self._shape_vals), fe.num_quads, coef, _lib.ptr(coef_f), _lib.host_doubles(cst),
                _lib.ptr(cst_f), _lib.ptr(self.plan.corner_pos), _lib.ptr(self._Ke) if jac else None, _lib.ptr(self._Re),
                _lib.stream_ptr()))
        res = torch.empty((fe.num_total_nodes, fe.vec), dtype=torch.float64, device=self.device)
        p = self.plan
        _lib.check(lib.fem_gather_residual(fe.vec, fe.num_nodes, fe.num_total_nodes, _lib.ptr(p.nc_ptr), _lib.ptr(p.nc),
                                           _lib.ptr(self._Re), _lib.ptr(self._f_ext), _lib.ptr(res), _lib.stream_ptr()))
        return self._add_face_residual(sol, res)

    def _launch_element(self, lib, fe, sol, iv, jac):
        _lib.check(lib.fem_element_residual_jacobian(
            _lib.ELE[self.ele_type], fe.vec, self._law.law_id, _lib.host_doubles(self._law.params()),
            _lib.ptr(self._points), _lib.ptr(self._cells), self.num_cells, _lib.ptr(sol), _lib.ptr(iv),
            _lib.ptr(self._ref), _lib.ptr(self.plan.corner_pos), _lib.ptr(self._Ke) if jac else None,
            _lib.ptr(self._Re), _lib.stream_ptr()))

    def compute_residual(self, sol_list):
        """sol_list: [ (num_total_nodes, vec) ] -> res_list of the same shapes (problem.py:462-475)."""
        return [self._run_element_kernel(self._as_sol(sol_list), jac=False)]

    def newton_update(self, sol_list):
        """Residual list; the tangent stays on the device for get_A (problem.py:477-491): as finished CSR values when
        the fused assembly is registered for this problem, as element tangents otherwise."""
        sol = self._as_sol(sol_list)
        self._last_sol = sol
        self._A_data = None
        mode = self.assembly_mode()
        if mode != 'staged':
            self._Ke_valid = False
            return [self._run_fused(sol) if mode == 'fused' else self._run_ring(sol)]
        self._Ke_valid = True
        return [self._run_element_kernel(sol, jac=True)]

    def staged_tangents(self):
        """Element tangents of the last newton_update in the staging layout of the two-kernel path; after a fused
        assembly they are produced on demand (problem.V, changed Dirichlet sets)."""
        if self._last_sol is None:
            raise AttributeError("element tangents are defined after newton_update()")
        if not getattr(self, '_Ke_valid', False):
            self._run_element_kernel(self._last_sol, jac=True)
            self._Ke_valid = True
        return self._Ke

    def block_tangents(self):
        """Element tangents as K blocks in the reference's V layout (row block of corner (c, a) = 8 blocks of 3 x 3): what
        problem.V needs.  When the hot path staged tile-major rows, the reference-layout kernel is run on demand."""
        if self._last_sol is None:
            raise AttributeError("element tangents are defined after newton_update()")
        if not getattr(self, '_Ke_valid', False) or getattr(self, '_Ke_tiles', False):
            self._run_element_kernel(self._last_sol, jac=True, tiles=False)
            self._Ke_valid = True
        return self._Ke

    # ---- reference attributes, materialised on demand ---------------------------------------------------
    def element_tangents(self):
        """(num_cells, ndof, ndof) element tangents in the reference's layout (row = test dof)."""
        Ke = self.block_tangents()
        fe = self.fes[0]
        N, v = fe.num_nodes, fe.vec
        rows = Ke[self.plan.corner_pos.long()][:, :N * v * v]              # (C*N, N*v*v) in (c, a) order
        return rows.reshape(self.num_cells, N, N, v, v).permute(0, 1, 3, 2, 4).reshape(self.num_cells, N * v, N * v)

    @property
    def V(self):
        """COO values aligned with I/J: cell blocks, then (zero) face blocks (problem.py:453-458)."""
        Ke = self.element_tangents()
        ndof = Ke.shape[1]
        parts = [Ke.reshape(-1)]
        laws_by_set = {fs['k']: fs for fs in self._face_sets}
        for k, b in enumerate(self.boundary_inds_list):
            parts.append(self._face_blocks(laws_by_set[k]).reshape(-1) if k in laws_by_set else
                         torch.zeros(len(b) * ndof * ndof, dtype=torch.float64, device=self.device))
        return torch.cat(parts)

    def _face_blocks(self, fs):
        """(F, ndof, ndof) face tangents of a registered surface law in the reference's V layout (host-side convenience:
        the hot path adds them straight into the CSR values, fem_face_tangent)."""
        fe = self.fes[0]
        law, vec = fs['law_host'], fe.vec
        coef = torch.tensor(law[:vec], dtype=torch.float64, device=self.device)
        uref = torch.tensor(law[3:3 + vec], dtype=torch.float64, device=self.device)
        N = fs['fvals'][fs['face_lid'].long()]                                      # (F, FQ, NN)
        u = torch.einsum('fqn,fnv->fqv', N, self._last_sol[self._cells[fs['face_cell'].long()].long()])
        d = coef * law[6] * (u - uref) ** (law[6] - 1.0)                            # (F, FQ, vec)
        K = torch.einsum('fqi,fqa,fqb,fq->faib', d, N, N, fs['nanson'])
        out = torch.zeros(K.shape[0], fe.num_nodes, vec, fe.num_nodes, vec, dtype=torch.float64, device=self.device)
        for i in range(vec):
            out[:, :, i, :, i] = K[:, :, i, :]
        return out.reshape(K.shape[0], fe.num_nodes * vec, fe.num_nodes * vec)

    def _coo(self):
        fe = self.fes[0]
        inds = (fe.vec * fe.cells[:, :, None].astype(np.int64) + np.arange(fe.vec)[None, None, :]).reshape(self.num_cells, -1)
        blocks = [inds] + [inds[b[:, 0]] for b in self.boundary_inds_list]
        n = inds.shape[1]
        I = np.concatenate([np.repeat(x[:, :, None], n, axis=2).reshape(-1) for x in blocks])
        J = np.concatenate([np.repeat(x[:, None, :], n, axis=1).reshape(-1) for x in blocks])
        return I, J

    @property
    def I(self):
        return self._coo()[0]

    @property
    def J(self):
        return self._coo()[1]

    def print_BC_info(self):
        fe = self.fes[0]
        for i, b in enumerate(self.boundary_inds_list):
            print(f"Surface boundary set {i + 1}: (num_selected_faces, 2) = {b.shape}")
        for i in range(len(fe.node_inds_list)):
            print(f"Dirichlet part {i + 1}: {len(fe.node_inds_list[i])} dofs on component {fe.vec_inds_list[i][:1]}")
