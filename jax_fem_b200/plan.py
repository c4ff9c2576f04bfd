"""Assembly plan: the sparsity pattern and the cell -> CSR-slot maps, built once per Problem.

Replaces the (I, J) COO index arrays of Problem.__post_init__ (jax_fem/problem.py:86-107) and
PETSc's setPreallocationCOO (jax_fem/solver.py:476-478).  The reference materialises C*ndof^2
integers twice (37-74 GB at 200^3); here only the node-block graph is built:

    brow_ptr, bcol : CSR of the node adjacency (node m is a neighbour of n iff they share a cell),
                     columns ascending;
    src_ptr, src   : for every node-block entry the element blocks contributing to it, as indices
                     corner_pos[c*N + a]*N + b into the element-tangent buffer, in ascending (c, a, b)
                     order (fixed summation order);
    corner_pos     : where the element kernel stores the row block of corner (c, a): the node-sorted
                     corner order, so that all row blocks of one mesh node are adjacent in memory;
    nc_ptr, nc     : for every node the (cell, a) corners touching it, coded c*N + a, ascending;
    indptr, indices: the scalar CSR pattern PETSc would produce from (I, J): every (I, J) pair is
                     kept (explicit zeros included), columns ascending, int32.

Everything is plain torch tensor plumbing (sort / unique / cumsum) and runs on whichever device the
connectivity lives on, so the same code is exercised by the CPU tests and on the B200.
"""
from dataclasses import dataclass

import torch

INT32_MAX = 2 ** 31 - 1
SPLIT_SOURCES = 4      # an entry with more sources is gathered as two half rows (csrc/sparse.cu phase C)


@dataclass
class AssemblyPlan:
    num_nodes: int
    num_cells: int
    nodes_per_cell: int
    vec: int
    brow_ptr: torch.Tensor
    bcol: torch.Tensor
    src_ptr: torch.Tensor
    src: torch.Tensor
    nc_ptr: torch.Tensor
    nc: torch.Tensor
    indptr: torch.Tensor
    indices: torch.Tensor
    corner_pos: torch.Tensor = None    # (C*N,) position of corner (c,a) in the node-sorted order (inverse of nc)
    gdesc: torch.Tensor = None         # gather work split: (first corner, first entry, first source, first row) per CTA
    m_sb: torch.Tensor = None          # gather rows (entries; entries with many sources are split in two halves) in
    m_se: torch.Tensor = None          # processing order: source range, entry index, second-half flag
    m_ent: torch.Tensor = None
    m_add: torch.Tensor = None
    eorder: torch.Tensor = None        # entries of every gather CTA sorted by descending source count
    edst: torch.Tensor = None          # offset in CSR data of (row vec*n, col vec*m) per entry
    erow: torch.Tensor = None          # row node of every entry (int32)
    _tperm: torch.Tensor = None

    @staticmethod
    def gather_config(nodes_per_cell):
        """(corners per work item, max corners of one node) -- csrc/sparse.cu::GatherCfg."""
        return (32, 16) if nodes_per_cell <= 8 else (8, 8)

    @property
    def row_block(self):
        """Doubles per corner row block in the element-tangent buffer (padded to a 16-byte multiple)."""
        return (self.nodes_per_cell * self.vec * self.vec + 1) // 2 * 2

    @property
    def n_items(self):
        return int(self.src.numel())

    @property
    def n_gather_blocks(self):
        return int(self.gdesc.numel()) // 4 - 1

    def entry_meta(self, bc_flag):
        """emeta of fem_gather_csr: (source begin, source end, CSR destination, info) per gather row in processing
        order; info = entry_info | bit 20 for the second half of a split entry (added to the first half's result)."""
        if getattr(self, 'native', None) is not None:
            return native_entry_meta(self, bc_flag)
        e = self.m_ent.long()
        info = self.entry_info(bc_flag).long()[e] | (self.m_add.long() << 20)
        return torch.stack([self.m_sb.long(), self.m_se.long(), self.edst.long()[e], info], dim=1).to(torch.int32).contiguous()

    def entry_info(self, bc_flag):
        """einfo of fem_gather_csr: vec*len(n) | diag << 16 | Dirichlet flags of the entry's rows << 17."""
        v = self.vec
        lens = (self.brow_ptr[1:] - self.brow_ptr[:-1]).long()
        erow = self.erow.long()
        info = v * lens[erow]
        info = info | ((self.bcol.long() == erow).long() << 16)
        if bc_flag is not None:
            f = bc_flag.long().reshape(self.num_nodes, v)
            for i in range(v):
                info = info | (f[erow, i] << (17 + i))
        return info.to(torch.int32)

    @property
    def nnzb(self):
        return int(self.bcol.numel())

    @property
    def nnz(self):
        return self.nnzb * self.vec * self.vec

    @property
    def n(self):
        return self.num_nodes * self.vec

    @property
    def tperm(self):
        """Entry (n, m) -> index of entry (m, n); the graph is structurally symmetric."""
        if self._tperm is None:
            nn = self.num_nodes
            counts = (self.brow_ptr[1:] - self.brow_ptr[:-1]).long()
            brow = torch.repeat_interleave(torch.arange(nn, device=self.bcol.device), counts)
            keys = brow * nn + self.bcol.long()
            tkeys = self.bcol.long() * nn + brow
            pos = torch.searchsorted(keys, tkeys)
            assert bool((keys[pos] == tkeys).all()), "pattern is not structurally symmetric"
            self._tperm = pos.to(torch.int32)
        return self._tperm


def _exclusive_ptr(counts):
    ptr = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=counts.device)
    torch.cumsum(counts, 0, out=ptr[1:])
    return ptr


def expand_scalar_pattern(brow_ptr, bcol, vec, chunk=1 << 22):
    """Node-block graph -> scalar CSR (indptr, indices), int32, rows vec*n+i, columns vec*m+k."""
    dev = bcol.device
    nn = brow_ptr.numel() - 1
    nnzb = bcol.numel()
    nnz = nnzb * vec * vec
    if nnz > INT32_MAX:
        raise ValueError(f"nnz = {nnz} exceeds int32 (the reference's PETSc.IntType); shard the mesh")
    lens = (brow_ptr[1:] - brow_ptr[:-1]).long()
    comp = torch.arange(vec, device=dev)
    # indptr[vec*n + i] = vec*vec*brow_ptr[n] + i*vec*len(n)
    row_start = (vec * vec * brow_ptr[:-1].long())[:, None] + comp[None, :] * (vec * lens)[:, None]
    indptr = torch.empty(nn * vec + 1, dtype=torch.int32, device=dev)
    indptr[:-1] = row_start.reshape(-1).to(torch.int32)
    indptr[-1] = nnz
    indices = torch.empty(nnz, dtype=torch.int32, device=dev)
    brow = torch.repeat_interleave(torch.arange(nn, device=dev), lens)
    for s in range(0, nnzb, chunk):
        e = min(nnzb, s + chunk)
        n_e = brow[s:e]
        slot = torch.arange(s, e, device=dev) - brow_ptr[n_e].long()
        base = (vec * vec * brow_ptr[n_e].long() + vec * slot)                  # (E,)
        pos = base[:, None, None] + comp[None, :, None] * (vec * lens[n_e])[:, None, None] + comp[None, None, :]
        val = (vec * bcol[s:e].long())[:, None, None] + comp[None, None, :]
        indices[pos.reshape(-1)] = val.expand(-1, vec, -1).reshape(-1).to(torch.int32)
    return indptr, indices


def build_plan(cells, num_nodes, vec):
    """cells: (C, N) integer tensor (any device)."""
    cells = cells.long()
    C, N = cells.shape
    if C * N * N > INT32_MAX:
        raise ValueError("C*N*N exceeds int32 source codes; shard the mesh across GPUs")
    dev = cells.device
    keys = (cells[:, :, None] * num_nodes + cells[:, None, :]).reshape(-1)
    skeys, order = torch.sort(keys, stable=True)
    del keys
    ukeys, counts = torch.unique_consecutive(skeys, return_counts=True)
    del skeys
    codes = order                      # (c*N + a)*N + b of every source, grouped by entry, ascending inside
    del order
    src_ptr = _exclusive_ptr(counts).to(torch.int32)
    brow = torch.div(ukeys, num_nodes, rounding_mode='floor')
    bcol = (ukeys - brow * num_nodes).to(torch.int32)
    brow_ptr = _exclusive_ptr(torch.bincount(brow, minlength=num_nodes)).to(torch.int32)
    del ukeys
    flat = cells.reshape(-1)
    nc64 = torch.sort(flat, stable=True)[1]
    nc = nc64.to(torch.int32)
    nc_ptr = _exclusive_ptr(torch.bincount(flat, minlength=num_nodes)).to(torch.int32)
    # element row blocks are stored in node-sorted corner order: corner (c,a) -> corner_pos[c*N + a]
    corner_pos64 = torch.empty_like(nc64)
    corner_pos64[nc64] = torch.arange(nc64.numel(), device=dev)
    corner_pos = corner_pos64.to(torch.int32)
    # src: index of the VEC x VEC block of every source inside the element-tangent buffer
    src = (corner_pos64[torch.div(codes, N, rounding_mode='floor')] * N + codes % N).to(torch.int32)
    del codes, nc64, corner_pos64
    indptr, indices = expand_scalar_pattern(brow_ptr, bcol, vec)
    # gather work decomposition (csrc/sparse.cu::gather_csr_kernel): CTA b owns the nodes whose first corner is in
    # [32 b, 32 (b+1)); their corners / entries / sources are contiguous ranges
    width, tail = AssemblyPlan.gather_config(N)
    deg = nc_ptr[1:] - nc_ptr[:-1]
    if deg.numel() and int(deg.max()) > tail:
        raise ValueError(f"a node belongs to {int(deg.max())} cells (> {tail}): mesh valence too high")
    n_corners = int(nc.numel())
    n_blocks = (n_corners + width - 1) // width
    starts = torch.arange(n_blocks + 1, device=dev, dtype=torch.int64) * width
    node0 = torch.searchsorted(nc_ptr[:-1].long().contiguous(), starts)            # first node of every CTA
    ent0 = brow_ptr.long()[node0]
    # gather rows: one per entry, two for an entry with more than SPLIT_SOURCES sources (first half stores, second half
    # adds: halves the longest per-thread loop of the gather).  Inside a CTA the rows are sorted by descending source
    # count (stable) so that the lanes of a warp loop equally long.
    n_ent = bcol.numel()
    cta_of_entry = torch.searchsorted(ent0[1:].contiguous(), torch.arange(n_ent, device=dev), right=True)
    sp = src_ptr.long()
    split = counts > SPLIT_SOURCES
    half = torch.div(counts + 1, 2, rounding_mode='floor')
    ent = torch.arange(n_ent, device=dev)
    r_ent = torch.cat([ent, ent[split]])
    r_sb = torch.cat([sp[:-1], (sp[:-1] + half)[split]])
    r_se = torch.cat([torch.where(split, sp[:-1] + half, sp[1:]), sp[1:][split]])
    r_add = torch.cat([torch.zeros(n_ent, dtype=torch.int64, device=dev), torch.ones(int(split.sum()), dtype=torch.int64, device=dev)])
    key = cta_of_entry[r_ent] * 64 + (63 - (r_se - r_sb).clamp(max=63))
    rorder = torch.sort(key, stable=True)[1]
    m_ent, m_sb, m_se, m_add = [t[rorder].to(torch.int32) for t in (r_ent, r_sb, r_se, r_add)]
    row0 = _exclusive_ptr(torch.bincount(cta_of_entry[r_ent], minlength=n_blocks))
    gdesc = torch.stack([nc_ptr.long()[node0], ent0, src_ptr.long()[ent0], row0], dim=1)
    gdesc = gdesc.reshape(-1).to(torch.int32)
    # (kept for reference / tests) balanced processing order of the unsplit entries
    key = cta_of_entry * 64 + (63 - counts.clamp(max=63))
    eorder = torch.sort(key, stable=True)[1].to(torch.int32)
    del cta_of_entry, key, rorder, r_ent, r_sb, r_se, r_add
    lens = (brow_ptr[1:] - brow_ptr[:-1]).long()
    erow = torch.repeat_interleave(torch.arange(num_nodes, device=dev), lens)
    slot = torch.arange(bcol.numel(), device=dev) - brow_ptr[:-1].long()[erow]
    edst = (vec * vec * brow_ptr[:-1].long()[erow] + vec * slot)
    assert int(edst.max()) <= INT32_MAX if edst.numel() else True
    edst = edst.to(torch.int32)
    erow = erow.to(torch.int32)
    del counts
    return AssemblyPlan(m_sb=m_sb, m_se=m_se, m_ent=m_ent, m_add=m_add, corner_pos=corner_pos, gdesc=gdesc, eorder=eorder, edst=edst, erow=erow, num_nodes=num_nodes, num_cells=C, nodes_per_cell=N, vec=vec, brow_ptr=brow_ptr, bcol=bcol,
                        src_ptr=src_ptr, src=src, nc_ptr=nc_ptr, nc=nc, indptr=indptr, indices=indices)


# ---- the same plan built by the library (csrc/plan.cu, fem_plan_create): no torch operators involved ------------------------
class _DeviceTable:
    """A device table owned by a fem_plan handle, exposed through __cuda_array_interface__ so that torch wraps it without a copy."""

    def __init__(self, ptr, count, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<i4", "data": (int(ptr), False), "version": 3, "strides": None}


class _PlanHandle:
    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        try:
            from . import _lib
            if self.handle:
                _lib.load().fem_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_TABLES = {"brow_ptr": 0, "bcol": 1, "indptr": 2, "indices": 3, "corner_pos": 4, "nc_ptr": 5, "nc": 6, "gdesc": 7, "src": 8,
           "src_ptr": 9, "_tperm": 10, "m_sb": 11, "m_se": 12, "m_ent": 13, "m_add": 14, "edst": 15, "erow": 16}


def build_plan_native(cells, num_nodes, vec):
    """AssemblyPlan whose tables are built by fem_plan_create on the device of `cells` (int32 CUDA tensor (C, N))."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    cells = cells.to(torch.int32).contiguous()
    C, N = cells.shape
    handle = ctypes.c_void_p()
    _lib.check(lib.fem_plan_create(_lib.ptr(cells), C, num_nodes, N, vec, _lib.stream_ptr(), ctypes.byref(handle)))
    owner = _PlanHandle(handle)
    tables = {}
    for name, which in _TABLES.items():
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        _lib.check(lib.fem_plan_table(handle, which, ctypes.byref(p), ctypes.byref(n)))
        if n.value == 0:
            tables[name] = torch.zeros(0, dtype=torch.int32, device=cells.device)
        else:
            tables[name] = torch.as_tensor(_DeviceTable(p.value, n.value, owner), device=cells.device)
    plan = AssemblyPlan(num_nodes=num_nodes, num_cells=C, nodes_per_cell=N, vec=vec, eorder=None, **tables)
    plan.native = owner
    return plan


def native_entry_meta(plan, bc_flag):
    """emeta of fem_gather_csr from the library (fem_plan_entry_meta) for a plan built by build_plan_native."""
    from . import _lib
    emeta = torch.empty((plan.m_ent.numel(), 4), dtype=torch.int32, device=plan.bcol.device)
    flag = None if bc_flag is None else bc_flag.to(torch.uint8).contiguous()
    _lib.check(_lib.load().fem_plan_entry_meta(plan.native.handle, _lib.ptr(flag), _lib.ptr(emeta), _lib.stream_ptr()))
    return emeta
