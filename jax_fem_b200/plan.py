"""Assembly plan: the sparsity pattern and the cell -> CSR-slot maps, built once per Problem.

Replaces the (I, J) COO index arrays of Problem.__post_init__ (jax_fem/problem.py:86-107) and
PETSc's setPreallocationCOO (jax_fem/solver.py:476-478).  The reference materialises C*ndof^2
integers twice (37-74 GB at 200^3); here only the node-block graph is built:

    brow_ptr, bcol : CSR of the node adjacency (node m is a neighbour of n iff they share a cell),
                     columns ascending;
    src_ptr, src   : for every node-block entry the (cell, a, b) triples contributing to it,
                     coded p = (c*N + a)*N + b, ascending (fixed summation order);
    nc_ptr, nc     : for every node the (cell, a) corners touching it, coded c*N + a, ascending;
    indptr, indices: the scalar CSR pattern PETSc would produce from (I, J): every (I, J) pair is
                     kept (explicit zeros included), columns ascending, int32.

Everything is plain torch tensor plumbing (sort / unique / cumsum) and runs on whichever device the
connectivity lives on, so the same code is exercised by the CPU tests and on the B200.
"""
from dataclasses import dataclass

import torch

INT32_MAX = 2 ** 31 - 1


def packed_pairs(nn):
    """(a, b) of every stored node-pair block, in the order of csrc/common.cuh::pair_index: for nn >= 8 the
    pairs with b < nn/2 come first (the element kernel emits the row block in two column halves)."""
    nb = nn // 2 if nn >= 8 else nn
    first = [(a, b) for a in range(nb) for b in range(a, nb)] if nb != nn else []
    rest = [(a, b) for a in range(nn) for b in range(max(a, nb if nb != nn else a), nn)]
    return first + rest


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
    blk_ent: torch.Tensor = None       # gather CTA b owns entries [blk_ent[b], blk_ent[b+1])
    edst: torch.Tensor = None          # offset in CSR data of (row vec*n, col vec*m) per entry
    erow: torch.Tensor = None          # row node of every entry (int32)
    _tperm: torch.Tensor = None

    GATHER_ITEMS = 256                 # csrc/sparse.cu::kGatherItems
    GATHER_TAIL = 64                   # csrc/sparse.cu::kGatherTail

    @property
    def n_items(self):
        return int(self.src.numel())

    @property
    def n_gather_blocks(self):
        return (self.n_items + self.GATHER_ITEMS - 1) // self.GATHER_ITEMS

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
    src = order.to(torch.int32)
    del order
    src_ptr = _exclusive_ptr(counts).to(torch.int32)
    brow = torch.div(ukeys, num_nodes, rounding_mode='floor')
    bcol = (ukeys - brow * num_nodes).to(torch.int32)
    brow_ptr = _exclusive_ptr(torch.bincount(brow, minlength=num_nodes)).to(torch.int32)
    del ukeys
    flat = cells.reshape(-1)
    nc = torch.sort(flat, stable=True)[1].to(torch.int32)
    nc_ptr = _exclusive_ptr(torch.bincount(flat, minlength=num_nodes)).to(torch.int32)
    indptr, indices = expand_scalar_pattern(brow_ptr, bcol, vec)
    # gather work decomposition (csrc/sparse.cu::gather_csr_kernel)
    max_src = int(counts.max()) if counts.numel() else 0
    if max_src > AssemblyPlan.GATHER_TAIL:
        raise ValueError(f"a node pair is shared by {max_src} cells (> {AssemblyPlan.GATHER_TAIL}): mesh valence too high")
    n_items = int(src.numel())
    n_blocks = (n_items + AssemblyPlan.GATHER_ITEMS - 1) // AssemblyPlan.GATHER_ITEMS
    starts = torch.arange(n_blocks + 1, device=dev, dtype=torch.int64) * AssemblyPlan.GATHER_ITEMS
    blk_ent = torch.searchsorted(src_ptr[:-1].long().contiguous(), starts).to(torch.int32)
    lens = (brow_ptr[1:] - brow_ptr[:-1]).long()
    erow = torch.repeat_interleave(torch.arange(num_nodes, device=dev), lens)
    slot = torch.arange(bcol.numel(), device=dev) - brow_ptr[:-1].long()[erow]
    edst = (vec * vec * brow_ptr[:-1].long()[erow] + vec * slot)
    assert int(edst.max()) <= INT32_MAX if edst.numel() else True
    edst = edst.to(torch.int32)
    erow = erow.to(torch.int32)
    del counts
    return AssemblyPlan(blk_ent=blk_ent, edst=edst, erow=erow, num_nodes=num_nodes, num_cells=C, nodes_per_cell=N, vec=vec, brow_ptr=brow_ptr, bcol=bcol,
                        src_ptr=src_ptr, src=src, nc_ptr=nc_ptr, nc=nc, indptr=indptr, indices=indices)
