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
    _tperm: torch.Tensor = None

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
    del brow, ukeys
    flat = cells.reshape(-1)
    nc = torch.sort(flat, stable=True)[1].to(torch.int32)
    nc_ptr = _exclusive_ptr(torch.bincount(flat, minlength=num_nodes)).to(torch.int32)
    indptr, indices = expand_scalar_pattern(brow_ptr, bcol, vec)
    return AssemblyPlan(num_nodes=num_nodes, num_cells=C, nodes_per_cell=N, vec=vec, brow_ptr=brow_ptr, bcol=bcol,
                        src_ptr=src_ptr, src=src, nc_ptr=nc_ptr, nc=nc, indptr=indptr, indices=indices)
