"""Patch plan of the fused owner-computes assembly (csrc/fused.cu).

The two-kernel assembly (element kernel -> element tangents in HBM -> CSR gather) moves C*N*N*vec^2 doubles
through HBM twice.  The fused kernel replaces that staging by ownership: the mesh nodes are partitioned into
spatially compact *patches* (<= 64 nodes, 4x4x4 on a structured HEX8 grid); one CTA owns a patch, evaluates every
cell that touches it (cells on the patch surface are evaluated by each patch they touch), accumulates the row
blocks of its own nodes in shared memory in ascending cell order (fixed order, no atomics) and writes the finished
CSR rows exactly once.  This is the precomputed cell -> CSR-slot map of the north star for
_PetscTangentCache.update / get_A (jax_fem/solver.py:469-553) and the scatter-add of
compute_residual_vars_helper (jax_fem/problem.py:426-437), organised by owner instead of by entry.

Tables (all int32, built once per Problem with torch sort/unique/searchsorted on the mesh's device):

    phdr (P+1, 8)     per patch: [0] first owned node (index into pn_*), [1] first local node (lnodes),
                      [2] first patch-cell (pc_*), [3] first chunk (ck_*), [4] accumulator doubles of the patch;
                      row P closes the ranges
    pn_node, pn_out,  per owned node, grouped by patch, ascending node id: global node id, offset of its rows in
    pn_acc, pn_info   the CSR `data` (= vec^2 brow_ptr[n]), offset of its rows in the patch's accumulator,
                      len(n) | (slot of the diagonal block << 8)
    lnodes            per patch the global ids of its local nodes: the owned ones first (same order as pn_node),
                      then the halo nodes of its cells, ascending
    pc_cell, pc_ln,   per (patch, cell) pair, ordered by (patch, chunk, cell id): global cell id, the N local node
    pc_lm             numbers of the cell's corners (uint8 packed in N/4 words), (first lane, mask of owned corners)
    ck_cell, ck_lane, the cells of a patch are processed in chunks of <= `chunk` cells built greedily so that no owned
    ck_rnd            node occurs more than `rmax` times in a chunk (csrc/plan_host.cpp); per chunk the first
                      patch-cell, the first lane and the number of accumulation rounds (<= rmax)
    ln_desc, ln_slot  per lane = owned corner (cell, a) of a chunk, cell-major: cell-in-chunk | a << 5 |
                      owned-node index << 8 | round << 16, and for every column corner b the slot of node
                      cells[c, b] in the row of the owning node (uint8 packed).  Two lanes that add into the same
                      row in the same chunk have different rounds (round = number of earlier cells of the chunk
                      holding that node), so the sum runs in a fixed (chunk, cell) order.
"""
from dataclasses import dataclass

import torch

@dataclass(frozen=True)
class FusedConfig:
    """One (patch size, chunk size) choice; the index in CONFIGS selects the matching kernel instantiation
    (csrc/fused.cu::fem_assemble_fused, `config`)."""
    edge: tuple          # node layers of a patch per direction on tensor-product grids
    max_owned: int       # owned nodes per patch
    chunk: int           # cells per chunk
    rmax: int            # accumulation rounds per chunk (an owned node occurs at most rmax times in a chunk)
    max_local: int       # local nodes per patch (uint8 numbering)

    @property
    def acc_doubles(self):
        return self.max_owned * 27 * 9


CONFIGS = (
    FusedConfig(edge=(4, 4, 4), max_owned=64, chunk=32, rmax=2, max_local=255),    # 0: 1 CTA/SM, 256 threads
    FusedConfig(edge=(2, 4, 4), max_owned=32, chunk=16, rmax=2, max_local=159),    # 1: 2 CTAs/SM, 256 threads, 4 tasks/corner
    FusedConfig(edge=(2, 4, 4), max_owned=32, chunk=16, rmax=2, max_local=159),    # 2: 2 CTAs/SM, 128 threads
    FusedConfig(edge=(2, 4, 4), max_owned=32, chunk=16, rmax=2, max_local=159),    # 3: 2 CTAs/SM, 256 threads, 2 tasks/corner
    FusedConfig(edge=(4, 4, 4), max_owned=64, chunk=8, rmax=1, max_local=255),     # 4: FP64 tensor-core (DMMA) phase 2, one warp per cell
)
DEFAULT_CONFIG = 0


@dataclass
class PatchPlan:
    config: int
    n_patches: int
    n_chunks: int
    n_lanes: int
    nodes_per_cell: int
    vec: int
    phdr: torch.Tensor
    pn_node: torch.Tensor
    pn_out: torch.Tensor
    pn_acc: torch.Tensor
    pn_info: torch.Tensor
    lnodes: torch.Tensor
    pc_cell: torch.Tensor
    pc_ln: torch.Tensor
    pc_lm: torch.Tensor
    ck_cell: torch.Tensor
    ck_lane: torch.Tensor
    ck_rnd: torch.Tensor
    ln_desc: torch.Tensor
    ln_slot: torch.Tensor
    patch_of_node: torch.Tensor

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.__dict__.values() if isinstance(t, torch.Tensor))


def _exclusive_ptr(counts):
    ptr = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=counts.device)
    torch.cumsum(counts, 0, out=ptr[1:])
    return ptr


def _segment_rank(sorted_keys):
    """Position of every element inside its run of equal keys (keys sorted)."""
    n = sorted_keys.numel()
    idx = torch.arange(n, device=sorted_keys.device)
    first = torch.ones(n, dtype=torch.bool, device=sorted_keys.device)
    first[1:] = sorted_keys[1:] != sorted_keys[:-1]
    start = torch.where(first, idx, torch.zeros_like(idx))
    start = torch.cummax(start, 0)[0]
    return idx - start


def assign_patches(points, num_nodes, max_owned, edge=(4, 4, 4)):
    """Spatially compact node clusters: patch id per node (int64, consecutive ids, ordered along x, y, z).

    Tensor-product grids (box_mesh / rectangle_mesh of jax_fem/generate_mesh.py:120-189, also graded ones) are cut
    into bricks of `edge` node layers per direction by coordinate rank; other meshes are binned with the mean node
    spacing.  Bins holding more than max_owned nodes are split along the node numbering."""
    pts = points.double()
    nn, dim = pts.shape
    dev = pts.device
    ext = (pts.max(0)[0] - pts.min(0)[0]).clamp_min(1e-300)
    tol = 1e-9 * float(ext.max())
    ranks, sizes = [], []
    for d in range(dim):
        xs, order = torch.sort(pts[:, d])
        new = torch.ones(nn, dtype=torch.int64, device=dev)
        new[1:] = (xs[1:] - xs[:-1] > tol).long()
        new[0] = 0
        r = torch.empty(nn, dtype=torch.int64, device=dev)
        r[order] = torch.cumsum(new, 0)
        ranks.append(r)
        sizes.append(int(r.max()) + 1)
    prod = 1
    for s in sizes:
        prod *= s
    if prod == nn:                                     # tensor-product grid
        bins = [r // edge[d] for d, r in enumerate(ranks)]
        nb = [(s + edge[d] - 1) // edge[d] for d, s in enumerate(sizes)]
    else:
        h = float(torch.prod(ext)) ** (1.0 / dim) / max(nn, 1) ** (1.0 / dim)
        bins = [torch.floor((pts[:, d] - pts[:, d].min()) / (edge[d] * h)).long() for d in range(dim)]
        nb = [int(b.max()) + 1 for b in bins]
    lin = torch.zeros(nn, dtype=torch.int64, device=dev)
    for d in range(dim):
        lin = lin * nb[d] + bins[d]
    # split bins that hold too many nodes (ascending node id inside a bin), then renumber consecutively
    order = torch.sort(lin, stable=True)[1]
    piece = torch.div(_segment_rank(lin[order]), max_owned, rounding_mode='floor')
    key = lin[order] * (nn // max_owned + 2) + piece
    first = torch.ones(nn, dtype=torch.int64, device=dev)
    first[1:] = (key[1:] != key[:-1]).long()
    first[0] = 0
    pid = torch.empty(nn, dtype=torch.int64, device=dev)
    pid[order] = torch.cumsum(first, 0)
    return pid


def _pack_u8(x):
    """(L, 4k) small non-negative integers -> (L, k) int32 words, byte j of word w = x[:, 4w + j]."""
    L, n = x.shape
    assert n % 4 == 0
    x = x.reshape(L, n // 4, 4).long()
    w = x[:, :, 0] | (x[:, :, 1] << 8) | (x[:, :, 2] << 16) | (x[:, :, 3] << 24)
    w = torch.where(w >= 2 ** 31, w - 2 ** 32, w)
    return w.to(torch.int32).contiguous()


def build_patch_plan(points, cells, num_nodes, vec, brow_ptr, bcol, config=DEFAULT_CONFIG):
    """points (nodes, dim), cells (C, N) with N a multiple of 4, node-block graph (brow_ptr, bcol) of plan.py."""
    cfg = CONFIGS[config]
    cells = cells.long()
    nn = num_nodes
    vv = vec * vec
    lens = (brow_ptr[1:] - brow_ptr[:-1]).long()
    maxlen = int(lens.max()) if nn else 1
    max_owned = max(1, min(cfg.max_owned, cfg.acc_doubles // (vv * max(maxlen, 1))))
    while True:
        plan = _build(points, cells, nn, vec, brow_ptr.long(), bcol.long(), lens, max_owned, cfg, config)
        if plan is not None:
            return plan
        if max_owned == 1:
            raise ValueError("mesh valence too high for the fused assembly (a single node's cells exceed the patch limits)")
        max_owned = max(1, max_owned // 2)


def _greedy_chunks(cell_ptr, owned_idx, max_owned, cfg):
    """csrc/plan_host.cpp::fem_patch_chunks_host on host copies -> (chunk of every patch-cell inside its patch,
    round of every corner, chunks per patch) on the tables' device."""
    import ctypes
    import numpy as np
    from . import _lib
    dev = owned_idx.device
    cp = np.ascontiguousarray(cell_ptr.cpu().numpy().astype(np.int64))
    own = np.ascontiguousarray(owned_idx.cpu().numpy().astype(np.uint8))
    M, N = own.shape
    P = len(cp) - 1
    cell_chunk = np.zeros(M, dtype=np.int32)
    rank = np.zeros((M, N), dtype=np.uint8)
    n_chunks = np.zeros(P, dtype=np.int32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    code = _lib.load().fem_patch_chunks_host(P, vp(cp), vp(own), N, max_owned, cfg.chunk, cfg.rmax, vp(cell_chunk),
                                             vp(rank), vp(n_chunks))
    if code != 0:
        raise RuntimeError(f"fem_patch_chunks_host failed ({code})")
    return (torch.from_numpy(cell_chunk).to(dev).long(), torch.from_numpy(rank).to(dev).long(),
            torch.from_numpy(n_chunks).to(dev).long())


def _build(points, cells, nn, vec, brow_ptr, bcol, lens, max_owned, cfg, config):
    C, N = cells.shape
    dev = cells.device
    vv = vec * vec
    pon = assign_patches(points.to(dev), nn, max_owned, cfg.edge)
    P = int(pon.max()) + 1 if nn else 0
    # owned nodes grouped by patch, ascending node id
    pn_node = torch.sort(pon, stable=True)[1]
    node_ptr = _exclusive_ptr(torch.bincount(pon, minlength=P))
    own_idx = torch.empty(nn, dtype=torch.int64, device=dev)
    own_idx[pn_node] = torch.arange(nn, device=dev) - node_ptr[pon[pn_node]]
    acc_len = vv * lens[pn_node]
    acc_cum = _exclusive_ptr(acc_len)
    pn_acc = acc_cum[:-1] - acc_cum[node_ptr[pon[pn_node]]]
    acc_total = acc_cum[node_ptr[1:]] - acc_cum[node_ptr[:-1]]
    if P and int(acc_total.max()) > cfg.acc_doubles:
        return None
    gkeys = torch.repeat_interleave(torch.arange(nn, device=dev), lens) * nn + bcol       # ascending
    diag = torch.searchsorted(gkeys, pn_node * nn + pn_node) - brow_ptr[pn_node]
    pn_info = lens[pn_node] | (diag << 8)
    pn_out = vv * brow_ptr[pn_node]
    # (patch, cell) pairs, ascending cell id inside a patch
    cp = pon[cells]                                                                        # (C, N)
    pk = torch.unique((cp * C + torch.arange(C, device=dev)[:, None]).reshape(-1))         # sorted
    pc_patch = torch.div(pk, C, rounding_mode='floor')
    pc_cell = pk - pc_patch * C
    del pk, cp
    M = pc_cell.numel()
    cell_ptr = _exclusive_ptr(torch.bincount(pc_patch, minlength=P))
    pcn = cells[pc_cell]                                                                   # (M, N) global nodes
    owned = pon[pcn] == pc_patch[:, None]
    # chunks: greedy first fit per patch, then reorder the patch-cells by (patch, chunk, cell)
    owned_idx = torch.where(owned, own_idx[pcn], torch.full_like(pcn, 255))
    cell_chunk, rank, n_chunks_p = _greedy_chunks(cell_ptr, owned_idx, max_owned, cfg)
    chunk_ptr = _exclusive_ptr(n_chunks_p)
    n_chunks = int(chunk_ptr[-1])
    pc_chunk = chunk_ptr[pc_patch] + cell_chunk
    order = torch.sort(pc_chunk, stable=True)[1]
    pc_patch, pc_cell, pcn, owned, rank, pc_chunk = pc_patch[order], pc_cell[order], pcn[order], owned[order], rank[order], pc_chunk[order]
    del order, owned_idx, cell_chunk
    ck_cell = _exclusive_ptr(torch.bincount(pc_chunk, minlength=n_chunks))
    cic = torch.arange(M, device=dev) - ck_cell[pc_chunk]                                  # cell index inside its chunk
    # local node numbering: owned first (ascending id), then halo (ascending id)
    lk = pc_patch[:, None] * (2 * nn) + torch.where(owned, torch.zeros_like(pcn), torch.full_like(pcn, nn)) + pcn
    # isolated owned nodes (no cell) still need a local number: add every (patch, owned node)
    lkeys = torch.unique(torch.cat([lk.reshape(-1), pon[pn_node] * (2 * nn) + pn_node]))
    lpatch = torch.div(lkeys, 2 * nn, rounding_mode='floor')
    lnode_ptr = _exclusive_ptr(torch.bincount(lpatch, minlength=P))
    lnodes = lkeys % nn
    if P and int((lnode_ptr[1:] - lnode_ptr[:-1]).max()) > cfg.max_local:
        return None
    lidx = torch.searchsorted(lkeys, lk.reshape(-1)).reshape(M, N) - lnode_ptr[pc_patch][:, None]
    del lk, lkeys, lpatch
    pc_ln = _pack_u8(lidx)
    # lanes: owned corners, cell-major inside a chunk
    lp, la = torch.nonzero(owned, as_tuple=True)                                           # row-major => (pair, a) ascending
    ln_node = pcn[lp, la]
    ln_chunk = pc_chunk[lp]
    ln_rank = rank[lp, la]
    ck_lane = _exclusive_ptr(torch.bincount(ln_chunk, minlength=n_chunks))
    n_own = owned.long().sum(1)
    pc_first = _exclusive_ptr(n_own)[:-1]
    pc_mask = (owned.long() << torch.arange(N, device=dev)[None, :]).sum(1)
    pc_lm = torch.stack([pc_first, pc_mask], dim=1)
    ck_rnd = torch.zeros(n_chunks, dtype=torch.int64, device=dev)
    if ln_chunk.numel():
        ck_rnd.scatter_reduce_(0, ln_chunk, ln_rank + 1, reduce='amax')
    if n_chunks and int(ck_rnd.max()) > cfg.rmax:
        raise ValueError("a cell lists the same node twice (degenerate connectivity): the fused assembly orders its "
                         f"accumulation by rounds and allows {cfg.rmax} per chunk")
    ln_desc = cic[lp] | (la << 5) | (own_idx[ln_node] << 8) | (ln_rank << 16)
    slots = torch.searchsorted(gkeys, (ln_node[:, None] * nn + pcn[lp]).reshape(-1)).reshape(-1, N) - brow_ptr[ln_node][:, None]
    ln_slot = _pack_u8(slots)
    i32 = lambda t: t.to(torch.int32).contiguous()
    phdr = torch.zeros((P + 1, 8), dtype=torch.int64, device=dev)
    phdr[:, 0], phdr[:, 1], phdr[:, 2], phdr[:, 3] = node_ptr, lnode_ptr, cell_ptr, chunk_ptr
    phdr[:-1, 4] = acc_total
    return PatchPlan(config=config, n_patches=P, n_chunks=n_chunks, n_lanes=int(ln_desc.numel()), nodes_per_cell=N, vec=vec,
                     phdr=i32(phdr), pn_node=i32(pn_node), pn_out=i32(pn_out), pn_acc=i32(pn_acc), pn_info=i32(pn_info),
                     lnodes=i32(lnodes), pc_cell=i32(pc_cell), pc_ln=pc_ln, pc_lm=i32(pc_lm), ck_cell=i32(ck_cell), ck_lane=i32(ck_lane),
                     ck_rnd=i32(ck_rnd), ln_desc=i32(ln_desc), ln_slot=ln_slot, patch_of_node=pon)
