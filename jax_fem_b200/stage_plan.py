"""Schedule of the staged one-kernel assembly (csrc/staged.cu): which order the cells are evaluated in, when the CSR
rows of a node group may be summed, and where every corner's row block lives in the L2-resident staging ring.

Replaces the role of the dense COO buffer ``problem.V`` (jax_fem/problem.py:453) + ``setValuesCOO``
(jax_fem/solver.py:525): the element tangents are produced and consumed inside one kernel and never reach HBM.

Tables (all int32, on the device of the connectivity):

  corder (C)        processing slot -> cell id.  Cells are swept along the axis that follows the node numbering,
                    tile by tile of the cross-section, so that the set of nodes with some but not all of their cells
                    evaluated (the "frontier", whose row blocks must stay in the ring) is a few thousand cells wide.
  cells_p (C, 8)    connectivity in processing order.
  E item i          = slots [16 i, 16 i + 16);  G item g = work item g of the CSR gather (AssemblyPlan.gdesc).
  tdesc (nE+nG, 32) one descriptor per ticket, in ticket order (csrc/staged.cu): E item i as 0x80000000 | i; G item g with
                    its first staging row, its corner / entry / source / emeta ranges (AssemblyPlan.gdesc) and the E items
                    holding the cells around its nodes (inline up to 20, else in gdep).  A G item's ticket comes `slack` E
                    items after the last E item it needs, so that it is normally complete when a CTA picks the ticket up.
  dest_row (C*8)    staging row of corner (slot, a): ring rows [0, R + 48) are recycled, spill rows behind them are not.
  prev_g (C*8)      G item that read the previous occupant of that row (-1: none); E waits for it before writing.

Ring allocation: G items whose rows live shorter than the ring allows get consecutive ring rows in order of birth
(first E item writing to them); the previous occupant of every ring row is found by sorting (row, birth position), and
the plan checks that this occupant's ticket precedes the overwriting E item's ticket by `margin` tickets -- the kernel
only ever waits for EARLIER tickets, which makes it deadlock-free; items that fail the check are moved to the spill
area and the allocation is repeated.  With no ring-eligible item the layout degenerates to the plain staging buffer.
"""
from dataclasses import dataclass

import torch

CELLS_PER_ITEM = 16
MAX_ITEM_CORNERS = 48          # csrc/staged.cu::kMaxC
E_FLAG = -(2 ** 31)            # 0x80000000 as int32
DESC_INTS, DESC_INLINE = 32, 20  # csrc/staged.cu::kDescInts, kDescInline


@dataclass
class StageConfig:
    ring_bytes: int = 112 << 20    # staging ring; must stay well below the 126 MB L2
    tile_cells: int = 2000         # cells per sweep layer of a tile (frontier width)
    slack: int = 512               # E items between the last E item a G item needs and the G item's ticket
    margin: int = 2048             # tickets between a G item and the E item that recycles its rows
    in_flight: int = 200           # E items that may be running or queued on resident CTAs (sizes the ring pre-filter)
    row_bytes: int = 576


@dataclass
class StagePlan:
    corder: torch.Tensor
    cells_p: torch.Tensor
    dest_row: torch.Tensor
    prev_g: torch.Tensor
    tdesc: torch.Tensor
    gdep: torch.Tensor
    n_e: int
    n_g: int
    ring_rows: int
    n_rows: int                    # rows of the staging buffer (ring + 48 + spill)
    spill_fraction: float

    @property
    def staging_bytes(self):
        return self.n_rows * 576


def _sweep_order(points, cells, tile_cells):
    """Processing order of the cells: strips of the cross-section, each swept along the axis the node numbering follows.

    The axes are ranked by how fast they vary along the node numbering (mean |coordinate step| between consecutive
    nodes): the slowest one is the sweep axis, the cross-section is cut into strips along the middle one only -- a cut
    across the fastest axis would split the runs of consecutive nodes that form a G item."""
    C = cells.shape[0]
    dim = points.shape[1]
    dev = cells.device
    if C <= tile_cells or dim < 2 or points.shape[0] < 2:
        return torch.arange(C, device=dev)
    cen = points[cells.long()].mean(1)                                   # (C, dim)
    lo, hi = points.min(0).values, points.max(0).values
    ext = (hi - lo).clamp_min(1e-300)
    step = ((points[1:] - points[:-1]).abs().mean(0) / ext).tolist()
    axes = sorted(range(dim), key=lambda d: step[d])                     # slowest ... fastest
    sweep, mid = axes[0], axes[1] if dim == 3 else axes[-1]
    h = float((ext.prod() / C) ** (1.0 / dim))                           # mean cell size
    layers = max(1, int(round(float(ext[sweep]) / h)))
    n_strips = max(1, int(round(C / layers / tile_cells)))
    n_strips = min(n_strips, max(1, int(float(ext[mid]) / (2 * h))))     # a strip is at least two cells wide
    strip = ((cen[:, mid] - lo[mid]) / ext[mid] * n_strips).floor().long().clamp_(0, n_strips - 1)
    layer = ((cen[:, sweep] - lo[sweep]) / ext[sweep] * layers).floor().long().clamp_(0, layers - 1)
    return torch.sort(strip * layers + layer, stable=True)[1]


def _seg_reduce(values, seg, n, how):
    out = torch.full((n,), -1 if how == 'amax' else 2 ** 62, dtype=torch.int64, device=values.device)
    return out.scatter_reduce(0, seg, values, how, include_self=True)


def build_stage_plan(plan, cells, points, config=None):
    """plan: AssemblyPlan of an 8-node-cell mesh; cells (C, 8), points (nodes, dim) on the same device."""
    cfg = config or StageConfig()
    dev = cells.device
    C, N = cells.shape
    assert N == 8 and plan.nodes_per_cell == 8, "the staged assembly is registered for 8-node cells"
    corder = _sweep_order(points, cells, cfg.tile_cells)
    slot_of_cell = torch.empty_like(corder)
    slot_of_cell[corder] = torch.arange(C, device=dev)
    n_e = -(-C // CELLS_PER_ITEM)

    gd = plan.gdesc.view(-1, 4).long()
    n_g = gd.shape[0] - 1
    c0 = gd[:, 0].contiguous()
    size = c0[1:] - c0[:-1]
    assert int(size.max()) <= MAX_ITEM_CORNERS
    n_corners = int(c0[-1])
    pos = torch.arange(n_corners, device=dev)
    item_of_pos = torch.bucketize(pos, c0[1:].contiguous(), right=True)
    e_of_pos = torch.div(slot_of_cell[torch.div(plan.nc.long(), N, rounding_mode='floor')], CELLS_PER_ITEM, rounding_mode='floor')
    maxdep = _seg_reduce(e_of_pos, item_of_pos, n_g, 'amax')
    mindep = _seg_reduce(e_of_pos, item_of_pos, n_g, 'amin')
    mindep = torch.where(maxdep < 0, torch.zeros_like(mindep), mindep)

    # merged ticket sequence: E item i has key 2 i, G item g has key 2 (maxdep + slack) + 1
    gkey = 2 * (maxdep + cfg.slack).clamp(max=n_e - 1) + 1
    gkey = torch.where(maxdep < 0, torch.zeros_like(gkey), gkey)
    keys = torch.cat([2 * torch.arange(n_e, device=dev), gkey])
    order = torch.sort(keys, stable=True)[1]
    ticket = torch.empty_like(order)
    ticket[order] = torch.arange(order.numel(), device=dev)
    tick_e, tick_g = ticket[:n_e], ticket[n_e:]
    # E items every G item waits for: the distinct E items of its corners
    pairs = torch.unique(item_of_pos * n_e + e_of_pos)
    dep_item, dep_e = torch.div(pairs, n_e, rounding_mode='floor'), pairs % n_e
    dep_cnt = torch.bincount(dep_item, minlength=n_g)
    dep_ptr = torch.cumsum(dep_cnt, 0) - dep_cnt

    # ---- ring allocation ----------------------------------------------------------------------------------------------
    ring_rows = max(0, cfg.ring_bytes // cfg.row_bytes)
    ring_rows = min(ring_rows, n_corners)
    life_limit = ring_rows // (CELLS_PER_ITEM * N) - cfg.slack - cfg.in_flight
    if ring_rows >= n_corners:
        life_limit = n_e                             # everything fits in one lap
    spill = (maxdep - mindep) > life_limit
    row0 = torch.zeros(n_g, dtype=torch.int64, device=dev)
    prev_of_pos = torch.full((n_corners,), -1, dtype=torch.int64, device=dev)
    for _ in range(6):
        ring_items = torch.nonzero(~spill).reshape(-1)
        prev_of_pos.fill_(-1)
        if ring_items.numel() == 0 or ring_rows == 0:
            spill[:] = True
            break
        birth = torch.sort(mindep[ring_items] * n_g + ring_items, stable=True)[1]
        ring_sorted = ring_items[birth]
        cum = torch.cumsum(size[ring_sorted], 0) - size[ring_sorted]
        cum_start = torch.zeros(n_g, dtype=torch.int64, device=dev)
        cum_start[ring_sorted] = cum
        row0[ring_sorted] = cum % ring_rows
        # previous occupant of every ring row: sort the ring corners by (row, birth position)
        rp = torch.nonzero(~spill[item_of_pos]).reshape(-1)              # ring corner positions
        it = item_of_pos[rp]
        off = rp - c0[it]
        prow = row0[it] + off
        cpos = cum_start[it] + off
        o = torch.sort(prow * (int(cum[-1]) + MAX_ITEM_CORNERS + 1) + cpos)[1]
        prow_s, it_s, e_s = prow[o], it[o], e_of_pos[rp][o]
        same = torch.zeros_like(prow_s, dtype=torch.bool)
        same[1:] = prow_s[1:] == prow_s[:-1]
        prev_item = torch.full_like(it_s, -1)
        prev_item[1:] = torch.where(same[1:], it_s[:-1], prev_item[1:])
        has = prev_item >= 0
        if not bool(has.any()):
            break
        pi = prev_item.clamp(min=0)
        bad = has & (tick_g[pi] + cfg.margin >= tick_e[e_s])
        prev_of_pos[rp[o]] = prev_item
        if not bool(bad.any()):
            break
        spill[it_s[bad]] = True                      # the overwriting item leaves the ring; allocate again
    else:
        spill[:] = True
        prev_of_pos.fill_(-1)
    n_ring = ring_rows + MAX_ITEM_CORNERS if bool((~spill).any()) else 0
    sp_items = torch.nonzero(spill).reshape(-1)
    sp_cum = torch.cumsum(size[sp_items], 0) - size[sp_items]
    row0[sp_items] = n_ring + sp_cum
    n_rows = n_ring + int(size[sp_items].sum())

    row_of_pos = row0[item_of_pos] + (pos - c0[item_of_pos])
    cells_p = cells.long()[corder]
    corner_pos_p = plan.corner_pos.long().view(C, N)[corder]             # node-sorted position of corner (slot, a)
    dest_row = row_of_pos[corner_pos_p.reshape(-1)]

    prev_g = prev_of_pos[corner_pos_p.reshape(-1)]
    prev_g = torch.where(spill[item_of_pos[corner_pos_p.reshape(-1)]], torch.full_like(prev_g, -1), prev_g)

    # ticket descriptors
    n_t = n_e + n_g
    tdesc = torch.full((n_t, DESC_INTS), -1, dtype=torch.int64, device=dev)
    tdesc[tick_e, 0] = torch.arange(n_e, device=dev) + E_FLAG
    g_all = torch.arange(n_g, device=dev)
    tdesc[tick_g, 0] = g_all
    tdesc[tick_g, 1] = row0
    tdesc[tick_g, 2:6] = gd[:-1]
    tdesc[tick_g, 6:10] = gd[1:]
    tdesc[tick_g, 10] = dep_cnt
    tdesc[tick_g, 11] = torch.where(dep_cnt > DESC_INLINE, dep_ptr, torch.full_like(dep_ptr, -1))
    li = torch.arange(dep_item.numel(), device=dev) - dep_ptr[dep_item]
    inl = li < DESC_INLINE
    tdesc[tick_g[dep_item[inl]], 12 + li[inl]] = dep_e[inl]
    assert n_rows * 72 < 2 ** 31 * 8 and int(dest_row.max()) < 2 ** 31
    i32 = lambda t: t.to(torch.int32).contiguous()
    gdep = i32(dep_e) if dep_e.numel() else torch.zeros(1, dtype=torch.int32, device=dev)
    return StagePlan(corder=i32(corder), cells_p=i32(cells_p), dest_row=i32(dest_row), prev_g=i32(prev_g), tdesc=i32(tdesc),
                     gdep=gdep, n_e=n_e, n_g=n_g, ring_rows=ring_rows if n_ring else 0, n_rows=max(n_rows, 1),
                     spill_fraction=float(size[sp_items].sum()) / max(n_corners, 1))
