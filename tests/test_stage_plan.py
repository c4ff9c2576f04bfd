"""The schedule of the staged one-kernel assembly (jax_fem_b200/stage_plan.py) on the CPU: a randomised emulation of the
persistent grid of csrc/staged.cu (workers taking tickets in order, two tickets of look-ahead, blocking on the two prefix
conditions) must never deadlock, and every G item must find exactly the row blocks of its own corners in its staging rows
-- also when the ring is tiny (many laps), when the tiles are tiny (many spilled items) and on renumbered meshes."""
import numpy as np
import pytest
import torch

from jax_fem_b200.generate_mesh import box_mesh
from jax_fem_b200.plan import build_plan
from jax_fem_b200.stage_plan import StageConfig, build_stage_plan, CELLS_PER_ITEM


def emulate(plan, sp, workers, seed, lookahead=3):
    rng = np.random.default_rng(seed)
    td = sp.tdesc.numpy().astype(np.int64)
    gdep = sp.gdep.numpy()
    prev_g = sp.prev_g.numpy().reshape(-1, 8)
    dest = sp.dest_row.numpy().reshape(-1, 8)
    corder = sp.corder.numpy()
    cpos = plan.corner_pos.numpy().reshape(-1, 8)
    C = len(corder)
    n_t = len(td)
    assert n_t == sp.n_e + sp.n_g
    rows = np.full(sp.n_rows, -1, dtype=np.int64)
    done_e, done_g = np.zeros(sp.n_e, bool), np.zeros(sp.n_g, bool)
    nxt_ticket = 0
    queues = [[] for _ in range(workers)]
    checked = 0

    def refill(q):
        nonlocal nxt_ticket
        while len(q) < lookahead and nxt_ticket < n_t:
            q.append(nxt_ticket)
            nxt_ticket += 1

    def deps(t):
        d = td[t]
        if d[11] >= 0:
            return gdep[d[11]:d[11] + d[10]]
        return d[12:12 + d[10]]

    def ready(t):
        code = td[t, 0]
        if code < 0:
            i = code + 2 ** 31
            p = prev_g[i * CELLS_PER_ITEM:(i + 1) * CELLS_PER_ITEM].reshape(-1)
            return bool(done_g[p[p >= 0]].all())
        return bool(done_e[deps(t)].all())

    for q in queues:
        refill(q)
    finished = 0
    while finished < n_t:
        runnable = [w for w, q in enumerate(queues) if q and ready(q[0])]
        assert runnable, f"deadlock after {finished} of {n_t} tickets"
        w = runnable[rng.integers(len(runnable))]
        t = queues[w].pop(0)
        code = td[t, 0]
        if code < 0:
            i = code + 2 ** 31
            for s in range(i * CELLS_PER_ITEM, min(C, (i + 1) * CELLS_PER_ITEM)):
                rows[dest[s]] = cpos[corder[s]]
            done_e[i] = True
        else:
            c0, c1, r0 = td[t, 2], td[t, 6], td[t, 1]
            assert np.array_equal(rows[r0:r0 + (c1 - c0)], np.arange(c0, c1)), f"G item {code} read foreign rows"
            checked += c1 - c0
            done_g[code] = True
        refill(queues[w])
        finished += 1
    assert checked == len(corder) * 8 and done_e.all() and done_g.all()


def _plan(nx, ny, nz, renumber=None):
    m = box_mesh(nx, ny, nz, 1., 1., 1.)
    cells = m.cells_dict['hexahedron'].astype(np.int64)
    pts = m.points
    if renumber is not None:
        perm = np.random.default_rng(renumber).permutation(len(pts))
        inv = np.empty_like(perm)
        inv[perm] = np.arange(len(perm))
        pts, cells = pts[perm], inv[cells]
    cells_t = torch.from_numpy(cells)
    return build_plan(cells_t, len(pts), 3), cells_t, torch.from_numpy(pts)


@pytest.mark.parametrize("cfg", [
    StageConfig(),                                                              # everything fits: one lap, no waits
    StageConfig(ring_bytes=700 * 576, tile_cells=60, slack=2, margin=8, in_flight=0),        # tiny ring: many laps + spilled tile edges
    StageConfig(ring_bytes=300 * 576, tile_cells=10 ** 9, slack=0, margin=0, in_flight=0),   # ring smaller than a layer: mostly spill
])
def test_schedule_never_deadlocks_and_rows_match(cfg):
    plan, cells, pts = _plan(14, 9, 8)
    sp = build_stage_plan(plan, cells, pts, cfg)
    assert sorted(sp.corder.tolist()) == list(range(cells.shape[0]))
    # every staging row is owned by at most one live corner at a time: checked by the emulation under random schedules
    for workers, seed in ((1, 0), (7, 1), (64, 2)):
        emulate(plan, sp, workers, seed)


def test_ring_is_used_and_recycled():
    plan, cells, pts = _plan(20, 8, 8)
    cfg = StageConfig(ring_bytes=1500 * 576, tile_cells=40, slack=2, margin=8, in_flight=0)
    sp = build_stage_plan(plan, cells, pts, cfg)
    assert sp.ring_rows > 0 and sp.spill_fraction < 0.9
    assert int(sp.prev_g.max()) >= 0, "rows must be recycled on this mesh"
    assert sp.n_rows < cells.shape[0] * 8, "the staging buffer must be smaller than the full element-tangent buffer"
    emulate(plan, sp, 16, 3)


def test_renumbered_mesh_degrades_to_spill_but_stays_correct():
    plan, cells, pts = _plan(8, 7, 6, renumber=5)
    sp = build_stage_plan(plan, cells, pts, StageConfig(ring_bytes=400 * 576, tile_cells=50, slack=1, margin=4, in_flight=0))
    emulate(plan, sp, 9, 4)
