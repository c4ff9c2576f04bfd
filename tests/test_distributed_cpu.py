"""World-size-2 (and 3) tests of the sharded path's host logic on CPU with the gloo backend: partition invariants,
halo exchange, all-reduce, and that owned rows + one ghost layer reproduce the global operator."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fem, laws as olaws
import jax_fem_b200 as jf
from jax_fem_b200.distributed import Halo, TorchDistComm, distributed_bicgstab, partition_mesh


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _global_problem():
    m = jf.box_mesh(6, 3, 2, 2.0, 1.0, 0.7)
    pts, cells = m.points, m.cells_dict['hexahedron']
    opb = fem.Problem(fem.Mesh(pts, cells), 3, 3, law=olaws.LinearElastic(70e3, 0.3),
                      dirichlet_bc_info=[[lambda p: np.isclose(p[0], 0.)] * 3, [0, 1, 2], [lambda p: 0.] * 3])
    opb.newton_update(np.zeros((len(pts), 3)))
    return pts, cells, fem.get_A(opb)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pts, cells, A = _global_problem()
        nn = len(pts)
        part = partition_mesh(cells, nn, rank, world)
        comm = TorchDistComm()
        # ownership is a disjoint cover; ghosts are exactly the non-owned nodes of the local cells
        counts = torch.zeros(nn, dtype=torch.int64)
        counts[torch.from_numpy(part.owned)] += 1
        comm.allreduce(counts)
        assert bool((counts == 1).all())
        assert set(np.unique(cells[part.local_cells])) == set(part.l2g)
        touching = np.flatnonzero(np.isin(cells, part.owned).any(axis=1))
        assert np.array_equal(part.local_cells, touching)
        assert np.array_equal(part.l2g[part.cells_local], cells[part.local_cells])
        # halo exchange: ghosts receive the owners' values, only neighbour ranks talk
        xg = np.sin(np.arange(nn * 3, dtype=np.float64)).reshape(nn, 3)
        x = torch.zeros(part.n_local, 3, dtype=torch.float64)
        x[:part.n_owned] = torch.from_numpy(xg[part.owned])
        halo = Halo(part, comm, 3, 'cpu')
        halo.update(x)
        assert np.array_equal(x.numpy(), xg[part.l2g])
        assert all(abs(s - rank) == 1 for s in part.neighbours)          # x-slabs: nearest neighbours only
        # owned rows of the global operator only need local columns (one ghost layer is enough)
        dof = lambda nodes: (3 * np.asarray(nodes)[:, None] + np.arange(3)).reshape(-1)
        rows = A[dof(part.owned)]
        assert set(np.unique(rows.indices)) <= set(dof(part.l2g))
        y_local = rows[:, dof(part.l2g)] @ x.numpy().reshape(-1)
        assert np.abs(y_local - (A @ xg.reshape(-1))[dof(part.owned)]).max() < 1e-9 * abs(A).max()
        # fused 4-double all-reduce used by the distributed CG
        s = torch.tensor([float(rank + 1), 1.0, 0.0, 2.0], dtype=torch.float64)
        comm.allreduce(s)
        assert s.tolist() == [world * (world + 1) / 2, float(world), 0.0, 2.0 * world]
        ret[rank] = part.n_owned
    finally:
        dist.destroy_process_group()


class _ScipyOps:
    """Owned rows of a global SciPy matrix in local column numbering: the matrix pieces of distributed_bicgstab on CPU."""

    def __init__(self, M, part):
        dof = lambda nodes: (3 * np.asarray(nodes)[:, None] + np.arange(3)).reshape(-1)
        self.rows = M[dof(part.owned)][:, dof(part.l2g)].tocsr()
        self.diag = torch.from_numpy(M.diagonal()[dof(part.owned)])
        self.n_owned = 3 * part.n_owned

    def matvec(self, x, out):
        out[:self.n_owned] = torch.from_numpy(self.rows @ x.numpy())
        return out

    def diagonal(self):
        return self.diag


def _bicgstab_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pts, cells, A = _global_problem()
        AT = A.T.tocsr()                                    # the adjoint operator: non-symmetric (Dirichlet columns)
        nn = len(pts)
        part = partition_mesh(cells, nn, rank, world)
        comm = TorchDistComm()
        halo = Halo(part, comm, 3, 'cpu')
        bg = np.random.default_rng(0).standard_normal(3 * nn)
        dof = lambda nodes: (3 * np.asarray(nodes)[:, None] + np.arange(3)).reshape(-1)
        b = torch.from_numpy(bg[dof(part.l2g)])
        x, info = distributed_bicgstab(None, b, None, part, halo, comm, 3, ops=_ScipyOps(AT, part))
        jac = AT.diagonal()
        xo, ko = fem.bicgstab(AT, bg, M=lambda v: v / jac)
        assert abs(info['iterations'] - ko) <= 2 and info['err'] < 1e-6
        assert np.abs(x.numpy() - xo[dof(part.l2g)]).max() <= 1e-8 * np.abs(xo).max()      # ghosts included
        ret[rank] = info['iterations']
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_bicgstab_recurrences_gloo(world):
    """The sharded Jacobi-BiCGSTAB (halo exchange + 4 all-reduces per iteration over gloo, matrix pieces from SciPy)
    reproduces the oracle's restatement of jax.scipy.sparse.linalg.bicgstab on the non-symmetric adjoint operator."""
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_bicgstab_worker, args=(world, port, ret), nprocs=world, join=True)
        assert len(set(ret.values())) == 1                  # every rank saw the same iteration count


@pytest.mark.parametrize("world", [2, 3])
def test_partition_and_halo_exchange_gloo(world):
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert sum(ret.values()) == 7 * 4 * 3


def test_partition_single_rank_has_no_ghosts():
    m = jf.box_mesh(3, 2, 2, 1, 1, 1)
    part = partition_mesh(m.cells_dict['hexahedron'], len(m.points), 0, 1)
    assert part.n_owned == len(m.points) and len(part.ghosts) == 0 and not part.neighbours
    assert np.array_equal(part.cells_local, m.cells_dict['hexahedron'])


def test_send_lists_from_local_cells_equal_the_other_ranks_ghost_lists():
    """partition_mesh derives what a rank sends from its own cells only; by definition it must be exactly the ghost
    nodes the neighbour lists for this owner, in the neighbour's order (ascending global id)."""
    import numpy as np
    from jax_fem_b200.distributed import _rank_view, node_ranges, partition_mesh
    from jax_fem_b200.generate_mesh import box_mesh
    m = box_mesh(9, 4, 3, 3.0, 1.0, 0.8)
    cells = m.cells_dict['hexahedron'].astype(np.int64)
    n = len(m.points)
    perm = np.random.default_rng(3).permutation(n)
    for conn in (cells, perm[cells]):                     # x-slabs, and a numbering without spatial order
        for world in (2, 3, 5):
            ranges = node_ranges(n, world)
            for rank in range(world):
                part = partition_mesh(conn, n, rank, world)
                for s in range(world):
                    if s == rank:
                        continue
                    _, g_s, o_s = _rank_view(conn, ranges, s)
                    mine = g_s[o_s == rank] - ranges[rank]
                    assert np.array_equal(part.send.get(s, np.zeros(0, np.int64)), mine)


def test_interior_node_range_reads_no_ghosts():
    """The run of owned nodes whose rows are multiplied while the halo exchange is in flight (fem_halo_set_interior) must not
    touch a ghost column and must be maximal."""
    import jax_fem_b200 as jf
    from jax_fem_b200.distributed import partition_mesh, interior_node_range
    m = jf.box_mesh(12, 5, 4, 3., 1., 1.)
    cells = m.cells_dict['hexahedron']
    for world in (2, 3, 4):
        for rank in range(world):
            part = partition_mesh(cells, len(m.points), rank, world)
            lo, hi = interior_node_range(part)
            assert 0 <= lo < hi <= part.n_owned
            cl = part.cells_local
            neigh_max = np.zeros(part.n_local, dtype=np.int64)
            np.maximum.at(neigh_max, cl.reshape(-1), np.repeat(cl.max(axis=1), cl.shape[1]))
            assert (neigh_max[lo:hi] < part.n_owned).all()                      # no ghost neighbour inside the run
            assert lo == 0 or neigh_max[lo - 1] >= part.n_owned                 # maximal on both sides
            assert hi == part.n_owned or neigh_max[hi] >= part.n_owned
