#!/usr/bin/env python
"""Benchmark of the hot path: residual + Jacobian assembly (DOFs/s) and CG SpMV (GB/s), HEX8 elasticity.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl reference]

Workload (BASELINE.json north_star target): box_mesh(S,S,S) HEX8, linear elasticity E=70e3 nu=0.3,
u=0 on x=0, traction on x=1; default S=200 (24,361,803 DOF, nnz 1,953,736,209: the >= 24M-DOF mesh of the north star,
cfg 2's problem at cfg 3's size).  The mesh is FIXED: with --gpus N its cells are sharded in x-slabs over the N ranks
("scaling": "strong").  A "step" is one full residual+Jacobian assembly: element kernels -> deterministic gather into CSR
with Dirichlet rows treated -> residual with apply_bc_vec.  `value` = DOFs/s with inputs resident in HBM; `e2e` = the
same step through the public API from pinned host memory (solution H2D, residual D2H inside the timed region).  Every
run also times the SpMV and a full Jacobi-CG solve (`cg_solve`: iterations, ms per iteration, true residual) and, on one
GPU, cfg 2 itself (100^3, `cfg2_100cube`).

One JSON line on stdout (rank 0).  --impl reference times the CPU oracle (NumPy/SciPy port of the reference's
algorithm; the reference itself needs jax/basix/petsc4py which are not installable here) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
_T0 = time.perf_counter()


def log(msg):
    """Progress on stderr (stdout carries exactly one JSON line)."""
    if os.environ.get("RANK", "0") == "0":
        print(f"[bench {time.perf_counter() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)

HBM_FALLBACK_GBS = 6650.0


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback"


def algorithmic_bytes(n_dofs, nnz, n_cells, n_nodes, nodes_per_cell=8, dim=3, vec=3):
    """compulsory traffic, scalar CSR fp64 + int32."""
    b_spmv = 12 * nnz + 20 * n_dofs
    b_asm = 8 * nnz + 8 * n_dofs + 4 * nodes_per_cell * n_cells + (8 * dim + 8 * vec) * n_nodes
    return b_asm, b_spmv


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
def cpu_oracle_assembly(size, repeats=1, threads=None):
    """Time the oracle's assembly (element einsum + COO->CSR + BC rows) on the host; returns (DOFs/s, dict)."""
    from oracle import fem, laws
    threads = threads or min(32, os.cpu_count() or 1)
    t0 = time.perf_counter()
    mesh = fem.box_mesh(size, size, size, 1., 1., 1.)
    left = lambda p: np.isclose(p[0], 0., atol=1e-5)
    bc = [[left] * 3, [0, 1, 2], [lambda p: 0.] * 3]
    pb = fem.Problem(mesh, 3, 3, dirichlet_bc_info=bc, law=laws.LinearElastic(70e3, 0.3))
    setup_s = time.perf_counter() - t0
    sol = 1e-3 * np.random.default_rng(0).standard_normal((len(mesh.points), 3))
    I, J = pb.coo_pattern()
    best = None
    # Element phase: cell chunks on a thread pool (NumPy releases the GIL in einsum/matmul) with BLAS pinned to one
    # thread per call -- many-core hosts otherwise spend their time in BLAS thread hand-offs on 24x9 GEMMs.
    # Global phase (COO -> CSR, BC rows) is single-threaded, like PETSc's setValuesCOO in the reference.
    from concurrent.futures import ThreadPoolExecutor
    from threadpoolctl import threadpool_limits
    C = pb.num_cells
    bounds = np.linspace(0, C, 4 * threads + 1).astype(int)
    chunks = [slice(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    for _ in range(repeats):
        t0 = time.perf_counter()
        with threadpool_limits(limits=1), ThreadPoolExecutor(threads) as pool:
            Ks = list(pool.map(lambda sl: pb.cell_jacobians(sol, sl), chunks))
            Rs = list(pool.map(lambda sl: pb.cell_residuals(sol, sl), chunks))
        pb.V_cells = np.concatenate(Ks)
        res = np.zeros((len(mesh.points), 3))
        np.add.at(res, pb.cells.reshape(-1), np.concatenate(Rs).reshape(-1, 3))
        t1 = time.perf_counter()
        A = fem.zero_rows(fem.coo_to_csr(I, J, pb.coo_values(), pb.num_total_dofs_all_vars), pb.bc_rows())
        fem.apply_bc_vec(res.reshape(-1), sol.reshape(-1), pb)
        t2 = time.perf_counter()
        dt = t2 - t0
        if best is None or dt < best[0]:
            best = (dt, t1 - t0, t2 - t1)
    x = np.random.default_rng(0).standard_normal(A.shape[0])
    t0 = time.perf_counter()
    for _ in range(5):
        A @ x
    spmv_s = (time.perf_counter() - t0) / 5
    n = pb.num_total_dofs_all_vars
    return n / best[0], {"n_dofs": n, "nnz": int(A.nnz), "element_s": best[1], "coo_to_csr_s": best[2],
                         "spmv_ms": spmv_s * 1e3, "spmv_gbs": (12 * A.nnz + 20 * n) / spmv_s / 1e9, "setup_s": setup_s,
                         "threads": threads}


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's algorithm on the host cores (bounded sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    size = args.ref_size
    for _ in range(args.warmup and 1):
        cpu_oracle_assembly(min(size, 20))
    vals = []
    t_all = time.perf_counter()
    info = None
    for _ in range(max(1, min(args.steps, 3))):
        v, info = cpu_oracle_assembly(size)
        vals.append(v)
    value = float(np.mean(vals))
    cores = info["threads"]
    line = {
        "impl": "reference", "metric": "assembled DOFs/s (residual+Jacobian), HEX8 linear elasticity", "value": value,
        "unit": "DOF/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": 1, "ms_per_step": 1e3 * info["n_dofs"] / value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"HEX8 box {args.size}x{args.size}x{args.size} linear elasticity E=70e3 nu=0.3, u=0 on x=0, traction on x=1 "
                               f"(the fixed mesh of the b200 arm); CPU sample = {size}^3 cells of the same workload, size-normalised DOF/s",
                   "size": args.size, "sample_size": size},
        "cpu_baseline": {"value": value, "unit": "DOF/s", "cores": cores, "kind": "port",
                         "sample": f"{size}^3 HEX8 cells ({info['n_dofs']} DOF): einsum element matrices {info['element_s']:.2f}s + "
                                   f"scipy COO->CSR+BC {info['coo_to_csr_s']:.2f}s; SpMV {info['spmv_gbs']:.1f} GB/s",
                         "note": "NumPy/SciPy restatement of the reference algorithm (oracle/), not the reference: "
                                 "jax/basix/petsc4py are not installable in this image"},
        "e2e": {"value": value, "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "spmv_gbs": info["spmv_gbs"], "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
def problem_class(workload="elasticity"):
    import jax_fem_b200 as jf
    from jax_fem_b200 import laws

    class Elasticity(jf.Problem):             # docs/source/learn/linear_elasticity/example.ipynb cells 7,12
        def get_tensor_map(self):
            return laws.LinearElasticity(70e3, 0.3)

        def get_surface_maps(self):
            return [lambda u, x: np.array([0., 0., 100.])]

    class NeoHookean(jf.Problem):             # applications/scalability/hyperelastic3d_common.py:15-42 (cfg 3 law)
        def get_tensor_map(self):
            return laws.NeoHookean(10.0, 0.3)

        def get_surface_maps(self):
            return [lambda u, x: np.array([0., 1e-3, 0.])]

    class SIMPCantilever(jf.Problem):         # cfg 5: topology_optimization example.ipynb cell 9 law, 3-D isotropic form
        def get_tensor_map(self):
            return laws.SIMP(70e3, 70.0, 0.3, 3.0)

        def get_surface_maps(self):
            return [lambda u, x: np.array([0., 0., 100.])]

    return {"neohookean": NeoHookean, "simp": SIMPCantilever}.get(workload, Elasticity)


def build_problem(size, world=1, comm=None, workload="elasticity"):
    """The fixed size^3 box (cfg 5: the 4s x s x 2s cantilever) on one GPU, or its cells sharded in x-slabs with one ghost
    layer over `world` ranks (jax_fem_b200/distributed.py) -- strong scaling.  Returns (problem, sharded)."""
    import jax_fem_b200 as jf
    Lx = 1.0
    hex27 = workload == "hex27"
    if workload == "simp":
        # cfg 5: cantilever Lx:Ly:Lz = 2:0.5:1 (applications/outdated/top_opt/box.py:29-30): 4s x s x 2s cells, s = size / 2
        s = max(2, size // 2)
        Lx = 2.0
        m = jf.box_mesh(4 * s, s, 2 * s, 2.0, 0.5, 1.0)
    else:
        m = (jf.box_mesh_hex27 if hex27 else jf.box_mesh)(size, size, size, Lx, 1., 1.)
    cells = m.cells_dict['hexahedron27' if hex27 else 'hexahedron']
    ele = 'HEX27' if hex27 else 'HEX8'
    left = lambda p: np.isclose(p[0], 0., atol=1e-5)
    right = lambda p: np.isclose(p[0], Lx, atol=1e-5)
    kw = dict(dirichlet_bc_info=[[left] * 3, [0, 1, 2], [lambda p: 0.] * 3], location_fns=[right], ele_type=ele)
    if workload == "neohookean":
        # cfg 3 (applications/scalability/hyperelastic3d_common.py:83-103): x=0 fixed, x=Lx pulled by 2 % in x with
        # u_y = u_z = 0, traction on y=1
        ymax = lambda p: np.isclose(p[1], 1., atol=1e-5)
        kw = dict(dirichlet_bc_info=[[left] * 3 + [right] * 3, [0, 1, 2] * 2,
                                     [lambda p: 0.] * 3 + [lambda p: 0.02 * Lx] + [lambda p: 0.] * 2],
                  location_fns=[ymax], ele_type=ele)
    cls = problem_class(workload)
    if world == 1:
        return cls(jf.Mesh(m.points, cells), vec=3, dim=3, **kw), None
    from jax_fem_b200.distributed import ShardedProblem
    sp = ShardedProblem(cls, m.points, cells, comm, vec=3, dim=3, **kw)
    return sp.problem, sp


def time_assembly(prob, sol, steps, jf, torch, barrier):
    """(ms per step, element ms, bc ms, gather ms) of `steps` assemblies, CUDA events on the current stream."""
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    barrier()
    for k in range(steps):
        ev[k][0].record()
        res = prob.newton_update([sol])[0]
        ev[k][1].record()
        jf.apply_bc_vec(res.reshape(-1), sol.reshape(-1), prob)
        ev[k][2].record()
        jf.get_A(prob)
        ev[k][3].record()
    barrier()
    t = [sum(e[i].elapsed_time(e[i + 1]) for e in ev) / steps for i in range(3)]
    return ev[0][0].elapsed_time(ev[-1][3]) / steps, t[0], t[1], t[2]


def measured_traffic(args, world):
    """DRAM bytes per assembly step from the committed ncu capture of this code (profiles/r02_traffic.json, written by
    tools/ncu_traffic.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`): only valid for the workload, size
    and assembly mode it was captured on; None otherwise (a bench run cannot host ncu itself)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_traffic.json")
    try:
        t = json.load(open(path))
    except (OSError, ValueError):
        return None
    for entry in t.get("captures", []):
        if (entry.get("workload") == args.workload and entry.get("size") == args.size and world == 1
                and entry.get("mode") == os.environ.get("FEM_ASSEMBLY", "staged")):
            return entry["assembly_bytes_per_step"]
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=200, help="cells per direction of the FIXED mesh (200: the north star's 24M-DOF mesh)")
    ap.add_argument("--cfg2-size", type=int, default=100, help="edge of the cfg 2 sub-run on one GPU (0 = skip)")
    ap.add_argument("--no-solve", action="store_true", help="skip the Jacobi-CG solve (it is part of every default run)")
    ap.add_argument("--ref-size", type=int, default=64,
                    help="edge length of the CPU sample (cells per direction): 64^3 is about 10-20 s of host work")
    ap.add_argument("--bicgstab-iters", type=int, default=0, help="also time this many Jacobi-BiCGSTAB iterations (cg_solve.bicgstab)")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--newton", action="store_true",
                    help="also time the full Newton solve with per-iteration re-assembly (cfg 3 with --workload neohookean)")
    ap.add_argument("--adjoint", action="store_true",
                    help="also time cfg 5's chain: forward solve + implicit adjoint + per-element density gradient (--workload simp)")
    ap.add_argument("--workload", default="elasticity", choices=["elasticity", "neohookean", "hex27", "simp"],
                    help="elasticity = cfg 2 (the headline, default); neohookean = cfg 3's law on the same mesh; "
                         "hex27 = cfg 4 (use --size 60)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("BENCH_WATCHDOG_S", "900")), exit=True)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import jax_fem_b200 as jf
    from jax_fem_b200 import _lib
    from jax_fem_b200.solver import jax_solve

    # stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL prints its version banner there)
    # is sent to stderr, and the line goes out through a private copy of the original descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    t0 = time.perf_counter()
    log(f"building problem {args.size}^3 (fixed mesh), {world} GPU(s)")
    comm = None
    if world > 1:
        from jax_fem_b200.distributed import NcclComm
        comm = NcclComm()                        # the library's own NCCL communicator (csrc/dist.cu)
    prob, sharded = build_problem(args.size, world, comm, args.workload)
    setup_s = time.perf_counter() - t0
    fe = prob.fes[0]
    n_local, nnz = prob.num_total_dofs_all_vars, prob.plan.nnz
    n = sharded.n_owned if sharded else n_local               # dofs this rank owns (what it is credited for)
    log(f"problem built in {setup_s:.1f}s: {n} owned dofs (+{n_local - n} ghost), local nnz {prob.plan.nnz}")
    own_frac = n / n_local
    cells_per_gpu = int(prob.num_cells * own_frac)
    b_asm, b_spmv = algorithmic_bytes(n, int(nnz * own_frac), int(prob.num_cells * own_frac), int(fe.num_total_nodes * own_frac),
                                      nodes_per_cell=fe.num_nodes)
    if args.workload == "simp":
        # theta = 0.5 + 0.1 U(-1, 1) per cell (default_rng(0) on the GLOBAL cell numbering), constant over a cell's points
        n_cells_global = prob.num_cells if not sharded else int(sharded.part.local_cells.max()) + 1
        if sharded:
            t = torch.tensor([n_cells_global], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            n_cells_global = int(t.item())
        theta_g = 0.5 + 0.1 * np.random.default_rng(0).uniform(-1, 1, n_cells_global)
        local_cells = sharded.part.local_cells if sharded else np.arange(n_cells_global)
        prob.internal_vars = [torch.from_numpy(np.repeat(theta_g[local_cells][:, None], fe.num_quads, axis=1)).to(dev)]
    # the same global field on every partition (seeded on the global node numbering)
    n_nodes_global = fe.num_total_nodes if not sharded else int(sharded.part.node_ranges[-1])
    sol_global = 1e-3 * np.random.default_rng(0).standard_normal((n_nodes_global, 3))
    sol_host = torch.from_numpy(sol_global[sharded.part.l2g] if sharded else sol_global).pin_memory()
    del sol_global
    sol = sol_host.to(dev)
    res_host = torch.empty(n_local, dtype=torch.float64).pin_memory()

    def step(sol_dev):
        res = prob.newton_update([sol_dev])[0]
        res_vec = jf.apply_bc_vec(res.reshape(-1), sol_dev.reshape(-1), prob)
        A = jf.get_A(prob)
        return res_vec, A

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        res_vec, A = step(sol)
    barrier()
    log("warm-up done")

    # ---- timed region: device-resident inputs -------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, t_elem, t_bc, t_gather = time_assembly(prob, sol, args.steps, jf, torch, barrier)
    res_vec, A = step(sol)

    # HEX27: the default element step takes the exact affine-cell pass wherever a cell's geometry map is affine (every cell of
    # this box mesh); the same assembly with the pass switched off (FEM_HEX27_AFFINE=0: FP64 tensor-core kernel for every cell,
    # what a mesh of curved cells costs) is timed beside it so that the line carries both
    hex27_info = None
    if args.workload == "hex27" and world == 1:
        n_general = int(len(prob.hex27_general_cells()))
        os.environ["FEM_HEX27_AFFINE"] = "0"
        try:
            pg, _ = build_problem(args.size, workload="hex27")
        finally:
            os.environ.pop("FEM_HEX27_AFFINE")
        for _ in range(args.warmup):
            pg.newton_update([sol])
            jf.get_A(pg)
        g_ms = time_assembly(pg, sol, max(2, args.steps // 2), jf, torch, barrier)[0]
        del pg
        torch.cuda.empty_cache()
        hex27_info = {"cells": int(prob.num_cells), "cells_on_the_affine_pass": int(prob.num_cells) - n_general,
                      "cells_on_the_general_dmma_kernel": n_general, "ms_per_step": ms_step,
                      "ms_per_step_general_kernel_for_every_cell": g_ms,
                      "note": "affine pass: K_e = E detJ J^-T Ghat J^-1 from reference Gram tables, exact up to rounding, chosen per cell "
                              "on every call by an exact affinity test (DESIGN.md 4.1)"}
        log(f"hex27: {hex27_info}")

    # ---- e2e: public API from pinned host buffers -------------------------------------------------------
    # Every step copies its input from pinned host memory and its result (the residual with Dirichlet rows) back.
    # The copies run on two more streams, one per direction (PCIe is full duplex), and are software-pipelined against the
    # compute: the H2D of step k+1 goes into the other input buffer as soon as step k-1 has consumed it (so it overlaps the
    # whole of step k), the D2H of step k overlaps get_A of step k and the start of step k+1.  All of it is inside the timed
    # region; a step is only as fast as the slower of its kernels and its 8 B/DOF in each direction over PCIe.
    main_s, h2d_s, d2h_s = torch.cuda.current_stream(), torch.cuda.Stream(), torch.cuda.Stream()
    sol_bufs = [torch.empty_like(sol), torch.empty_like(sol)]

    def e2e_pass(n_steps):
        """-> (ms per step, last result, last matrix)"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        res_ready, d2h_done = torch.cuda.Event(), torch.cuda.Event()
        barrier()
        e0.record()
        h2d_s.wait_stream(main_s)
        d2h_s.wait_stream(main_s)
        with torch.cuda.stream(h2d_s):
            sol_bufs[0].copy_(sol_host, non_blocking=True)
            h2d_done[0].record(h2d_s)
        for k in range(n_steps):
            if k + 1 < n_steps:
                with torch.cuda.stream(h2d_s):
                    if k >= 1:
                        h2d_s.wait_event(consumed[(k + 1) % 2])       # step k-1 was the last reader of that buffer
                    sol_bufs[(k + 1) % 2].copy_(sol_host, non_blocking=True)
                    h2d_done[(k + 1) % 2].record(h2d_s)
            main_s.wait_event(h2d_done[k % 2])
            res = prob.newton_update([sol_bufs[k % 2]])[0]
            res_vec = jf.apply_bc_vec(res.reshape(-1), sol_bufs[k % 2].reshape(-1), prob)
            consumed[k % 2].record(main_s)
            res_ready.record(main_s)
            res_vec.record_stream(d2h_s)
            with torch.cuda.stream(d2h_s):
                d2h_s.wait_event(res_ready)
                res_host.copy_(res_vec, non_blocking=True)
                d2h_done.record(d2h_s)
            A = jf.get_A(prob)
        main_s.wait_event(d2h_done)
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / n_steps, res_vec, A

    e2e_pass(max(args.warmup, 3))            # untimed: the caching allocator settles on the blocks the pipelined copies need
    e2e_ms, res_vec, A = e2e_pass(args.steps)
    # the PCIe rates these copies get on this box (outside the timed region): with ~8 B/DOF each way they bound e2e
    p0, p1, p2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    p0.record()
    for _ in range(3):
        sol_bufs[0].copy_(sol_host, non_blocking=True)
    p1.record()
    for _ in range(3):
        res_host.copy_(res_vec, non_blocking=True)
    p2.record()
    torch.cuda.synchronize()
    pcie = {"h2d_gbs": 3 * n_local * 8 / (p0.elapsed_time(p1) * 1e6), "d2h_gbs": 3 * n_local * 8 / (p1.elapsed_time(p2) * 1e6)}
    del sol_bufs
    log(f"assembly {ms_step:.3f} ms/step (element {t_elem:.3f}, bc {t_bc:.3f}, gather {t_gather:.3f}); e2e {e2e_ms:.3f} ms")

    # ---- SpMV (the CG kernel) on the assembled matrix; with N > 1 it includes the halo exchange of x ---------
    x = torch.from_numpy(np.random.default_rng(0).standard_normal(n_local)).to(dev)
    y = torch.empty_like(x)
    lib = _lib.load()
    indptr, indices, data = A.getValuesCSR()
    ws = torch.zeros(lib.fem_krylov_workspace(n_local), dtype=torch.float64, device=dev)

    def spmv():
        if sharded:
            sharded.halo.update(x)
            _lib.check(lib.fem_dcg_spmv_dot(n, n_local, _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data), 3,
                                            _lib.ptr(prob.plan.brow_ptr), _lib.ptr(prob.plan.bcol), _lib.ptr(x), _lib.ptr(y),
                                            0, _lib.ptr(ws), _lib.stream_ptr()))
        else:
            A.mult(x, y)
    for _ in range(10 if sharded else 3):          # the first halo exchanges also set up the NCCL peer connections
        spmv()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50 if sharded else 20
    barrier()
    s0.record()
    for _ in range(reps):
        spmv()
    s1.record()
    barrier()
    spmv_ms = s0.elapsed_time(s1) / reps
    del x, y, ws
    log(f"spmv {spmv_ms:.3f} ms = {b_spmv / spmv_ms / 1e6:.0f} GB/s per GPU")

    # ---- the Jacobi-CG solve of the linear problem (tol = atol = 1e-10), every run ------------------------------------
    solve = None
    if not args.no_solve and args.workload in ("elasticity", "simp"):
        dofs = torch.zeros(n_local, dtype=torch.float64, device=dev)
        r0 = jf.apply_bc_vec(prob.newton_update([dofs.reshape(-1, 3)])[0].reshape(-1), dofs, prob)
        A0 = jf.get_A(prob)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sharded:
            from jax_fem_b200.distributed import distributed_cg
            run_cg = lambda **kw: distributed_cg(A0, -r0, torch.zeros_like(dofs), sharded.part, sharded.halo, comm, 3, **kw)
        else:
            run_cg = lambda **kw: jax_solve(A0, -r0, torch.zeros_like(dofs), True, method='cg', return_info=True, **kw)
        try:
            run_cg(maxiter=5)                         # warm-up: workspace, occupancy queries, NCCL channels of the all-reduce
        except AssertionError:
            pass                                      # 5 iterations do not meet the reference's err < 0.1 post-check
        barrier()
        w0 = time.perf_counter()
        c0.record()
        xs, info = run_cg()
        c1.record()
        barrier()
        cs_wall = time.perf_counter() - w0
        t_cg = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
        chk = torch.stack([xs[:n].abs().sum(), (xs[:n] ** 2).sum()])
        if world > 1:
            dist.all_reduce(t_cg, op=dist.ReduceOp.MAX)
            dist.all_reduce(chk, op=dist.ReduceOp.SUM)
        cg_ms = float(t_cg.item())
        its = max(int(info['iterations']), 1)
        log(f"cg: {info}, {cg_ms / 1e3:.2f}s device time ({cs_wall:.2f}s wall)")
        solve = {"method": "Jacobi-CG, tol = atol = 1e-10" + (", whole loop in the library: 1 halo exchange (ncclSend/ncclRecv) + 2 "
                 "all-reduces per iteration" if sharded else ", device-resident scalars"),
                 "iterations": int(info['iterations']), "seconds": cg_ms / 1e3, "ms_per_iteration": cg_ms / its,
                 "err": info.get('err'), "final_rr": info.get('rr'),
                 "solution_l1": float(chk[0]), "solution_l2": float(chk[1].sqrt()),
                 "halo_bytes_per_rank_per_exchange": sharded.halo.bytes_per_exchange if sharded else 0,
                 "effective_spmv_gbs": world * b_spmv / (cg_ms * 1e-3 / its) / 1e9}
        # optional: K iterations of Jacobi-BiCGSTAB (the reference's default solver and its adjoint solver) on the same system
        if args.bicgstab_iters > 0:
            if sharded:
                from jax_fem_b200.distributed import distributed_bicgstab
                run_bi = lambda k: distributed_bicgstab(A0, -r0, torch.zeros_like(dofs), sharded.part, sharded.halo, comm, 3, maxiter=k)
            else:
                run_bi = lambda k: jax_solve(A0, -r0, torch.zeros_like(dofs), True, method='bicgstab', return_info=True, maxiter=k)

            def guarded(k):
                try:
                    return run_bi(k)[1]
                except AssertionError:                # a capped run does not meet the reference's err < 0.1 post-check
                    return {"iterations": k}
            guarded(5)
            barrier()
            c0.record()
            binfo = guarded(args.bicgstab_iters)
            c1.record()
            barrier()
            t_bi = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_bi, op=dist.ReduceOp.MAX)
            kb = max(int(binfo.get('iterations', args.bicgstab_iters)), 1)
            solve["bicgstab"] = {"iterations_timed": kb, "ms_per_iteration": float(t_bi.item()) / kb,
                                 "note": "capped run (2 SpMV, 2 halo exchanges and 4 all-reduces per iteration), includes the final true-residual check"}
            log(f"bicgstab: {solve['bicgstab']}")
        del A0, r0, xs, dofs
    newton = None
    if args.newton:
        barrier()
        c0 = time.perf_counter()
        if sharded:
            sharded.solve()
            ninfo = sharded.last_info
            ninfo = {"newton_iterations": ninfo["newton_iterations"], "krylov_iterations": ninfo["cg_iterations"],
                     "residuals": ninfo["residuals"]}
        else:
            jf.solver(prob, {"jax_solver": {"method": "cg"}})
            li = prob.last_newton_info
            ninfo = {"newton_iterations": li["iterations"], "final_residual": li["res_val"]}
        barrier()
        ns = time.perf_counter() - c0
        log(f"newton: {ninfo}, {ns:.2f}s")
        newton = dict(ninfo, seconds=ns, method="Newton (tol 1e-6, rel 1e-8), tangent re-assembled every iteration, Jacobi-CG 1e-10"
                      + (", cells sharded in x-slabs" if sharded else ""))
    adjoint = None
    if args.adjoint:
        from jax_fem_b200.solver import implicit_vjp
        barrier()
        c0 = time.perf_counter()
        if sharded:
            u = sharded.solve_linear()
            it_fwd = sharded.last_info["iterations"]
            f_ext = prob._f_ext if prob._f_ext is not None else torch.zeros_like(u)
            Jloc = -(f_ext[:sharded.part.n_owned] * u[:sharded.part.n_owned]).sum().reshape(1)
            comm.allreduce(Jloc)
            barrier()
            c1 = time.perf_counter()
            grad = sharded.adjoint_gradient(u, -f_ext)
            it_adj = sharded.last_info["iterations"]
            rr_adj = sharded.last_info["rr"]
            err_adj = sharded.last_info.get("err")
            # compliance is self-adjoint: lambda = u on the free rows -- a check that needs no second solve
            lam, rows = sharded.last_lambda, prob.bc_data()[0].long()
            d = (lam - u.reshape(-1))
            d[rows] = 0.
            sa = torch.stack([(d[:n] ** 2).sum(), (u.reshape(-1)[:n] ** 2).sum()])
            comm.allreduce(sa)
            self_adj = float((sa[0] / sa[1]).sqrt())
        else:
            u = jf.solver(prob, {"jax_solver": {"method": "cg"}})[0]
            it_fwd = None
            Jloc = -(prob._f_ext * u).sum().reshape(1)
            torch.cuda.synchronize()
            c1 = time.perf_counter()
            grad = implicit_vjp(prob, [u], None, [-prob._f_ext], {"jax_solver": {}})
            it_adj = rr_adj = err_adj = self_adj = None
        barrier()
        c2 = time.perf_counter()
        gsum = grad.sum(1)
        # global checksums of dJ/dtheta: every cell once (a cell in a ghost layer is counted by the owner of its first node)
        mine = torch.ones_like(gsum, dtype=torch.bool) if not sharded else \
            torch.from_numpy(sharded.part.cells_local[:, 0] < sharded.part.n_owned).to(dev)
        gl = torch.stack([gsum[mine].abs().sum(), (gsum[mine] ** 2).sum(), gsum[mine].sum()])
        if sharded:
            comm.allreduce(gl)
        log(f"adjoint: J={float(Jloc):.6e}, forward {c1 - c0:.2f}s ({it_fwd} its), adjoint+gradient {c2 - c1:.2f}s ({it_adj} its), "
            f"|dJ/dtheta|_max={float(gsum.abs().max()):.4e}, ||lambda - u|| / ||u|| = {self_adj}")
        adjoint = {"objective": "compliance int t.u ds", "J": float(Jloc), "forward_seconds": c1 - c0, "forward_iterations": it_fwd,
                   "adjoint_seconds": c2 - c1, "adjoint_iterations": it_adj, "adjoint_final_rr": rr_adj, "adjoint_true_residual": err_adj,
                   "self_adjoint_check_rel": self_adj, "grad_abs_max": float(gsum.abs().max()),
                   "grad_l1_global": float(gl[0]), "grad_l2_global": float(gl[1].sqrt()), "grad_sum_global": float(gl[2]),
                   "method": "forward Jacobi-CG 1e-10; adjoint A^T lambda = dJ/du by Jacobi-BiCGSTAB 1e-10; gradient -lambda^T dc/dtheta per cell"
                             + (", cells sharded in x-slabs, whole Krylov loops in the library" if sharded else "")}
    clocks = sampler.stop() if rank == 0 else None

    # ---- cfg 2 itself (BASELINE.json configs[1], 100^3) on one GPU, for continuity with round 1 --------------------------
    cfg2 = None
    if world == 1 and args.cfg2_size and args.cfg2_size != args.size and args.workload == "elasticity":
        del A, res_vec, sol, prob
        torch.cuda.empty_cache()
        p2, _ = build_problem(args.cfg2_size)
        f2 = p2.fes[0]
        s2 = torch.from_numpy(1e-3 * np.random.default_rng(0).standard_normal((f2.num_total_nodes, 3))).to(dev)
        for _ in range(args.warmup):
            p2.newton_update([s2])
            jf.get_A(p2)
        ms2, te2, tb2, tg2 = time_assembly(p2, s2, args.steps, jf, torch, barrier)
        ba2, _ = algorithmic_bytes(p2.num_total_dofs_all_vars, p2.plan.nnz, p2.num_cells, f2.num_total_nodes)
        peaks2, _ = measured_peaks()
        cfg2 = {"workload": f"HEX8 box {args.cfg2_size}^3 linear elasticity (cfg 2)", "n_dofs": p2.num_total_dofs_all_vars,
                "ms_per_step": ms2, "value": p2.num_total_dofs_all_vars / (ms2 * 1e-3), "unit": "DOF/s",
                "kernels_ms": {"element_kernel+gather_residual": te2, "apply_bc_vec": tb2, "gather_csr": tg2},
                "roofline_frac": ba2 / ((te2 + tb2 + tg2) * 1e-3) / 1e9 / float(peaks2["hbm_gbs"])}
        log(f"cfg 2 ({args.cfg2_size}^3): {ms2:.3f} ms/step")
    # max over ranks (device times)
    t = torch.tensor([ms_step, e2e_ms, spmv_ms, t_elem, t_gather, t_bc], dtype=torch.float64, device=dev)
    n_total = torch.tensor([float(n)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_total, op=dist.ReduceOp.SUM)
    n_total = int(n_total.item())
    ms_step, e2e_ms, spmv_ms, t_elem, t_gather, t_bc = [float(v) for v in t.tolist()]

    if rank == 0:
        peaks, peak_src = measured_peaks()
        hbm = float(peaks["hbm_gbs"])
        value = n_total / (ms_step * 1e-3)
        asm_kernel_ms = t_elem + t_gather + t_bc
        achieved = b_asm / (asm_kernel_ms * 1e-3) / 1e9
        spmv_gbs = b_spmv / (spmv_ms * 1e-3) / 1e9
        mode = os.environ.get("FEM_ASSEMBLY", "staged")
        nx = "4s x s x 2s" if args.workload == "simp" else f"{args.size}x{args.size}x{args.size}"
        line = {
            "metric": "assembled DOFs/s (residual+Jacobian), HEX8 linear elasticity", "value": value, "unit": "DOF/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": {"neohookean": "[NON-HEADLINE: Neo-Hookean E=10 nu=0.3] ", "hex27": "[NON-HEADLINE: HEX27, 216-point quadrature] ",
                                    "simp": "[NON-HEADLINE: SIMP Emax=70e3 Emin=70 p=3, theta = 0.5 + 0.1 U(-1,1)] ",
                                    "elasticity": ""}[args.workload] +
                                   (f"HEX8 cantilever 2 x 0.5 x 1, {nx} cells with s = {max(2, args.size // 2)}, u=0 on x=0, traction on x=2 (cfg 5)"
                                    if args.workload == "simp" else
                                    f"HEX8 box {nx} linear elasticity E=70e3 nu=0.3, u=0 on x=0, traction on x=1: the north star's "
                                    f">= 24M-DOF mesh (cfg 2's problem on cfg 3's 200^3 mesh), FIXED for every GPU count"),
                       "n_dofs_total": n_total, "n_dofs_per_gpu": n, "nnz_per_gpu": int(nnz * own_frac), "cells_per_gpu": cells_per_gpu,
                       "per_gpu": "whole mesh" if world == 1 else f"x-slab of the fixed mesh (1/{world} of the nodes) + 1 ghost cell layer per interface; "
                                  "assembly needs no communication; SpMV = halo ncclSend/ncclRecv with <= 2 neighbours; Krylov dots = ncclAllReduce",
                       "assembly_mode": mode,
                       "l2": "inputs larger than L2 (element matrices + CSR >> 126 MB), no flush needed"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": measured_traffic(args, world), "traffic_source": "profiles/r02_traffic.json (ncu dram__bytes of this code, same workload/size/mode) or null",
                         "peak_source": peak_src,
                         "kernel": "assembly = element_kernel + gather_residual + apply_bc_vec + gather_csr",
                         "algorithmic_bytes": b_asm, "kernel_ms": asm_kernel_ms,
                         "kernels_ms": {"element_kernel+gather_residual": t_elem, "apply_bc_vec": t_bc, "gather_csr": t_gather}},
            "roofline_spmv": {"bound": "hbm", "achieved": spmv_gbs, "peak": hbm, "unit": "GB/s", "frac": spmv_gbs / hbm,
                              "kernel_bytes": int(8 * nnz * own_frac + 4 * nnz * own_frac / 9 + 24 * n),
                              "kernel_gbs": (8 * nnz * own_frac + 4 * nnz * own_frac / 9 + 24 * n) / (spmv_ms * 1e-3) / 1e9,
                              "frac_of_kernel_bytes": (8 * nnz * own_frac + 4 * nnz * own_frac / 9 + 24 * n) / (spmv_ms * 1e-3) / 1e9 / hbm,
                              "algorithmic_bytes": b_spmv, "kernel_ms": spmv_ms, "kernel": "spmv_block_fused_kernel<3,8,0> (node-block CSR: one column index per 3x3 block, 8 lanes per node)"
                                        + (" + halo exchange" if world > 1 else ""),
                              "note": "algorithmic bytes are those of scalar CSR (12 B/nnz, SURVEY 8d); the kernel reads 8.44 B/nnz, "
                                      "so `frac` can exceed 1; `frac_of_kernel_bytes` is the fraction on the kernel's own traffic"},
            "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": "DOF/s", "h2d_bytes_per_step": n_local * 8,
                    "d2h_bytes_per_step": n_local * 8, "ms_per_step": e2e_ms,
                    "pcie_gbs_this_box": pcie,
                    "pcie_bound_ms": max(n_local * 8 / (pcie["h2d_gbs"] * 1e6), n_local * 8 / (pcie["d2h_gbs"] * 1e6))},
            "gpu_launches": args.steps * world * (6 if args.workload == "hex27" and os.environ.get("FEM_HEX27_AFFINE", "1") != "0" else 4), "clocks": clocks, "setup_s": setup_s,
        }
        if solve:
            line["cg_solve"] = solve
        if hex27_info:
            line["hex27"] = hex27_info
        if newton:
            line["newton_solve"] = newton
        if adjoint:
            line["adjoint"] = adjoint
        if cfg2:
            line["cfg2_100cube"] = cfg2
        if not args.no_cpu_baseline and world == 1:
            log(f"timing the CPU oracle on {args.ref_size}^3")
            v, info = cpu_oracle_assembly(args.ref_size)
            line["cpu_baseline"] = {"value": v, "unit": "DOF/s", "cores": info["threads"], "kind": "port",
                                    "sample": f"{args.ref_size}^3 HEX8 cells ({info['n_dofs']} DOF) of the same workload: element "
                                              f"{info['element_s']:.2f}s + COO->CSR+BC {info['coo_to_csr_s']:.2f}s; "
                                              f"SpMV {info['spmv_gbs']:.1f} GB/s"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
