"""A/B timing of the assembly modes (FEM_ASSEMBLY = ring | staged | fused) on one GPU: python tools/ab_assembly.py [size] [modes]."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jax_fem_b200 as jf
from bench import build_problem

size = int(sys.argv[1]) if len(sys.argv) > 1 else 100
modes = sys.argv[2].split(',') if len(sys.argv) > 2 else ['staged', 'ring']
workload = sys.argv[3] if len(sys.argv) > 3 else 'elasticity'
steps = int(os.environ.get('AB_STEPS', '10'))
out = {}
ref = None
for mode in modes:
    os.environ['FEM_ASSEMBLY'] = mode
    prob, _ = build_problem(size, workload=workload)
    fe = prob.fes[0]
    sol = torch.from_numpy(1e-3 * np.random.default_rng(0).standard_normal((fe.num_total_nodes, 3))).cuda()
    if workload == 'simp':
        th = 0.5 + 0.1 * np.random.default_rng(0).uniform(-1, 1, prob.num_cells)
        prob.internal_vars = [torch.from_numpy(np.repeat(th[:, None], 8, axis=1)).cuda()]
    t0 = time.perf_counter()
    for _ in range(3):
        res = prob.newton_update([sol])[0]
        rv = jf.apply_bc_vec(res.reshape(-1), sol.reshape(-1), prob)
        A = jf.get_A(prob)
    torch.cuda.synchronize()
    prob.check_assembly_status()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        res = prob.newton_update([sol])[0]
        rv = jf.apply_bc_vec(res.reshape(-1), sol.reshape(-1), prob)
        A = jf.get_A(prob)
    e1.record()
    torch.cuda.synchronize()
    prob.check_assembly_status()
    ms = e0.elapsed_time(e1) / steps
    info = {'ms': ms}
    if mode == 'ring':
        sp = prob.stage_plan
        info.update(ring_rows=sp.ring_rows, staging_MB=sp.staging_bytes / 2 ** 20, spill=sp.spill_fraction)
    d = A.data
    if ref is None:
        ref = (d.clone(), rv.clone())
    else:
        info['data_relmax_vs_first'] = float((d - ref[0]).abs().max() / ref[0].abs().max())
        info['res_relmax_vs_first'] = float((rv - ref[1]).abs().max() / ref[1].abs().max())
    out[mode] = info
    print(mode, json.dumps(info), flush=True)
    del prob, A, d
    torch.cuda.empty_cache()
