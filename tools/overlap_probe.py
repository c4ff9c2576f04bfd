"""Upper bound of running the element kernel and the CSR gather concurrently (two streams, no dependency: the gather
reads the previous step's element tangents).  Timing experiment only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jax_fem_b200 as jf
from bench import build_problem

prob, _ = build_problem(100)
dev = prob.device
sol = torch.from_numpy(1e-3 * np.random.default_rng(0).standard_normal((prob.fes[0].num_total_nodes, 3))).to(dev)
for _ in range(3):
    prob.newton_update([sol]); A = jf.get_A(prob)
torch.cuda.synchronize()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(concurrent, reps=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        if concurrent:
            s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s1):
                prob._run_element_kernel(sol, jac=True)
            with torch.cuda.stream(s2):
                jf.get_A(prob)
            torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        else:
            prob._run_element_kernel(sol, jac=True); jf.get_A(prob)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("sequential %.3f ms, concurrent %.3f ms" % (run(False), run(True)))
