"""Nonlinear / linear solvers and the implicit adjoint, with the reference's API.

Mirror of jax_fem/solver.py: solver (:1112-1356), newton_step (:390-421), get_A (:540-553),
apply_bc_vec / assign_bc / copy_bc (:290-363), linear_solver (:261-284), jax_solve (:63-92),
implicit_vjp (:1362-1418), ad_wrapper (:1421-1455), option resolution (:1079-1106).

One linear back-end exists: the device-resident Jacobi-preconditioned BiCGSTAB / CG of
csrc/krylov.cu (``jax_solver``; option ``method`` in {'bicgstab' (reference default), 'cg'}) plus
the documented ``custom_solver`` hook.  ``petsc_solver`` / ``amgx_solver`` / ``spsolve_solver``,
arc-length and dynamic relaxation are outside the hot path and raise NotImplementedError.
"""
import math
import threading
import time

import numpy as np
import torch

from . import _lib, laws, logger
from .sparse import CSRMatrix

################################################################################
# logging helpers (same buckets as the reference: local_assembly / global_matrix / linear)


class _Stopwatch:
    """Wall time per bucket of a Newton solve; the buckets are the ones the reference reports (local_assembly: element
    kernels, global_matrix: CSR assembly, linear: Krylov solve)."""
    BUCKETS = ('local_assembly', 'global_matrix', 'linear')

    def __init__(self):
        self.seconds = dict.fromkeys(self.BUCKETS, 0.)
        self._t0 = time.perf_counter()

    def add(self, bucket, dt):
        self.seconds[bucket] += dt

    def report_iteration(self, k, res, rel, spent):
        parts = ", ".join(f"{name} {spent[name]:.3f} s" for name in self.BUCKETS if name in spent)
        logger.info("Newton %d: |res| = %.3g (%.3g of the first one); %s", k, res, rel, parts)

    def report_total(self, n_iters):
        wall = time.perf_counter() - self._t0
        share = "; ".join(f"{name} {dt:.3f} s ({100. * dt / wall if wall > 0 else 0.:.0f}%)" for name, dt in self.seconds.items())
        logger.info("Newton finished after %d iteration(s) in %.3f s: %s", n_iters, wall, share)


def _sync_time():
    torch.cuda.synchronize()
    return time.perf_counter()


################################################################################
# small device helpers

def _norm(v):
    """L2 norm through the library's deterministic dot."""
    ws = torch.empty(2048, dtype=torch.float64, device=v.device)
    out = (_lib.ctypes.c_double * 1)()
    _lib.check(_lib.load().fem_dot(v.numel(), _lib.ptr(v), _lib.ptr(v), out, _lib.ptr(ws), _lib.stream_ptr()))
    return math.sqrt(out[0])


################################################################################
# Dirichlet boundary conditions ("row elimination")

def apply_bc_vec(res_vec, dofs, problem, scale=1.):
    """res[bc rows] = dofs[bc rows] - value*scale  (solver.py:290-304); returns a new vector."""
    rows, vals, _ = problem.bc_data()
    res = res_vec.reshape(-1).clone()
    _lib.check(_lib.load().fem_apply_bc_vec(rows.numel(), _lib.ptr(rows), _lib.ptr(vals), float(scale),
                                            _lib.ptr(dofs.reshape(-1).contiguous()), _lib.ptr(res), _lib.stream_ptr()))
    return res


def apply_bc(res_fn, problem, scale=1.):
    def res_fn_bc(dofs):
        return apply_bc_vec(res_fn(dofs), dofs, problem, scale)
    return res_fn_bc


def assign_bc(dofs, problem):
    rows, vals, _ = problem.bc_data()
    out = dofs.reshape(-1).clone()
    out[rows.long()] = vals
    return out


def assign_ones_bc(dofs, problem):
    rows, _, _ = problem.bc_data()
    out = dofs.reshape(-1).clone()
    out[rows.long()] = 1.
    return out


def assign_zeros_bc(dofs, problem):
    rows, _, _ = problem.bc_data()
    out = dofs.reshape(-1).clone()
    out[rows.long()] = 0.
    return out


def copy_bc(dofs, problem):
    rows, _, _ = problem.bc_data()
    out = torch.zeros_like(dofs.reshape(-1))
    out[rows.long()] = dofs.reshape(-1)[rows.long()]
    return out


def get_flatten_fn(fn_sol_list, problem):
    def fn_dofs(dofs):
        return torch.cat([v.reshape(-1) for v in fn_sol_list(problem.unflatten_fn_sol_list(dofs))])
    return fn_dofs


################################################################################
# Tangent stiffness matrix

def get_A(problem):
    """Assemble the element tangents of the last newton_update into CSR, Dirichlet rows zeroed with unit
    diagonal and the full pattern kept (solver.py:469-553).  Returns a device CSRMatrix."""
    if hasattr(problem, 'P_mat'):
        raise NotImplementedError("P_mat (multipoint constraints) is outside the B200 hot path")
    if problem._last_sol is None:
        raise RuntimeError("get_A() needs problem.newton_update(sol_list) first")
    p = problem.plan
    fe = problem.fes[0]
    data = problem.assembled_values()          # fused owner-computes assembly already produced the CSR values
    if data is not None:
        return CSRMatrix(p, data)
    Ke = problem.staged_tangents()
    emeta = problem.entry_meta()
    data = torch.empty(p.nnz, dtype=torch.float64, device=problem.device)
    if getattr(problem, '_Ke_tiles', False):       # tile-major rows of fem_element_tiles; isotropic map applied after the sum
        _lib.check(_lib.load().fem_gather_csr_tiles(p.n_gather_blocks, _lib.ptr(p.gdesc), _lib.ptr(emeta), _lib.ptr(p.src),
                                                    _lib.ptr(Ke), _lib.ptr(data), problem._Ke_post, _lib.stream_ptr()))
    else:
        _lib.check(_lib.load().fem_gather_csr(fe.vec, fe.num_nodes, p.n_gather_blocks, _lib.ptr(p.gdesc), _lib.ptr(emeta),
                                              _lib.ptr(p.src), _lib.ptr(Ke), _lib.ptr(data), _lib.stream_ptr()))
    problem.add_face_tangent(problem._last_sol, data)      # registered u-dependent surface maps (problem.py:456-458)
    return CSRMatrix(p, data)


################################################################################
# Linear solvers

_workspaces = {}


def _krylov_workspace(n, device):
    """Scratch of the device-resident Krylov solvers (scalars, ticket counter, 8 vectors), one per (thread, stream, size):
    concurrent solves from different threads or streams never share it.  Only the most recent size of a thread is kept."""
    key = (threading.get_ident(), torch.cuda.current_stream(device).cuda_stream, device)
    hit = _workspaces.get(key)
    if hit is None or hit[0] != n:
        hit = (n, torch.zeros(_lib.load().fem_krylov_workspace(n), dtype=torch.float64, device=device))
        _workspaces[key] = hit
    return hit[1]


def jax_solve(A, b, x0, precond, method='bicgstab', tol=1e-10, atol=1e-10, maxiter=10000, check_every=25,
              return_info=False):
    """Jacobi-preconditioned Krylov solve on the device; stopping rule and post-check of solver.py:63-92."""
    n = A.getSize()[0]
    indptr, indices, data = A.getValuesCSR()
    b = b.reshape(-1).contiguous()
    x = torch.zeros_like(b) if x0 is None else x0.reshape(-1).clone().contiguous()
    diag = A.diagonal() if precond else None
    ws = _krylov_workspace(n, b.device)
    info = (_lib.ctypes.c_double * 4)()
    fn = {'bicgstab': _lib.load().fem_pbicgstab, 'cg': _lib.load().fem_pcg}[method]
    plan = getattr(A, 'plan', None)             # FE matrices carry their node-block structure; others use plain CSR
    vec, brow_ptr, bcol = (plan.vec, plan.brow_ptr, plan.bcol) if plan is not None else (1, None, None)
    _lib.check(fn(n, _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data), vec, _lib.ptr(brow_ptr), _lib.ptr(bcol),
                  _lib.ptr(diag), _lib.ptr(b), _lib.ptr(x), float(tol), float(atol), int(maxiter), int(check_every),
                  _lib.ptr(ws), info, _lib.stream_ptr()))
    iters, err = int(info[0]), float(info[2])
    logger.debug("jax_solver(%s) - %d iterations, linear solve res = %.3g", method, iters, err)
    assert err < 0.1, f"JAX linear solver failed to converge with err = {err}"
    if return_info:
        return x, {'iterations': iters, 'rr': float(info[1]), 'err': err}
    return x


def linear_solver(A, b, x0, linear_options):
    """Dispatch of solver.py:261-284, restricted to the single device back-end + the custom hook."""
    known = {'jax_solver', 'amgx_solver', 'spsolve_solver', 'petsc_solver', 'custom_solver'}
    if len(linear_options.keys() & known) == 0:
        linear_options['jax_solver'] = {}
    if 'jax_solver' in linear_options:
        opts = linear_options['jax_solver']
        return jax_solve(A, b, x0, opts.get('precond', True), method=opts.get('method', 'bicgstab'),
                         tol=opts.get('tol', 1e-10), atol=opts.get('atol', 1e-10), maxiter=opts.get('maxiter', 10000))
    if 'custom_solver' in linear_options:
        return linear_options['custom_solver'](A, b, x0, linear_options)
    raise NotImplementedError("only 'jax_solver' (device Jacobi-BiCGSTAB/CG) and 'custom_solver' exist on the B200 "
                              "hot path; petsc/amgx/spsolve back-ends are out of scope and do not fall back")


################################################################################
# Newton

# option schema of solver(problem, solver_options) (jax_fem/solver.py:1079-1106): either one nonlinear method holding its
# own configuration, or -- the older flat form -- Newton options and linear back-ends side by side
NONLINEAR_METHODS = ('newton', 'arc_length', 'dynamic_relax')
LINEAR_BACKENDS = ('jax_solver', 'amgx_solver', 'spsolve_solver', 'petsc_solver', 'custom_solver')
NEWTON_SETTINGS = ('tol', 'rel_tol', 'line_search_flag', 'initial_guess')


def _resolve_solver_options(solver_options):
    """-> (nonlinear method, its configuration).  A flat dictionary is read as a Newton configuration whose linear
    back-ends are collected under 'linear'."""
    given = dict(solver_options or {})
    chosen = [name for name in NONLINEAR_METHODS if name in given]
    if len(chosen) > 1:
        raise ValueError(f"solver_options names {len(chosen)} nonlinear methods ({', '.join(chosen)}); exactly one is allowed")
    if chosen:
        cfg = given[chosen[0]]
        if not isinstance(cfg, dict):
            raise ValueError(f"solver_options[{chosen[0]!r}] has to be a dictionary of settings, got {type(cfg).__name__}")
        return chosen[0], cfg
    cfg = {name: given[name] for name in NEWTON_SETTINGS if name in given}
    backends = {name: given[name] for name in LINEAR_BACKENDS if name in given}
    if backends:
        cfg['linear'] = backends
    return 'newton', cfg


def newton_step(problem, res_vec, A, dofs, newton_cfg, timing):
    """Solve A inc = -res with the BC-exact initial guess x0 = assign_bc(0) - copy_bc(dofs) (solver.py:390-421)."""
    b = -res_vec
    rows, vals, _ = problem.bc_data()
    x0 = torch.empty_like(dofs)
    _lib.check(_lib.load().fem_bc_initial_guess(dofs.numel(), rows.numel(), _lib.ptr(rows), _lib.ptr(vals),
                                                _lib.ptr(dofs), _lib.ptr(x0), _lib.stream_ptr()))
    t0 = _sync_time()
    inc = linear_solver(A, b, x0, newton_cfg.get('linear', {}))
    linear_s = _sync_time() - t0
    timing['linear'] += linear_s
    if newton_cfg.get('line_search_flag', False):
        return line_search(problem, dofs, inc), linear_s
    return dofs + inc, linear_s


def line_search(problem, dofs, inc, halvings=3):
    """Backtracking of the reference (solver.py:424-462): try the full Newton step, then halve the step up to three times
    and keep halving only while the norm of the residual (Dirichlet rows included) still decreases.  Uses residual-only
    evaluations (no tangent)."""
    def norm_at(step):
        trial = dofs + step * inc
        res = problem.compute_residual(problem.unflatten_fn_sol_list(trial))[0].reshape(-1)
        return _norm(apply_bc_vec(res, trial, problem))

    step, best = 1., norm_at(1.)
    for attempt in range(halvings):
        candidate = norm_at(0.5 * step)
        logger.debug("line search %d: step %g -> |res| %g, step %g -> |res| %g", attempt, step, best, 0.5 * step, candidate)
        if candidate > best:
            break
        step, best = 0.5 * step, candidate
    return dofs + step * inc


def solver(problem, solver_options={}):
    """Newton solve; returns sol_list = [ (num_total_nodes, vec) CUDA float64 tensor ]."""
    method, cfg = _resolve_solver_options(solver_options)
    if method != 'newton':
        raise NotImplementedError(f"'{method}' is outside the B200 hot path (SURVEY.md 2: rows 9-10)")
    unsupported = set(cfg.get('linear', {})) & {'petsc_solver', 'amgx_solver', 'spsolve_solver'}
    if unsupported:
        raise NotImplementedError(f"linear back-end(s) {sorted(unsupported)} are outside the B200 hot path: only "
                                  "'jax_solver' (device Jacobi-BiCGSTAB/CG) and 'custom_solver' exist; no fallback")
    logger.info("Solving the nonlinear problem...")
    watch = _Stopwatch()
    timing = watch.seconds
    n = problem.num_total_dofs_all_vars
    if 'initial_guess' in cfg:
        dofs = torch.cat([problem._as_sol([g]).reshape(-1) for g in cfg['initial_guess']]).clone()
    else:
        dofs = torch.zeros(n, dtype=torch.float64, device=problem.device)
    rel_tol = cfg.get('rel_tol', 1e-8)
    tol = cfg.get('tol', 1e-6)

    def newton_update_helper(dofs):
        t0 = _sync_time()
        res_list = problem.newton_update(problem.unflatten_fn_sol_list(dofs))
        local_s = _sync_time() - t0
        watch.add('local_assembly', local_s)
        res_vec = apply_bc_vec(res_list[0].reshape(-1), dofs, problem)
        t0 = _sync_time()
        A = get_A(problem)
        global_s = _sync_time() - t0
        watch.add('global_matrix', global_s)
        return res_vec, A, local_s, global_s

    res_vec, A, local_s, global_s = newton_update_helper(dofs)
    res_val = _norm(res_vec)
    res_val_initial = res_val
    rel_res_val = res_val / res_val_initial if res_val_initial > 0 else 0.
    watch.report_iteration(0, res_val, rel_res_val, {'local_assembly': local_s, 'global_matrix': global_s})
    n_iters = 0
    while (rel_res_val > rel_tol) and (res_val > tol):
        n_iters += 1
        dofs, linear_s = newton_step(problem, res_vec, A, dofs, cfg, timing)
        res_vec, A, local_s, global_s = newton_update_helper(dofs)
        res_val = _norm(res_vec)
        rel_res_val = res_val / res_val_initial
        watch.report_iteration(n_iters, res_val, rel_res_val, {'linear': linear_s, 'local_assembly': local_s, 'global_matrix': global_s})
    assert math.isfinite(res_val), "res_val contains NaN, stop the program!"
    assert bool(torch.isfinite(dofs).all()), "dofs contains NaN, stop the program!"
    problem.last_newton_info = {'iterations': n_iters, 'res_val': res_val, 'timing': dict(timing)}
    watch.report_total(n_iters)
    return problem.unflatten_fn_sol_list(dofs)


################################################################################
# Implicit differentiation (adjoint method)

def implicit_vjp(problem, sol_list, params, v_list, adjoint_solver_options):
    """-lambda^T dc/dp with A^T lambda = v (solver.py:1362-1418) for per-quadrature-point parameters.

    Returns d/d(internal_vars[0]) of shape (num_cells, num_quads); the chain through the user's
    ``set_params`` (params -> internal_vars) is left to torch.autograd by ad_wrapper."""
    if params is not None:
        problem.set_params(params)
    problem.newton_update(sol_list)
    A = get_A(problem)
    v_vec = torch.cat([v.reshape(-1) for v in v_list]).contiguous()
    A_T = A.transpose()
    adjoint_vec = linear_solver(A_T, v_vec, None, dict(adjoint_solver_options))
    # c = apply_bc(residual): Dirichlet rows of c do not depend on the parameters
    lam = assign_zeros_bc(adjoint_vec, problem)
    fe = problem.fes[0]
    iv = problem._internal_var()
    if iv is None:
        raise ValueError("implicit_vjp needs a per-quadrature-point parameter in problem.internal_vars")
    grad = torch.empty_like(iv)
    law = problem._law
    if problem.ele_type == 'HEX27':
        _lib.check(_lib.load().fem_hex27_adjoint_param_grad(
            law.law_id, _lib.host_doubles(law.params()), _lib.ptr(problem._points), _lib.ptr(problem._cells), problem.num_cells,
            _lib.ptr(problem._as_sol(sol_list)), _lib.ptr(iv), _lib.ptr(lam), _lib.ptr(problem._ref), fe.num_quads,
            _lib.ptr(grad), _lib.stream_ptr()))
        return grad
    _lib.check(_lib.load().fem_adjoint_param_grad(
        _lib.ELE[problem.ele_type], fe.vec, law.law_id, _lib.host_doubles(law.params()), _lib.ptr(problem._points),
        _lib.ptr(problem._cells), problem.num_cells, _lib.ptr(problem._as_sol(sol_list)), _lib.ptr(iv), _lib.ptr(lam),
        _lib.ptr(problem._ref), _lib.ptr(grad), _lib.stream_ptr()))
    return grad


def ad_wrapper(problem, solver_options={}, adjoint_solver_options={}):
    """Differentiable forward map params -> sol_list (solver.py:1421-1455).

    The reference registers a jax.custom_vjp; here ``fwd_pred`` is differentiable by torch.autograd:
    ``set_params(params)`` (user code, torch ops) builds ``problem.internal_vars``; the solve is a custom
    autograd Function whose backward is the implicit adjoint, so ``objective(fwd_pred(params)).backward()``
    fills ``params.grad`` exactly as ``jax.grad`` does in the reference."""

    if problem.ele_type == 'HEX27' and not isinstance(problem._law, laws.SIMP):
        raise NotImplementedError("ad_wrapper on HEX27: the parameter-gradient kernel is registered for SIMP only; it does not fall back")

    class _Solve(torch.autograd.Function):
        @staticmethod
        def forward(ctx, theta):
            problem.internal_vars = [theta.detach()] + list(problem.internal_vars[1:])
            sol = solver(problem, solver_options)[0]
            ctx.save_for_backward(theta.detach(), sol)
            return sol

        @staticmethod
        def backward(ctx, v):
            theta, sol = ctx.saved_tensors
            logger.info("Running backward and solving the adjoint problem...")
            problem.internal_vars = [theta] + list(problem.internal_vars[1:])
            return implicit_vjp(problem, [sol], None, [v.contiguous()], adjoint_solver_options)

    def fwd_pred(params):
        problem.set_params(params)
        iv = problem.internal_vars
        if len(iv) == 0 or not isinstance(iv[0], torch.Tensor):
            raise ValueError("set_params must store a torch tensor of shape (num_cells, num_quads) in internal_vars[0]")
        theta = iv[0].to(device=problem.device, dtype=torch.float64)
        return [_Solve.apply(theta)]

    return fwd_pred
