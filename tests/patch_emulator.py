"""NumPy walk through the fused-assembly tables in the kernel's order (csrc/fused.cu): test infrastructure for
jax_fem_b200/patch_plan.py -- it lets the CPU suite check the patch plan against the oracle's CSR without a GPU."""
import numpy as np


def emulate(pp, Ke, Re, bc_flag, f_ext, nnz):
    """Ke (C, N, vec, N, vec) element tangents, Re (C, N, vec) element residuals, bc_flag (n,) -> (CSR data, nodal residual)."""
    N, v = pp.nodes_per_cell, pp.vec
    vv = v * v
    g = lambda t: t.cpu().numpy().astype(np.int64)
    phdr, pn_node, pn_out, pn_acc, pn_info = g(pp.phdr), g(pp.pn_node), g(pp.pn_out), g(pp.pn_acc), g(pp.pn_info)
    pc_cell, ck_cell, ck_lane, ck_rnd, ln_desc = g(pp.pc_cell), g(pp.ck_cell), g(pp.ck_lane), g(pp.ck_rnd), g(pp.ln_desc)
    ln_slot = pp.ln_slot.cpu().numpy().view(np.uint8).reshape(-1, N).astype(np.int64)
    data = np.full(nnz, np.nan)
    res = np.full((len(pn_node), v), np.nan)
    for p in range(pp.n_patches):
        n0, n1 = phdr[p, 0], phdr[p + 1, 0]
        k0, k1 = phdr[p, 3], phdr[p + 1, 3]
        acc = np.zeros(phdr[p, 4])
        racc = np.zeros((n1 - n0, v))
        for k in range(k0, k1):
            for r in range(ck_rnd[k]):
                for l in range(ck_lane[k], ck_lane[k + 1]):
                    d = ln_desc[l]
                    cl, a, nl, rk = d & 31, (d >> 5) & 7, (d >> 8) & 255, (d >> 16) & 255
                    if rk != r:
                        continue
                    assert cl < ck_cell[k + 1] - ck_cell[k]
                    c = pc_cell[ck_cell[k] + cl]
                    ln = pn_info[n0 + nl] & 255
                    base = pn_acc[n0 + nl]
                    for b in range(N):
                        s = ln_slot[l, b]
                        for i in range(v):
                            acc[base + i * v * ln + v * s: base + i * v * ln + v * s + v] += Ke[c, a, i, b, :]
                    racc[nl] += Re[c, a]
        for i in range(n1 - n0):
            n = pn_node[n0 + i]
            ln, dg = pn_info[n0 + i] & 255, pn_info[n0 + i] >> 8
            rows = acc[pn_acc[n0 + i]: pn_acc[n0 + i] + vv * ln].reshape(v, v * ln).copy()
            for j in range(v):
                if bc_flag[v * n + j]:
                    rows[j] = 0.0
                    rows[j, v * dg + j] = 1.0
            data[pn_out[n0 + i]: pn_out[n0 + i] + vv * ln] = rows.reshape(-1)
            res[n] = racc[i] + (f_ext[n] if f_ext is not None else 0.0)
    return data, res
