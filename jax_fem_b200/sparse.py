"""Device-resident CSR matrix handle returned by get_A.

Duck-types the part of petsc4py's Mat that the reference's call sites use:
getValuesCSR (jax_fem/solver.py:65), getSize (:66), transpose(out) (:1405-1408), mult (:127).
The arrays are torch CUDA tensors (int32 indptr/indices, float64 data) and stay in HBM; nothing is
copied to the host unless the caller asks (to_scipy).
"""
import torch

from . import _lib


class CSRMatrix:
    def __init__(self, plan, data, transposed=False):
        self.plan = plan
        self.data = data
        self.transposed = transposed
        self._diag = None

    # -- PETSc-like surface -------------------------------------------------------------------------
    def getSize(self):
        return (self.plan.n, self.plan.n)

    def getValuesCSR(self):
        return self.plan.indptr, self.plan.indices, self.data

    def transpose(self, out=None):
        """Non in-place transpose (A.transpose(A_T) form).  The pattern is structurally symmetric, so the
        transpose shares indptr/indices and only the values are permuted (on the device)."""
        p = self.plan
        data_t = torch.empty_like(self.data)
        lib = _lib.load()
        _lib.check(lib.fem_csr_transpose_values(p.vec, p.num_nodes, _lib.ptr(p.brow_ptr), _lib.ptr(p.bcol),
                                                _lib.ptr(p.tperm), _lib.ptr(self.data), _lib.ptr(data_t),
                                                _lib.stream_ptr()))
        result = CSRMatrix(p, data_t, transposed=not self.transposed)
        if out is not None and isinstance(out, CSRMatrix):
            out.plan, out.data, out.transposed, out._diag = p, data_t, result.transposed, None
            return out
        return result

    def mult(self, x, y=None):
        y = torch.empty_like(x) if y is None else y
        p = self.plan
        _lib.check(_lib.load().fem_spmv(p.n, _lib.ptr(p.indptr), _lib.ptr(p.indices), _lib.ptr(self.data), p.vec,
                                        _lib.ptr(p.brow_ptr), _lib.ptr(p.bcol), _lib.ptr(x), _lib.ptr(y),
                                        _lib.stream_ptr()))
        return y

    def __matmul__(self, x):
        return self.mult(x.contiguous())

    def diagonal(self):
        if self._diag is None:
            p = self.plan
            d = torch.empty(p.n, dtype=torch.float64, device=self.data.device)
            _lib.check(_lib.load().fem_csr_diagonal(p.n, _lib.ptr(p.indptr), _lib.ptr(p.indices),
                                                    _lib.ptr(self.data), _lib.ptr(d), _lib.stream_ptr()))
            self._diag = d
        return self._diag

    # -- host export (tests, custom solvers) --------------------------------------------------------
    def to_scipy(self):
        import scipy.sparse as sp
        n = self.plan.n
        return sp.csr_matrix((self.data.cpu().numpy(), self.plan.indices.cpu().numpy(),
                              self.plan.indptr.cpu().numpy()), shape=(n, n))
