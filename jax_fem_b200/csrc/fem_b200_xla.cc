// XLA FFI custom calls over the C ABI of libfem_b200 (include/fem_b200.h): the "thin jax.ffi C-ABI custom calls" of the north
// star.  Each handler unpacks XLA buffers + the CUDA stream of the call frame and forwards to ONE C entry point; there is
// no logic here.  Built only where the FFI headers exist:
//
//     nvcc / g++ -std=c++17 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//         fem_b200_xla.cc -L../lib -lfem_b200 -o ../lib/libfem_b200_xla.so
//
// (`python -m jax_fem_b200.build` does this when `import jax` works and skips it otherwise; `--check-xla-shim` compiles
// against the syntax stub under tools/xla_ffi_stub in images without jax, as this one.)  jax_fem_b200/xla_ffi.py registers
// the symbols with jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(fn), platform="CUDA"); INTEGRATION.md shows how the
// reference's Problem.newton_update / get_A / jax_solve call them through jax.ffi.ffi_call.
//
// The assembly plan (fem_plan_create) is a handle created once per Problem outside jit, through ctypes; handlers receive it
// as the int64 attribute `plan`, the usual way of passing an opaque pointer to an XLA custom call.
#if __has_include("xla/ffi/api/ffi.h")
#include <cstdint>
#include <string>
#include <cuda_runtime_api.h>
#include "xla/ffi/api/ffi.h"
#include "../../include/fem_b200.h"

namespace ffi = xla::ffi;

namespace {
ffi::Error Status(int code) {
  if (code == FEM_OK) return ffi::Error::Success();
  return ffi::Error(code == FEM_EINVAL ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    std::string("libfem_b200: ") + fem_last_error());
}

const int32_t* Table(int64_t plan, int which) {
  const int32_t* p = nullptr;
  fem_plan_table(reinterpret_cast<const void*>(plan), which, &p, nullptr);
  return p;
}

// Problem.newton_update + get_A (jax_fem/problem.py:447-491, jax_fem/solver.py:469-553): element kernel -> residual gather
// -> CSR gather, Dirichlet rows treated through `emeta` (fem_plan_entry_meta, computed once per Dirichlet set).
// params: the 8 law parameters (host attribute values are not arrays in every jaxlib, so they travel as 8 scalars).
ffi::Error AssembleImpl(cudaStream_t stream, int64_t plan, int64_t ele_type, int64_t vec, int64_t law_id, double p0, double p1,
                        double p2, double p3, double p4, ffi::Buffer<ffi::F64> points, ffi::Buffer<ffi::S32> cells,
                        ffi::Buffer<ffi::F64> sol, ffi::Buffer<ffi::F64> internal_var, ffi::Buffer<ffi::F64> ref_tables,
                        ffi::Buffer<ffi::S32> emeta, ffi::Buffer<ffi::F64> f_ext, ffi::ResultBuffer<ffi::F64> data,
                        ffi::ResultBuffer<ffi::F64> res, ffi::ResultBuffer<ffi::F64> Ke, ffi::ResultBuffer<ffi::F64> Re) {
  const double params[8] = {p0, p1, p2, p3, p4, 0., 0., 0.};
  const int64_t n_cells = cells.dimensions()[0], nn = cells.dimensions()[1], n_nodes = points.dimensions()[0];
  int64_t sizes[8];
  if (int e = fem_plan_sizes(reinterpret_cast<const void*>(plan), sizes)) return Status(e);
  const double* iv = internal_var.element_count() ? internal_var.typed_data() : nullptr;
  const double* fx = f_ext.element_count() ? f_ext.typed_data() : nullptr;
  if (int e = fem_element_residual_jacobian((int)ele_type, (int)vec, (int)law_id, params, points.typed_data(), cells.typed_data(),
                                            n_cells, sol.typed_data(), iv, ref_tables.typed_data(),
                                            Table(plan, FEM_PLAN_CORNER_POS), Ke->typed_data(), Re->typed_data(), stream))
    return Status(e);
  if (int e = fem_gather_residual((int)vec, (int)nn, n_nodes, Table(plan, FEM_PLAN_NC_PTR), Table(plan, FEM_PLAN_NC),
                                  Re->typed_data(), fx, res->typed_data(), stream))
    return Status(e);
  return Status(fem_gather_csr((int)vec, (int)nn, sizes[2], Table(plan, FEM_PLAN_GDESC), emeta.typed_data(),
                               Table(plan, FEM_PLAN_SRC), Ke->typed_data(), data->typed_data(), stream));
}

// apply_bc_vec (jax_fem/solver.py:290-304): res[rows] = sol[rows] - vals * scale, in place on a copy made by XLA
ffi::Error ApplyBcImpl(cudaStream_t stream, double scale, ffi::Buffer<ffi::S32> rows, ffi::Buffer<ffi::F64> vals,
                       ffi::Buffer<ffi::F64> sol, ffi::Buffer<ffi::F64> res_in, ffi::ResultBuffer<ffi::F64> res) {
  const size_t bytes = sizeof(double) * res_in.element_count();
  if (cudaMemcpyAsync(res->typed_data(), res_in.typed_data(), bytes, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemcpyAsync failed");
  return Status(fem_apply_bc_vec((int64_t)rows.element_count(), rows.typed_data(), vals.typed_data(), scale, sol.typed_data(),
                                 res->typed_data(), stream));
}

// A @ x on the plan's pattern (BCOO matvec of jax_fem/solver.py:67,87)
ffi::Error SpmvImpl(cudaStream_t stream, int64_t plan, int64_t vec, ffi::Buffer<ffi::F64> data, ffi::Buffer<ffi::F64> x,
                    ffi::ResultBuffer<ffi::F64> y) {
  return Status(fem_spmv((int64_t)x.element_count(), Table(plan, FEM_PLAN_INDPTR), Table(plan, FEM_PLAN_INDICES), data.typed_data(),
                         (int)vec, Table(plan, FEM_PLAN_BROW_PTR), Table(plan, FEM_PLAN_BCOL), x.typed_data(), y->typed_data(),
                         stream));
}

// A^T values (A.transpose(), jax_fem/solver.py:1405-1408): the pattern is structurally symmetric
ffi::Error TransposeImpl(cudaStream_t stream, int64_t plan, int64_t vec, ffi::Buffer<ffi::F64> data,
                         ffi::ResultBuffer<ffi::F64> data_t) {
  int64_t sizes[8];
  if (int e = fem_plan_sizes(reinterpret_cast<const void*>(plan), sizes)) return Status(e);
  return Status(fem_csr_transpose_values((int)vec, sizes[5] / vec, Table(plan, FEM_PLAN_BROW_PTR), Table(plan, FEM_PLAN_BCOL),
                                         Table(plan, FEM_PLAN_TPERM), data.typed_data(), data_t->typed_data(), stream));
}

// jax_solve (jax_fem/solver.py:63-92): Jacobi-preconditioned BiCGSTAB (method 0, the reference's default) or CG (method 1).
// x0 -> x; info = (iterations, ||r||^2, ||A x - b||); workspace = fem_krylov_workspace(n) doubles, allocated by XLA as a result.
ffi::Error KrylovImpl(cudaStream_t stream, int64_t plan, int64_t vec, int64_t method, double tol, double atol, int64_t maxiter,
                      ffi::Buffer<ffi::F64> data, ffi::Buffer<ffi::F64> b, ffi::Buffer<ffi::F64> x0,
                      ffi::ResultBuffer<ffi::F64> x, ffi::ResultBuffer<ffi::F64> info, ffi::ResultBuffer<ffi::F64> diag,
                      ffi::ResultBuffer<ffi::F64> workspace) {
  const int64_t n = (int64_t)b.element_count();
  if ((int64_t)workspace->element_count() < fem_krylov_workspace(n))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "workspace smaller than fem_krylov_workspace(n)");
  if (cudaMemcpyAsync(x->typed_data(), x0.typed_data(), sizeof(double) * n, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemcpyAsync failed");
  const int32_t *indptr = Table(plan, FEM_PLAN_INDPTR), *indices = Table(plan, FEM_PLAN_INDICES);
  if (int e = fem_csr_diagonal(n, indptr, indices, data.typed_data(), diag->typed_data(), stream)) return Status(e);
  double info_host[4] = {0., 0., 0., 0.};
  auto solve = method == 1 ? fem_pcg : fem_pbicgstab;
  if (int e = solve(n, indptr, indices, data.typed_data(), (int)vec, Table(plan, FEM_PLAN_BROW_PTR), Table(plan, FEM_PLAN_BCOL),
                    diag->typed_data(), b.typed_data(), x->typed_data(), tol, atol, (int)maxiter, 25, workspace->typed_data(),
                    info_host, stream))
    return Status(e);
  if (cudaMemcpyAsync(info->typed_data(), info_host, sizeof(double) * 3, cudaMemcpyHostToDevice, stream) != cudaSuccess ||
      cudaStreamSynchronize(stream) != cudaSuccess)      // info_host is a stack buffer
    return ffi::Error(ffi::ErrorCode::kInternal, "copy of the solver info failed");
  return ffi::Error::Success();
}

// -lambda^T dc/dtheta per quadrature point (jax_fem/solver.py:1386-1394,1414-1416)
ffi::Error AdjointGradImpl(cudaStream_t stream, int64_t ele_type, int64_t vec, int64_t law_id, double p0, double p1, double p2,
                           double p3, double p4, ffi::Buffer<ffi::F64> points, ffi::Buffer<ffi::S32> cells,
                           ffi::Buffer<ffi::F64> sol, ffi::Buffer<ffi::F64> internal_var, ffi::Buffer<ffi::F64> lam,
                           ffi::Buffer<ffi::F64> ref_tables, ffi::ResultBuffer<ffi::F64> grad) {
  const double params[8] = {p0, p1, p2, p3, p4, 0., 0., 0.};
  return Status(fem_adjoint_param_grad((int)ele_type, (int)vec, (int)law_id, params, points.typed_data(), cells.typed_data(),
                                       cells.dimensions()[0], sol.typed_data(), internal_var.typed_data(), lam.typed_data(),
                                       ref_tables.typed_data(), grad->typed_data(), stream));
}
}  // namespace

using F64 = ffi::Buffer<ffi::F64>;
using S32 = ffi::Buffer<ffi::S32>;
#define FEM_STREAM Ctx<ffi::PlatformStream<cudaStream_t>>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(fem_b200_xla_assemble, AssembleImpl,
                              ffi::Ffi::Bind().FEM_STREAM.Attr<int64_t>("plan").Attr<int64_t>("ele_type").Attr<int64_t>("vec")
                                  .Attr<int64_t>("law_id").Attr<double>("p0").Attr<double>("p1").Attr<double>("p2").Attr<double>("p3")
                                  .Attr<double>("p4").Arg<F64>().Arg<S32>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<S32>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fem_b200_xla_apply_bc_vec, ApplyBcImpl,
                              ffi::Ffi::Bind().FEM_STREAM.Attr<double>("scale").Arg<S32>().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fem_b200_xla_spmv, SpmvImpl,
                              ffi::Ffi::Bind().FEM_STREAM.Attr<int64_t>("plan").Attr<int64_t>("vec").Arg<F64>().Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fem_b200_xla_transpose, TransposeImpl,
                              ffi::Ffi::Bind().FEM_STREAM.Attr<int64_t>("plan").Attr<int64_t>("vec").Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fem_b200_xla_krylov, KrylovImpl,
                              ffi::Ffi::Bind().FEM_STREAM.Attr<int64_t>("plan").Attr<int64_t>("vec").Attr<int64_t>("method")
                                  .Attr<double>("tol").Attr<double>("atol").Attr<int64_t>("maxiter").Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fem_b200_xla_adjoint_grad, AdjointGradImpl,
                              ffi::Ffi::Bind().FEM_STREAM.Attr<int64_t>("ele_type").Attr<int64_t>("vec").Attr<int64_t>("law_id")
                                  .Attr<double>("p0").Attr<double>("p1").Attr<double>("p2").Attr<double>("p3").Attr<double>("p4")
                                  .Arg<F64>().Arg<S32>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>());
#else
// no XLA FFI headers in this environment: nothing to build (jax_fem_b200/xla_ffi.py raises when asked to register)
#endif
