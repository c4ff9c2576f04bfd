"""ctypes binding of the C ABI declared in include/fem_b200.h.

The product path has NO CPU fallback: if the shared library is missing it is an error, and every
compute entry point returns FEM_ENODEV (raised here as RuntimeError) when no CUDA device exists.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libfem_b200.so")
HEADER = os.path.join(HERE, "..", "include", "fem_b200.h")

ELE = {"HEX8": 0, "QUAD4": 1, "HEX27": 2}
c_i32p = ctypes.c_void_p
c_f64p = ctypes.c_void_p
_vp, _i, _i64, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double

_SIGNATURES = {
    "fem_last_error": (ctypes.c_char_p, []),
    "fem_version": (_i, []),
    "fem_device_count": (_i, []),
    "fem_plan_create": (_i, [_vp, _i64, _i64, _i, _i, _vp, _vp]),
    "fem_plan_destroy": (_i, [_vp]),
    "fem_plan_sizes": (_i, [_vp, _vp]),
    "fem_plan_table": (_i, [_vp, _i, _vp, _vp]),
    "fem_plan_entry_meta": (_i, [_vp, _vp, _vp, _vp]),
    "fem_element_residual_jacobian": (_i, [_i, _i, _i, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fem_halo_set_interior": (_i, [_vp, _i64, _i64]),
    "fem_halo_p2p_alloc": (_i, [_vp, _i64, _i64, _vp]),
    "fem_halo_p2p_connect": (_i, [_vp, _i, _vp, _i64, _i64, _i]),
    "fem_halo_p2p_enable": (_i, [_vp, _i]),
    "fem_hex27_adjoint_param_grad": (_i, [_i, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "fem_mass_term": (_i, [_i, _i, _vp, _vp, _i64, _vp, _vp, _vp, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fem_hex27_residual_jacobian": (_i, [_i, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fem_assemble_fused": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i64] + [_vp] * 18 + [_i, _vp]),
    "fem_staged_ctrl_ints": (_i64, [_i64, _i64]),
    "fem_assemble_staged": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64] + [_vp] * 9),
    "fem_staged_status": (_i, [_vp, _vp, _vp]),
    "fem_patch_chunks_host": (_i, [_i64, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "fem_gather_csr": (_i, [_i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fem_element_tiles": (_i, [_i, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fem_gather_csr_tiles": (_i, [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fem_gather_residual": (_i, [_i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fem_face_residual": (_i, [_i, _i, _i, _i64] + [_vp] * 13),
    "fem_face_tangent": (_i, [_i, _i, _i, _i64] + [_vp] * 16),
    "fem_apply_bc_vec": (_i, [_i64, _vp, _vp, _d, _vp, _vp, _vp]),
    "fem_bc_initial_guess": (_i, [_i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "fem_spmv": (_i, [_i64, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "fem_csr_diagonal": (_i, [_i64, _vp, _vp, _vp, _vp, _vp]),
    "fem_csr_transpose_values": (_i, [_i, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fem_krylov_workspace": (_i64, [_i64]),
    "fem_pcg": (_i, [_i64, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _d, _d, _i, _i, _vp, _vp, _vp]),
    "fem_pbicgstab": (_i, [_i64, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _d, _d, _i, _i, _vp, _vp, _vp]),
    "fem_adjoint_param_grad": (_i, [_i, _i, _i, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fem_dot": (_i, [_i64, _vp, _vp, _vp, _vp, _vp]),
    "fem_dcg_begin": (_i, [_vp, _d, _d, _i, _vp]),
    "fem_dcg_spmv_dot": (_i, [_i64, _i64, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "fem_dcg_init": (_i, [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fem_dcg_scalars": (_i, [_i, _vp, _vp]),
    "fem_dcg_update": (_i, [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fem_dcg_direction": (_i, [_i64, _vp, _vp, _vp, _vp, _vp]),
    "fem_axpy": (_i, [_i64, _d, _vp, _vp, _vp]),
    "fem_nccl_unique_id": (_i, [_vp]),
    "fem_nccl_comm_create": (_i, [_i, _i, _vp, _vp]),
    "fem_nccl_comm_destroy": (_i, [_vp]),
    "fem_halo_create": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fem_halo_destroy": (_i, [_vp]),
    "fem_halo_exchange": (_i, [_vp, _vp, _vp]),
    "fem_allreduce_sum": (_i, [_vp, _vp, _i, _vp]),
    "fem_dist_pcg": (_i, [_vp, _i64, _i64, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _d, _d, _i, _i, _vp, _vp, _vp]),
    "fem_dist_pbicgstab": (_i, [_vp, _i64, _i64, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _d, _d, _i, _i, _vp, _vp, _vp]),
}

_lib = None


def declared_symbols():
    """Every function name include/fem_b200.h declares (used by the CPU tests)."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fem_[a-z0-9_]+)\s*\(", text)))


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m jax_fem_b200.build` "
                "(the hot path is CUDA-only; there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(code):
    if code != 0:
        msg = load().fem_last_error().decode()
        raise RuntimeError(f"libfem_b200 error {code}: {msg}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous(), "libfem_b200 needs contiguous tensors"
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def host_doubles(values, n=8):
    arr = (ctypes.c_double * n)(*([float(v) for v in values] + [0.0] * (n - len(values))))
    return arr
