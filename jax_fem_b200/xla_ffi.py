"""Registration of the XLA FFI custom calls of csrc/fem_b200_xla.cc with JAX (the "thin jax.ffi C-ABI custom calls" of the
hot path).  Needs jax + jaxlib at BUILD time (jax.ffi.include_dir()) and at run time; neither is installable in this image,
so everything here fails loudly when they are missing -- there is no fallback.

    from jax_fem_b200 import xla_ffi
    xla_ffi.build()        # compiles lib/libfem_b200_xla.so against jax.ffi.include_dir()
    xla_ffi.register()     # jax.ffi.register_ffi_target(name, capsule, platform="CUDA") for every handler
    y = jax.ffi.ffi_call("fem_b200_xla_spmv", jax.ShapeDtypeStruct(x.shape, x.dtype))(data, x, plan=plan_address, vec=3)

INTEGRATION.md lists the call sites in jax_fem/problem.py and jax_fem/solver.py that these calls replace.
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM_SRC = os.path.join(HERE, "csrc", "fem_b200_xla.cc")
SHIM_LIB = os.path.join(HERE, "lib", "libfem_b200_xla.so")
STUB_INCLUDE = os.path.join(HERE, "..", "tools", "xla_ffi_stub")
TARGETS = ("fem_b200_xla_assemble", "fem_b200_xla_apply_bc_vec", "fem_b200_xla_spmv", "fem_b200_xla_transpose",
           "fem_b200_xla_krylov", "fem_b200_xla_adjoint_grad")


def include_dir():
    """jax.ffi.include_dir(), or None when jax is not importable."""
    try:
        import jax
        return jax.ffi.include_dir()
    except Exception:
        return None


def check_syntax(cuda_include="/usr/local/cuda/include"):
    """Compile the shim against the FFI stub headers (tools/xla_ffi_stub): catches typos where jax cannot be installed."""
    cmd = ["g++", "-std=c++17", "-fsyntax-only", f"-I{STUB_INCLUDE}", f"-I{cuda_include}", SHIM_SRC]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("fem_b200_xla.cc does not compile:\n" + proc.stderr)
    return True


def build(cuda_home="/usr/local/cuda"):
    inc = include_dir()
    if inc is None:
        raise RuntimeError("jax is not importable: the XLA FFI shim cannot be built here (no fallback); "
                           "install jax + jaxlib, or bind include/fem_b200.h through ctypes as jax_fem_b200/_lib.py does")
    cmd = ["g++", "-std=c++17", "-O2", "-shared", "-fPIC", f"-I{inc}", f"-I{cuda_home}/include", SHIM_SRC,
           f"-L{os.path.dirname(SHIM_LIB)}", "-lfem_b200", f"-L{cuda_home}/lib64", "-lcudart",
           f"-Wl,-rpath,{os.path.dirname(SHIM_LIB)}", "-o", SHIM_LIB]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building libfem_b200_xla.so failed:\n" + proc.stderr)
    return SHIM_LIB


def register():
    """Register every handler of the shim as a CUDA FFI target; returns the list of names."""
    try:
        import jax
    except Exception as e:
        raise RuntimeError("jax is not importable: the XLA FFI path is unavailable in this environment") from e
    if not os.path.exists(SHIM_LIB):
        build()
    lib = ctypes.CDLL(SHIM_LIB)
    for name in TARGETS:
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name)), platform="CUDA")
    return list(TARGETS)
