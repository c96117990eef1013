"""Loader / builder of the CUDA shared object behind the C ABI (include/dcbf_mpc.h).

The library is built IN-TREE (csrc/libdcbf_mpc.so) with nvcc for sm_100a.  There is no CPU fallback: if the
library is missing or cannot be loaded, every product entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(CSRC, "libdcbf_mpc.so")
SOURCES = ["dcbf_kernels.cu", "dcbf_core.cuh", "dcbf_lanes.cuh", "dcbf_warp.cuh", "dcbf_math.cuh", "dcbf_gen.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


class DcbfParams(C.Structure):
    """ctypes mirror of `dcbf_params` (include/dcbf_mpc.h)."""
    _fields_ = [("formulation", C.c_int32), ("max_iter", C.c_int32), ("select_obs", C.c_int32), ("goal_shift", C.c_int32),
                ("has_fen", C.c_int32), ("close_any", C.c_int32), ("tiny_count", C.c_int32), ("reserved1", C.c_int32),
                ("w_p", C.c_double), ("w_q", C.c_double), ("w_r", C.c_double), ("w_t", C.c_double),
                ("gamma", C.c_double), ("s_turn", C.c_double),
                ("bvx_min", C.c_double), ("bvx_max", C.c_double), ("bvy_min", C.c_double), ("bvy_max", C.c_double),
                ("leg_sq", C.c_double), ("ang_max", C.c_double), ("detect_sq", C.c_double), ("close_radius", C.c_double),
                ("tol", C.c_double), ("constr_viol_tol", C.c_double), ("mu_init", C.c_double), ("tiny_alpha", C.c_double),
                ("mu_warm", C.c_double), ("mu_shift", C.c_double), ("resto_window", C.c_double), ("kappa_eps", C.c_double)]


# every symbol include/dcbf_mpc.h declares
EXPORTS = ["dcbf_abi_version", "dcbf_default_params", "dcbf_create", "dcbf_destroy", "dcbf_last_error",
           "dcbf_set_fields", "dcbf_num_rows", "dcbf_num_vars", "dcbf_eval", "dcbf_solve", "dcbf_rollout",
           "dcbf_set_fields_host", "dcbf_solve_host", "dcbf_launch_count", "dcbf_fp64_peak_tflops", "dcbf_tick", "dcbf_alip_foot", "dcbf_math_probe",
           "dcbf_gen_fields", "dcbf_gen_states", "dcbf_heading_input", "dcbf_setup_info", "dcbf_veldes_foot", "dcbf_solve_host_async", "dcbf_wait"]


def needs_build() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(_HERE, "..", "include", "dcbf_mpc.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/dcbf_kernels.cu for sm_100a (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return SO_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO_PATH, os.path.join(CSRC, "dcbf_kernels.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return SO_PATH


_LIB = None


def load():
    """Load libdcbf_mpc.so and declare the prototypes.  Raises if the library is not there."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("DCBF_LIB", SO_PATH)   # experiments with alternative builds of the same sources
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, dp, ip = C.c_void_p, C.c_void_p, C.c_void_p   # device/host pointers travel as integers
    lib.dcbf_abi_version.restype = C.c_int
    lib.dcbf_default_params.argtypes = [C.c_int, C.POINTER(DcbfParams)]
    lib.dcbf_create.argtypes = [C.POINTER(DcbfParams), C.c_int, C.POINTER(vp)]
    lib.dcbf_destroy.argtypes = [vp]
    lib.dcbf_destroy.restype = None
    lib.dcbf_last_error.argtypes = [vp]
    lib.dcbf_last_error.restype = C.c_char_p
    lib.dcbf_set_fields.argtypes = [vp, C.c_int32, C.c_int32, dp, C.c_int32, dp, vp]
    lib.dcbf_num_rows.argtypes = [vp]
    lib.dcbf_num_vars.argtypes = [vp]
    lib.dcbf_eval.argtypes = [vp, C.c_int32] + [dp] * 14 + [vp]
    lib.dcbf_solve.argtypes = [vp, C.c_int32] + [dp] * 14 + [vp]
    lib.dcbf_rollout.argtypes = [vp, C.c_int32, C.c_int32] + [dp] * 9 + [vp]
    lib.dcbf_set_fields_host.argtypes = [vp, C.c_int32, C.c_int32, ip, C.c_int32, ip]
    lib.dcbf_solve_host.argtypes = [vp, C.c_int32] + [ip] * 14
    lib.dcbf_solve_host_async.argtypes = [vp, C.c_int32] + [ip] * 14
    lib.dcbf_wait.argtypes = [vp]
    lib.dcbf_tick.argtypes = [vp, C.c_int32] + [dp] * 21 + [vp]
    lib.dcbf_alip_foot.argtypes = [vp, C.c_int32] + [dp] * 5 + [C.c_int32] + [C.c_double] * 4 + [dp] * 3 + [vp]
    lib.dcbf_math_probe.argtypes = [vp, C.c_int32, dp, dp, dp, vp]
    lib.dcbf_gen_fields.argtypes = [vp, C.c_int32, C.c_uint64, C.c_int32, C.c_int32] + [C.c_double] * 4 + [dp, dp, vp, vp]
    lib.dcbf_gen_states.argtypes = [vp, C.c_int32, C.c_uint64, vp] + [C.c_double] * 3 + [dp, dp, vp, dp, dp, vp, vp]
    lib.dcbf_heading_input.argtypes = [vp, C.c_int32, dp, dp, dp, C.c_int32, C.c_int32, dp, C.c_int32, vp]
    lib.dcbf_setup_info.argtypes = [vp, C.c_int32] + [dp] * 6 + [vp]
    lib.dcbf_veldes_foot.argtypes = [vp, C.c_int32, dp, dp, dp, C.c_double, C.c_double, dp, dp, vp]
    lib.dcbf_launch_count.argtypes = [vp]
    lib.dcbf_launch_count.restype = C.c_int64
    lib.dcbf_fp64_peak_tflops.argtypes = [vp, C.c_int32]
    lib.dcbf_fp64_peak_tflops.restype = C.c_double
    _LIB = lib
    return lib
