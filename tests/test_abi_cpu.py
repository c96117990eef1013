"""CPU tests: the C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol the header declares."""
import ctypes as C
import os
import re

from mujoco_lip_mpc_simulation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols(built):
    hdr = open(os.path.join(ROOT, "include", "dcbf_mpc.h")).read()
    declared = set(re.findall(r"\b(dcbf_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dcbf_abi_version() == 4


def test_default_params_match_reference_constants(built):
    import math
    lib = _lib.load()
    P = _lib.DcbfParams()
    assert lib.dcbf_default_params(0, C.byref(P)) == 0   # MPC_LIP_sig_step.py:34-40,340-353
    assert (P.w_p, P.w_q, P.w_r, P.gamma, P.bvy_max, P.close_radius) == (2.0, 1.0, 15.0, 0.4, 0.3, 0.35)
    assert lib.dcbf_default_params(1, C.byref(P)) == 0   # MPC_LIP_modi.py:35-41,397-411
    assert (P.w_p, P.w_r, P.gamma, P.bvy_max, P.has_fen, P.select_obs, P.close_radius) == (0.0, 50.0, 0.2, 0.35, 1, 1, 0.15)
    assert abs(P.s_turn - 0.024 * 180 / math.pi) < 1e-15 and abs(P.ang_max - math.pi / 16) < 1e-15
    assert lib.dcbf_default_params(2, C.byref(P)) == 0   # MPC_DD_sig_step.py:33-37,323-338
    assert (P.w_t, P.w_r, P.gamma, P.bvx_min, P.bvx_max) == (2.0, 50.0, 0.2, 0.4, 0.8)
    assert lib.dcbf_default_params(7, C.byref(P)) < 0
    # the python mirror used by the host-sim tests must agree field by field
    import hostsim_binding as H
    for form in range(3):
        lib.dcbf_default_params(form, C.byref(P))
        Q = H.default_params(form)
        for name, _ in _lib.DcbfParams._fields_:
            assert getattr(P, name) == getattr(Q, name), (form, name)


def test_sass_is_sm100_fp64(built):
    """the kernels are compiled for sm_100a and the solve kernel runs on the FP64 pipe (DFMA in SASS)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([cuobjdump, "-sass", _lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "solve_lip_kernel" in out and "DFMA" in out


def test_integration_stub_mirrors_the_params_struct():
    """the ctypes stub INTEGRATION.md shows a maintainer declares dcbf_params field for field (a shorter struct would let
    dcbf_default_params write past it)"""
    import os
    import re
    from mujoco_lip_mpc_simulation_b200 import _lib
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "INTEGRATION.md")).read()
    block = text[text.index("class dcbf_params(C.Structure)"):text.index("class GpuSolve")]
    ints = re.findall(r'\("(\w+)", C\.c_int32\)', block)
    doubles = " ".join(re.findall(r'"([a-z_ ]+)"', block[block.index("C.c_double"):])).split()
    assert ints + doubles == [n for n, _ in _lib.DcbfParams._fields_]
