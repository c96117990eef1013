"""ORACLE (test infrastructure) -- ctypes binding of oracle/libdcbf_oracle.so (see dcbf_oracle.c header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FORMS = {"sig_step": 0, "modi": 1, "dd": 2}


class OrcParams(C.Structure):
    _fields_ = [("form", C.c_int),
                ("p", C.c_double), ("q", C.c_double), ("r", C.c_double), ("gamma", C.c_double),
                ("s_turn", C.c_double), ("t_smooth", C.c_double),
                ("bvx_min", C.c_double), ("bvx_max", C.c_double), ("bvy_min", C.c_double), ("bvy_max", C.c_double),
                ("leg_sq", C.c_double), ("ang_max", C.c_double),
                ("has_fen", C.c_int), ("max_iter", C.c_int), ("tol", C.c_double),
                ("select_obs", C.c_int), ("goal_shift", C.c_int), ("close_radius", C.c_double),
                ("split_abs", C.c_int)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libdcbf_oracle.so")
    src = os.path.join(_HERE, "dcbf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        assert _LIB.orc_sizeof_params() == C.sizeof(OrcParams)
    return _LIB


def params(form, **over) -> OrcParams:
    P = OrcParams()
    lib().orc_default_params(FORMS[form] if isinstance(form, str) else int(form), C.byref(P))
    for k, v in over.items():
        setattr(P, k, v)
    return P


def _d(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _p(a, typ=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(typ))


def setup_info(P, xk, goal, leg, cir, elp):
    cir = _d(np.zeros((0, 3)) if cir is None else cir).reshape(-1, 3)
    elp = _d(np.zeros((0, 5)) if elp is None else elp).reshape(-1, 5)
    out = np.zeros(4, dtype=np.int32)
    ge = np.zeros(2)
    xk, goal = _d(xk), _d(np.ravel(goal))
    lib().orc_setup_info(C.byref(P), _p(xk), _p(goal), int(leg), len(cir), _p(cir), len(elp), _p(elp), _p(out, C.c_int), _p(ge))
    return dict(n=int(out[0]), m=int(out[1]), nc=int(out[2]), ne=int(out[3]), goal=ge)


def evaluate(P, xk, goal, leg, cir, elp, u, last_u=None):
    """-> f, grad[n], c[m], jac[m,n], cl[m], cu[m] in the reference's u-space and row order."""
    info = setup_info(P, xk, goal, leg, cir, elp)
    n, m = info["n"], info["m"]
    cir = _d(np.zeros((0, 3)) if cir is None else cir).reshape(-1, 3)
    elp = _d(np.zeros((0, 5)) if elp is None else elp).reshape(-1, 5)
    xk, goal, u, last_u = _d(xk), _d(np.ravel(goal)), _d(u), _d(last_u)
    f = C.c_double()
    grad, c, jac, cl, cu = np.zeros(n), np.zeros(m), np.zeros((m, n)), np.zeros(m), np.zeros(m)
    lib().orc_eval(C.byref(P), _p(xk), _p(goal), int(leg), len(cir), _p(cir), len(elp), _p(elp), _p(last_u), _p(u),
                   C.byref(f), _p(grad), _p(c), _p(jac), _p(cl), _p(cu))
    return f.value, grad, c, jac, cl, cu


def solve(P, xk, goal, leg, cir, elp, u0, last_u=None):
    form = P.form
    n, nx = (6, 3) if form == 2 else (15, 5)
    cir = _d(np.zeros((0, 3)) if cir is None else cir).reshape(-1, 3)
    elp = _d(np.zeros((0, 5)) if elp is None else elp).reshape(-1, 5)
    xk, goal, u0, last_u = _d(xk), _d(np.ravel(goal)), _d(u0), _d(last_u)
    u, xp, pp = np.zeros(n), np.zeros((3, nx)), np.zeros((3, 3))
    f, viol = C.c_double(), C.c_double()
    st, it, cl = C.c_int(), C.c_int(), C.c_int()
    lib().orc_solve(C.byref(P), _p(xk), _p(goal), int(leg), len(cir), _p(cir), len(elp), _p(elp), _p(last_u), _p(u0),
                    _p(u), _p(xp), _p(pp), C.byref(f), C.byref(st), C.byref(it), C.byref(viol), C.byref(cl))
    return dict(u=u, x_plan=xp, p_plan=pp, f=f.value, status=st.value, iters=it.value, viol=viol.value, close2goal=bool(cl.value))


def solve_batch(P, xk, goal, leg, cir, elp, u0, field=None, last_u=None, threads=1):
    """cir: [F,nc,3], elp: [F,ne,5] (or None); field[B] int32 or None (identity)."""
    form = P.form
    n, nx = (6, 3) if form == 2 else (15, 5)
    xk = _d(xk).reshape(-1, nx)
    B = len(xk)
    goal = _d(np.broadcast_to(np.asarray(goal, dtype=np.float64).reshape(-1, 2), (B, 2)))
    leg = np.ascontiguousarray(np.broadcast_to(np.asarray(leg, dtype=np.int32), (B,)), dtype=np.int32)
    cir = _d(np.zeros((1, 0, 3)) if cir is None else cir)
    elp = _d(np.zeros((cir.shape[0], 0, 5)) if elp is None else elp)
    nc, ne = cir.shape[1], elp.shape[1]
    field_a = None if field is None else np.ascontiguousarray(field, dtype=np.int32)
    u0, last_u = _d(u0).reshape(B, n), _d(last_u)
    u, xp, pp = np.zeros((B, n)), np.zeros((B, 3, nx)), np.zeros((B, 3, 3))
    f, viol = np.zeros(B), np.zeros(B)
    st, it, cl = np.zeros(B, np.int32), np.zeros(B, np.int32), np.zeros(B, np.int32)
    lib().orc_solve_batch(C.byref(P), B, int(threads), _p(xk), _p(goal), _p(leg, C.c_int), nc, _p(cir), ne,
                          _p(elp) if ne else None, _p(field_a, C.c_int), _p(last_u), _p(u0), _p(u), _p(xp), _p(pp), _p(f),
                          _p(st, C.c_int), _p(it, C.c_int), _p(viol), _p(cl, C.c_int))
    return dict(u=u, x_plan=xp, p_plan=pp, f=f, status=st, iters=it, viol=viol, close2goal=cl.astype(bool))
