"""TEST-ONLY ctypes binding of tests/hostsim/libhostsim.so: the per-lane CUDA solver code compiled for the host so
that the algorithm can be checked against the oracle without a GPU.  Not a product path (see hostsim.cpp)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "..", "mujoco_lip_mpc_simulation_b200", "csrc")


class DcbfParams(C.Structure):
    _fields_ = [("formulation", C.c_int32), ("max_iter", C.c_int32), ("select_obs", C.c_int32), ("goal_shift", C.c_int32),
                ("has_fen", C.c_int32), ("close_any", C.c_int32), ("tiny_count", C.c_int32), ("reserved1", C.c_int32),
                ("w_p", C.c_double), ("w_q", C.c_double), ("w_r", C.c_double), ("w_t", C.c_double),
                ("gamma", C.c_double), ("s_turn", C.c_double),
                ("bvx_min", C.c_double), ("bvx_max", C.c_double), ("bvy_min", C.c_double), ("bvy_max", C.c_double),
                ("leg_sq", C.c_double), ("ang_max", C.c_double), ("detect_sq", C.c_double), ("close_radius", C.c_double),
                ("tol", C.c_double), ("constr_viol_tol", C.c_double), ("mu_init", C.c_double), ("tiny_alpha", C.c_double),
                ("mu_warm", C.c_double), ("mu_shift", C.c_double), ("resto_window", C.c_double), ("kappa_eps", C.c_double)]


FORMS = {"sig_step": 0, "modi": 1, "dd": 2}
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "hostsim", "libhostsim.so")
        srcs = [os.path.join(_HERE, "hostsim", "hostsim.cpp")] + [os.path.join(_CSRC, f) for f in ("dcbf_core.cuh", "dcbf_lanes.cuh", "dcbf_math.cuh", "dcbf_warp.cuh", "dcbf_gen.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["sh", os.path.join(_HERE, "hostsim", "build.sh")])
        _LIB = C.CDLL(so)
    return _LIB


def default_params(form) -> DcbfParams:
    """Python mirror of dcbf_default_params (csrc/dcbf_kernels.cu); checked against the C ABI in the gpu tests."""
    import math
    f = FORMS[form] if isinstance(form, str) else int(form)
    P = DcbfParams()
    P.formulation = f
    P.w_q = 1.0
    P.bvx_min, P.bvx_max, P.bvy_min, P.leg_sq, P.ang_max = 0.4, 0.8, 0.15, 0.09, math.pi / 16
    P.detect_sq = 16.0
    P.tol, P.constr_viol_tol, P.mu_init = 1e-8, 1e-4, 0.1
    P.max_iter = 200
    P.tiny_alpha, P.tiny_count = (5e-2, 2) if f == 2 else (1e-2, 3)
    P.mu_warm, P.mu_shift = 1e-4, 2.5e-3
    P.resto_window = 1e-3 if f == 1 else 0.1
    P.kappa_eps = (30.0, 30.0, 10.0)[f]
    if f == 0:
        P.w_p, P.w_r, P.gamma, P.s_turn, P.bvy_max = 2.0, 15.0, 0.4, 0.014 * 180 / math.pi, 0.3
        P.goal_shift, P.close_radius, P.close_any = 1, 0.35, 1
    elif f == 1:
        P.w_p, P.w_r, P.gamma, P.s_turn, P.bvy_max = 0.0, 50.0, 0.2, 0.024 * 180 / math.pi, 0.35
        P.has_fen, P.select_obs, P.goal_shift, P.close_radius = 1, 1, 1, 0.15
    else:
        P.w_p, P.w_r, P.gamma, P.s_turn, P.bvy_max = 0.0, 50.0, 0.2, 0.024 * 180 / math.pi, 0.35
        P.has_fen, P.w_t, P.close_radius = 1, 2.0, 0.35
    return P


def _d(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _p(a, typ=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(typ))


def _fields(cir, elp):
    cir = _d(np.zeros((1, 0, 3)) if cir is None else cir)
    if cir.ndim == 2:
        cir = cir[None]
    F = cir.shape[0]
    elp = _d(np.zeros((F, 0, 5)) if elp is None else elp)
    if elp.ndim == 2:
        elp = elp[None]
    return F, cir, elp


def solve(P, x0, goal, leg, cir, elp, warm, field=None, last_u=None):
    dd = P.formulation == 2
    nx, nu = (3, 6) if dd else (5, 15)
    x0 = _d(x0).reshape(-1, nx)
    B = len(x0)
    goal = _d(np.broadcast_to(np.asarray(goal, dtype=np.float64).reshape(-1, 2), (B, 2)))
    leg = np.ascontiguousarray(np.broadcast_to(np.asarray(leg, dtype=np.int32), (B,)), dtype=np.int32)
    F, cir, elp = _fields(cir, elp)
    field = None if field is None else np.ascontiguousarray(field, dtype=np.int32)
    warm, last_u = _d(warm).reshape(B, nu), _d(last_u)
    u, xp, pp = np.zeros((B, nu)), np.zeros((B, 3, nx)), np.zeros((B, 3, 3))
    obj, viol = np.zeros(B), np.zeros(B)
    st, it, cl = np.zeros(B, np.int32), np.zeros(B, np.int32), np.zeros(B, np.uint8)
    lib().hostsim_solve(C.byref(P), B, _p(x0), _p(goal), _p(leg, C.c_int32), _p(field, C.c_int32), F, cir.shape[1], _p(cir),
                        elp.shape[1], _p(elp), _p(warm), _p(last_u), _p(u), _p(xp), _p(pp), _p(st, C.c_int32),
                        _p(it, C.c_int32), _p(obj), _p(viol), _p(cl, C.c_uint8))
    return dict(u=u, x_plan=xp, p_plan=pp, f=obj, status=st, iters=it, viol=viol, close2goal=cl.astype(bool))


def evaluate(P, x0, goal, leg, cir, elp, z, lam=None, field=None, last_u=None):
    dd = P.formulation == 2
    nx, n = (3, 6) if dd else (5, 9)
    x0 = _d(x0).reshape(-1, nx)
    B = len(x0)
    goal = _d(np.broadcast_to(np.asarray(goal, dtype=np.float64).reshape(-1, 2), (B, 2)))
    leg = np.ascontiguousarray(np.broadcast_to(np.asarray(leg, dtype=np.int32), (B,)), dtype=np.int32)
    F, cir, elp = _fields(cir, elp)
    K = cir.shape[1] + elp.shape[1]
    m = 3 * (K + 1) if dd else 3 * (4 + K + (1 if P.has_fen else 0))
    field = None if field is None else np.ascontiguousarray(field, dtype=np.int32)
    z, lam, last_u = _d(z).reshape(B, n), _d(lam), _d(last_u)
    f, grad, c, jac = np.zeros(B), np.zeros((B, n)), np.zeros((B, m)), np.zeros((B, m, n))
    cl, cu, hess = np.zeros((B, m)), np.zeros((B, m)), np.zeros((B, n, n))
    lib().hostsim_eval(C.byref(P), B, _p(x0), _p(goal), _p(leg, C.c_int32), _p(field, C.c_int32), F, cir.shape[1], _p(cir),
                       elp.shape[1], _p(elp), _p(last_u), _p(z), _p(lam), m, _p(f), _p(grad), _p(c), _p(jac), _p(cl), _p(cu), _p(hess))
    return dict(f=f, grad=grad, c=c, jac=jac, cl=cl, cu=cu, hess=hess)


def rollout(P, steps, x0, goal, leg, cir, elp, field=None):
    x0 = _d(x0).reshape(-1, 5)
    B = len(x0)
    goal = _d(np.broadcast_to(np.asarray(goal, dtype=np.float64).reshape(-1, 2), (B, 2)))
    leg = np.ascontiguousarray(np.broadcast_to(np.asarray(leg, dtype=np.int32), (B,)), dtype=np.int32)
    F, cir, elp = _fields(cir, elp)
    field = None if field is None else np.ascontiguousarray(field, dtype=np.int32)
    xf, traj = np.zeros((B, 5)), np.zeros((B, steps, 8))
    sd, ni, ti = np.zeros(B, np.int32), np.zeros(B, np.int32), np.zeros(B, np.int32)
    lib().hostsim_rollout(C.byref(P), B, steps, _p(x0), _p(goal), _p(leg, C.c_int32), _p(field, C.c_int32), F, cir.shape[1],
                          _p(cir), elp.shape[1], _p(elp), _p(xf), _p(sd, C.c_int32), _p(ni, C.c_int32), _p(ti, C.c_int32), _p(traj))
    return dict(x_final=xf, steps_done=sd, n_infeasible=ni, total_iters=ti, traj=traj)


def lean_math(a, y, x):
    """fsincos / fatan2 of csrc/dcbf_math.cuh evaluated on the host."""
    a, y, x = _d(a).ravel(), _d(y).ravel(), _d(x).ravel()
    n = len(a)
    sn, cs, at = np.zeros(n), np.zeros(n), np.zeros(n)
    lib().hostsim_math(n, _p(a), _p(sn), _p(cs), _p(y), _p(x), _p(at))
    return sn, cs, at


def lean_log(x):
    """flog of csrc/dcbf_math.cuh evaluated on the host."""
    x = _d(x).ravel()
    out = np.zeros(len(x))
    lib().hostsim_log(len(x), _p(x), _p(out))
    return out


def warp_tables():
    """Constant tables of the warp kernels (csrc/dcbf_warp.cuh: build_warp_tables) and the dense feature map."""
    desc, hs = np.zeros(96, np.int32), np.zeros((6, 48), np.int32)
    hc, cab, T = np.zeros((6, 48)), np.zeros((10, 6)), np.zeros((24, 9))
    rc = lib().hostsim_warp_tables(_p(desc, C.c_int32), _p(hc), _p(hs, C.c_int32), _p(cab), _p(T))
    assert rc == 0
    return dict(desc=desc, hc=hc, hs=hs, cab=cab, T=T)


def lin_descriptors():
    """closed-form descriptors of the typed linear slots (WarpTables::desc_lin_lip, desc_lin) and the DD dot-product descriptors"""
    lip, dd, ddd = np.zeros(96, np.int32), np.zeros(64, np.int32), np.zeros(64, np.int32)
    assert lib().hostsim_lin_descriptors(_p(lip, C.c_int32), _p(dd, C.c_int32), _p(ddd, C.c_int32)) == 0
    return dict(lin_lip=lip, lin_dd=dd, desc_dd=ddd)


def philox(counter, k0, k1):
    c = np.ascontiguousarray(counter, dtype=np.uint32).copy()
    lib().hostsim_philox(_p(c, C.c_uint32), C.c_uint32(k0), C.c_uint32(k1))
    return c


def gen_fields(F, seed, num, mix, margin=8.5, radius=1.0, half_gap=0.8, safe_dis=0.4, stall=2000, max_restarts=64):
    """make_field of csrc/dcbf_gen.cuh for fields 0..F-1 (the function gen_fields_kernel calls per thread)."""
    Kc, Ke = ((num + 1) // 2, num // 2) if mix else (num, 0)
    cir, elp, draws = np.zeros((F, Kc, 3)), np.zeros((F, max(Ke, 1), 5)), np.zeros(F, np.int32)
    lib().hostsim_gen_fields(F, C.c_uint64(seed), num, int(mix), C.c_double(margin), C.c_double(radius), C.c_double(half_gap),
                             C.c_double(safe_dis), stall, max_restarts, _p(cir), _p(elp), _p(draws, C.c_int32))
    return cir, elp[:, :Ke], draws


def gen_states(B, seed, cir, elp, field=None, dd=False, goal=(10.0, 10.0), bvy_max=0.3):
    F, cir, elp = _fields(cir, elp)
    nx, nw = (3, 6) if dd else (5, 15)
    field = None if field is None else np.ascontiguousarray(field, dtype=np.int32)
    x0, g, warm, last_u = np.zeros((B, nx)), np.zeros((B, 2)), np.zeros((B, nw)), np.zeros((B, 2))
    leg, att = np.zeros(B, np.int32), np.zeros(B, np.int32)
    lib().hostsim_gen_states(B, C.c_uint64(seed), int(dd), C.c_double(goal[0]), C.c_double(goal[1]), C.c_double(bvy_max),
                             _p(field, C.c_int32), F, cir.shape[1], _p(cir), elp.shape[1], _p(elp), _p(x0), _p(g), _p(leg, C.c_int32),
                             _p(warm), _p(last_u), _p(att, C.c_int32))
    return dict(x0=x0, goal=g, leg=leg, warm=warm, last_u=last_u if dd else None, attempts=att)
