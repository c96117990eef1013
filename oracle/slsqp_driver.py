"""ORACLE (test infrastructure) -- SciPy stand-in for the cyipopt solve.

The reference solves each NLP with cyipopt -> Ipopt 3.14.x -> HSL MA57 (un-vendored, unpinned third-party code,
absent from this image; call sites MPC_LIP_sig_step.py:256-277, MPC_LIP_modi.py:274-295,
MPC_DD_sig_step.py:171-191).  This driver reaches the same KKT points with scipy.optimize SLSQP on the same
four callbacks and classifies (local) infeasibility with an elastic phase-1 problem, which is the mathematical
content of Ipopt status 2 ("restoration converged to a point of local infeasibility").

PARITY UNPINNED with respect to Ipopt's iteration-capped exits (-1, 1, -2): those depend on Ipopt's internal
L-BFGS path, which no file in the reference records.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import minimize

FEAS_TOL = 1e-6      # a point counts as feasible below this max row violation
INFEAS_TOL = 1e-4    # Ipopt constr_viol_tol; above this after phase-1 => status 2


def violation(c, cl, cu):
    return float(max(0.0, np.max(np.maximum(cl - c, 0.0), initial=0.0), np.max(np.maximum(c - cu, 0.0), initial=0.0)))


def _ineq(c_fun, j_fun, cl, cu):
    lo = np.isfinite(cl)
    hi = np.isfinite(cu)

    def fun(u):
        c = c_fun(u)
        return np.concatenate([c[lo] - cl[lo], cu[hi] - c[hi]])

    def jac(u):
        J = j_fun(u)
        return np.concatenate([J[lo], -J[hi]], axis=0)
    return fun, jac


def _slsqp(f, g, c_fun, j_fun, u0, cl, cu, lb, ub, maxiter=400):
    fun, jac = _ineq(c_fun, j_fun, cl, cu)
    bounds = None if lb is None else list(zip(lb, ub))
    res = minimize(f, u0, jac=g, method="SLSQP", bounds=bounds,
                   constraints=[{"type": "ineq", "fun": fun, "jac": jac}],
                   options={"ftol": 1e-15, "maxiter": maxiter})
    return res


def phase1(c_fun, j_fun, u0, cl, cu, lb, ub):
    """min sum(t) s.t. cl - t <= c(u) <= cu + t, t >= 0  (one elastic per row)."""
    n, m = len(u0), len(cl)
    lo = np.isfinite(cl)
    hi = np.isfinite(cu)
    c0 = c_fun(u0)
    t0 = np.maximum(0.0, np.maximum(np.where(lo, cl - c0, 0.0), np.where(hi, c0 - cu, 0.0))) + 1e-3

    def fun(w):
        c = c_fun(w[:n])
        t = w[n:]
        return np.concatenate([(c + t - cl)[lo], (cu + t - c)[hi]])

    def jac(w):
        J = j_fun(w[:n])
        I = np.eye(m)
        return np.concatenate([np.hstack([J, I])[lo], np.hstack([-J, I])[hi]], axis=0)

    bnds = ([(None, None)] * n if lb is None else list(zip(lb, ub))) + [(0.0, None)] * m
    grad = np.concatenate([np.zeros(n), np.ones(m)])
    res = minimize(lambda w: float(np.sum(w[n:])), np.concatenate([u0, t0]), jac=lambda w: grad,
                   method="SLSQP", bounds=bnds, constraints=[{"type": "ineq", "fun": fun, "jac": jac}],
                   options={"ftol": 1e-15, "maxiter": 400})
    u = res.x[:n]
    return u, violation(c_fun(u), cl, cu)


def solve_callbacks(f, g, c_fun, j_fun, u0, cl, cu, lb=None, ub=None):
    """-> dict(u, f, status, viol, nit, how).  status uses Ipopt's integers: 0 solved, 2 locally infeasible,
    -1 otherwise (not converged)."""
    u0 = np.asarray(u0, dtype=np.float64)
    res = _slsqp(f, g, c_fun, j_fun, u0, cl, cu, lb, ub)
    u = res.x
    viol = violation(c_fun(u), cl, cu)
    how = "slsqp"
    nit = int(res.nit)
    if not (res.success and viol <= FEAS_TOL):
        # try to restore feasibility first, then re-optimise from the restored point
        best_u, best_v = u, viol
        for start in (u0, u):
            ur, vr = phase1(c_fun, j_fun, start, cl, cu, lb, ub)
            if vr < best_v:
                best_u, best_v = ur, vr
            if vr <= FEAS_TOL:
                res2 = _slsqp(f, g, c_fun, j_fun, ur, cl, cu, lb, ub)
                v2 = violation(c_fun(res2.x), cl, cu)
                nit += int(res2.nit)
                if res2.success and v2 <= FEAS_TOL:
                    return dict(u=res2.x, f=float(f(res2.x)), status=0, viol=v2, nit=nit, how="phase1+slsqp")
        if best_v > INFEAS_TOL:
            return dict(u=best_u, f=float(f(best_u)), status=2, viol=best_v, nit=nit, how="phase1-infeasible")
        return dict(u=best_u, f=float(f(best_u)), status=-1, viol=best_v, nit=nit, how="not-converged")
    return dict(u=u, f=float(f(u)), status=0, viol=viol, nit=nit, how=how)
