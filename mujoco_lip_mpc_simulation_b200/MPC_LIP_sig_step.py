"""Drop-in for /root/reference/MPC_LIP_sig_step.py: same class names, constructor, methods and return values; the
solve runs on the GPU (libdcbf_mpc.so) instead of cyipopt.

    from mujoco_lip_mpc_simulation_b200.MPC_LIP_sig_step import MPCCBF
    mpc = MPCCBF(goal, obs_list, obs_safe, margin)
    x_list, p0, hd_list, close_2_goal = mpc.gen_control_test(state, leg_ind, init_guess)
"""
from __future__ import annotations

import numpy as np

from ._planner_base import LipPlannerBase


class MPCCBF(LipPlannerBase):
    FORM = "sig_step"

    def __init__(self, goals, obs_param, obs_cbf, margin, step=3, device=None, **solver_overrides):
        """MPC_LIP_sig_step.py:14-86.  obs_param: raw circles (plot only); obs_cbf: inflated circles used by the D-CBF rows."""
        self.obs_list, self.obs_safe, self.power = obs_param, obs_cbf, 4
        self.bvy_max = 0.3
        self._init_common(goals, obs_cbf, None, margin, step, device, **solver_overrides)

    def solveMPCCBF(self, xk, od_ev, init_guess):
        """-> u (R^15).  Warm-start rule of MPC_LIP_sig_step.py:185-189: None -> [xk,xk,xk], else the shifted previous
        plan [g1, g2, g2].  The returned vector uses the representative u_k := x_{k+1} (same p_k and x_{k+1} as any
        other member of the reference's 6-dimensional solution family)."""
        x = np.ravel(xk).astype(np.float64)
        if init_guess is None:
            u0 = np.concatenate([x, x, x])
        else:
            u0 = np.concatenate([np.ravel(init_guess[1]), np.ravel(init_guess[2]), np.ravel(init_guess[2])]).astype(np.float64)
        self._last = self._solve_one(x, od_ev, u0)
        return self._last.u[0].copy()

    def gen_control_test(self, state, leg_ind, init_guess, plot=False, trajec=[]):
        """MPC_LIP_sig_step.py:89-133 -> (xk_list[1:], p_list[0], hd_list, close_2_goal)."""
        self.init_state = np.asarray(state, dtype=np.float64).reshape(5, 1)
        self.solveMPCCBF(self.init_state, leg_ind, init_guess)
        r = self._last
        x_list = [r.x_plan[0, i].copy() for i in range(3)]
        hd_list = [float(r.x_plan[0, i, 4]) for i in range(3)]
        return x_list, r.p_plan[0, 0].copy(), hd_list, bool(r.close2goal[0])

    @property
    def last_status(self):
        """Ipopt-style status of the most recent solve (the reference discards it, MPC_LIP_sig_step.py:277-278)."""
        return int(self._last.status[0])


class LIP_Prob:
    """cyipopt callback protocol of MPC_LIP_sig_step.py:337-548, evaluated by the K1 kernel (dcbf_eval) in the reduced
    space and mapped back to the reference's u-space with dP_du (grad_u = dP_du^T grad_z, jac_u = jac_z dP_du)."""
    FORM = "sig_step"

    def __init__(self, xk, M_A, M_B, A, W, dx, dp, obs_safe, goal, step, device=None):
        from . import _lipmodel
        from .batch import DcbfSolver
        self._k = _lipmodel.constants()
        self.xk = np.asarray(xk, dtype=np.float64).reshape(5)
        self.goal = np.asarray(goal, dtype=np.float64).reshape(2)
        self.N = step
        self._solver = DcbfSolver(self.FORM, device=device)
        self._set_obstacles(obs_safe)
        self._leg = 1   # only selects which bound vector dcbf_eval would report; the callbacks do not depend on it

    def _set_obstacles(self, obs_safe):
        self._solver.set_fields(np.asarray(obs_safe, dtype=np.float64).reshape(1, -1, 3))

    def _eval(self, u):
        z = self._k.p_from_u(self.xk, u)
        r = self._solver.evaluate(self.xk[None], self.goal[None], [self._leg], z[None], want_hess=False)
        return {k: v[0].cpu().numpy() for k, v in r.items() if v is not None}

    def objective(self, u):
        return float(self._eval(u)["f"])

    def gradient(self, u):
        return self._k.dP_du.T @ self._eval(u)["grad"]

    def constraints(self, u):
        return self._eval(u)["c"]

    def jacobian(self, u):
        return self._eval(u)["jac"] @ self._k.dP_du
