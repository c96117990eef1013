"""Drop-in for /root/reference/MPC_LIP_modi.py (circles + ellipses, obstacle selection, speed/turn coupling row)."""
from __future__ import annotations

import numpy as np

from . import MPC_LIP_sig_step as _sig
from ._planner_base import LipPlannerBase, _as_obs


class MPCCBF(LipPlannerBase):
    FORM = "modi"

    def __init__(self, goals, cir_param, cir_cbf, elp_param, elp_cbf, margin, step=3, device=None, **solver_overrides):
        """MPC_LIP_modi.py:14-87."""
        self.cir_list, self.elp_list = cir_param, elp_param
        self.cir_safe, self.elp_safe = cir_cbf, elp_cbf
        self.bvy_max = 0.35
        self.sel_cir, self.sel_elp = [], []
        self._init_common(goals, cir_cbf, elp_cbf, margin, step, device, **solver_overrides)

    def select_obs(self, xk):
        """Detection-range selection (MPC_LIP_modi.py:325-338).  Kept for callers that read sel_cir / sel_elp; the
        solver applies the same rule on the GPU (dcbf_params.select_obs)."""
        x = np.ravel(xk)
        self.sel_cir = [list(c) for c in self._cir if (x[0] - c[0]) ** 2 + (x[1] - c[1]) ** 2 - c[2] ** 2 <= 16.0]
        self.sel_elp = [list(e) for e in self._elp if (x[0] - e[0]) ** 2 + (x[1] - e[1]) ** 2 - max(e[2], e[3]) ** 2 <= 16.0]

    def solveMPCCBF(self, xk, od_ev, init_guess):
        """-> (u, feasi) with u0 = init_guess verbatim (MPC_LIP_modi.py:197-301); feasi keeps Ipopt's integers
        (2 = infeasible problem detected)."""
        self._last = self._solve_one(np.ravel(xk), od_ev, np.ravel(init_guess))
        return self._last.u[0].copy(), int(self._last.status[0])

    def gen_control_test(self, state, leg_ind, init_guess, plot=False, trajec=[]):
        """MPC_LIP_modi.py:90-146 -> (xk_list[1:], p_list[0], hd_list, close_2_goal, feasi, pos_det[126,2])."""
        self.init_state = np.asarray(state, dtype=np.float64).reshape(5, 1)
        self.select_obs(self.init_state)
        _, feasi = self.solveMPCCBF(self.init_state, leg_ind, init_guess)
        r = self._last
        x_list = [r.x_plan[0, i].copy() for i in range(3)]
        hd_list = [float(r.x_plan[0, i, 4]) for i in range(3)]
        starts = [np.ravel(state).astype(np.float64)] + x_list[:2]
        pos_det = np.concatenate([self.xk_track_det(starts[j], r.p_plan[0, j], self.dt) for j in range(3)])
        return x_list, r.p_plan[0, 0].copy(), hd_list, bool(r.close2goal[0]), feasi, pos_det


class LIP_Prob(_sig.LIP_Prob):
    """MPC_LIP_modi.py:394-655."""
    FORM = "modi"

    def __init__(self, xk, M_A, M_B, A, W, dx, dp, cir_safe, elp_safe, goal, step, device=None):
        self._elp0 = elp_safe
        super().__init__(xk, M_A, M_B, A, W, dx, dp, cir_safe, goal, step, device)

    def _set_obstacles(self, cir_safe):
        elp = _as_obs(self._elp0, 5)
        self._solver.set_fields(_as_obs(cir_safe, 3)[None], elp[None] if len(elp) else None)
