"""Drop-in for /root/reference/MPC_DD_sig_step.py (differential-drive / unicycle formulation)."""
from __future__ import annotations

import numpy as np

from . import _lipmodel
from ._planner_base import _as_obs
from .batch import DcbfSolver


class MPCCBF:
    def __init__(self, goals, cir_param, cir_cbf, elp_param, elp_cbf, margin, step=3, device=None, **solver_overrides):
        """MPC_DD_sig_step.py:12-67."""
        if step != 3:
            raise ValueError("the horizon of the reference formulations is hard-wired to N = 3 steps")
        k = _lipmodel.constants()
        self.goal = np.asarray(goals, dtype=np.float64).reshape(-1, 2)[0].reshape(2, 1)
        self.beta, self.dt, self.N, self.margin = k.beta, k.dt, step, margin
        self.cir_list, self.elp_list, self.cir_safe, self.elp_safe = cir_param, elp_param, cir_cbf, elp_cbf
        self.leg, self.x_max, self.v_max, self.v_min, self.ang_max, self.tot_time = 0.09, 5, 0.8, 0.4, np.pi / 16, 80
        self.A = np.eye(3)
        self.A_L, self.B_L, self.W, self.M_A, self.M_B = k.A, k.B, k.W, k.M_A, k.M_B
        self.B_pos_shr, self.inv_B_pos_shr = k.B_pos_shr, k.inv_B_pos_shr
        self._solver = DcbfSolver("dd", device=device, **solver_overrides)
        self._cir, self._elp = _as_obs(cir_cbf, 3), _as_obs(elp_cbf, 5)
        self._solver.set_fields_host(self._cir[None], self._elp[None] if len(self._elp) else None)

    def solveMPCCBF(self, xk, init_guess, last_u):
        """-> (u[6], fesi)  (MPC_DD_sig_step.py:123-193)."""
        r = self._solver.solve_host(np.ravel(xk).astype(np.float64)[None], self.goal.ravel()[None], None,
                                    np.ravel(init_guess).astype(np.float64)[None], last_u=np.ravel(last_u).astype(np.float64)[None])
        self._last = r
        return r.u[0].copy(), int(r.status[0])

    def gen_dd_control(self, state, init_guess, last_u, plot=False, trajec=[]):
        """MPC_DD_sig_step.py:70-120 -> (states[4][3], heading[3], control[3] of (2,1), close2goal, fesi)."""
        self.init_state = np.asarray(state, dtype=np.float64).reshape(3, 1)
        u, fesi = self.solveMPCCBF(self.init_state, init_guess, last_u)
        r = self._last
        states = [list(np.ravel(state).astype(float))] + [list(map(float, r.x_plan[0, i])) for i in range(3)]
        heading = [float(r.x_plan[0, i, 2]) for i in range(3)]
        control = [u[2 * i:2 * i + 2].reshape(2, 1).copy() for i in range(3)]
        return states, heading, control, bool(r.close2goal[0]), fesi

    def set_fields(self, cir, elp=None):
        self._solver.set_fields(cir, elp)

    def solve_batch(self, x0, warm, last_u, goal=None, field=None):
        goal = self.goal.ravel() if goal is None else goal
        return self._solver.solve(x0, goal, None, warm, field=field, last_u=last_u)

    def get_next_states(self, glo_pos, glo_vel, glo_hd, glo_p, t_rest, plot=False):
        A, B = _lipmodel.flow_matrices(t_rest, t_rest * (1.0 / self.dt))
        xk = np.concatenate([np.ravel(glo_pos), np.ravel(glo_vel), [float(glo_hd)]]).astype(np.float64)
        p = np.asarray(glo_p, dtype=np.float64).ravel()
        return A @ xk + B @ p, _lipmodel.track_det(xk, p, t_rest, self.dt)

    def xk_track_det(self, xk, contr, t_rest):
        return _lipmodel.track_det(xk, contr, t_rest, self.dt)

    def cal_foot_with_posdes(self, x_state, pos_des_glo):
        """MPC_DD_sig_step.py:265-270."""
        ax = self.A_L @ np.asarray(x_state, dtype=np.float64).ravel()
        return self.inv_B_pos_shr @ (np.asarray(pos_des_glo, dtype=np.float64).ravel() - ax[0:2])

    def select_obs(self, xk):
        x = np.ravel(xk)
        self.sel_cir = [list(c) for c in self._cir if (x[0] - c[0]) ** 2 + (x[1] - c[1]) ** 2 - c[2] ** 2 <= 16.0]
        self.sel_elp = [list(e) for e in self._elp if (x[0] - e[0]) ** 2 + (x[1] - e[1]) ** 2 - max(e[2], e[3]) ** 2 <= 16.0]

    def tube_func(self, heading_list, init_tube_value):
        return _lipmodel.tube(heading_list, init_tube_value, 0.2, 0.3)


class LIP_Prob:
    """cyipopt callback protocol of MPC_DD_sig_step.py:320-572, evaluated by the K1 kernel (z = u for DD)."""

    def __init__(self, xk, A, dt, cir_safe, elp_safe, goal, step, last_u, device=None):
        self.xk = np.asarray(xk, dtype=np.float64).reshape(3)
        self.goal = np.asarray(goal, dtype=np.float64).reshape(2)
        self.last_u = np.asarray(last_u, dtype=np.float64).reshape(2)
        self._solver = DcbfSolver("dd", device=device)
        elp = _as_obs(elp_safe, 5)
        self._solver.set_fields(_as_obs(cir_safe, 3)[None], elp[None] if len(elp) else None)

    def _eval(self, u):
        r = self._solver.evaluate(self.xk[None], self.goal[None], None, np.asarray(u, dtype=np.float64)[None],
                                  last_u=self.last_u[None], want_hess=False)
        return {k: v[0].cpu().numpy() for k, v in r.items() if v is not None}

    def objective(self, u):
        return float(self._eval(u)["f"])

    def gradient(self, u):
        return self._eval(u)["grad"]

    def constraints(self, u):
        return self._eval(u)["c"]

    def jacobian(self, u):
        return self._eval(u)["jac"]
