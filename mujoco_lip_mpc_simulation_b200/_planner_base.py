"""Shared machinery of the three drop-in MPCCBF classes: one DcbfSolver per planner object, obstacle upload, the
B = 1 legacy call path and the batched entry points."""
from __future__ import annotations

import numpy as np

from . import _lipmodel
from .batch import DcbfSolver


def _as_obs(a, width):
    a = np.asarray(a if a is not None and len(a) else np.zeros((0, width)), dtype=np.float64)
    return a.reshape(-1, width)


class LipPlannerBase:
    """Common part of MPC_LIP_sig_step.MPCCBF and MPC_LIP_modi.MPCCBF."""
    FORM = "sig_step"

    def _init_common(self, goals, cir_cbf, elp_cbf, margin, step, device=None, **solver_overrides):
        if step != 3:
            raise ValueError("the horizon of the reference formulations is hard-wired to N = 3 foot steps")
        k = _lipmodel.constants()
        self.goal = np.asarray(goals, dtype=np.float64).reshape(-1, 2)[0].reshape(2, 1)
        self.beta, self.dt, self.N, self.margin = k.beta, k.dt, step, margin
        self.leg, self.x_max = 0.09, 5
        self.bvx_max, self.bvx_min, self.bvy_min = 0.8, 0.4, 0.15
        self.ang_max = np.pi / 16
        self.step_gap, self.sigma = 0.3, k.sigma
        self.A, self.B, self.W, self.M_A, self.M_B = k.A, k.B, k.W, k.M_A, k.M_B
        self.B_vel_shr, self.inv_B_vel_shr = k.B_vel_shr, k.inv_B_vel_shr
        self.dx_du, self.dP_du = k.dx_du, k.dP_du
        self._k = k
        self._solver = DcbfSolver(self.FORM, device=device, **solver_overrides)
        self._cir = _as_obs(cir_cbf, 3)
        self._elp = _as_obs(elp_cbf, 5)
        self._solver.set_fields_host(self._cir[None], self._elp[None] if len(self._elp) else None)

    # ---- closed-form helpers of the reference call surface (host side, not on the hot path) ----------------------------
    def get_next_states(self, glo_pos, glo_vel, glo_hd, glo_p, t_rest, plot=False):
        """LIP flow to the end of the current step (MPC_LIP_sig_step.py:136-165)."""
        A, B = _lipmodel.flow_matrices(t_rest, t_rest * (1.0 / self.dt))
        xk = np.concatenate([np.ravel(glo_pos), np.ravel(glo_vel), [float(glo_hd)]]).astype(np.float64)
        p = np.asarray(glo_p, dtype=np.float64).ravel()
        return A @ xk + B @ p, _lipmodel.track_det(xk, p, t_rest, self.dt)

    def alip_des_vel(self, vx_max, leg_ind):
        """MPC_LIP_sig_step.py:168-173."""
        import math
        vdes_x = self.sigma * vx_max * self.dt / 2
        vdes_y = 0.5 * (-0.5 * leg_ind * self.step_gap) * (self.beta * math.sinh(self.beta * self.dt)) / (math.cosh(self.beta * self.dt) + 1)
        return np.array([vdes_x, vdes_y])

    def cal_foot_with_veldes(self, x_state, vel_des_glo):
        """MPC_LIP_sig_step.py:176-181."""
        ax = self.A @ np.asarray(x_state, dtype=np.float64).ravel()
        return self.inv_B_vel_shr @ (np.asarray(vel_des_glo, dtype=np.float64).ravel() - ax[2:4])

    def solve_footdisp(self, xk, u):
        """p = W (u - A xk)  (MPC_LIP_sig_step.py:302-306); returns a (3,1) array like the reference."""
        xk = np.asarray(xk, dtype=np.float64).reshape(5, 1)
        u = np.asarray(u, dtype=np.float64).reshape(5, 1)
        return self.W @ (u - self.A @ xk)

    def xk_track_det(self, xk, contr, t_rest):
        return _lipmodel.track_det(xk, contr, t_rest, self.dt)

    def tube_func(self, heading_list, init_tube_value):
        return _lipmodel.tube(heading_list, init_tube_value, 0.15, 0.5)

    # ---- batched entry points (device tensors in, device tensors out) ---------------------------------------------------
    def set_fields(self, cir, elp=None):
        """Obstacle fields for solve_batch / rollout_batch: cir [F,Kc,3], elp [F,Ke,5] (already inflated)."""
        self._solver.set_fields(cir, elp)

    def solve_batch(self, x0, leg, warm, goal=None, field=None):
        """Many re-plans at once.  `warm` is the reference's u0 per scenario ([B,15])."""
        goal = self.goal.ravel() if goal is None else goal
        return self._solver.solve(x0, goal, leg, warm, field=field)

    def rollout_batch(self, steps, x0, leg, goal=None, field=None, want_traj=True):
        goal = self.goal.ravel() if goal is None else goal
        return self._solver.rollout(steps, x0, goal, leg, field=field, want_traj=want_traj)

    # ---- B = 1 path used by the legacy methods ------------------------------------------------------------------------------
    def _solve_one(self, xk, od_ev, u0):
        xk = np.asarray(xk, dtype=np.float64).ravel()
        r = self._solver.solve_host(xk[None], self.goal.ravel()[None], np.array([1 if od_ev > 0 else -1], dtype=np.int32),
                                    np.asarray(u0, dtype=np.float64).reshape(1, 15))
        return r
