"""Host-side constants of the LIP step-to-step model shared by the drop-in planner classes.

Mirrors what MPCCBF.__init__ precomputes in the reference (MPC_LIP_sig_step.py:16-86, MPC_LIP_modi.py:16-87,
MPC_DD_sig_step.py:14-67); plain float64 ndarrays instead of np.matrix.  The solve itself never uses these: they
exist so that the small closed-form helpers of the call surface (get_next_states, solve_footdisp, xk_track_det, ...)
and attribute access (planner.A, planner.dx_du, ...) keep working.
"""
from __future__ import annotations

import math

import numpy as np

HEIGHT, GRAV, DT = 1.0, 9.81, 0.4
BETA = math.sqrt(GRAV / HEIGHT)


def flow_matrices(t: float, heading_gain: float = 1.0):
    """A(t), B(t) of the LIP flow over a duration t (MPC_LIP_sig_step.py:47-56 for t = dt; :138-147 with the heading
    column scaled by t/dt for a partial step)."""
    ch, sh = math.cosh(BETA * t), math.sinh(BETA * t)
    A = np.eye(5)
    A[0, 0] = A[1, 1] = A[2, 2] = A[3, 3] = ch
    A[0, 2] = A[1, 3] = sh / BETA
    A[2, 0] = A[3, 1] = sh * BETA
    B = np.zeros((5, 3))
    B[0, 0] = B[1, 1] = 1.0 - ch
    B[2, 0] = B[3, 1] = -sh * BETA
    B[4, 2] = heading_gain
    return A, B


class LipConstants:
    def __init__(self):
        self.beta, self.dt = BETA, DT
        self.A, self.B = flow_matrices(DT)
        ch, sh = math.cosh(BETA * DT), math.sinh(BETA * DT)
        wa, wb = 5.0, 1.0
        den = wa * (ch - 1.0) ** 2 + wb * (sh * BETA) ** 2
        W = np.zeros((3, 5))
        W[0, 0] = W[1, 1] = -wa * (ch - 1.0) / den
        W[0, 2] = W[1, 3] = -wb * sh * BETA / den
        W[2, 4] = 1.0
        self.W = W
        self.M_A = self.A - self.B @ W @ self.A
        self.M_B = self.B @ W
        self.B_vel_shr = self.B[2:4, 0:2]
        self.inv_B_vel_shr = np.linalg.inv(self.B_vel_shr)
        self.B_pos_shr = self.B[0:2, 0:2]
        self.inv_B_pos_shr = np.linalg.inv(self.B_pos_shr)
        pre = [self.M_B, self.M_A @ self.M_B, self.M_A @ self.M_A @ self.M_B]
        pl = [W, -W @ self.A @ self.M_B, -W @ self.A @ self.M_A @ self.M_B]
        self.dx_du = np.zeros((20, 15))
        self.dP_du = np.zeros((9, 15))
        for r in range(1, 4):
            for c in range(r):
                self.dx_du[5 * r:5 * r + 5, 5 * c:5 * c + 5] = pre[r - 1 - c]
        for r in range(3):
            for c in range(r + 1):
                self.dP_du[3 * r:3 * r + 3, 5 * c:5 * c + 5] = pl[r - c]
        # sigma = beta * coth(beta dt / 2)   (MPC_LIP_sig_step.py:44, mpmath.coth in the reference)
        self.sigma = BETA / math.tanh(DT * BETA / 2.0)

    def p_from_u(self, xk, u):
        """z = (p0, p1, p2) of a reference decision vector u (R^15)."""
        x = np.asarray(xk, dtype=np.float64).ravel().copy()
        u = np.asarray(u, dtype=np.float64).ravel()
        out = np.zeros(9)
        for i in range(3):
            p = self.W @ (u[5 * i:5 * i + 5] - self.A @ x)
            out[3 * i:3 * i + 3] = p
            x = self.A @ x + self.B @ p
        return out


_CONST = None


def constants() -> LipConstants:
    global _CONST
    if _CONST is None:
        _CONST = LipConstants()
    return _CONST


def track_det(xk, contr, t_rest, dt=DT):
    """Dense position samples of one step: the start position followed by the LIP flow at t = 0, 0.01, ... (<= t_rest),
    as MPCCBF.xk_track_det builds them (MPC_LIP_sig_step.py:281-299)."""
    xk = np.asarray(xk, dtype=np.float64).ravel()
    p = np.asarray(contr, dtype=np.float64).ravel()
    ts = np.arange(0, t_rest + 0.01, 0.01)
    out = [xk[0:2].copy()]
    for t in ts:
        A, B = flow_matrices(float(t), float(t) * (1.0 / dt))
        out.append((A @ xk + B @ p)[0:2])
    return np.array(out)


def tube(heading_list, init_tube_value, width, gain_in, gain_out=0.7):
    """Heading low-pass with a dead-band (MPCCBF.tube_func: MPC_LIP_sig_step.py:309-327 uses 0.15/0.5/0.7,
    MPC_DD_sig_step.py:290-308 uses 0.2/0.3/0.7)."""
    new = np.zeros_like(np.asarray(heading_list, dtype=np.float64))
    v = init_tube_value
    for i, h in enumerate(heading_list):
        d = h - v
        if d > 0:
            v += (gain_in if width > d else gain_out) * d
        elif d < 0:
            v += (gain_in if -width < d else gain_out) * d
        new[i] = v
    return new


def angle_a_minus_b(a, b):
    """Logger.angle_A_minus_B (data_procs/logger_mpc.py:169-175), elementwise."""
    r = np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)
    r = np.where((r < 0) & (np.abs(r) > math.pi), r + 2 * math.pi, np.where((r > 0) & (np.abs(r) > math.pi), r - 2 * math.pi, r))
    return r


def logger_tube(turning, cur_hd, width=0.15, gain_in=0.4, gain_out=0.7):
    """Logger.tube_func (data_procs/logger_mpc.py:284-300): the pending turn scaled by 0.4 inside the tube and 0.7 outside."""
    t, cur = np.asarray(turning, dtype=np.float64), np.asarray(cur_hd, dtype=np.float64)
    gain = np.where(t > 0, np.where(width > t, gain_in, gain_out), np.where(-width < t, gain_in, gain_out))
    tube = np.where(t != 0, cur + gain * t, cur)
    return angle_a_minus_b(tube, cur)


def avg_hd(cur_hd, nex_turn, mpc_hds):
    """Logger.avg_hd (data_procs/logger_mpc.py:208-215): (nex_turn + three heading increments of the last plan) / 4."""
    cur, h = np.asarray(cur_hd, dtype=np.float64), np.asarray(mpc_hds, dtype=np.float64)
    s = np.asarray(nex_turn, dtype=np.float64) + angle_a_minus_b(h[..., 0], cur)
    s = s + angle_a_minus_b(h[..., 1], h[..., 0])
    s = s + angle_a_minus_b(h[..., 2], h[..., 1])
    return s / 4.0
