"""Closed-form part of /root/reference/ALIP_plan/planner.py (class ALIP): angular-momentum LIP one-step foot
placement.  ~40 flop per call, used by the DD pipeline to turn (v, omega) into a foot target
(data_procs/logger_dd.py:356-363).  The full-order-model half of the reference class (FOM2LIP, Bezier outputs,
FROST kinematics) is out of scope.

The moving-platform ("DRS") terms need a controller object the reference never constructs (planner.py:7,62 are
commented out, amplitudes default to 0: planner.py:47-50); they are zero here, as SURVEY.md row 4 prescribes.
All methods accept scalars or numpy arrays (batched evaluation by broadcasting).
"""
from __future__ import annotations

import math

import numpy as np


class ALIPParam:
    def __init__(self, H=1.0, T=0.4, m=45.0):
        self.H, self.T, self.m = H, T, m


class ALIP:
    def __init__(self, params):
        """planner.py:15-61 (walking parameters only)."""
        self.H, self.T, self.m = params.H, params.T, params.m
        self.g, self.W = 9.81, 0.2
        self.lambda_lip = math.sqrt(self.g / self.H)
        self.mhl = self.m * self.H * self.lambda_lip
        self.t_abs = 0.0
        self.t_begining_current_Step = 0.0
        self.amp_x = self.amp_y = 0.0

    def getTimedState(self, x0, y0, t):
        """planner.py:188-208: [p, L](t) for the sagittal (x) and lateral (y) planes."""
        t = np.minimum(t, self.T)
        l, mhl = self.lambda_lip, self.mhl
        ch, sh = np.cosh(l * t), np.sinh(l * t)
        x0, y0 = np.asarray(x0, dtype=np.float64), np.asarray(y0, dtype=np.float64)
        xt = np.stack([ch * x0[..., 0] + sh / mhl * x0[..., 1], mhl * sh * x0[..., 0] + ch * x0[..., 1]], axis=-1)
        yt = np.stack([ch * y0[..., 0] - sh / mhl * y0[..., 1], -mhl * sh * y0[..., 0] + ch * y0[..., 1]], axis=-1)
        return xt, yt

    def AMprediction(self, xt, yt, t):
        """planner.py:210-230: angular momentum at the end of the step."""
        t = np.minimum(t, self.T)
        l, T = self.lambda_lip, self.T
        xt, yt = np.asarray(xt, dtype=np.float64), np.asarray(yt, dtype=np.float64)
        Ly_est = self.mhl * np.sinh(l * (T - t)) * xt[..., 0] + np.cosh(l * (T - t)) * xt[..., 1]
        Lx_est = -self.mhl * np.sinh(l * (T - t)) * yt[..., 0] + np.cosh(l * (T - t)) * yt[..., 1]
        return Ly_est, Lx_est

    def computeSw2CoM(self, Ly_est, Lx_est, Ly_des, support):
        """planner.py:232-248."""
        l, T = self.lambda_lip, self.T
        den = self.mhl * math.sinh(l * T)
        px = Ly_des / den - math.cosh(l * T) / den * Ly_est
        base = 0.5 * self.m * self.H * self.W * (l * math.sinh(l * T)) / (1 + math.cosh(l * T))
        Lx_des = np.where(np.asarray(support) == 1, base, -base)
        py = -Lx_des / den + math.cosh(l * T) / den * Lx_est
        return px, py

    def regulate_lateral_step(self, foot_index, u_lateral):
        """planner.py:346-370."""
        fi, u = np.asarray(foot_index), np.asarray(u_lateral, dtype=np.float64)
        out = np.where(fi == 1, np.clip(u, 0.1, 0.45), np.where(fi == -1, np.clip(u, -0.45, -0.1), u))
        return out if out.ndim else float(out)

    def computeStepping(self, p_sp2CoM, Ly_est, Lx_est, v_des, support):
        """planner.py:250-261 -> (px_sp2sw, py_sp2sw)."""
        Ly_des = self.m * self.H * v_des
        px_sw, py_sw = self.computeSw2CoM(Ly_est, Lx_est, Ly_des, support)
        p = np.asarray(p_sp2CoM, dtype=np.float64)
        return p[..., 0] - px_sw, self.regulate_lateral_step(support, p[..., 1] - py_sw)

    def getFootPlacement(self, speed, support, time, x_alip, y_alip):
        """(px, py, Ly_est, Lx_est) as data_procs/logger_dd.py:359 calls it.  The five-argument variant lives in the reference's
        missing top-level ALIP.py (only an orphan .pyc ships); the composition follows planner.py:263-320, which is the same
        chain on a full-order state: AMprediction at `time`, then computeStepping with the stance-to-CoM offset (x[0], y[0])."""
        x_alip, y_alip = np.asarray(x_alip, dtype=np.float64), np.asarray(y_alip, dtype=np.float64)
        Ly_est, Lx_est = self.AMprediction(x_alip, y_alip, time)
        ux, uy = self.computeStepping(np.stack([x_alip[..., 0], y_alip[..., 0]], axis=-1), Ly_est, Lx_est, speed, support)
        return ux, uy, Ly_est, Lx_est
