"""Reader / writer of the reference's run logs (data_procs/logger_mpc.py:449-474 writes them, plot_data_cir.py reads them): one
pickle per quantity with a common path prefix,

    <prefix>pos.pkl, time.pkl, foot.pkl, heading.pkl, body_vel.pkl          arrays sampled along the run
    <prefix>cir.pkl, ellp.pkl                                              obstacle lists [[cx, cy, r], ...], [[cx, cy, a, b, phi], ...]
    <prefix>real_end.pkl, pred_end.pkl                                     realised / predicted step ends
    <prefix>pred_feasi_end.pkl, pred_fail_end.pkl, pred_full_end.pkl       lists of plan trajectories (126 x 2 each): plans whose
                                                                           status was != 2 / == 2 (main_sim_mpc.py:118-121) / all

plus the inversion of a recorded plan trajectory back to the plan itself: every 42-row segment of pos_det is the start position
followed by the LIP flow  p + cosh(beta t) (x - p) + sinh(beta t)/beta v  at t = 0, 0.01, ..., 0.40 (MPC_LIP_modi.py:117-122,
304-322), so (x_k, v_k, p_k) follow from a 3-parameter least-squares fit per coordinate."""
from __future__ import annotations

import pickle

import numpy as np

from . import _lipmodel

KEYS_ARRAY = ("pos", "time", "foot", "heading", "body_vel", "real_end")
KEYS_LIST = ("cir", "ellp", "pred_end", "pred_feasi_end", "pred_fail_end", "pred_full_end")


def read_run(prefix: str) -> dict:
    """All quantities of one recorded run, e.g. read_run('data_log/LIP_me1_'); missing files are skipped."""
    out = {}
    for k in KEYS_ARRAY + KEYS_LIST + ("turning",):
        try:
            with open(prefix + k + ".pkl", "rb") as f:
                out[k] = pickle.load(f)
        except FileNotFoundError:
            pass
    return out


def write_run(prefix: str, run: dict) -> None:
    """Write a run in the reference's layout and types (arrays stay arrays, obstacle / plan collections are lists)."""
    for k, v in run.items():
        if k in KEYS_ARRAY or k == "turning":
            v = np.asarray(v, dtype=np.float64)
        elif k in ("cir", "ellp"):
            v = [list(map(float, o)) for o in v]
        elif k in KEYS_LIST:
            v = [np.asarray(a, dtype=np.float64) for a in v]
        else:
            raise KeyError(k)
        with open(prefix + k + ".pkl", "wb") as f:
            pickle.dump(v, f)


def _basis():
    k = _lipmodel.constants()
    t = np.arange(0, k.dt + 0.01, 0.01)
    return np.stack([np.ones_like(t), np.cosh(k.beta * t), np.sinh(k.beta * t) / k.beta], axis=1)   # [41, 3]


def plan_from_pos_det(pos_det):
    """(x[3,2], v[3,2], p[3,2], residual) of a recorded 126 x 2 plan trajectory: start position, start velocity and foot placement
    of each of the three planned steps, and the largest deviation of the fit from the record."""
    pd = np.asarray(pos_det, dtype=np.float64).reshape(3, 42, 2)
    Bm = _basis()
    x, v, p = np.zeros((3, 2)), np.zeros((3, 2)), np.zeros((3, 2))
    res = 0.0
    for j in range(3):
        coef, *_ = np.linalg.lstsq(Bm, pd[j, 1:], rcond=None)    # rows: p, x - p, v
        p[j], x[j], v[j] = coef[0], coef[0] + coef[1], coef[2]
        res = max(res, float(np.max(np.abs(Bm @ coef - pd[j, 1:]))), float(np.max(np.abs(pd[j, 0] - x[j]))))
    return x, v, p, res


def run_from_rollout(traj, cir, elp, plans=None, status=None, dt=None):
    """A run dict from one scenario of a closed-loop rollout (dcbf_rollout `traj[steps, 8]` = px, py, vx, vy, theta, foot_x, foot_y,
    status after each step; NaN once stopped), sampled once per step.  `plans` (optional, [steps, 126, 2]) are filed under
    pred_feasi_end / pred_fail_end by `status` like main_sim_mpc.py:118-121."""
    k = _lipmodel.constants()
    dt = k.dt if dt is None else dt
    tr = np.asarray(traj, dtype=np.float64)
    tr = tr[~np.isnan(tr[:, 0])]
    n = len(tr)
    c, s = np.cos(tr[:, 4]), np.sin(tr[:, 4])
    run = dict(pos=tr[:, 0:2], time=dt * np.arange(1, n + 1), foot=tr[:, 5:7], heading=tr[:, 4],
               body_vel=np.stack([c * tr[:, 2] + s * tr[:, 3], -s * tr[:, 2] + c * tr[:, 3]], axis=1),
               cir=cir, ellp=elp, real_end=tr[:, 0:2])
    if plans is not None:
        st = tr[:, 7].astype(int) if status is None else np.asarray(status)[:n]
        plans = [np.asarray(a) for a in plans[:n]]
        run.update(pred_full_end=plans, pred_feasi_end=[a for a, q in zip(plans, st) if q != 2],
                   pred_fail_end=[a for a, q in zip(plans, st) if q == 2], pred_end=[a[[0, 41]] for a in plans])
    return run
