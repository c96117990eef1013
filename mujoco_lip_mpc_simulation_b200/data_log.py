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


# ---------------------------------------------------------------------------------------------------------------------------
# sup_learn/*.csv: one row per control tick of a run, written by the learning logger (data_procs/logger_iml.py:377-401) with
# numpy.savetxt's default '%.18e' format.
#   X_data.csv      obstacle list raveled (K x [cx, cy, r]) | CoM position (2) | CoM velocity (2) | base heading | stance foot (2)
#                   | goal (2) | leg_ind | rest_t                                             -- all in the map frame
#   y_mpc_data.csv  planned foot (2) | 0 | hd_input_pr + hd_input_cos | predicted position x_nex[0:2] | desired velocity (2)
#   y_act_data.csv  the same eight columns for what the robot did (constant over a step)
SUP_FILES = ("X_data.csv", "y_mpc_data.csv", "y_act_data.csv")


def read_sup_learn(folder: str) -> dict:
    """{'X': [T, 3K+11], 'y_mpc': [T, 8], 'y_act': [T, 8]} plus the named views of sup_learn_fields."""
    import os
    out = {}
    for key, name in zip(("X", "y_mpc", "y_act"), SUP_FILES):
        path = os.path.join(folder, name)
        if os.path.exists(path):
            out[key] = np.atleast_2d(np.loadtxt(path, delimiter=","))
    if "X" in out:
        out.update(sup_learn_fields(out["X"], out.get("y_mpc")))
    return out


def write_sup_learn(folder: str, X, y_mpc, y_act=None) -> None:
    import os
    os.makedirs(folder, exist_ok=True)
    for name, a in zip(SUP_FILES, (X, y_mpc, y_act)):
        if a is not None:
            np.savetxt(os.path.join(folder, name), np.atleast_2d(np.asarray(a, dtype=np.float64)), delimiter=",")


def sup_learn_fields(X, y_mpc=None) -> dict:
    """Named columns of the feature rows (logger_iml.py:378-383) and, if given, of the MPC rows (:392-395).  The heading input of
    the prediction is recovered as hd_input_pr = y_mpc[:, 3] - heading (hd_input_cos is the base heading; the map and robot frames
    coincide when hd_init = 0, main_sim_mpc.py:24-25)."""
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    K = (X.shape[1] - 11) // 3
    assert X.shape[1] == 3 * K + 11, "feature rows are 3K + 11 wide"
    o = 3 * K
    out = dict(obs=X[:, :o].reshape(-1, K, 3), pos=X[:, o:o + 2], vel=X[:, o + 2:o + 4], heading=X[:, o + 4], stance=X[:, o + 5:o + 7],
               goal=X[:, o + 7:o + 9], leg_ind=X[:, o + 9].astype(np.int32), rest_t=X[:, o + 10])
    if y_mpc is not None:
        y = np.atleast_2d(np.asarray(y_mpc, dtype=np.float64))
        out.update(foot=y[:, 0:2], hd_target=y[:, 3], x_nex_pos=y[:, 4:6], v_des=y[:, 6:8], hd_input_pr=y[:, 3] - out["heading"])
    return out


def sup_learn_rows(obs, pos, vel, heading, stance, goal, leg_ind, rest_t, foot, hd_input_pr, x_nex_pos, v_des=None):
    """Feature and MPC rows of B ticks in the reference layout, from the inputs and outputs of DcbfSolver.tick (foot =
    p_plan[:, 0, :2], x_nex_pos = x_next[:, :2]); obs is the un-inflated obstacle list [K, 3] of the run or [B, K, 3]."""
    pos = np.atleast_2d(np.asarray(pos, dtype=np.float64))
    B = pos.shape[0]
    obs = np.asarray(obs, dtype=np.float64)
    obs = np.broadcast_to(obs.reshape(-1, obs.shape[-2] * 3) if obs.ndim == 3 else obs.reshape(1, -1), (B, obs.shape[-2] * 3))
    col = lambda a, w: np.broadcast_to(np.asarray(a, dtype=np.float64).reshape(-1, w), (B, w))   # noqa: E731
    X = np.concatenate([obs, pos, col(vel, 2), col(heading, 1), col(stance, 2), col(goal, 2), col(leg_ind, 1), col(rest_t, 1)], axis=1)
    hd = col(heading, 1) + col(hd_input_pr, 1)
    y = np.concatenate([col(foot, 2), np.zeros((B, 1)), hd, col(x_nex_pos, 2), col(np.zeros(2) if v_des is None else v_des, 2)], axis=1)
    return X, y
