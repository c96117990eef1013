"""Replay of the reference's recorded runs (tests/golden/recorded_runs.npz, made by oracle/gen_recorded.py from
/root/reference/data_log) through a solver backend -- test infrastructure shared by tests/test_recorded_cpu.py (oracle backend),
tests/test_gpu_recorded.py and tools/recorded_report.py (CUDA backend through the C ABI).

The replay is the control loop of main_sim_mpc.py:73-121 + Logger.gen_nex_foot_input (data_procs/logger_mpc.py:318-341) run
OPEN LOOP on the logged robot states: per step, Logger.set_stf_head (hd_input_pr from the previous step's last plan:
tube_func + avg_hd), then at every logged re-plan tick the LIP prediction to the end of the step (get_next_states), the
warm-start rule (previous plan verbatim; [x_nex] * 3 the first time) and the re-plan.  The plan the reference FILED for a step is
the last re-plan of that step; it is compared with ours: Ipopt's verdict (pred_fail <=> status 2) and the foot placements.
"""
from __future__ import annotations

import os

import numpy as np

from mujoco_lip_mpc_simulation_b200 import _lipmodel

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "recorded_runs.npz")
SAFE_DIS = 0.4
GOAL = (10.0, 10.0)
KC, KE = 6, 4
FAR = 1.0e4   # padding obstacles sit here: MPCCBF.select_obs (MPC_LIP_modi.py:325-338) never selects them, so they are no rows


def load():
    return np.load(G)


def padded_fields(g):
    """[R, KC, 3] circles and [R, KE, 5] ellipses, inflated, padded with far-away obstacles the selection drops."""
    R = len(g["lip_name"])
    cir = np.tile([FAR, FAR, 0.1], (R, KC, 1)).astype(np.float64)
    elp = np.tile([FAR, FAR, 0.1, 0.1, 0.0], (R, KE, 1)).astype(np.float64)
    for r in range(R):
        nc, ne = int(g["lip_n_cir"][r]), int(g["lip_n_elp"][r])
        cir[r, :nc] = g["lip_cir"][r, :nc] + [0, 0, SAFE_DIS]
        elp[r, :ne] = g["lip_elp"][r, :ne] + [0, 0, SAFE_DIS, SAFE_DIS, 0]
    return cir, elp


def tick_tables(g):
    """per-run tables of the logged quantities at the re-plan ticks: dict of [R, S, T, ...] arrays (T = re-plan ticks per step,
    padded by repeating the last valid entry), plus n_steps[R], n_ticks[R] and i[R, T] (tick number inside the step)."""
    R = len(g["lip_name"])
    nst = g["lip_n_steps"].astype(int)
    ntk = np.array([40 if e == 1 else 4 for e in g["lip_ticks_per_step"]])
    S, T = int(nst.max()), int(ntk.max())
    tab = dict(pos=np.zeros((R, S, T, 2)), body_vel=np.zeros((R, S, T, 2)), heading=np.zeros((R, S, T)), foot=np.zeros((R, S, T, 2)),
               turning=np.full((R, S, T), np.nan), i=np.zeros((R, T), dtype=int))
    run, step, ii = g["tick_run"], g["tick_step"], g["tick_i"]
    for r in range(R):
        every = int(g["lip_ticks_per_step"][r])
        sel = run == r
        k = (ii[sel] // every if every == 10 else ii[sel]).astype(int)
        for name in ("pos", "body_vel", "heading", "foot", "turning"):
            tab[name][r, step[sel], k] = g["tick_" + name][sel]
        tab["i"][r, :ntk[r]] = np.arange(ntk[r]) * every
        for name in ("pos", "body_vel", "heading", "foot", "turning"):    # padding: repeat the last valid entry
            tab[name][r, :, ntk[r]:] = tab[name][r, :, ntk[r] - 1:ntk[r]]
            tab[name][r, nst[r]:] = tab[name][r, nst[r] - 1:nst[r]]
        tab["i"][r, ntk[r]:] = tab["i"][r, ntk[r] - 1]
    return tab, nst, ntk


def map_velocity(body_vel, heading):
    """Logger.vel_fot_loc_2_map_glo (logger_mpc.py:158-163): foot-frame CoM velocity rotated by the base heading"""
    c, s = np.cos(heading), np.sin(heading)
    return np.stack([c * body_vel[..., 0] - s * body_vel[..., 1], s * body_vel[..., 0] + c * body_vel[..., 1]], axis=-1)


def planner_leg(pos, foot, heading):
    """the od_ev argument of solveMPCCBF: -leg_ind (logger_mpc.py:336), leg_ind < 0 = the stance foot is the left one
    (logger_mpc.py:193-202); read off the geometry: +1 when the stance foot lies to the left of the heading direction"""
    r = foot - pos
    return np.where(np.cos(heading) * r[..., 1] - np.sin(heading) * r[..., 0] > 0, 1, -1).astype(np.int32)


class OracleBackend:
    """CPU: host mirrors of the prediction / heading rules + oracle/dcbf_oracle.c for the re-plan (checker side)."""

    def __init__(self, cir, elp, threads=None, **over):
        from oracle import c_oracle
        self.co, self.cir, self.elp = c_oracle, cir, elp
        self.P = c_oracle.params("modi", max_iter=300, **{{"w_p": "p"}.get(k, k): v for k, v in over.items()})
        self.threads = threads or os.cpu_count() or 4

    def heading_input(self, cur_hd, nex_turn, hds):
        nt = _lipmodel.logger_tube(nex_turn, cur_hd)
        return nt, _lipmodel.avg_hd(cur_hd, nt, hds)

    def tick(self, pos, vel, hd, foot, hd_pr, t_rest, leg, field, prev, first):
        B = len(pos)
        xn = np.zeros((B, 5))
        for b in range(B):
            A, Bm = _lipmodel.flow_matrices(float(t_rest[b]), float(t_rest[b]) / _lipmodel.DT)
            xn[b] = A @ np.array([pos[b, 0], pos[b, 1], vel[b, 0], vel[b, 1], hd[b]]) + Bm @ np.array([foot[b, 0], foot[b, 1], hd_pr[b]])
        warm = np.where(first[:, None], np.tile(xn, (1, 3)), prev.reshape(B, 15))
        o = self.co.solve_batch(self.P, xn, np.tile(GOAL, (B, 1)), leg, self.cir, self.elp, warm, field=field, threads=self.threads)
        return xn, o["x_plan"], o["p_plan"], o["status"], o["viol"]


class CudaBackend:
    """GPU: dcbf_heading_input + dcbf_tick through the C ABI (the product)."""

    def __init__(self, cir, elp, device=0, **over):
        import torch
        from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
        self.torch = torch
        self.s = DcbfSolver("modi", device=device, max_iter=300, **over)
        self.s.set_fields(cir, elp)
        self.dev = self.s.tdev

    def heading_input(self, cur_hd, nex_turn, hds):
        t = self.torch
        nt = t.as_tensor(nex_turn, dtype=t.float64, device=self.dev).clone()
        out = self.s.heading_input(cur_hd, nt, mpc_hds=hds)
        return nt.cpu().numpy(), out.cpu().numpy()

    def tick(self, pos, vel, hd, foot, hd_pr, t_rest, leg, field, prev, first):
        B = len(pos)
        gp = np.concatenate([foot, hd_pr[:, None]], axis=1)
        mode = np.where(first, 2, 0).astype(np.uint8)
        r = self.s.tick(pos, vel, hd, gp, t_rest, np.tile(GOAL, (B, 1)), leg, prev_plan=prev.reshape(B, 15), mode=mode, field=field,
                        want_pos_det=False)
        pl = r["plan"]
        return (r["x_next"].cpu().numpy(), pl.x_plan.cpu().numpy(), pl.p_plan.cpu().numpy(), pl.status.cpu().numpy(), pl.viol.cpu().numpy())


# Constants of the recorded runs.  LIP_mexy is the run main_sim_mpc.py writes as shipped, and LIP_mexx / LIP_dcbf behave like it:
# the shipped MPC_LIP_modi.py constants reproduce their plans (median |dp0| 3e-10 .. 6e-7).  The twenty LIP_me<k> runs were made
# with an earlier state of that file -- the two constants in which MPC_LIP_modi.py differs from MPC_LIP_sig_step.py still had the
# sig_step values (bvy_max 0.30, MPC_LIP_sig_step.py:38; p = 2, :340): their recorded plans hold node speeds
# sqrt(0.4^2 + 0.30^2) = 0.5000 and never sqrt(0.4^2 + 0.35^2) = 0.5315, and with p = 2 the median |dp0| drops from 1e-3 to 6e-10.
def run_params(name):
    return {} if name in ("mexy", "mexx", "dcbf") else dict(bvy_max=0.30, w_p=2.0)


def replay_all(make_backend, g=None):
    """replay every LIP run with the constants it was recorded with; make_backend(cir, elp, **params) -> backend"""
    g = load() if g is None else g
    cir, elp = padded_fields(g)
    names = [str(n) for n in g["lip_name"]]
    outs = []
    for key in sorted({tuple(sorted(run_params(n).items())) for n in names}):
        runs = [r for r, n in enumerate(names) if tuple(sorted(run_params(n).items())) == key]
        outs.append(replay_lip(make_backend(cir, elp, **dict(key)), g, runs))
    out = {k: np.concatenate([o[k] for o in outs]) for k in outs[0]}
    order = np.lexsort((out["step"], out["run"]))
    return {k: v[order] for k, v in out.items()}


def replay_lip(backend, g=None, runs=None):
    """-> dict of per-plan arrays aligned with g['plan_*'] (restricted to `runs` if given): status, p[.,3,3], x_nex[.,5],
    x_plan[.,3,5], viol, hd_pr (the heading input used), and for the run with a logged heading input also hd_pr_logged."""
    g = load() if g is None else g
    tab, nst, ntk = tick_tables(g)
    R = len(nst)
    use = np.arange(R) if runs is None else np.asarray(runs)
    n = len(use)
    S, T = int(nst[use].max()), int(ntk[use].max())
    nex_turn, hds = np.zeros(n), np.zeros((n, 3))
    prev, first = np.zeros((n, 3, 5)), np.ones(n, dtype=bool)
    hd_pr = np.zeros(n)
    field = use.astype(np.int32)
    res = {k: [] for k in ("run", "step", "status", "p", "x_nex", "x_plan", "viol", "hd_pr", "hd_pr_logged")}
    for s in range(S):
        for k in range(T):
            live = np.nonzero((nst[use] > s) & (ntk[use] > k))[0]
            if len(live) == 0:
                continue
            u_ = use[live]
            pos, bv, hd, foot = tab["pos"][u_, s, k], tab["body_vel"][u_, s, k], tab["heading"][u_, s, k], tab["foot"][u_, s, k]
            if k == 0:     # Logger.set_stf_head: first tick of a step
                nex_turn[live], hd_pr[live] = backend.heading_input(hd, nex_turn[live], hds[live])
            t_rest = _lipmodel.DT - 0.01 * tab["i"][u_, k]
            leg = planner_leg(pos, foot, hd)
            xn, xp, pp, st, viol = backend.tick(pos, map_velocity(bv, hd), hd, foot, hd_pr[live], t_rest, leg, field[live], prev[live], first[live])
            prev[live], nex_turn[live], hds[live], first[live] = xp, pp[:, 0, 2], xp[:, :, 4], False
            for jj, j in enumerate(live):
                if k == ntk[use[j]] - 1:      # the re-plan the reference filed for this step
                    res["run"].append(use[j]); res["step"].append(s); res["status"].append(st[jj]); res["p"].append(pp[jj]); res["x_nex"].append(xn[jj])
                    res["x_plan"].append(xp[jj]); res["viol"].append(viol[jj]); res["hd_pr"].append(hd_pr[j]); res["hd_pr_logged"].append(tab["turning"][use[j], s, k])
    out = {k: np.array(v) for k, v in res.items()}
    # align with the fixture's plan order (run-major, step-minor)
    order = np.lexsort((out["step"], out["run"]))
    out = {k: v[order] for k, v in out.items()}
    sel = np.isin(g["plan_run"], use)
    assert np.array_equal(out["run"], g["plan_run"][sel]) and np.array_equal(out["step"], g["plan_step"][sel])
    out["label"] = g["plan_label"][sel]
    out["p_rec"], out["x_rec"], out["v_rec"] = g["plan_p"][sel], g["plan_x"][sel], g["plan_v"][sel]
    return out


def summarize_lip(out, names=None):
    """agreement numbers of a replay: start-state reproduction, Ipopt verdict vs status 2, first foot placement"""
    lab, st = out["label"] == 2, out["status"] == 2
    dx0 = np.abs(np.concatenate([out["x_rec"][:, 0], out["v_rec"][:, 0]], axis=1) - out["x_nex"][:, :4]).max(axis=1)
    both = ~lab & (out["status"] == 0)
    dp0 = np.abs(out["p"][:, 0, :2] - out["p_rec"][:, 0]).max(axis=1)
    return dict(n=len(lab), start_state_err=float(dx0.max()), class_agree=float(np.mean(lab == st)),
                rec_fail=int(lab.sum()), rec_fail_ours_infeasible=int((lab & st).sum()), rec_fail_ours_solved=int((lab & (out["status"] == 0)).sum()),
                rec_ok_ours_infeasible=int((~lab & st).sum()), ours_other=int(np.sum((out["status"] != 0) & (out["status"] != 2))),
                both_feasible=int(both.sum()), dp0_median=float(np.median(dp0[both])) if both.any() else float("nan"),
                dp0_le_1e4=float(np.mean(dp0[both] <= 1e-4)) if both.any() else float("nan"),
                dp0_le_1e3=float(np.mean(dp0[both] <= 1e-3)) if both.any() else float("nan"),
                dp0_le_1e2=float(np.mean(dp0[both] <= 1e-2)) if both.any() else float("nan"))


# ---- differential drive -------------------------------------------------------------------------------------------------------
def dd_inputs(g):
    """start state, recovered controls (used as warm start and as previous control: the logged runs do not record u_{-1}) and the
    inflated fields of every recorded DD plan, grouped by obstacle count: list of dict(x0, u, cir, elp, field, label, idx)"""
    st = g["ddp_states"]                                         # [n, 4, 3]
    dth = st[:, 1:, 2] - st[:, :-1, 2]
    d = st[:, 1:, :2] - st[:, :-1, :2]
    c, s = np.cos(st[:, :-1, 2]), np.sin(st[:, :-1, 2])
    v = (d[:, :, 0] * c + d[:, :, 1] * s) / _lipmodel.DT          # x+ = x + dt v cos(th), y+ = y + dt v sin(th)  (MPC_DD_sig_step.py:360-363)
    u = np.stack([v, dth], axis=2).reshape(len(st), 6)
    groups = []
    names = np.array([str(n) for n in g["dd_name"]])
    # DD_mexx follows the shipped MPC_DD_sig_step.py constants (median |du| 2e-9); the twenty DD_me<k> runs were made with p = 2
    # like their LIP counterparts (median |du| 2e-2 with the shipped p = 0, 1e-5 with p = 2; see run_params)
    keys = sorted({(int(a), int(b), n == "mexx") for a, b, n in zip(g["dd_n_cir"], g["dd_n_elp"], names)})
    for nc, ne, shipped in keys:
        runs = np.nonzero((g["dd_n_cir"] == nc) & (g["dd_n_elp"] == ne) & ((names == "mexx") == shipped))[0]
        remap = -np.ones(len(g["dd_name"]), dtype=int); remap[runs] = np.arange(len(runs))
        idx = np.nonzero(np.isin(g["ddp_run"], runs))[0]
        groups.append(dict(x0=st[idx, 0], u=u[idx], cir=g["dd_cir"][runs, :nc] + [0, 0, SAFE_DIS],
                           elp=g["dd_elp"][runs, :ne] + [0, 0, SAFE_DIS, SAFE_DIS, 0], field=remap[g["ddp_run"][idx]].astype(np.int32),
                           label=g["ddp_label"][idx], idx=idx, params={} if shipped else dict(w_p=2.0), resid=np.abs(d[idx, :, 0] * s[idx] - d[idx, :, 1] * c[idx]).max()))
    return groups


def summarize_dd(label, status, u, u_rec):
    lab, st = label == 2, status == 2
    both = ~lab & (status == 0)
    du = np.abs(u - u_rec).max(axis=1)
    return dict(n=len(lab), class_agree=float(np.mean(lab == st)), rec_fail=int(lab.sum()), rec_fail_ours_infeasible=int((lab & st).sum()),
                rec_fail_ours_solved=int((lab & (status == 0)).sum()), rec_ok_ours_infeasible=int((~lab & st).sum()),
                ours_other=int(np.sum((status != 0) & (status != 2))), both_feasible=int(both.sum()), du_median=float(np.median(du[both])),
                du_le_1e4=float(np.mean(du[both] <= 1e-4)), du_le_1e3=float(np.mean(du[both] <= 1e-3)), du_le_1e2=float(np.mean(du[both] <= 1e-2)))
