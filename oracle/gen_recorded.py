"""ORACLE (test infrastructure) -- freeze EVERY recorded re-plan of the reference's run logs as a compact fixture.

Run in the build container only (needs /root/reference):   python -m oracle.gen_recorded
Writes tests/golden/recorded_runs.npz.

The reference ships the logs of 23 LIP runs and 22 differential-drive runs under data_log/ (writer:
data_procs/logger_mpc.py:449-474, data_procs/logger_dd.py:445-466; the feasible / failed label of a plan is Ipopt's status
!= 2 / == 2, main_sim_mpc.py:118-121): 1 778 LIP plans and the DD plans, produced by the real pipeline -- MuJoCo + Digit + the
planner classes + cyipopt -- on the authors' machine.  Nothing here is an output of this repository.

What is recoverable, and how (established by probing the logs):
  * a recorded LIP plan is pos_det of gen_control_test (126 x 2, MPC_LIP_modi.py:117-122): three 42-row segments, each the
    start position followed by the LIP flow at t = 0 .. 0.40 s; a 3-parameter fit per coordinate returns (x_k, v_k, p_k) of every
    planned step to 1e-14, i.e. the START STATE x_nex[0:4] of the re-plan and its three foot placements -- exactly;
  * the plan filed at a foot change is the LAST re-plan of that step.  In the LIP_me*/mexx/dcbf runs that re-plan happened at tick
    30 of the 40-tick step (rest_t = 0.10 s: the LIP prediction of the logged CoM position, body velocity, heading and stance
    foot of tick 40 s + 30 reproduces the fitted x_nex[0:4] of all 1 697 plans to 5e-14), in LIP_mexy -- the run main_sim_mpc.py
    writes as shipped, one re-plan per tick -- at tick 39 (rest_t = 0.01 s, 81 plans, 3e-15);
  * heading of the start state: x_nex[4] = heading + rest_t / dt * hd_input_pr (MPC_LIP_modi.py:149-178).  hd_input_pr is logged
    only in LIP_mexy (turning.pkl); elsewhere it is a function of the previous step's last plan (Logger.set_stf_head,
    logger_mpc.py:264-277) and the replay in the tests chains it through dcbf_heading_input;
  * stance side: the sign of cross(heading direction, stance foot - CoM); the planner is called with -leg_ind
    (logger_mpc.py:336), +1 when the stance foot is the left one;
  * goal (10, 10), inflation 0.4 m (plot_data_cir.py:44-47,108; main_sim_mpc.py:11-19).
  * a recorded DD plan is the four states [x, y, theta] of gen_dd_control (MPC_DD_sig_step.py:83-99): start state and the
    three controls follow exactly; the previous control u_{-1} of the smoothness cost is not logged.

Stored per LIP run: the logged quantities at the re-plan ticks of every step (ticks 0, 10, 20, 30; every tick for LIP_mexy) and
per plan the fitted (x, v, p), the label, run and step.  Per DD plan: the four states, label, run.
"""
from __future__ import annotations

import glob
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mujoco_lip_mpc_simulation_b200 import data_log  # noqa: E402

REF = "/root/reference/data_log"
OUT = os.path.join(ROOT, "tests", "golden", "recorded_runs.npz")
KMAX = 6   # obstacle lists are padded to KMAX circles / ellipses (radius 0 far away is NOT used: counts are stored)


def _label(plan, fails):
    return 2 if any(np.array_equal(plan, f) for f in fails) else 0


def main():
    lip = dict(name=[], n_steps=[], ticks_per_step=[], cir=[], n_cir=[], elp=[], n_elp=[])
    tick = dict(run=[], step=[], i=[], pos=[], body_vel=[], heading=[], foot=[], turning=[])
    plan = dict(run=[], step=[], label=[], x=[], v=[], p=[], fit_res=[])
    for ri, path in enumerate(sorted(glob.glob(os.path.join(REF, "LIP_*_pos.pkl")))):
        pre = path[:-len("pos.pkl")]
        name = os.path.basename(pre)[4:-1]
        run = data_log.read_run(pre)
        full, fails = run["pred_full_end"], run["pred_fail_end"]
        every = 1 if "turning" in run else 10      # LIP_mexy: main_sim_mpc.py as shipped re-plans on every tick
        cir = np.asarray(run["cir"], dtype=np.float64).reshape(-1, 3)
        elp = np.asarray(run["ellp"], dtype=np.float64).reshape(-1, 5)
        lip["name"].append(name); lip["n_steps"].append(len(full)); lip["ticks_per_step"].append(every)
        pc = np.zeros((KMAX, 3)); pc[:len(cir)] = cir
        pe = np.zeros((KMAX, 5)); pe[:len(elp)] = elp
        lip["cir"].append(pc); lip["n_cir"].append(len(cir)); lip["elp"].append(pe); lip["n_elp"].append(len(elp))
        for s in range(len(full)):
            for i in range(0, 40, every):
                j = 40 * s + i
                tick["run"].append(ri); tick["step"].append(s); tick["i"].append(i)
                tick["pos"].append(run["pos"][j]); tick["body_vel"].append(run["body_vel"][j]); tick["heading"].append(run["heading"][j])
                tick["foot"].append(run["foot"][j]); tick["turning"].append(run["turning"][j] if "turning" in run else np.nan)
            x, v, p, res = data_log.plan_from_pos_det(full[s])
            plan["run"].append(ri); plan["step"].append(s); plan["label"].append(_label(full[s], fails))
            plan["x"].append(x); plan["v"].append(v); plan["p"].append(p); plan["fit_res"].append(res)
        assert len(full) == len(run["pred_feasi_end"]) + len(fails)
    dd = dict(name=[], cir=[], n_cir=[], elp=[], n_elp=[])
    ddp = dict(run=[], step=[], label=[], states=[])
    for ri, path in enumerate(sorted(glob.glob(os.path.join(REF, "DD_*_pos.pkl")))):
        pre = path[:-len("pos.pkl")]
        load = lambda k: pickle.load(open(pre + k + ".pkl", "rb"))   # noqa: E731
        full, fails = load("pred_full_end"), load("pred_fail_end")
        cir = np.asarray(load("cir"), dtype=np.float64).reshape(-1, 3)
        elp = np.asarray(load("ellp"), dtype=np.float64).reshape(-1, 5)
        pc = np.zeros((KMAX, 3)); pc[:len(cir)] = cir
        pe = np.zeros((KMAX, 5)); pe[:len(elp)] = elp
        dd["name"].append(os.path.basename(pre)[3:-1]); dd["cir"].append(pc); dd["n_cir"].append(len(cir)); dd["elp"].append(pe); dd["n_elp"].append(len(elp))
        for s, a in enumerate(full):
            ddp["run"].append(ri); ddp["step"].append(s); ddp["label"].append(_label(a, fails)); ddp["states"].append(np.asarray(a, dtype=np.float64))
    out = {}
    out.update({"lip_" + k: np.array(v) for k, v in lip.items()})
    out.update({"tick_" + k: np.array(v) for k, v in tick.items()})
    out.update({"plan_" + k: np.array(v) for k, v in plan.items()})
    out.update({"dd_" + k: np.array(v) for k, v in dd.items()})
    out.update({"ddp_" + k: np.array(v) for k, v in ddp.items()})
    np.savez_compressed(OUT, **out)
    print(f"LIP: {len(lip['name'])} runs, {len(plan['run'])} plans ({int(np.sum(np.array(plan['label']) == 2))} filed under pred_fail), "
          f"{len(tick['run'])} re-plan ticks, worst fit residual {max(plan['fit_res']):.1e}")
    print(f"DD : {len(dd['name'])} runs, {len(ddp['run'])} plans ({int(np.sum(np.array(ddp['label']) == 2))} filed under pred_fail)")
    print(f"{OUT}: {os.path.getsize(OUT) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
