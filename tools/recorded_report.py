#!/usr/bin/env python
"""Agreement of the CUDA solver with EVERY recorded cyipopt re-plan of the reference (tests/golden/recorded_runs.npz), per run.

Run on the GPU box:  python tools/recorded_report.py > profiles/r03_recorded.md      (add --oracle for the CPU oracle's table)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import recorded_replay as rr  # noqa: E402

use_oracle = "--oracle" in sys.argv
g = rr.load()
make = (lambda c, e, **kw: rr.OracleBackend(c, e, **kw)) if use_oracle else (lambda c, e, **kw: rr.CudaBackend(c, e, **kw))
out = rr.replay_all(make, g)
who = "oracle (oracle/dcbf_oracle.c)" if use_oracle else "CUDA solver (dcbf_heading_input + dcbf_tick through the C ABI)"
print(f"# Recorded cyipopt re-plans of the reference, re-solved by the {who}\n")
print("Source: /root/reference/data_log (23 LIP runs, 21 DD runs; real MuJoCo + Digit + planner + cyipopt pipeline), frozen by")
print("oracle/gen_recorded.py.  Replay: tests/recorded_replay.py (open loop on the logged robot states; the heading input and the warm")
print("start are chained from re-plan to re-plan as Logger does).  `label` = the reference's own filing: pred_fail <=> Ipopt status 2.\n")
print("## LIP (MPC_LIP_modi), all 1 778 plans\n")
hdr = ("| run | constants | plans | start state max err | same verdict (as shipped) | filed infeasible | ... status 2 here | ... solved here | "
       "filed feasible, status 2 here | other exits here (-1/-2/1) | both feasible | median abs dp0 [m] | dp0 <= 1e-4 | <= 1e-3 | <= 1e-2 |")
print(hdr)
print("|" + "---|" * 15)


def row(name, const, s):
    print(f"| {name} | {const} | {s['n']} | {s['start_state_err']:.1e} | {s['class_agree']:.4f} | {s['rec_fail']} | {s['rec_fail_ours_infeasible']} | "
          f"{s['rec_fail_ours_solved']} | {s['rec_ok_ours_infeasible']} | {s['ours_other']} | {s['both_feasible']} | {s['dp0_median']:.1e} | "
          f"{s['dp0_le_1e4']:.3f} | {s['dp0_le_1e3']:.3f} | {s['dp0_le_1e2']:.3f} |")


names = [str(n) for n in g["lip_name"]]
for r, n in enumerate(names):
    m = out["run"] == r
    const = ", ".join(f"{k}={v}" for k, v in rr.run_params(n).items()) or "as shipped"
    row("LIP_" + n, const, rr.summarize_lip({k: v[m] for k, v in out.items()}))
row("**all**", "", rr.summarize_lip(out))
m = ~np.isnan(out["hd_pr_logged"])
d = np.abs(out["hd_pr"][m] - out["hd_pr_logged"][m])
print(f"\nHeading input of the prediction (Logger.tube_func + avg_hd), chained over the 3 240 re-plans of LIP_mexy, against the logged "
      f"`turning.pkl` at the 81 filed plans: median abs error {np.median(d):.1e}, {np.mean(d <= 1e-6):.3f} within 1e-6, max {d.max():.1e} "
      f"(steps after a re-plan that took a different local optimum).")
lab, st = out["label"] == 2, out["status"] == 2
both = ~lab & (out["status"] == 0)
dp0 = np.abs(out["p"][:, 0, :2] - out["p_rec"][:, 0]).max(axis=1)
bad = both & (dp0 > 1e-4)
chained = sum(1 for i in np.nonzero(bad)[0] if i > 0 and out["run"][i - 1] == out["run"][i] and (dp0[i - 1] > 1e-4 or out["status"][i - 1] != 0))
print(f"\nBuckets, as shipped -> converged: of the {int((~lab).sum())} plans the reference filed as feasible, {int((~lab & st).sum())} are "
      f"problems this solver proves locally infeasible (status 2: restoration converged to a stationary point of the violation above 1e-4) -- "
      f"the reference's iteration-capped exits (max_iter 30, L-BFGS; -1 / 1 / -2 are filed as feasible, main_sim_mpc.py:118-121).  Of the "
      f"{int(both.sum())} plans both call feasible, {int(bad.sum())} differ by more than 1e-4 m in the first foot placement; {chained} of those "
      f"follow a step whose plan already differed (the heading input of a step is a function of the previous step's plan, so one different "
      f"local optimum or capped exit displaces the start heading of the following problems).")

if not use_oracle:
    from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
    print("\n## Differential drive (MPC_DD_sig_step), all 1 377 plans\n")
    print("| runs | constants | plans | same verdict (as shipped) | filed infeasible | ... status 2 here | filed feasible, status 2 here | other exits here | "
          "both feasible | median abs du | du <= 1e-4 | <= 1e-3 | <= 1e-2 |")
    print("|" + "---|" * 13)
    L, S, U, UR = [], [], [], []
    for grp in rr.dd_inputs(g):
        B = len(grp["x0"])
        s = DcbfSolver("dd", device=0, max_iter=300, **grp["params"])
        s.set_fields(grp["cir"], grp["elp"])
        r = s.solve(grp["x0"], np.tile(rr.GOAL, (B, 1)), None, grp["u"], field=grp["field"], last_u=grp["u"][:, :2].copy())
        a = (grp["label"], r.status.cpu().numpy(), r.u.cpu().numpy(), grp["u"])
        for acc, v in zip((L, S, U, UR), a):
            acc.append(v)
        sm = rr.summarize_dd(*a)
        const = ", ".join(f"{k}={v}" for k, v in grp["params"].items()) or "as shipped"
        print(f"| {len(grp['cir'])} ({grp['cir'].shape[1]} circles + {grp['elp'].shape[1]} ellipses) | {const} | {sm['n']} | {sm['class_agree']:.4f} | {sm['rec_fail']} | "
              f"{sm['rec_fail_ours_infeasible']} | {sm['rec_ok_ours_infeasible']} | {sm['ours_other']} | {sm['both_feasible']} | {sm['du_median']:.1e} | "
              f"{sm['du_le_1e4']:.3f} | {sm['du_le_1e3']:.3f} | {sm['du_le_1e2']:.3f} |")
    sm = rr.summarize_dd(*(np.concatenate(a) for a in (L, S, U, UR)))
    print(f"| **all** | | {sm['n']} | {sm['class_agree']:.4f} | {sm['rec_fail']} | {sm['rec_fail_ours_infeasible']} | {sm['rec_ok_ours_infeasible']} | {sm['ours_other']} | "
          f"{sm['both_feasible']} | {sm['du_median']:.1e} | {sm['du_le_1e4']:.3f} | {sm['du_le_1e3']:.3f} | {sm['du_le_1e2']:.3f} |")
    print("\nThe previous control u_{-1} of the smoothness cost (MPC_DD_sig_step.py:351-369) is not logged; the recorded plan's own first control "
          "stands in for it, which bounds the agreement of the controls at the 1e-3 level.  Every plan the reference filed as infeasible is "
          "status 2 here; the plans it filed as feasible for problems that are infeasible are its iteration-capped exits (max_iter 40).")
