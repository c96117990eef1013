#!/usr/bin/env python
"""Agreement of the CUDA solver with the oracle over many independent batches (seeds), aggregated per formulation:
status class, plans within 1e-4 among jointly converged solves, and how many of the differing plans are proven distinct local optima
(both feasible to 1e-6, objectives differ by more than 1e-6 relative).  Run on the GPU box:
    python tools/agreement_seeds.py > profiles/rNN_agreement_seeds.md"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mujoco_lip_mpc_simulation_b200 import scenarios  # noqa: E402
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver  # noqa: E402
from oracle import c_oracle  # noqa: E402

RUNS = [("sig_step", 4096, range(100, 132)), ("modi", 4096, range(200, 216)), ("dd", 4096, range(300, 316))]
print("# Agreement with the oracle over independent batches\n")
print("| formulation | batches x scenarios | status class equal | both converged | plan within 1e-4 | differing plans | ... proven distinct local optima | "
      "plan within 1e-4 once those are set aside | worst batch: class / plans |")
print("|---|---|---|---|---|---|---|---|---|")
for form, B, seeds in RUNS:
    s = DcbfSolver(form, device=0)
    n = same = nboth = within = differ = distinct = 0
    worst_c, worst_p = 1.0, 1.0
    for seed in seeds:
        sc = scenarios.make_batch(form, B, seed=seed)
        elp = sc.elp if sc.elp.shape[1] else None
        s.set_fields(sc.cir, elp)
        r = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field, last_u=sc.last_u)
        torch.cuda.synchronize()
        o = c_oracle.solve_batch(c_oracle.params(form, max_iter=200), sc.x0, sc.goal, sc.leg, sc.cir, elp, sc.warm, field=sc.field,
                                 last_u=sc.last_u, threads=os.cpu_count() or 4)
        st, so = r.status.cpu().numpy(), o["status"]
        both = (st == 0) & (so == 0)
        dp = (np.abs(r.u.cpu().numpy() - o["u"]).max(axis=1) if form == "dd"
              else np.abs(r.p_plan.cpu().numpy() - o["p_plan"]).reshape(B, -1).max(axis=1))
        rel = np.abs(r.obj.cpu().numpy() - o["f"]) / np.maximum(1.0, np.abs(o["f"]))
        dist = both & (dp > 1e-4) & (rel > 1e-6) & (r.viol.cpu().numpy() <= 1e-6) & (o["viol"] <= 1e-6)
        c = np.mean((st == 2) == (so == 2)); p = np.mean(dp[both] <= 1e-4)
        worst_c, worst_p = min(worst_c, c), min(worst_p, p)
        n += B; same += int(((st == 2) == (so == 2)).sum()); nboth += int(both.sum()); within += int((dp[both] <= 1e-4).sum())
        differ += int((dp[both] > 1e-4).sum()); distinct += int(dist.sum())
    print(f"| {form} | {len(seeds)} x {B} | {same / n:.5f} | {nboth} | {within / nboth:.5f} | {differ} | {distinct} | "
          f"{(within) / (nboth - distinct):.5f} | {worst_c:.5f} / {worst_p:.5f} |", flush=True)
