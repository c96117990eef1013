"""GPU probe (not a test): time both kernel families on the bench shapes."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver

def run(form, B, mode):
    os.environ["DCBF_KERNEL"] = mode
    sc = scenarios.make_batch(form, B, seed=0 if form == "sig_step" else 1)
    s = DcbfSolver(form, device=0)
    s.set_fields(sc.cir, sc.elp if sc.elp.shape[1] else None)
    d = lambda a, t: torch.as_tensor(a, dtype=t, device="cuda")
    x0, goal, leg, fld, warm = d(sc.x0, torch.float64), d(sc.goal, torch.float64), d(sc.leg, torch.int32), d(sc.field, torch.int32), d(sc.warm, torch.float64)
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = s.solve(x0, goal, leg, warm, field=fld); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    st = r.status.cpu().numpy(); it = r.iters.cpu().numpy()
    print(f"{form:8s} B={B:8d} {mode:6s} {best:9.3f} ms  {B/best*1e3:12.0f} solves/s  iters mean {it.mean():.1f} max {it.max()}  status "
          + str({int(k): int((st == k).sum()) for k in np.unique(st)}), flush=True)
    return r

import sys
SHAPES = (("sig_step", 1), ("sig_step", 4096), ("sig_step", 65536), ("modi", 65536)) if len(sys.argv) > 1 else (("sig_step", 1), ("sig_step", 4096), ("sig_step", 65536), ("sig_step", 1 << 20), ("modi", 65536))
for form, B in SHAPES:
    rt = run(form, B, "thread")
    rw = run(form, B, "warp")
    both = (rt.status == 0) & (rw.status == 0)
    dp = (rt.p_plan - rw.p_plan).abs().reshape(B, -1).max(dim=1).values
    print(f"   class agree {float(((rt.status == 2) == (rw.status == 2)).float().mean()):.5f}  both ok {int(both.sum())}  |dp|<=1e-4 {float((dp[both] <= 1e-4).float().mean()):.5f}  max dp {float(dp[both].max()):.2e}", flush=True)
