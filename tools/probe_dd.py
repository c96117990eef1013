"""GPU probe (not a test): DD formulation, per-thread kernel vs warp kernel."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver

def run(B, mode):
    os.environ["DCBF_KERNEL"] = mode
    sc = scenarios.make_batch("dd", B, seed=2)
    s = DcbfSolver("dd", device=0)
    s.set_fields(sc.cir, sc.elp)
    d = lambda a, t: torch.as_tensor(a, dtype=t, device="cuda")
    x0, goal, fld, warm, lu = d(sc.x0, torch.float64), d(sc.goal, torch.float64), d(sc.field, torch.int32), d(sc.warm, torch.float64), d(sc.last_u, torch.float64)
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = s.solve(x0, goal, None, warm, field=fld, last_u=lu); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    st = r.status.cpu().numpy(); it = r.iters.cpu().numpy()
    print(f"dd B={B:8d} {mode:6s} {best:9.3f} ms  {B/best*1e3:12.0f} solves/s  iters mean {it.mean():.2f} max {it.max()}  status "
          + str({int(k): int((st == k).sum()) for k in np.unique(st)}), flush=True)
    return r

for B in (4096, 65536):
    rt = run(B, "thread"); rw = run(B, "warp")
    both = (rt.status == 0) & (rw.status == 0)
    du = (rt.u - rw.u).abs().max(dim=1).values
    dx = (rt.x_plan - rw.x_plan).abs().reshape(B, -1).max(dim=1).values
    print(f"   class agree {float(((rt.status == 2) == (rw.status == 2)).float().mean()):.5f}  status equal {float((rt.status == rw.status).float().mean()):.5f}  both ok {int(both.sum())}  |du|<=1e-4 {float((du[both] <= 1e-4).float().mean()):.5f}  max du {float(du[both].max()):.2e} max dx {float(dx[both].max()):.2e}  iters equal {float((rt.iters == rw.iters).float().mean()):.4f}", flush=True)
