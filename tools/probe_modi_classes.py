"""GPU probe (not a test): throughput of the obstacle-selecting formulation (modi) by selected-obstacle count.
Scenarios of a 131072 batch are grouped by the number of obstacles select_obs keeps; each group is timed alone."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver

sc = scenarios.make_batch("modi", 131072, seed=1)
F = sc.field
px, py = sc.x0[:, 0], sc.x0[:, 1]
c, e = sc.cir[F], sc.elp[F]
dc = (px[:, None] - c[:, :, 0]) ** 2 + (py[:, None] - c[:, :, 1]) ** 2 - c[:, :, 2] ** 2
rm = np.maximum(e[:, :, 2], e[:, :, 3])
de = (px[:, None] - e[:, :, 0]) ** 2 + (py[:, None] - e[:, :, 1]) ** 2 - rm ** 2
ks = (dc <= 16).sum(1) + (de <= 16).sum(1)
s = DcbfSolver("modi", device=0)
s.set_fields(sc.cir, sc.elp)
d = lambda a, t: torch.as_tensor(a, dtype=t, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
groups = [("<=4", ks <= 4), ("==5", ks == 5), ("==6", ks == 6), (">=7", ks >= 7), ("all", ks >= 0)]
for name, m in groups:
    idx = np.nonzero(m)[0][:32768]
    B = len(idx)
    x0, goal, leg, fld, warm = d(sc.x0[idx], torch.float64), d(sc.goal[idx], torch.float64), d(sc.leg[idx], torch.int32), d(sc.field[idx], torch.int32), d(sc.warm[idx], torch.float64)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = s.solve(x0, goal, leg, warm, field=fld); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts[1:])[len(ts[1:]) // 2]
    it = r.iters.cpu().numpy(); st = r.status.cpu().numpy()
    print(f"Ks {name:4s} B={B:6d} {t:8.3f} ms {B / t * 1e3:12.0f} solves/s iters mean {it.mean():.2f} infeasible {np.mean(st == 2):.3f}", flush=True)
