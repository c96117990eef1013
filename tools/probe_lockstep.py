"""GPU probe (not a test): is the warp kernel bound by instruction supply?  A batch of B copies of ONE scenario makes every warp
follow the same control flow at (nearly) the same time, so the instruction-cache lines one warp misses are hits for the others;
compare the rate of interior-point iterations with a batch of B different scenarios."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for form in ("sig_step", "dd"):
    sc = scenarios.make_batch(form, B, seed=0)
    s = DcbfSolver(form, device=0)
    s.set_fields(sc.cir, sc.elp if sc.elp.shape[1] else None)
    d = lambda a, t: None if a is None else torch.as_tensor(a, dtype=t, device="cuda")
    for name, idx in (("different scenarios", np.arange(B)), ("copies of scenario 5", np.full(B, 5)), ("copies of scenario 11", np.full(B, 11))):
        x0, goal, leg, fld, warm = d(sc.x0[idx], torch.float64), d(sc.goal[idx], torch.float64), d(sc.leg[idx], torch.int32), d(sc.field[idx], torch.int32), d(sc.warm[idx], torch.float64)
        lu = d(None if sc.last_u is None else sc.last_u[idx], torch.float64)
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = s.solve(x0, goal, leg, warm, field=fld, last_u=lu); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts[1:])[len(ts[1:]) // 2]
        its = float(r.iters.sum())
        print(f"{form:8s} B={B} {name:22s} {t:8.3f} ms  iterations {its:.0f} (mean {its / B:.2f})  {its / t * 1e-3:8.1f} M iterations/s", flush=True)
