"""GPU probe (not a test): size-class split of the obstacle-selecting formulation against the unsplit launch."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
sc = scenarios.make_batch("modi", B, seed=41)
def run(split):
    os.environ["DCBF_SPLIT"] = str(split)
    s = DcbfSolver("modi", device=0)
    s.set_fields(sc.cir, sc.elp)
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return r, min(ts)
ref, t0 = run(0)
res, t1 = run(B)
st0, st1 = ref.status.cpu().numpy(), res.status.cpu().numpy()
it0, it1 = ref.iters.cpu().numpy(), res.iters.cpu().numpy()
dp = (res.p_plan - ref.p_plan).abs().reshape(B, -1).max(dim=1).values.cpu().numpy()
both = (st0 == 0) & (st1 == 0)
print(f"B={B} unsplit {t0:.3f} ms ({B/t0*1e3:.0f}/s)  split {t1:.3f} ms ({B/t1*1e3:.0f}/s)")
print("status equal", (st0 == st1).mean(), "iters equal", (it0 == it1).mean(), "mean iters", it0.mean(), it1.mean())
print("plans (both converged): max", dp[both].max(), "frac <= 1e-9", (dp[both] <= 1e-9).mean(), "frac <= 1e-4", (dp[both] <= 1e-4).mean())
print("status hist", {int(k): int((st1 == k).sum()) for k in np.unique(st1)}, {int(k): int((st0 == k).sum()) for k in np.unique(st0)})
