import time, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import __graft_entry__ as g
g.smoke()
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
s = DcbfSolver("sig_step", device=0)
print("fp64 peak TFLOP/s", s.fp64_peak_tflops(3))
for form, B in (("sig_step", 4096), ("sig_step", 65536), ("sig_step", 1<<20), ("modi", 65536), ("dd", 65536)):
    sc = scenarios.make_batch(form, B, seed=3)
    sol = DcbfSolver(form, device=0)
    sol.set_fields(sc.cir, sc.elp if sc.elp.shape[1] else None)
    dev = lambda a, dt: torch.as_tensor(a, dtype=dt, device="cuda")
    x0, goal, leg, field, warm = dev(sc.x0, torch.float64), dev(sc.goal, torch.float64), dev(sc.leg, torch.int32), dev(sc.field, torch.int32), dev(sc.warm, torch.float64)
    lu = None if sc.last_u is None else dev(sc.last_u, torch.float64)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        r = sol.solve(x0, goal, leg, warm, field=field, last_u=lu)
        torch.cuda.synchronize(); dt = time.time() - t0
    st = r.status.cpu().numpy(); it = r.iters.cpu().numpy()
    print(form, B, "time %.4fs  %.3e solves/s" % (dt, B / dt), "status", {int(k): int((st == k).sum()) for k in np.unique(st)},
          "iters mean %.1f max %d" % (it.mean(), it.max()), flush=True)
sc = scenarios.make_batch("sig_step", 65536, seed=5)
sol = DcbfSolver("sig_step", device=0); sol.set_fields(sc.cir)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    r = sol.rollout(50, sc.x0, sc.goal, sc.leg, field=sc.field, want_traj=False)
    torch.cuda.synchronize(); dt = time.time() - t0
tot = int(r["steps_done"].sum()); print("rollout 65536x50: %.3fs, %d solves, %.3e solves/s, iters/solve %.1f, infeasible frac %.3f" % (dt, tot, tot / dt, float(r["total_iters"].sum()) / tot, float(r["n_infeasible"].sum()) / tot))
