"""GPU probe (not a test): warm-started control tick of 65 536 modi scenarios (the bench sweep's tick_warm_modi_65536) for several
first barrier parameters `mu_warm`: ticks/s and the iteration distribution."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
B = 65536
s3 = scenarios.make_batch("modi", B, seed=1)
dev = torch.device("cuda", 0)
t = lambda a, dt: torch.as_tensor(a, dtype=dt, device=dev)
for mw in [float(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1e-4,3e-5").split(",")]:
    sv = DcbfSolver("modi", device=0, mu_warm=mw)
    sv.set_fields(s3.cir, s3.elp)
    targs = [t(s3.x0[:, 0:2], torch.float64), t(s3.x0[:, 2:4], torch.float64), t(s3.x0[:, 4], torch.float64),
             t(np.concatenate([s3.x0[:, 0:2], np.zeros((B, 1))], axis=1), torch.float64), t(np.full(B, 0.1), torch.float64)]
    goal, leg, fld = t(s3.goal, torch.float64), t(s3.leg, torch.int32), t(s3.field, torch.int32)
    tk = sv.tick(targs[0], targs[1], targs[2], targs[3], targs[4], goal, leg, field=fld)
    gen = torch.Generator(device=dev); gen.manual_seed(5)
    pos2 = targs[0] + 0.005 * torch.randn(targs[0].shape, generator=gen, device=dev, dtype=torch.float64)
    vel2 = targs[1] + 0.02 * torch.randn(targs[1].shape, generator=gen, device=dev, dtype=torch.float64)
    prev, md0 = tk["plan"].x_plan.reshape(B, 15).clone(), torch.zeros(B, dtype=torch.uint8, device=dev)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tw = sv.tick(pos2, vel2, targs[2], targs[3], targs[4], goal, leg, prev_plan=prev, mode=md0, field=fld)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    it = tw["plan"].iters.cpu().numpy(); st = tw["plan"].status.cpu().numpy()
    print(f"mu_warm {mw:g}: {B / best * 1e3 / 1e6:.2f} M ticks/s  {best:.3f} ms  iters mean {it.mean():.2f} p99 {np.percentile(it, 99):.0f} max {it.max()}  "
          f"n>40 {(it > 40).sum()} n>100 {(it > 100).sum()}  status {dict(zip(*np.unique(st, return_counts=True)))}", flush=True)
