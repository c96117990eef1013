"""GPU probe (not a test): step time of the config-2 batch for the seeds the ranks of a multi-GPU run use, with the slowest problems."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
s = DcbfSolver("sig_step", device=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for seed in range(8):
    sc = scenarios.make_batch("sig_step", 4096, seed=seed)
    s.set_fields(sc.cir)
    d = lambda a, t: torch.as_tensor(a, dtype=t, device="cuda")
    a = (d(sc.x0, torch.float64), d(sc.goal, torch.float64), d(sc.leg, torch.int32), d(sc.warm, torch.float64), d(sc.field, torch.int32))
    ts = []
    for _ in range(9):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = s.solve(a[0], a[1], a[2], a[3], field=a[4]); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    it = r.iters.cpu().numpy()
    print(f"seed {seed}: {sorted(ts[2:])[3]:.3f} ms  iters mean {it.mean():.2f}  top5 {sorted(it)[-5:]}  sum/1776 {it.sum() / 1776:.1f}")
