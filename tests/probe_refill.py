"""GPU probe (not a test): per-thread kernels with and without lane refill."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
os.environ["DCBF_KERNEL"] = "thread"
for form, B in (("sig_step", 65536), ("sig_step", 1 << 20), ("modi", 65536), ("dd", 65536)):
    sc = scenarios.make_batch(form, B, seed=1)
    d = lambda a, t: torch.as_tensor(a, dtype=t, device="cuda")
    x0, goal, leg, fld, warm = d(sc.x0, torch.float64), d(sc.goal, torch.float64), d(sc.leg, torch.int32), d(sc.field, torch.int32), d(sc.warm, torch.float64)
    lu = None if sc.last_u is None else d(sc.last_u, torch.float64)
    res = {}
    for mode in ("-1", "0"):
        os.environ["DCBF_REFILL_MIN_BATCH"] = mode
        s = DcbfSolver(form, device=0)
        s.set_fields(sc.cir, sc.elp if sc.elp.shape[1] else None)
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = s.solve(x0, goal, leg, warm, field=fld, last_u=lu); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[mode] = r
        print(f"{form:8s} B={B:8d} refill={'off' if mode == '-1' else 'on ':3s} {best:9.3f} ms {B / best * 1e3:12.0f} solves/s", flush=True)
    print("   identical:", bool(torch.equal(res["-1"].u, res["0"].u) and torch.equal(res["-1"].status, res["0"].status)))
