"""GPU probe (not a test): time alternative builds of the same sources (build/variants/lib_*.so) on the bench shapes.
Each variant runs in its own process because the library path is fixed at import (DCBF_LIB)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, %r)
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
shapes = [s.split(":") for s in sys.argv[1].split(",")]
for form, B, mode in shapes:
    B = int(B); os.environ["DCBF_KERNEL"] = mode
    sc = scenarios.make_batch(form, B, seed={"sig_step": 0, "modi": 1, "dd": 2}[form])
    s = DcbfSolver(form, device=0)
    s.set_fields(sc.cir, sc.elp if sc.elp.shape[1] else None)
    d = lambda a, t: None if a is None else torch.as_tensor(a, dtype=t, device="cuda")
    x0, goal, leg, fld, warm, lu = d(sc.x0, torch.float64), d(sc.goal, torch.float64), d(sc.leg, torch.int32), d(sc.field, torch.int32), d(sc.warm, torch.float64), d(sc.last_u, torch.float64)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = s.solve(x0, goal, leg, warm, field=fld, last_u=lu); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    st = r.status.cpu().numpy(); it = r.iters.cpu().numpy()
    print(f"  {form:8s} B={B:8d} {mode:6s} med {ts[len(ts)//2]:9.3f} ms min {ts[0]:9.3f}  {B/ts[len(ts)//2]*1e3:12.0f} solves/s  iters mean {it.mean():.2f} max {it.max()}  status "
          + str({int(k): int((st == k).sum()) for k in np.unique(st)}), flush=True)
''' % ROOT
shapes = sys.argv[1] if len(sys.argv) > 1 else "sig_step:1:warp,sig_step:4096:warp,sig_step:65536:warp,modi:4096:warp"
libs = sys.argv[2:] or ["default"] + sorted(f for f in os.listdir(os.path.join(ROOT, "build", "variants")) if f.endswith(".so"))
for lib in libs:
    env = dict(os.environ)
    if lib != "default":
        env["DCBF_LIB"] = os.path.join(ROOT, "build", "variants", lib)
    print(lib, flush=True)
    subprocess.run([sys.executable, "-c", CHILD, shapes], env=env)
