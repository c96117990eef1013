"""GPU probe (not a test): what the start order is worth at 65 536 scenarios -- the built-in order (predicted clearance, hardest
first), the natural order, and the order by the TRUE iteration count (the most any predictor could give)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, %r)
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
form, B, mode = sys.argv[1], int(sys.argv[2]), sys.argv[3]
sc = scenarios.make_batch(form, B, seed={"sig_step": 1, "modi": 1, "dd": 2}[form])
s = DcbfSolver(form, device=0)
s.set_fields(sc.cir, sc.elp if sc.elp.shape[1] else None)
d = lambda a, t: None if a is None else torch.as_tensor(a, dtype=t, device="cuda")
full = dict(x0=d(sc.x0, torch.float64), goal=d(sc.goal, torch.float64), leg=d(sc.leg, torch.int32), fld=d(sc.field, torch.int32),
            warm=d(sc.warm, torch.float64), lu=d(sc.last_u, torch.float64))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(perm):
    a = {k: (None if v is None else v[perm].contiguous()) for k, v in full.items()}
    ts = []
    for _ in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = s.solve(a["x0"], a["goal"], a["leg"], a["warm"], field=a["fld"], last_u=a["lu"]); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts[1:])[2], r
t0, r = timed(torch.arange(B, device="cuda"))
it = r.iters.double()
line = f"{form:8s} B={B} {mode:22s} {t0:8.3f} ms   iters mean {float(it.mean()):.2f} max {int(it.max())}"
if mode == "natural order":
    t1, _ = timed(torch.argsort(it, descending=True))
    line += f"    | true longest first {t1:8.3f} ms"
print(line, flush=True)
''' % ROOT
for form in ("sig_step", "dd", "modi"):
    for mode, env in (("built-in order", {}), ("natural order", {"DCBF_ORDER": "0"})):
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, "-c", CHILD, form, "65536", mode], env=e)
