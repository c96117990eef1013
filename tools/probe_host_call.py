"""GPU probe (not a test): where the time of one synchronous host-buffer call goes (config 2, page-locked buffers):
Python wrapper (batch.solve_host) vs the bare C entry point vs device-resident event time."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver, SolveResult, _ptr
B = 4096
sc = scenarios.make_batch("sig_step", B, seed=0)
s = DcbfSolver("sig_step", device=0)
s.set_fields_host(sc.cir)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
x0, goal, leg, warm, fld = pin(sc.x0), pin(sc.goal), pin(sc.leg.astype(np.int32)), pin(sc.warm), pin(sc.field.astype(np.int32))
out = SolveResult(pin(np.empty((B, 15))), pin(np.empty((B, 3, 5))), pin(np.empty((B, 3, 3))), pin(np.empty(B, np.int32)),
                  pin(np.empty(B, np.int32)), pin(np.empty(B)), pin(np.empty(B)), pin(np.empty(B, np.uint8)))
for _ in range(5):
    s.solve_host(x0, goal, leg, warm, field=fld, out=out)
N = 200
t0 = time.perf_counter()
for _ in range(N):
    s.solve_host(x0, goal, leg, warm, field=fld, out=out)
t_wrap = (time.perf_counter() - t0) / N
args = [_ptr(a) for a in (x0, goal, leg, fld, warm, None, out.u, out.x_plan, out.p_plan, out.status, out.iters, out.obj, out.viol, out.close2goal)]
t0 = time.perf_counter()
for _ in range(N):
    s.lib.dcbf_solve_host(s._ctx, B, *args)
t_c = (time.perf_counter() - t0) / N
d = lambda a, t: torch.as_tensor(a, dtype=t, device="cuda")
a = (d(sc.x0, torch.float64), d(sc.goal, torch.float64), d(sc.leg, torch.int32), d(sc.field, torch.int32), d(sc.warm, torch.float64))
o = SolveResult(*[torch.empty(x.shape, dtype=torch.from_numpy(x).dtype, device="cuda") for x in (out.u, out.x_plan, out.p_plan, out.status, out.iters, out.obj, out.viol, out.close2goal)])
flush = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
es = []
for _ in range(20):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.solve_into(B, a[0], a[1], a[2], a[3], a[4], None, o); e1.record(); torch.cuda.synchronize()
    es.append(e0.elapsed_time(e1))
print(f"python wrapper {t_wrap * 1e6:.1f} us   bare C call {t_c * 1e6:.1f} us   device-resident (events) {np.median(es) * 1e3:.1f} us")
