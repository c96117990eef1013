"""GPU probe (not a test): fixed overhead of the host-buffer entry point (wall clock per call at several batch sizes, page-locked buffers)."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver, SolveResult
s = DcbfSolver("sig_step", device=0)
sc = scenarios.make_batch("sig_step", 4096, seed=0)
s.set_fields(sc.cir)
pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory().numpy()
for B in (256, 1024, 4096):
    x0, goal, leg, warm, fld = pin(sc.x0[:B]), pin(sc.goal[:B]), pin(sc.leg[:B]), pin(sc.warm[:B]), pin(sc.field[:B])
    pe = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    out = SolveResult(pe((B, 15), torch.float64), pe((B, 3, 5), torch.float64), pe((B, 3, 3), torch.float64), pe((B,), torch.int32), pe((B,), torch.int32),
                      pe((B,), torch.float64), pe((B,), torch.float64), pe((B,), torch.uint8))
    for zc in ("1", "0"):
        os.environ["DCBF_ZEROCOPY"] = zc
        s2 = DcbfSolver("sig_step", device=0); s2.set_fields(sc.cir)
        for _ in range(5): s2.solve_host(x0, goal, leg, warm, field=fld, out=out)
        ts = []
        for _ in range(30):
            t0 = time.perf_counter(); s2.solve_host(x0, goal, leg, warm, field=fld, out=out); ts.append(time.perf_counter() - t0)
        # device-only time of the same batch
        d = lambda a, t: torch.as_tensor(a, dtype=t, device="cuda")
        a = (d(x0, torch.float64), d(goal, torch.float64), d(leg, torch.int32), d(warm, torch.float64), d(fld, torch.int32))
        es = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s2.solve(a[0], a[1], a[2], a[3], field=a[4]); e1.record(); torch.cuda.synchronize(); es.append(e0.elapsed_time(e1))
        print(f"B={B:5d} zero_copy={zc}: host call {np.median(ts) * 1e6:7.1f} us (min {min(ts) * 1e6:7.1f})   device-resident {np.median(es) * 1e3:7.1f} us")
