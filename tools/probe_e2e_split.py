"""GPU probe (not a test): where the gap between the device-timed step and the synchronous host-buffer call comes from.  One
4096-scenario sig_step batch through dcbf_solve with inputs / outputs in device memory or in page-locked host memory (mapped: the
kernels read / write it over PCIe), wall clock per call incl. the final synchronisation, L2 flushed before each call."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver, SolveResult
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sc = scenarios.make_batch("sig_step", B, seed=seed)
s = DcbfSolver("sig_step", device=0); s.set_fields(sc.cir)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def mk(where):
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt)
    ins = [t(sc.x0, torch.float64), t(sc.goal, torch.float64), t(sc.leg, torch.int32), t(sc.field, torch.int32), t(sc.warm, torch.float64)]
    return [a.cuda() if where == "dev" else a.pin_memory() for a in ins]
def mko(where):
    e = lambda shape, dt: torch.empty(shape, dtype=dt, device="cuda") if where == "dev" else torch.empty(shape, dtype=dt).pin_memory()
    return SolveResult(e((B, 15), torch.float64), e((B, 3, 5), torch.float64), e((B, 3, 3), torch.float64), e((B,), torch.int32), e((B,), torch.int32),
                       e((B,), torch.float64), e((B,), torch.float64), e((B,), torch.uint8))
for wi, wo, cold in (("dev", "dev", False), ("dev", "host", False), ("host", "dev", False), ("host", "host", False), ("dev", "dev", True), ("host", "dev", True), ("host", "host", True)):
    if True:
        a, o = mk(wi), mko(wo)
        ts, es = [], []
        for rep in range(25):
            flush.zero_(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(); s.solve_into(B, a[0], a[1], a[2], a[3], None if cold else a[4], None, o); e1.record()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0); es.append(e0.elapsed_time(e1))
        ts, es = np.array(ts[5:]) * 1e6, np.array(es[5:]) * 1e3
        print(f"inputs {wi:4s} outputs {wo:4s} start vector {'on device' if cold else 'passed   '}: wall {np.median(ts):7.1f} us (min {ts.min():7.1f})   events {np.median(es):7.1f} us (min {es.min():7.1f})", flush=True)
