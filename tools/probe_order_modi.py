"""GPU probe (not a test): which start order helps the obstacle-selecting formulation (config 3)?"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver

for B in (4096, 8192):
    sc = scenarios.make_batch("modi", B, seed=1)
    s = DcbfSolver("modi", device=0)
    s.set_fields(sc.cir, sc.elp)
    d = lambda a, t: torch.as_tensor(a, dtype=t, device="cuda")
    full = dict(x0=d(sc.x0, torch.float64), goal=d(sc.goal, torch.float64), leg=d(sc.leg, torch.int32), fld=d(sc.field, torch.int32), warm=d(sc.warm, torch.float64))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(perm):
        a = {k: v[perm].contiguous() for k, v in full.items()}
        ts = []
        for _ in range(9):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = s.solve(a["x0"], a["goal"], a["leg"], a["warm"], field=a["fld"]); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts[2:])[3], r

    t0, r = timed(torch.arange(B, device="cuda"))
    it, st = r.iters.double(), r.status
    pos, vel = full["x0"][:, :2], full["x0"][:, 2:4]
    cir = d(sc.cir, torch.float64)[full["fld"].long()]
    elp = d(sc.elp, torch.float64)[full["fld"].long()]
    cen = torch.cat([cir[:, :, :2], elp[:, :, :2]], 1)
    rad = torch.cat([cir[:, :, 2], torch.maximum(elp[:, :, 2], elp[:, :, 3])], 1)
    dist0 = torch.linalg.norm(pos[:, None, :] - cen, dim=2)
    nsel = ((dist0 ** 2 - rad ** 2) <= 16.0).sum(1).double()
    key = torch.full((B,), float("inf"), device="cuda", dtype=torch.float64)
    for k in range(4):
        pk = pos + vel * (0.4 * k)
        key = torch.minimum(key, (torch.linalg.norm(pk[:, None, :] - cen, dim=2) - rad).min(1).values)
    print(f"modi B={B}: natural {t0:.3f} ms; iters mean {float(it.mean()):.2f}; infeasible {float((st == 2).double().mean()):.2f}; "
          f"mean iters feasible {float(it[st == 0].mean()):.1f} infeasible {float(it[st == 2].mean()):.1f}; corr(iters, nsel) {float(torch.corrcoef(torch.stack([nsel, it]))[0, 1]):.2f}")
    for name, k_ in (("true iterations desc", -it), ("iterations x rows desc", -(it * (5 + nsel))), ("infeasible first", -(st == 2).double()), ("nsel desc", -nsel),
                     ("clearance asc", key), ("|clearance| asc", key.abs()), ("nsel desc then clearance", -nsel * 10 + key.clamp(-2, 2)),
                     ("nsel asc", nsel)):
        t1, _ = timed(torch.argsort(k_, stable=True))
        print(f"   {name:28s} {t1:.3f} ms   {t0 / t1:.3f}x")
