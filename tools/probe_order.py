"""GPU probe (not a test): how much of the small-batch step time is scheduling tail?  Times the config-2 batch in its natural order,
sorted by the true iteration count (longest first: the best a predictor could do) and sorted by cheap geometric predictors."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver

for form, B in (("sig_step", 4096), ("sig_step", 8192), ("modi", 4096), ("dd", 4096)):
    sc = scenarios.make_batch(form, B, seed=0)
    s = DcbfSolver(form, device=0)
    s.set_fields(sc.cir, sc.elp if sc.elp.shape[1] else None)
    d = lambda a, t: None if a is None else torch.as_tensor(a, dtype=t, device="cuda")
    full = dict(x0=d(sc.x0, torch.float64), goal=d(sc.goal, torch.float64), leg=d(sc.leg, torch.int32), fld=d(sc.field, torch.int32),
                warm=d(sc.warm, torch.float64), lu=d(sc.last_u, torch.float64))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(perm):
        a = {k: (None if v is None else v[perm].contiguous()) for k, v in full.items()}
        ts = []
        for _ in range(9):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = s.solve(a["x0"], a["goal"], a["leg"], a["warm"], field=a["fld"], last_u=a["lu"]); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts[2:])[3], r

    ident = torch.arange(B, device="cuda")
    t0, r = timed(ident)
    it = r.iters.double()
    print(f"{form} B={B}: natural {t0:.3f} ms  (iters mean {float(it.mean()):.2f} max {int(it.max())}; work-conserving bound {float(it.sum()) / 1776 * t0 / float(it.sum()) :.3f})")
    t1, _ = timed(torch.argsort(it, descending=True))
    print(f"   longest first (true iteration counts) {t1:.3f} ms   {t0 / t1:.3f}x")
    t1, _ = timed(torch.argsort(it, descending=False))
    print(f"   shortest first                        {t1:.3f} ms   {t0 / t1:.3f}x")
    if form != "dd":
        pos, vel = full["x0"][:, :2], full["x0"][:, 2:4]
    else:
        th = full["x0"][:, 2]
        pos, vel = full["x0"][:, :2], 0.8 * torch.stack([torch.cos(th), torch.sin(th)], 1)
    cir = torch.as_tensor(sc.cir, device="cuda")[full["fld"].long()]
    key = torch.full((B,), float("inf"), device="cuda", dtype=torch.float64)
    for k in range(4):
        pk = pos + vel * (0.4 * k)
        key = torch.minimum(key, (torch.linalg.norm(pk[:, None, :] - cir[:, :, :2], dim=2) - cir[:, :, 2]).min(dim=1).values)
    if sc.elp.shape[1]:
        elp = torch.as_tensor(sc.elp, device="cuda")[full["fld"].long()]
        for k in range(4):
            pk = pos + vel * (0.4 * k)
            key = torch.minimum(key, (torch.linalg.norm(pk[:, None, :] - elp[:, :, :2], dim=2) - torch.maximum(elp[:, :, 2], elp[:, :, 3])).min(dim=1).values)
    t2, _ = timed(torch.argsort(key))
    print(f"   smallest predicted clearance first    {t2:.3f} ms   {t0 / t2:.3f}x   (corr with iters {float(torch.corrcoef(torch.stack([-key, it]))[0, 1]):.2f})")
    for nb in (4, 8):
        qs = torch.quantile(key, torch.linspace(0, 1, nb + 1, device="cuda", dtype=torch.float64)[1:-1])
        cls = torch.bucketize(key, qs)
        t3, _ = timed(torch.argsort(cls, stable=True))
        print(f"   {nb} buckets of predicted clearance      {t3:.3f} ms   {t0 / t3:.3f}x")
