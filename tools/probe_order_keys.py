"""GPU probe (not a test): variants of the start-order key on the config-2 batches (device ordering off: DCBF_ORDER=0)."""
import os, sys, numpy as np, torch
os.environ["DCBF_ORDER"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
s = DcbfSolver("sig_step", device=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tot = {}
for seed in (0, 1, 2, 4):
    sc = scenarios.make_batch("sig_step", 4096, seed=seed)
    s.set_fields(sc.cir)
    d = lambda a, t: torch.as_tensor(a, dtype=t, device="cuda")
    full = dict(x0=d(sc.x0, torch.float64), goal=d(sc.goal, torch.float64), leg=d(sc.leg, torch.int32), fld=d(sc.field, torch.int32), warm=d(sc.warm, torch.float64))

    def timed(perm):
        a = {k: v[perm].contiguous() for k, v in full.items()}
        ts = []
        for _ in range(9):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.solve(a["x0"], a["goal"], a["leg"], a["warm"], field=a["fld"]); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts[2:])[3]
    pos, vel, th = full["x0"][:, :2], full["x0"][:, 2:4], full["x0"][:, 4]
    cir = d(sc.cir, torch.float64)[full["fld"].long()]
    gdir = full["goal"] - pos; gdir = gdir / torch.linalg.norm(gdir, dim=1, keepdim=True)

    def key(vv, ks, step=0.4):
        k_ = torch.full((4096,), float("inf"), device="cuda", dtype=torch.float64)
        for k in ks:
            pk = pos + vv * (step * k)
            k_ = torch.minimum(k_, (torch.linalg.norm(pk[:, None, :] - cir[:, :, :2], dim=2) - cir[:, :, 2]).min(1).values)
        return k_
    variants = {"natural": None, "vel k0..3 (current)": key(vel, range(4)), "vel k0..5": key(vel, range(6)), "vel k0..2": key(vel, range(3)),
                "vel k0 only": key(vel, [0]), "goal dir 0.6 m/s k0..3": key(0.6 * gdir, range(4)), "vel k0..3 half steps": key(vel, np.arange(0, 3.5, 0.5)),
                "min(vel, goal) k0..3": torch.minimum(key(vel, range(4)), key(0.6 * gdir, range(4)))}
    for name, k_ in variants.items():
        if k_ is None:
            t = timed(torch.arange(4096, device="cuda"))
        else:
            c = torch.clamp(torch.floor((k_ + 0.5) * 8.0), 0, 15)
            t = timed(torch.argsort(c, stable=True))
        tot.setdefault(name, []).append(t)
for name, ts in tot.items():
    print(f"{name:30s} " + "  ".join(f"{t:.3f}" for t in ts) + f"   mean {np.mean(ts):.3f} ms")
