"""Small invocation of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_probe.py
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py

Batches are tiny (the sanitizer slows the warp-cooperative kernels down by two to three orders of magnitude); every result is
still checked for sanity so that a silently skipped kernel is noticed."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver

B = int(os.environ.get("SAN_B", "96"))
for form in ("sig_step", "modi", "dd"):
    for mode in ("warp", "thread"):
        os.environ["DCBF_KERNEL"] = mode
        sc = scenarios.make_batch(form, B, seed=5, n_fields=8)
        s = DcbfSolver(form, device=0)
        s.set_fields(sc.cir, sc.elp if sc.elp.shape[1] else None)
        r = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field, last_u=sc.last_u)
        torch.cuda.synchronize()
        st = r.status.cpu().numpy()
        assert np.isin(st, (0, 1, 2, -1, -2)).all(), st
        info = s.setup_info(sc.x0, sc.goal, field=sc.field)
        ev = s.evaluate(sc.x0, sc.goal, sc.leg, r.u if form == "dd" else r.p_plan.reshape(B, 9), field=sc.field, last_u=sc.last_u)
        torch.cuda.synchronize()
        print(form, mode, "solve ok:", {int(k): int((st == k).sum()) for k in np.unique(st)}, "selected", int(info["count"].sum()), flush=True)
        if form != "dd":
            ro = s.rollout(3, sc.x0[:32], sc.goal[:32], sc.leg[:32], field=sc.field[:32])
            tk = s.tick(sc.x0[:, 0:2], sc.x0[:, 2:4], sc.x0[:, 4], np.concatenate([sc.x0[:, 0:2], np.zeros((B, 1))], axis=1), np.full(B, 0.1),
                        sc.goal, sc.leg, field=sc.field)
            torch.cuda.synchronize()
            print(form, mode, "rollout / tick ok:", int(ro["steps_done"].sum()), int((tk["plan"].status == 0).sum()), flush=True)
        del s
os.environ.pop("DCBF_KERNEL")
# batches large enough for the start order (sched_* kernels) and the size-class split (classify kernel) -- warp kernels only
for form, n in (("sig_step", 2048), ("dd", 2048)):
    sc = scenarios.make_batch(form, n, seed=6, n_fields=64)
    s = DcbfSolver(form, device=0)
    s.set_fields(sc.cir, sc.elp if sc.elp.shape[1] else None)
    r = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field, last_u=sc.last_u)
    torch.cuda.synchronize()
    print(form, n, "ordered batch ok:", int((r.status == 0).sum()), flush=True)
os.environ["DCBF_SPLIT"] = "512"
sc = scenarios.make_batch("modi", 512, seed=7, n_fields=32)
s = DcbfSolver("modi", device=0)
s.set_fields(sc.cir, sc.elp)
r = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field)
torch.cuda.synchronize()
print("modi split batch ok:", int((r.status == 0).sum()), flush=True)
# scenario generation + host-buffer entry point (page-locked, zero copy)
s = DcbfSolver("sig_step", device=0)
g = scenarios.make_batch_device(s, 512, seed=3, n_fields=64)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
sc = scenarios.make_batch("sig_step", 512, seed=8, n_fields=16)
s.set_fields_host(sc.cir)
h = s.solve_host(pin(sc.x0), pin(sc.goal), pin(sc.leg.astype(np.int32)), pin(sc.warm), field=pin(sc.field.astype(np.int32)))
print("generation / host entry ok:", int(g["attempts"].min()), int((h.status == 0).sum()), flush=True)
# host-buffer entry point on a batch the start order applies to: the classify pass stages the inputs on the device and writes the
# cold-start vector (warm = None); small batch without a start vector: cold_start_kernel
sc = scenarios.make_batch("sig_step", 2048, seed=9, n_fields=16)
s = DcbfSolver("sig_step", device=0)
s.set_fields_host(sc.cir)
ref = s.solve_host(pin(sc.x0), pin(sc.goal), pin(sc.leg.astype(np.int32)), pin(np.tile(sc.x0, (1, 3))), field=pin(sc.field.astype(np.int32)))
h = s.solve_host(pin(sc.x0), pin(sc.goal), pin(sc.leg.astype(np.int32)), None, field=pin(sc.field.astype(np.int32)))
small = s.solve(sc.x0[:100], sc.goal[:100], sc.leg[:100], None, field=sc.field[:100])
torch.cuda.synchronize()
assert np.array_equal(h.u, ref.u) and np.array_equal(small.u.cpu().numpy(), ref.u[:100])
print("staged host entry / cold-start rule ok:", int((h.status == 0).sum()), int((small.status == 0).sum()), flush=True)
