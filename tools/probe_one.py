"""GPU probe (not a test): one configuration, for ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
form, B, mode = sys.argv[1], int(sys.argv[2]), sys.argv[3]
os.environ["DCBF_KERNEL"] = mode
sc = scenarios.make_batch(form, B, seed=0)
s = DcbfSolver(form, device=0)
s.set_fields(sc.cir, sc.elp if sc.elp.shape[1] else None)
for _ in range(3):
    r = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field, last_u=sc.last_u)
torch.cuda.synchronize()
print("ok", float(r.iters.float().mean()))
