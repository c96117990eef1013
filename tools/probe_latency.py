"""GPU probe (not a test): latency of single solves (B = 1) through the host-buffer entry point, as bench.py measures p50_solve_us."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
sc = scenarios.make_batch("sig_step", 4096, seed=0)
s = DcbfSolver("sig_step", device=0); s.set_fields(sc.cir)
one = s.solve_host(sc.x0[:1], sc.goal[:1], sc.leg[:1], sc.warm[:1], field=sc.field[:1])
for rep in range(3):
    lat, its = [], []
    for j in range(200):
        t0 = time.perf_counter()
        s.solve_host(sc.x0[j:j + 1], sc.goal[j:j + 1], sc.leg[j:j + 1], sc.warm[j:j + 1], field=sc.field[j:j + 1], out=one)
        lat.append((time.perf_counter() - t0) * 1e6); its.append(int(one.iters[0]))
    lat, its = np.array(lat), np.array(its)
    print(f"p50 {np.median(lat):.1f} us  p95 {np.percentile(lat, 95):.1f}  mean {lat.mean():.1f}  us per iteration (fit) {np.polyfit(its, lat, 1)[0]:.2f}  fixed {np.polyfit(its, lat, 1)[1]:.1f}  mean iters {its.mean():.2f}", flush=True)
