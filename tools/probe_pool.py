"""GPU probe (not a test): alternative builds (build/variants/lib_*.so) on the bench's pool of eight 4096-scenario sig_step batches
(device-timed per step, L2 flushed in between, like bench.py), one 65 536-scenario batch and the dd / modi shapes; prints a checksum of
the results so that builds meant to be bit-identical can be told apart from builds that are not."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, hashlib, numpy as np, torch
sys.path.insert(0, %r)
from mujoco_lip_mpc_simulation_b200 import scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
d = lambda a, t: None if a is None else torch.as_tensor(a, dtype=t, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def prep(form, B, seed):
    sc = scenarios.make_batch(form, B, seed=seed)
    return sc, (d(sc.x0, torch.float64), d(sc.goal, torch.float64), d(sc.leg, torch.int32), d(sc.warm, torch.float64), d(sc.field, torch.int32), d(sc.last_u, torch.float64))
def run(s, a):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = s.solve(a[0], a[1], a[2], a[3], field=a[4], last_u=a[5]); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), r
def digest(r):
    h = hashlib.sha1()
    for t in (r.u, r.status, r.iters): h.update(t.cpu().numpy().tobytes())
    return h.hexdigest()[:10]
# pool of eight batches sharing one solver (fields of seed 0..7 differ: one solver each)
pool = []
for seed in range(8):
    sc, a = prep("sig_step", 4096, seed)
    s = DcbfSolver("sig_step", device=0); s.set_fields(sc.cir, None)
    pool.append((s, a))
for s, a in pool: run(s, a)
ts = np.zeros((5, 8)); dg = []
for rep in range(5):
    for i, (s, a) in enumerate(pool):
        ts[rep, i], r = run(s, a)
        if rep == 0: dg.append(digest(r))
med = np.median(ts, axis=0)
print("  pool 4096: mean of per-batch medians %%.4f ms (%%.3f M solves/s)  per batch %%s  digest %%s" %% (med.mean(), 4096 / med.mean() / 1e3, np.round(med, 3).tolist(), hashlib.sha1("".join(dg).encode()).hexdigest()[:10]), flush=True)
for form, B, seed in (("sig_step", 65536, 0), ("dd", 4096, 2), ("dd", 65536, 2), ("modi", 65536, 1)):
    sc, a = prep(form, B, seed)
    s = DcbfSolver(form, device=0); s.set_fields(sc.cir, sc.elp if sc.elp.shape[1] else None)
    run(s, a); run(s, a)
    t = sorted(run(s, a)[0] for _ in range(5)); r = run(s, a)[1]
    print("  %%-8s B=%%6d med %%.3f ms  %%.3f M solves/s  iters %%.3f  digest %%s" %% (form, B, t[2], B / t[2] / 1e3, r.iters.float().mean().item(), digest(r)), flush=True)
''' % ROOT
libs = sys.argv[1:] or ["default"] + sorted(f for f in os.listdir(os.path.join(ROOT, "build", "variants")) if f.endswith(".so"))
for lib in libs:
    env = dict(os.environ)
    if lib != "default":
        env["DCBF_LIB"] = os.path.join(ROOT, "build", "variants", lib)
    print(lib, flush=True)
    subprocess.run([sys.executable, "-c", CHILD], env=env)
