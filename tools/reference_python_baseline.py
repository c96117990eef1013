#!/usr/bin/env python
"""CPU baseline of the REFERENCE'S OWN Python path, for the record next to the GPU numbers.

The reference classes (MPC_LIP_sig_step.MPCCBF.solveMPCCBF, unmodified, imported from /root/reference) are timed around
solveMPCCBF only, single process and multiprocessing.Pool(P), on the first 64 scenarios of the bench workload (config 2, seed 0).
cyipopt / Ipopt / HSL are not installable here, so the solve behind the reference's cyipopt.Problem(...) call is the SciPy-SLSQP
stand-in of oracle/ref_loader.py driving the reference's objective / gradient / constraints / jacobian callbacks -- every number
is labelled "cyipopt unavailable".  /root/reference exists in the build container only (it cannot travel to the GPU box), so
this runs there and writes profiles/r03_reference_python.json, which bench.py attaches to its line as
cpu_baseline.reference_python (with the host it was measured on).

    OMP_NUM_THREADS=1 python tools/reference_python_baseline.py
"""
import json
import multiprocessing as mp
import os
import platform
import sys
import time

import numpy as np

os.environ.setdefault("OMP_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N = 64
_PLANNERS = {}


def _solve_chunk(args):
    from mujoco_lip_mpc_simulation_b200 import scenarios
    from oracle import ref_loader
    lo, hi = args
    sc = scenarios.make_batch("sig_step", 4096, seed=0)
    mod = ref_loader.load("MPC_LIP_sig_step")
    out = []
    for b in range(lo, hi):
        f = int(sc.field[b])
        if f not in _PLANNERS:
            cbf = sc.cir[f]
            raw = cbf - np.array([0.0, 0.0, sc.safe_dis])
            _PLANNERS[f] = mod.MPCCBF([list(sc.goal[b])], raw.tolist(), cbf.tolist(), [-0.5, 10.5])
        pl = _PLANNERS[f]
        xk = np.matrix(sc.x0[b]).T
        t0 = time.perf_counter()
        u = pl.solveMPCCBF(xk, int(sc.leg[b]), None)
        dt = time.perf_counter() - t0
        res = ref_loader.LAST_PROBLEM["result"]
        out.append((dt, int(res["status"]), int(res["nit"]), np.asarray(u, dtype=np.float64).ravel()))
    return out


def main():
    cores = os.cpu_count() or 1
    _solve_chunk((0, 2))                                   # warm-up (imports, constant matrices)
    t0 = time.perf_counter()
    single = _solve_chunk((0, N))
    t_single = time.perf_counter() - t0
    chunks = [(i * N // cores, (i + 1) * N // cores) for i in range(cores) if (i + 1) * N // cores > i * N // cores]
    with mp.Pool(cores) as pool:
        pool.map(_solve_chunk, [(0, 1)] * cores)           # warm-up of every worker
        t0 = time.perf_counter()
        parts = pool.map(_solve_chunk, chunks)
        t_pool = time.perf_counter() - t0
    lat = np.array([r[0] for r in single]) * 1e6
    st = np.array([r[1] for r in single])
    cpu = platform.processor() or ""
    try:
        cpu = [ln.split(":", 1)[1].strip() for ln in open("/proc/cpuinfo") if ln.startswith("model name")][0]
    except Exception:
        pass
    out = {"what": "reference MPC_LIP_sig_step.MPCCBF.solveMPCCBF (unmodified, /root/reference), timed around solveMPCCBF only",
           "solver_behind_cyipopt_call": "SciPy SLSQP on the reference callbacks (oracle/ref_loader.py) -- cyipopt unavailable",
           "scenarios": f"first {N} of config 2 (sig_step, K = 6 circles, seed 0), cold start",
           "host": {"cpu": cpu, "cores": cores, "where": "build container (the reference cannot travel to the GPU box)"},
           "single_process": {"solves_per_s": N / t_single, "p50_us": float(np.median(lat)), "p95_us": float(np.percentile(lat, 95))},
           "pool": {"processes": cores, "solves_per_s": sum(len(p) for p in parts) / t_pool},
           "status_hist": {str(int(k)): int((st == k).sum()) for k in np.unique(st)},
           "slsqp_iterations": {"mean": float(np.mean([r[2] for r in single])), "max": int(max(r[2] for r in single))}}
    path = os.path.join(ROOT, "profiles", "r03_reference_python.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
