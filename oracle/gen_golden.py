"""ORACLE (test infrastructure) -- freeze golden vectors from the *reference's own code*.

Run in the build container only (needs /root/reference):   python -m oracle.gen_golden
Writes small .npz fixtures under tests/golden/:

  callbacks_{sig_step,modi,dd}.npz
      Random points pushed through the reference classes' LIP_Prob.objective / gradient / constraints / jacobian
      (MPC_LIP_sig_step.py:372-496, MPC_LIP_modi.py:430-583, MPC_DD_sig_step.py:351-477) and through
      MPCCBF.solveMPCCBF with a recording cyipopt stub, which pins the bound vectors cl/cu, the detour goal, the
      warm-start rule and (modi) the obstacle selection exactly as the reference builds them.
  solves_{sig_step,modi,dd}.npz
      Optima for scenarios of the bench distribution: solved by the C oracle, then *verified with the reference's
      callbacks*: feasibility of every row and the KKT residual  min_{lam>=0} |grad f - J_act^T lam| / |grad f|
      evaluated with the reference's gradient and Jacobian at the returned point.  Only verified KKT points
      (and certified-infeasible problems) are kept.  Real cyipopt is not installed, so Ipopt's own iterates are
      not part of the pin ("cyipopt unavailable").
  config1_closed_loop.npz
      The reference's __main__ scenario (MPC_LIP_sig_step.py:553-575), 5 closed-loop re-plans.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
from scipy.optimize import nnls

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mujoco_lip_mpc_simulation_b200 import scenarios  # noqa: E402
from oracle import c_oracle, lip_np, ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _lists(a):
    return [list(map(float, r)) for r in a]


def gen_callbacks(form: str, n_cases: int, seed: int):
    rng = np.random.default_rng(seed)
    mod = ref_loader.load({"sig_step": "MPC_LIP_sig_step", "modi": "MPC_LIP_modi", "dd": "MPC_DD_sig_step"}[form])
    rec = {k: [] for k in ("xk", "goal", "leg", "cir", "elp", "u", "last_u", "f", "grad", "c", "jac", "cl", "cu",
                           "goal_eff", "u0", "warm_in", "sel_c", "sel_e")}
    sc = scenarios.make_batch(form, n_cases, seed=seed, n_fields=n_cases, num_obs=4 if form == "sig_step" else 6)
    for b in range(n_cases):
        cir, elp = sc.cir[sc.field[b]], sc.elp[sc.field[b]]
        xk, goal, leg = sc.x0[b], sc.goal[b], int(sc.leg[b])
        if form == "sig_step":
            planner = mod.MPCCBF([list(goal)], cir, cir, [-0.5, 10.5])
            guess = None if b % 2 == 0 else [xk + rng.normal(size=5) * 0.05 for _ in range(3)]
            planner.solveMPCCBF(np.matrix(xk).T, leg, guess)
            warm_in = np.zeros(15) if guess is None else np.concatenate(guess)
        elif form == "modi":
            planner = mod.MPCCBF([list(goal)], _lists(cir), _lists(cir), _lists(elp), _lists(elp), [-0.5, 10.5])
            planner.select_obs(np.matrix(xk).T)
            warm_in = np.tile(xk, 3) + rng.normal(size=15) * 0.05
            planner.solveMPCCBF(np.matrix(xk).T, leg, warm_in)
        else:
            planner = mod.MPCCBF([list(goal)], _lists(cir), _lists(cir), _lists(elp), _lists(elp), [-0.5, 10.5])
            warm_in = sc.warm[b]
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                planner.solveMPCCBF(np.matrix(xk).T, warm_in, list(sc.last_u[b]))
        LP = ref_loader.LAST_PROBLEM
        prob = LP["obj"]
        n = LP["n"]
        u = LP["u0"] + rng.normal(size=n) * (0.05 if form == "dd" else 0.2)
        rec["xk"].append(xk); rec["goal"].append(goal); rec["leg"].append(leg)
        rec["cir"].append(cir); rec["elp"].append(elp); rec["u"].append(u)
        rec["last_u"].append(sc.last_u[b] if sc.last_u is not None else np.zeros(2))
        rec["f"].append(float(prob.objective(u)))
        rec["grad"].append(np.ravel(np.asarray(prob.gradient(u), dtype=np.float64)))
        c = np.ravel(np.asarray(prob.constraints(u), dtype=np.float64))
        J = np.asarray(prob.jacobian(u), dtype=np.float64).reshape(len(c), n)
        # pad rows to the unselected row count so that the arrays stack (modi selection drops rows)
        rec["c"].append(c); rec["jac"].append(J)
        rec["cl"].append(LP["cl"]); rec["cu"].append(LP["cu"])
        rec["goal_eff"].append(LP["goal"]); rec["u0"].append(LP["u0"]); rec["warm_in"].append(warm_in)
        if form == "modi":
            rec["sel_c"].append(np.array([any(np.allclose(o, s) for s in planner.sel_cir) for o in cir]))
            rec["sel_e"].append(np.array([any(np.allclose(o, s) for s in planner.sel_elp) for o in elp]))
        else:
            rec["sel_c"].append(np.ones(len(cir), bool)); rec["sel_e"].append(np.ones(len(elp), bool))
    out = {}
    for k, v in rec.items():
        try:
            out[k] = np.stack([np.asarray(x) for x in v])
        except ValueError:
            out[k] = np.array(v, dtype=object)
    np.savez_compressed(os.path.join(OUT, f"callbacks_{form}.npz"), **out, allow_pickle=True)
    print(form, "callbacks:", n_cases, "cases")


def kkt_with_reference(form, prob_obj, u, cl, cu):
    """feasibility + KKT residual of u, evaluated with the reference's own callbacks."""
    g = np.ravel(np.asarray(prob_obj.gradient(u), dtype=np.float64))
    c = np.ravel(np.asarray(prob_obj.constraints(u), dtype=np.float64))
    J = np.asarray(prob_obj.jacobian(u), dtype=np.float64).reshape(len(c), len(u))
    viol = max(0.0, float(np.max(cl - c)), float(np.max(np.where(np.isfinite(cu), c - cu, -1.0))))
    lo = np.isfinite(cl) & (c - cl <= 1e-6)
    hi = np.isfinite(cu) & (cu - c <= 1e-6)
    A = np.concatenate([J[lo], -J[hi]]).T
    if A.shape[1] == 0:
        return viol, float(np.linalg.norm(g) / max(1.0, np.linalg.norm(g)))
    _, rn = nnls(A, g, maxiter=2000)
    return viol, float(rn / max(1.0, np.linalg.norm(g)))


def gen_solves(form: str, n_cases: int, seed: int):
    mod = ref_loader.load({"sig_step": "MPC_LIP_sig_step", "modi": "MPC_LIP_modi", "dd": "MPC_DD_sig_step"}[form])
    sc = scenarios.make_batch(form, n_cases, seed=seed, n_fields=n_cases)
    P = c_oracle.params(form, max_iter=500)
    elp = sc.elp if sc.elp.shape[1] else None
    res = c_oracle.solve_batch(P, sc.x0, sc.goal, sc.leg, sc.cir, elp, sc.warm, field=sc.field, last_u=sc.last_u, threads=8)
    keep, kkt, viol = [], [], []
    for b in range(n_cases):
        cir, e = sc.cir[sc.field[b]], sc.elp[sc.field[b]]
        xk, goal, leg = sc.x0[b], sc.goal[b], int(sc.leg[b])
        # build the reference problem object exactly as the reference does (recording stub; the SLSQP result is unused)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            if form == "sig_step":
                planner = mod.MPCCBF([list(goal)], cir, cir, [-0.5, 10.5])
                planner.solveMPCCBF(np.matrix(xk).T, leg, None)
            elif form == "modi":
                planner = mod.MPCCBF([list(goal)], _lists(cir), _lists(cir), _lists(e), _lists(e), [-0.5, 10.5])
                planner.select_obs(np.matrix(xk).T)
                planner.solveMPCCBF(np.matrix(xk).T, leg, sc.warm[b])
            else:
                planner = mod.MPCCBF([list(goal)], _lists(cir), _lists(cir), _lists(e), _lists(e), [-0.5, 10.5])
                planner.solveMPCCBF(np.matrix(xk).T, sc.warm[b], list(sc.last_u[b]))
        LP = ref_loader.LAST_PROBLEM
        if res["status"][b] == 0:
            v, r = kkt_with_reference(form, LP["obj"], res["u"][b], LP["cl"], LP["cu"])
            if form == "dd":
                v = max(v, float(np.max(np.asarray(LP["lb"]) - res["u"][b])), float(np.max(res["u"][b] - np.asarray(LP["ub"]))))
            ok = v <= 1e-6 and r <= 1e-5
        elif res["status"][b] == 2:
            v, r, ok = float(res["viol"][b]), np.nan, True
        else:
            v, r, ok = np.nan, np.nan, False
        keep.append(ok); kkt.append(r); viol.append(v)
    keep = np.array(keep)
    print(form, "solves: kept", int(keep.sum()), "of", n_cases, "| status", {int(k): int((res['status'][keep] == k).sum()) for k in np.unique(res['status'][keep])},
          "| max kkt", np.nanmax(np.array(kkt)[keep & (res['status'] == 0)]))
    sel = np.where(keep)[0]
    np.savez_compressed(os.path.join(OUT, f"solves_{form}.npz"),
                        x0=sc.x0[sel], goal=sc.goal[sel], leg=sc.leg[sel], cir=sc.cir[sc.field[sel]], elp=sc.elp[sc.field[sel]],
                        warm=sc.warm[sel], last_u=(sc.last_u[sel] if sc.last_u is not None else np.zeros((len(sel), 2))),
                        u=res["u"][sel], p_plan=res["p_plan"][sel], x_plan=res["x_plan"][sel], f=res["f"][sel],
                        status=res["status"][sel], viol_ref=np.array(viol)[sel], kkt_ref=np.array(kkt)[sel])


def gen_config1():
    sc = scenarios.config1()
    P = c_oracle.params("sig_step", max_iter=500)
    mod = ref_loader.load("MPC_LIP_sig_step")
    planner = mod.MPCCBF([[10, 10]], sc.cir[0], sc.cir[0], [-0.5, 10.5])
    state, leg, guess = sc.x0[0].copy(), 1, None
    rows = []
    for _ in range(5):
        u0 = lip_np.sig_step_warm_start(state, guess)
        r = c_oracle.solve(P, state, [10, 10], leg, sc.cir[0], None, u0)
        planner.solveMPCCBF(np.matrix(state).T, leg, guess)   # records the reference's own problem object
        LP = ref_loader.LAST_PROBLEM
        assert np.allclose(LP["u0"], u0)
        v, k = kkt_with_reference("sig_step", LP["obj"], r["u"], LP["cl"], LP["cu"])
        rows.append(dict(state=state.copy(), leg=leg, u0=u0, p0=r["p_plan"][0].copy(), f=r["f"], x_plan=r["x_plan"].copy(),
                         status=r["status"], viol_ref=v, kkt_ref=k))
        guess = [r["x_plan"][0], r["x_plan"][1], r["x_plan"][2]]
        state, leg = r["x_plan"][0].copy(), -leg
    np.savez_compressed(os.path.join(OUT, "config1_closed_loop.npz"), cir=sc.cir[0],
                        **{k: np.array([row[k] for row in rows]) for k in rows[0]})
    print("config1:", [np.round(r["p0"], 6).tolist() for r in rows], "max kkt", max(r["kkt_ref"] for r in rows))


def gen_helpers():
    """closed-form helpers of the call surface, straight from the reference classes."""
    import types
    rng = np.random.default_rng(7)
    sig = ref_loader.load("MPC_LIP_sig_step")
    dd = ref_loader.load("MPC_DD_sig_step")
    obs = np.array([[1, 1, 0.82]])
    ps = sig.MPCCBF([[10, 10]], obs, obs, [-0.5, 10.5])
    pd = dd.MPCCBF([[10, 10]], obs, obs, [], [], [-0.5, 10.5])
    out = {}
    pos, vel, hd, gp, tr = rng.normal(size=2), rng.normal(size=2), 0.3, np.array([0.13, -0.2, 0.1]), 0.25
    xn, traj = ps.get_next_states(pos, vel, hd, gp, tr)
    out.update(gns_in=np.concatenate([pos, vel, [hd], gp, [tr]]), gns_x=xn, gns_traj=traj)
    out["alip_des_vel"] = np.array([ps.alip_des_vel(0.7, 1), ps.alip_des_vel(0.5, -1)])
    xs, vd = rng.normal(size=5), rng.normal(size=2)
    out.update(cfv_in=np.concatenate([xs, vd]), cfv_out=ps.cal_foot_with_veldes(xs, vd))
    u5 = rng.normal(size=5)
    out.update(sfd_in=np.concatenate([xs, u5]), sfd_out=np.ravel(ps.solve_footdisp(np.matrix(xs).T, np.matrix(u5).T)))
    out.update(xtd_in=np.concatenate([xs, gp]), xtd_out=ps.xk_track_det(xs, gp, 0.4))
    hl = rng.normal(size=12) * 0.2
    out.update(tube_in=hl, tube_sig=ps.tube_func(hl, 0.05), tube_dd=pd.tube_func(hl, 0.05))
    pdz = rng.normal(size=2)
    out.update(cfp_in=np.concatenate([xs, pdz]), cfp_out=pd.cal_foot_with_posdes(xs, pdz))
    # ALIP closed form: the reference module imports robot-specific files that are broken / irrelevant here
    for name in ("fromFROST", "forwardKinematics", "helper", "pdb"):
        sys.modules.setdefault(name, types.ModuleType(name))
    import importlib.util
    spec = importlib.util.spec_from_file_location("_dcbf_ref_alip", ref_loader.REFERENCE_ROOT + "/ALIP_plan/planner.py")
    alip_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(alip_mod)
    prm = types.SimpleNamespace(H=1.0, T=0.4, m=45.0)
    al = alip_mod.ALIP(prm)
    al.DRS_motion_int = lambda a, b: (np.zeros(2), np.zeros(2))   # amplitudes are 0 (planner.py:47-50); idqp never exists
    x0, y0 = np.array([0.05, 2.0]), np.array([0.1, -1.0])
    xt, yt = al.getTimedState(x0, y0, 0.1)
    Ly, Lx = al.AMprediction(xt, yt, 0.1)
    st_r = al.computeStepping(np.array([0.1, 0.1]), Ly, Lx, 0.5, 1)
    st_l = al.computeStepping(np.array([0.1, -0.1]), Ly, Lx, 0.5, -1)
    out.update(alip_in=np.concatenate([x0, y0, [0.1]]), alip_xt=xt, alip_yt=yt, alip_am=np.array([Ly, Lx]),
               alip_step_r=np.array(st_r, dtype=float), alip_step_l=np.array(st_l, dtype=float),
               alip_reg=np.array([al.regulate_lateral_step(1, 0.05), al.regulate_lateral_step(1, 0.5), al.regulate_lateral_step(-1, -0.05),
                                  al.regulate_lateral_step(-1, -0.3), al.regulate_lateral_step(0, 0.7)]))
    # heading input of the LIP prediction: Logger.tube_func / avg_hd / angle_A_minus_B (data_procs/logger_mpc.py:169-175,208-215,284-300)
    import contextlib, io
    spec = importlib.util.spec_from_file_location("_dcbf_ref_logger_mpc", ref_loader.REFERENCE_ROOT + "/data_procs/logger_mpc.py")
    lg_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lg_mod)
    L = lg_mod.Logger
    n_h = 64
    cur = rng.uniform(-3.2, 3.2, n_h)
    turn = np.concatenate([rng.uniform(-0.4, 0.4, n_h - 4), [0.0, 0.15, -0.15, 0.149]])
    hds = cur[:, None] + np.cumsum(rng.uniform(-0.3, 0.3, (n_h, 3)), axis=1)
    hds[:8] += rng.choice([-2 * math.pi, 2 * math.pi], size=(8, 1))     # headings on the other side of the +-pi cut
    nt, hp = np.zeros(n_h), np.zeros(n_h)
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(n_h):
            me = types.SimpleNamespace(nex_turn=float(turn[i]), mpc_hds_list=[float(v) for v in hds[i]])
            me.angle_A_minus_B = lambda a, b, me=me: L.angle_A_minus_B(me, a, b)
            me.nex_turn = L.tube_func(me, me.nex_turn, float(cur[i]))
            nt[i], hp[i] = me.nex_turn, L.avg_hd(me, float(cur[i]))
    out.update(hdin_cur=cur, hdin_turn=turn, hdin_hds=hds, hdin_nex_turn=nt, hdin_pr=hp)
    np.savez_compressed(os.path.join(OUT, "helpers.npz"), **out)
    print("helpers:", len(out), "arrays")


def gen_data_log():
    """A sample of the reference's RECORDED runs (data_log/LIP_me*.pkl, written by data_procs/logger_mpc.py:449-474 on the authors'
    machine): plan trajectories with their feasible / failed label and the obstacle field of the run.  These are outputs of the
    real reference pipeline (cyipopt included), not of anything in this repository."""
    import pickle
    rec = {k: [] for k in ("plan", "label", "run")}
    fields_c, fields_e = [], []
    for ri, name in enumerate(("me1", "me2", "me3", "me7", "me12")):
        pre = os.path.join(ref_loader.REFERENCE_ROOT, "data_log", f"LIP_{name}_")
        load = lambda k: pickle.load(open(pre + k + ".pkl", "rb"))   # noqa: E731
        fields_c.append(np.asarray(load("cir"), dtype=np.float64)); fields_e.append(np.asarray(load("ellp"), dtype=np.float64))
        for lab, key in ((0, "pred_feasi_end"), (2, "pred_fail_end")):
            plans = load(key)
            take = plans[:: max(1, len(plans) // 12)][:12]
            for a in take:
                rec["plan"].append(np.asarray(a, dtype=np.float64)); rec["label"].append(lab); rec["run"].append(ri)
    # differential-drive runs (data_procs/logger_dd.py:445-466): plans are the four states [x, y, theta] of gen_dd_control
    dd = {k: [] for k in ("plan", "label", "run")}
    dd_c, dd_e = [], []
    for ri, name in enumerate(("me1", "me4", "me9", "me15")):
        pre = os.path.join(ref_loader.REFERENCE_ROOT, "data_log", f"DD_{name}_")
        load = lambda k: pickle.load(open(pre + k + ".pkl", "rb"))   # noqa: E731
        dd_c.append(np.asarray(load("cir"), dtype=np.float64)); dd_e.append(np.asarray(load("ellp"), dtype=np.float64))
        for lab, key in ((0, "pred_feasi_end"), (2, "pred_fail_end")):
            plans = load(key)
            for a in plans[:: max(1, len(plans) // 16)][:16]:
                dd["plan"].append(np.asarray(a, dtype=np.float64)); dd["label"].append(lab); dd["run"].append(ri)
    np.savez_compressed(os.path.join(OUT, "data_log_plans.npz"), plan=np.array(rec["plan"]), label=np.array(rec["label"]),
                        run=np.array(rec["run"]), cir=np.array(fields_c), elp=np.array(fields_e),
                        dd_plan=np.array(dd["plan"]), dd_label=np.array(dd["label"]), dd_run=np.array(dd["run"]),
                        dd_cir=np.array(dd_c), dd_elp=np.array(dd_e))
    print("data_log DD:", len(dd["plan"]), "recorded plans,", int(np.sum(np.array(dd["label"]) == 2)), "labelled infeasible")
    print("data_log:", len(rec["plan"]), "recorded plans,", int(np.sum(np.array(rec["label"]) == 2)), "labelled infeasible")


def gen_rand_obs():
    """Obstacle fields drawn by the reference generator itself (rand_obs.gen_ran_obs_list, rand_obs.py:74-81) under
    random.seed(k): the reference never seeds, so these are samples of its distribution, used to pin the PROPERTIES the batched
    generator must reproduce (ranges, rounding, separation) -- tests/test_scenario_gen_cpu.py."""
    import random
    ref_loader._install_stubs()
    if ref_loader.REFERENCE_ROOT not in sys.path:
        sys.path.append(ref_loader.REFERENCE_ROOT)
    import importlib
    ro = importlib.import_module("rand_obs")
    import signal

    class _Stall(Exception):
        pass

    def _alarm(*_):
        raise _Stall()
    signal.signal(signal.SIGALRM, _alarm)

    def draw(seed, num, typ):
        # random_circle has no exit when the circles placed so far leave no room (rand_obs.py:33-52): give each seed 2 s
        random.seed(seed)
        signal.setitimer(signal.ITIMER_REAL, 2.0)
        try:
            return ro.gen_ran_obs_list(num, typ)
        except _Stall:
            return None
        finally:
            signal.setitimer(signal.ITIMER_REAL, 0.0)
    cir6, mixc, mixe, stalled = [], [], [], 0
    k = 0
    while len(cir6) < 48 and k < 400:
        r = draw(1000 + k, 6, "cir"); k += 1
        if r is None:
            stalled += 1
        else:
            cir6.append(r[0])
    k = 0
    while len(mixc) < 48 and k < 400:
        r = draw(2000 + k, 6, "mix"); k += 1
        if r is None:
            stalled += 1
        else:
            mixc.append(r[0]); mixe.append(r[1])
    np.savez_compressed(os.path.join(OUT, "rand_obs_fields.npz"), cir6=np.array(cir6, dtype=np.float64),
                        mix_cir=np.array(mixc, dtype=np.float64), mix_elp=np.array(mixe, dtype=np.float64), stalled=np.int64(stalled))
    print("rand_obs:", len(cir6), "circle fields and", len(mixc), "mixed fields from the reference generator;", stalled, "seeds never finished")


def gen_sup_learn():
    """The reference's recorded learning set sup_learn/*.csv (640 control ticks of a main_sim_mpc.py run: MPC_LIP_modi with six
    circles, real cyipopt): features, the MPC's answers and the robot's.  Outputs of the real pipeline, not of this repository."""
    d = os.path.join(ref_loader.REFERENCE_ROOT, "sup_learn")
    X = np.loadtxt(os.path.join(d, "X_data.csv"), delimiter=",")
    y = np.loadtxt(os.path.join(d, "y_mpc_data.csv"), delimiter=",")
    a = np.loadtxt(os.path.join(d, "y_act_data.csv"), delimiter=",")
    with open(os.path.join(d, "X_data.csv")) as f:
        first = f.readline().strip()
    np.savez_compressed(os.path.join(OUT, "sup_learn.npz"), X=X, y_mpc=y, y_act=a, first_line=np.array(first))
    print("sup_learn:", X.shape, y.shape, a.shape)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--sup-learn-only" in sys.argv:
        gen_sup_learn()
        sys.exit(0)
    if "--rand-obs-only" in sys.argv:
        gen_rand_obs()
        sys.exit(0)
    if "--data-log-only" in sys.argv:
        gen_data_log()
        sys.exit(0)
    gen_helpers()
    gen_data_log()
    gen_rand_obs()
    gen_sup_learn()
    if "--helpers-only" in sys.argv:
        sys.exit(0)
    for form, seed in (("sig_step", 101), ("modi", 102), ("dd", 103)):
        gen_callbacks(form, 24, seed)
    gen_config1()
    for form, seed in (("sig_step", 201), ("modi", 202), ("dd", 203)):
        gen_solves(form, 96, seed)
