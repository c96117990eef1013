"""CPU tests against the reference's RECORDED runs (tests/golden/data_log_plans.npz: plan trajectories, feasible / failed labels and
obstacle fields copied out of /root/reference/data_log/LIP_me*.pkl by oracle/gen_golden.py:gen_data_log).  These arrays were produced
by the real reference pipeline -- MPCCBF.gen_control_test with cyipopt on the authors' machine -- so they pin what nothing else in the
repository can: the dense plan trajectory (xk_track_det) bit for bit, and the meaning of the pred_fail label (Ipopt status 2) in terms
of the D-CBF rows."""
import os

import numpy as np

from mujoco_lip_mpc_simulation_b200 import _lipmodel, data_log
from oracle import lip_np

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "data_log_plans.npz"))
SAFE_DIS, GAMMA = 0.4, 0.2      # main_sim_mpc.py:11, MPC_LIP_modi.py:405


def test_recorded_plan_trajectories_are_lip_flows_of_a_plan():
    for a in G["plan"]:
        x, v, p, res = data_log.plan_from_pos_det(a)
        assert res <= 1e-12
        # the restated xk_track_det reproduces the record from the recovered plan
        seg = np.concatenate([_lipmodel.track_det(np.concatenate([x[j], v[j], [0.0]]), np.concatenate([p[j], [0.0]]), 0.4) for j in range(3)])
        np.testing.assert_allclose(seg, a, rtol=0, atol=1e-12)
        # consecutive segments chain through the step-to-step map  x+ = A x + B p
        k = _lipmodel.constants()
        for j in range(2):
            nxt = k.A @ np.concatenate([x[j], v[j], [0.0]]) + k.B @ np.concatenate([p[j], [0.0]])
            np.testing.assert_allclose(nxt[0:2], x[j + 1], atol=1e-9)
            np.testing.assert_allclose(nxt[2:4], v[j + 1], atol=1e-8)


def _cbf_rows(plan, cir, elp):
    x, v, p, _ = data_log.plan_from_pos_det(plan)
    k = _lipmodel.constants()
    pos = [x[0], x[1], x[2], (k.A @ np.concatenate([x[2], v[2], [0.0]]) + k.B @ np.concatenate([p[2], [0.0]]))[0:2]]
    cs = cir + np.array([0.0, 0.0, SAFE_DIS])
    es = elp + np.array([0.0, 0.0, SAFE_DIS, SAFE_DIS, 0.0])
    rows = []
    for i in range(3):
        for c in cs:
            rows.append(lip_np.h_circle(c, *pos[i + 1]) + (GAMMA - 1.0) * lip_np.h_circle(c, *pos[i]))
        for e in es:
            rows.append(lip_np.h_ellipse(e, *pos[i + 1]) + (GAMMA - 1.0) * lip_np.h_ellipse(e, *pos[i]))
    return np.array(rows)


def test_recorded_labels_match_the_dcbf_rows():
    """the plans the reference filed under pred_fail (Ipopt status 2) violate a D-CBF row, the others do not."""
    worst = np.array([_cbf_rows(a, G["cir"][r], G["elp"][r]).min() for a, r in zip(G["plan"], G["run"])])
    fail, feasi = G["label"] == 2, G["label"] == 0
    assert np.mean(worst[fail] < -1e-4) >= 0.9, worst[fail]
    assert np.mean(worst[feasi] >= -1e-4) >= 0.95, worst[feasi]


def _dd_controls(plan, dt=0.4):
    """(v_i, w_i) of a recorded DD plan: x+ = x + dt v cos th, y+ = y + dt v sin th, th+ = th + w (MPC_DD_sig_step.py:356-363)"""
    d = np.diff(plan, axis=0)
    v = (d[:, 0] * np.cos(plan[:3, 2]) + d[:, 1] * np.sin(plan[:3, 2])) / dt
    return v, d[:, 2]


def test_recorded_dd_plans_follow_the_unicycle_model_and_their_labels():
    worst, defect = [], 0.0
    for a, r in zip(G["dd_plan"], G["dd_run"]):
        v, w = _dd_controls(a)
        roll = lip_np.dd_rollout(a[0], np.stack([v, w], axis=1).ravel())
        defect = max(defect, float(np.max(np.abs(np.asarray(roll)[1:] - a[1:]))))
        cs = G["dd_cir"][r] + np.array([0.0, 0.0, SAFE_DIS])
        es = G["dd_elp"][r] + np.array([0.0, 0.0, SAFE_DIS, SAFE_DIS, 0.0])
        rows = []
        for i in range(3):
            rows += [lip_np.h_circle(c, *a[i + 1, :2]) + (GAMMA - 1.0) * lip_np.h_circle(c, *a[i, :2]) for c in cs]
            rows += [lip_np.h_ellipse(e, *a[i + 1, :2]) + (GAMMA - 1.0) * lip_np.h_ellipse(e, *a[i, :2]) for e in es]
        worst.append(min(rows))
    assert defect <= 1e-12          # the recorded states are exactly the unicycle rollout of the recovered controls
    worst = np.array(worst)
    fail, feasi = G["dd_label"] == 2, G["dd_label"] == 0
    assert np.mean(worst[fail] < -1e-4) >= 0.95 and np.mean(worst[feasi] >= -1e-4) >= 0.9   # 100 % / 95 %


def test_run_files_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    traj = np.concatenate([rng.normal(size=(6, 7)), np.zeros((6, 1))], axis=1)
    traj[4, 7] = 2.0
    traj = np.concatenate([traj, np.full((2, 8), np.nan)])
    plans = rng.normal(size=(8, 126, 2))
    run = data_log.run_from_rollout(traj, G["cir"][0], G["elp"][0], plans=plans)
    assert len(run["pos"]) == 6 and len(run["pred_fail_end"]) == 1 and len(run["pred_feasi_end"]) == 5
    pre = str(tmp_path / "LIP_test_")
    data_log.write_run(pre, run)
    back = data_log.read_run(pre)
    assert set(back) == set(run)
    assert isinstance(back["cir"], list) and isinstance(back["cir"][0], list) and isinstance(back["pred_full_end"], list)
    for kk in ("pos", "time", "foot", "heading", "body_vel", "real_end"):
        assert isinstance(back[kk], np.ndarray)
        np.testing.assert_array_equal(back[kk], run[kk])
    np.testing.assert_array_equal(back["pred_fail_end"][0], plans[4])
    np.testing.assert_array_equal(back["pred_end"][2], plans[2][[0, 41]])


def _dd_resolve_inputs():
    """start state, recovered controls (warm start and previous control) and inflated obstacles of every recorded DD plan"""
    plans, run = G["dd_plan"], G["dd_run"]
    u = np.array([np.stack(_dd_controls(a), axis=1).ravel() for a in plans])
    cir = G["dd_cir"] + np.array([0.0, 0.0, SAFE_DIS])
    elp = G["dd_elp"] + np.array([0.0, 0.0, SAFE_DIS, SAFE_DIS, 0.0])
    return plans[:, 0, :], u, cir, elp, run.astype(np.int32)


def test_dd_solvers_reproduce_ipopt_feasibility_verdicts():
    """Re-solving the 110 recorded differential-drive re-plans from their start state: the oracle and the host build of the CUDA
    lane code must call infeasible (status 2) exactly the ones the reference's Ipopt run filed under pred_fail (status 2,
    logger_dd.py) -- the goal and the previous control of the recorded runs are not logged, so the plans themselves are not compared."""
    import hostsim_binding as hs
    from oracle import c_oracle
    x0, u, cir, elp, run = _dd_resolve_inputs()
    B = len(x0)
    goal = np.tile([10.0, 10.0], (B, 1))
    want = G["dd_label"] == 2
    r_host = hs.solve(hs.default_params("dd"), x0, goal, np.ones(B, np.int32), cir, elp, u, field=run, last_u=u[:, :2].copy())
    r_orc = c_oracle.solve_batch(c_oracle.params("dd"), x0, goal, np.ones(B, np.int32), cir, elp, u, field=run, last_u=u[:, :2].copy())
    for name, r in (("host build of the CUDA lanes", r_host), ("oracle", r_orc)):
        got = r["status"] == 2
        assert (got == want).mean() >= 0.98, (name, (got == want).mean())       # measured 109 / 110
        assert not (want & ~got).any(), name                                    # nothing Ipopt gave up on is called solved
