"""CPU tests: the ORACLE against every re-plan the reference recorded with the real cyipopt (tests/golden/recorded_runs.npz =
all 1 778 LIP and 1 377 differential-drive plans of /root/reference/data_log, frozen by oracle/gen_recorded.py).  This is the pin
of the solver oracle to Ipopt's own output: the same replay runs through the CUDA path in tests/test_gpu_recorded.py."""
import numpy as np

import recorded_replay as rr
from oracle import c_oracle


def test_fixture_holds_every_recorded_plan():
    g = rr.load()
    assert len(g["plan_label"]) == 1778 and int((g["plan_label"] == 2).sum()) == 238 and len(g["lip_name"]) == 23
    assert len(g["ddp_label"]) == 1377 and int((g["ddp_label"] == 2).sum()) == 294 and len(g["dd_name"]) == 21
    assert float(g["plan_fit_res"].max()) <= 1e-13      # a recorded plan trajectory IS the LIP flow of its fitted (x, v, p)
    for grp in rr.dd_inputs(g):
        assert grp["resid"] <= 1e-14                    # a recorded DD plan IS the unicycle rollout of its recovered controls


def test_oracle_reproduces_recorded_cyipopt_plans_lip():
    """Replay of all 23 LIP runs (open loop on the logged robot states, chained through the heading input and the warm start)."""
    g = rr.load()
    out = rr.replay_all(lambda c, e, **kw: rr.OracleBackend(c, e, **kw), g)
    s = rr.summarize_lip(out)
    print(s)
    assert s["n"] == 1778
    assert s["start_state_err"] <= 1e-12                          # x_nex of every recorded plan from the per-tick logs
    assert s["class_agree"] >= 0.98                               # Ipopt's verdict (pred_fail <=> status 2)
    assert s["rec_fail_ours_infeasible"] >= 0.98 * s["rec_fail"]
    assert s["dp0_median"] <= 1e-7 and s["dp0_le_1e4"] >= 0.80 and s["dp0_le_1e3"] >= 0.90
    # the run with a logged heading input (LIP_mexy, 3 240 chained re-plans): the chain reproduces it
    m = ~np.isnan(out["hd_pr_logged"])
    assert m.sum() == 81 and np.median(np.abs(out["hd_pr"][m] - out["hd_pr_logged"][m])) <= 1e-7


def test_oracle_reproduces_recorded_cyipopt_plans_dd():
    g = rr.load()
    lab, st, u, ur = [], [], [], []
    for grp in rr.dd_inputs(g):
        B = len(grp["x0"])
        P = c_oracle.params("dd", max_iter=300, **{{"w_p": "p"}.get(k, k): v for k, v in grp["params"].items()})
        o = c_oracle.solve_batch(P, grp["x0"], np.tile(rr.GOAL, (B, 1)), np.ones(B, np.int32), grp["cir"], grp["elp"], grp["u"],
                                 field=grp["field"], last_u=grp["u"][:, :2].copy(), threads=8)
        lab.append(grp["label"]); st.append(o["status"]); u.append(o["u"]); ur.append(grp["u"])
    s = rr.summarize_dd(*(np.concatenate(a) for a in (lab, st, u, ur)))
    print(s)
    assert s["n"] == 1377
    assert s["rec_fail_ours_infeasible"] == s["rec_fail"]         # everything Ipopt gave up on is infeasible here too
    assert s["class_agree"] >= 0.95                               # the rest: "feasi" plans of the iteration-capped reference (status -1)
    assert s["du_le_1e3"] >= 0.90 and s["du_le_1e2"] >= 0.98      # u_{-1} of the smoothness cost is not logged
