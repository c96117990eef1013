"""GPU tests: the CUDA solver (through the C ABI: dcbf_heading_input, dcbf_tick, dcbf_solve) against EVERY re-plan the reference
recorded with the real cyipopt -- 1 778 LIP plans of 23 runs and 1 377 differential-drive plans of 21 runs
(tests/golden/recorded_runs.npz, frozen from /root/reference/data_log by oracle/gen_recorded.py; replay in recorded_replay.py).

Reported twice, as SURVEY.md 8(d) asks: against the labels AS SHIPPED (the reference files a plan under pred_fail iff Ipopt's
status is 2, main_sim_mpc.py:118-121 -- iteration-capped exits -1 / 1 / -2 count as feasible), and with those exits bucketed:
a recorded "feasible" plan for a problem that is infeasible (status 2 here, confirmed by the oracle in the CPU test) is a
capped-iteration exit of the reference, not a solution."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import recorded_replay as rr  # noqa: E402


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.fail("the gpu tests need a CUDA device; there is no CPU fallback")
    from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
    return DcbfSolver


def test_cuda_replay_reproduces_recorded_cyipopt_plans_lip(gpu):
    g = rr.load()
    out = rr.replay_all(lambda c, e, **kw: rr.CudaBackend(c, e, **kw), g)
    s = rr.summarize_lip(out)
    print(s)
    assert s["n"] == 1778
    assert s["start_state_err"] <= 1e-11                          # tick_prepare_kernel: x_nex of every recorded plan
    assert s["class_agree"] >= 0.98                               # as shipped
    assert s["rec_fail_ours_infeasible"] >= 0.98 * s["rec_fail"]
    assert s["dp0_median"] <= 1e-7 and s["dp0_le_1e4"] >= 0.80 and s["dp0_le_1e3"] >= 0.90
    m = ~np.isnan(out["hd_pr_logged"])                            # heading_input_kernel chained over 3 240 re-plans of LIP_mexy
    assert m.sum() == 81 and np.median(np.abs(out["hd_pr"][m] - out["hd_pr_logged"][m])) <= 1e-7
    # every plan returned as solved is feasible
    ok = out["status"] == 0
    assert float(out["viol"][ok].max()) <= 1e-6


def test_cuda_solver_reproduces_recorded_cyipopt_plans_dd(gpu):
    g = rr.load()
    lab, st, u, ur = [], [], [], []
    for grp in rr.dd_inputs(g):
        B = len(grp["x0"])
        s = gpu("dd", device=0, max_iter=300, **grp["params"])
        s.set_fields(grp["cir"], grp["elp"])
        r = s.solve(grp["x0"], np.tile(rr.GOAL, (B, 1)), None, grp["u"], field=grp["field"], last_u=grp["u"][:, :2].copy())
        lab.append(grp["label"]); st.append(r.status.cpu().numpy()); u.append(r.u.cpu().numpy()); ur.append(grp["u"])
    sm = rr.summarize_dd(*(np.concatenate(a) for a in (lab, st, u, ur)))
    print(sm)
    assert sm["n"] == 1377
    assert sm["rec_fail_ours_infeasible"] == sm["rec_fail"]
    assert sm["class_agree"] >= 0.95
    assert sm["du_le_1e3"] >= 0.90 and sm["du_le_1e2"] >= 0.98
