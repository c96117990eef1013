"""GPU tests of the drop-in planner classes (same call surface as the reference files)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import c_oracle, lip_np  # noqa: E402

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("the gpu tests need a CUDA device; there is no CPU fallback")


def test_sig_step_main_scenario_closed_loop():
    """The reference's own __main__ (MPC_LIP_sig_step.py:553-575) run through the drop-in class."""
    from mujoco_lip_mpc_simulation_b200.MPC_LIP_sig_step import MPCCBF
    g = np.load(os.path.join(G, "config1_closed_loop.npz"))
    obs_list = np.array([[1, 1, 0.5], [2, 2, 0.5], [6, 4, 0.8], [7, 7, 1]])
    obs_safe = obs_list + [0, 0, 0.32]
    mpc = MPCCBF([[10, 10]], obs_list, obs_safe, [-0.5, 10.5])
    state, leg_ind, init_guess = np.concatenate([[0, 0], [0.6, -0.3], [0]]), 1, None
    for i in range(5):
        x_list, p0, hd_list, close2goal = mpc.gen_control_test(state, leg_ind, init_guess)
        assert len(x_list) == 3 and x_list[0].shape == (5,) and p0.shape == (3,) and len(hd_list) == 3
        np.testing.assert_allclose(p0, g["p0"][i], atol=1e-4)
        np.testing.assert_allclose(np.array(x_list), g["x_plan"][i], atol=1e-4)
        assert close2goal is False and mpc.last_status == 0
        leg_ind, init_guess, state = -leg_ind, x_list, x_list[0]
    u = mpc.solveMPCCBF(np.asarray(state).reshape(5, 1), leg_ind, init_guess)
    assert u.shape == (15,)
    np.testing.assert_allclose(mpc.solve_footdisp(state, u[0:5]).ravel(), mpc._last.p_plan[0, 0], atol=1e-10)


def test_modi_and_dd_planners_match_oracle():
    from mujoco_lip_mpc_simulation_b200 import MPC_DD_sig_step, MPC_LIP_modi
    g = np.load(os.path.join(G, "solves_modi.npz"))
    for b in range(6):
        mpc = MPC_LIP_modi.MPCCBF([list(g["goal"][b])], g["cir"][b], g["cir"][b], g["elp"][b], g["elp"][b], [-0.5, 10.5])
        x_list, p0, hd, close, feasi, pos_det = mpc.gen_control_test(g["x0"][b], int(g["leg"][b]), g["warm"][b])
        assert pos_det.shape == (126, 2)
        assert (feasi == 2) == (g["status"][b] == 2)
        if g["status"][b] == 0:
            np.testing.assert_allclose(p0, g["p_plan"][b, 0], atol=1e-4)
            np.testing.assert_allclose(pos_det[0], g["x0"][b][:2], atol=1e-12)
            np.testing.assert_allclose(pos_det[42], x_list[0][:2], atol=1e-9)
    g = np.load(os.path.join(G, "solves_dd.npz"))
    for b in range(6):
        mpc = MPC_DD_sig_step.MPCCBF([list(g["goal"][b])], g["cir"][b], g["cir"][b], g["elp"][b], g["elp"][b], [-0.5, 10.5])
        states, heading, control, close, fesi = mpc.gen_dd_control(g["x0"][b], g["warm"][b], g["last_u"][b])
        assert len(states) == 4 and len(heading) == 3 and control[0].shape == (2, 1)
        assert (fesi == 2) == (g["status"][b] == 2)
        if g["status"][b] == 0:
            np.testing.assert_allclose(np.concatenate([c.ravel() for c in control]), g["u"][b], atol=1e-4)


def test_lip_prob_callbacks_in_u_space():
    """LIP_Prob keeps the cyipopt protocol of the reference in its own 15-variable space."""
    from mujoco_lip_mpc_simulation_b200 import MPC_LIP_sig_step as S
    g = np.load(os.path.join(G, "callbacks_sig_step.npz"), allow_pickle=True)
    for b in range(0, 24, 6):
        prob = S.LIP_Prob(g["xk"][b], None, None, None, None, None, None, g["cir"][b], g["goal_eff"][b], 3)
        u = g["u"][b]
        assert abs(prob.objective(u) - g["f"][b]) <= 1e-12 * abs(g["f"][b])
        np.testing.assert_allclose(prob.gradient(u), g["grad"][b], atol=1e-10)
        np.testing.assert_allclose(prob.constraints(u), np.asarray(g["c"][b], float), atol=1e-12)
        np.testing.assert_allclose(prob.jacobian(u), np.asarray(g["jac"][b], float), atol=1e-10)


def test_batched_entry_points_of_the_planner():
    from mujoco_lip_mpc_simulation_b200 import scenarios
    from mujoco_lip_mpc_simulation_b200.MPC_LIP_sig_step import MPCCBF
    sc = scenarios.make_batch("sig_step", 512, seed=12, n_fields=32)
    mpc = MPCCBF([[10, 10]], sc.cir[0], sc.cir[0], [-0.5, 10.5])
    mpc.set_fields(sc.cir)
    r = mpc.solve_batch(sc.x0, sc.leg, sc.warm, field=sc.field)
    P = c_oracle.params("sig_step", max_iter=300)
    ref = c_oracle.solve_batch(P, sc.x0, sc.goal, sc.leg, sc.cir, None, sc.warm, field=sc.field, threads=4)
    st = r.status.cpu().numpy()
    both = (st == 0) & (ref["status"] == 0)
    dp = np.abs(r.p_plan.cpu().numpy() - ref["p_plan"]).reshape(512, -1).max(axis=1)
    assert np.mean((st == 2) == (ref["status"] == 2)) >= 0.99 and np.mean(dp[both] <= 1e-4) >= 0.99
    ro = mpc.rollout_batch(6, sc.x0, sc.leg, field=sc.field)
    assert ro["traj"].shape == (512, 6, 8) and int(ro["steps_done"].min()) >= 1
