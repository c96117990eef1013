"""CPU tests: closed-form helpers of the drop-in call surface against golden vectors frozen from the reference
classes (tests/golden/helpers.npz, oracle/gen_golden.py:gen_helpers)."""
import os

import numpy as np

from mujoco_lip_mpc_simulation_b200 import _lipmodel
from mujoco_lip_mpc_simulation_b200.ALIP_plan.planner import ALIP, ALIPParam

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "helpers.npz"))


class _Host(object):
    """the helper methods do not touch the GPU: bind them to a bare object carrying the constants"""
    def __init__(self):
        from mujoco_lip_mpc_simulation_b200._planner_base import LipPlannerBase
        k = _lipmodel.constants()
        self.__class__ = type("HostPlanner", (LipPlannerBase,), {})
        self.beta, self.dt, self.sigma, self.step_gap = k.beta, k.dt, k.sigma, 0.3
        self.A, self.B, self.W, self.inv_B_vel_shr = k.A, k.B, k.W, k.inv_B_vel_shr


def test_lip_helpers_match_reference():
    h = _Host()
    i = G["gns_in"]
    xn, traj = h.get_next_states(i[0:2], i[2:4], i[4], i[5:8], i[8])
    np.testing.assert_allclose(xn, G["gns_x"], atol=1e-13)
    np.testing.assert_allclose(traj, G["gns_traj"], atol=1e-13)
    np.testing.assert_allclose([h.alip_des_vel(0.7, 1), h.alip_des_vel(0.5, -1)], G["alip_des_vel"], atol=1e-14)
    np.testing.assert_allclose(h.cal_foot_with_veldes(G["cfv_in"][:5], G["cfv_in"][5:]), G["cfv_out"], atol=1e-13)
    np.testing.assert_allclose(np.ravel(h.solve_footdisp(G["sfd_in"][:5], G["sfd_in"][5:])), G["sfd_out"], atol=1e-13)
    np.testing.assert_allclose(h.xk_track_det(G["xtd_in"][:5], G["xtd_in"][5:], 0.4), G["xtd_out"], atol=1e-13)
    np.testing.assert_allclose(h.tube_func(G["tube_in"], 0.05), G["tube_sig"], atol=1e-15)
    np.testing.assert_allclose(_lipmodel.tube(G["tube_in"], 0.05, 0.2, 0.3), G["tube_dd"], atol=1e-15)
    k = _lipmodel.constants()
    ax = k.A @ G["cfp_in"][:5]
    np.testing.assert_allclose(k.inv_B_pos_shr @ (G["cfp_in"][5:] - ax[0:2]), G["cfp_out"], atol=1e-13)


def test_alip_closed_form_matches_reference():
    a = ALIP(ALIPParam(H=1.0, T=0.4, m=45.0))
    i = G["alip_in"]
    xt, yt = a.getTimedState(i[0:2], i[2:4], i[4])
    np.testing.assert_allclose(xt, G["alip_xt"], atol=1e-13)
    np.testing.assert_allclose(yt, G["alip_yt"], atol=1e-13)
    Ly, Lx = a.AMprediction(xt, yt, i[4])
    np.testing.assert_allclose([Ly, Lx], G["alip_am"], atol=1e-12)
    np.testing.assert_allclose(a.computeStepping(np.array([0.1, 0.1]), Ly, Lx, 0.5, 1), G["alip_step_r"], atol=1e-13)
    np.testing.assert_allclose(a.computeStepping(np.array([0.1, -0.1]), Ly, Lx, 0.5, -1), G["alip_step_l"], atol=1e-13)
    got = [a.regulate_lateral_step(1, 0.05), a.regulate_lateral_step(1, 0.5), a.regulate_lateral_step(-1, -0.05),
           a.regulate_lateral_step(-1, -0.3), a.regulate_lateral_step(0, 0.7)]
    np.testing.assert_allclose(got, G["alip_reg"], atol=0)
    # batched evaluation by broadcasting gives the same numbers
    xb, yb = a.getTimedState(np.tile(i[0:2], (4, 1)), np.tile(i[2:4], (4, 1)), i[4])
    np.testing.assert_allclose(xb[2], G["alip_xt"], atol=1e-13)
    np.testing.assert_allclose(yb[3], G["alip_yt"], atol=1e-13)


def test_model_constants_match_survey():
    k = _lipmodel.constants()
    assert abs(k.sigma - 5.637507) < 1e-6
    assert abs(k.A[0, 0] - 1.892976) < 1e-6 and abs(k.A[0, 2] - 0.513166) < 1e-6 and abs(k.A[2, 0] - 5.034157) < 1e-6
    assert abs(k.W[0, 0] + 0.15223) < 1e-5 and abs(k.W[0, 2] + 0.17164) < 1e-5
    np.testing.assert_allclose(k.W @ k.B, np.eye(3), atol=1e-12)


def test_heading_input_rules_match_reference_logger():
    """Logger.tube_func / avg_hd / angle_A_minus_B (data_procs/logger_mpc.py:169-175,208-215,284-300), bit for bit."""
    nt = _lipmodel.logger_tube(G["hdin_turn"], G["hdin_cur"])
    assert np.array_equal(nt, G["hdin_nex_turn"])
    assert np.array_equal(_lipmodel.avg_hd(G["hdin_cur"], nt, G["hdin_hds"]), G["hdin_pr"])
    assert (np.abs(G["hdin_nex_turn"]) <= 0.7 * 0.4 + 1e-12).all()
