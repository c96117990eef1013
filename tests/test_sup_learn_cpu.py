"""The reference's recorded learning set (sup_learn/*.csv, 640 control ticks of a main_sim_mpc.py run with the real cyipopt;
fixture tests/golden/sup_learn.npz made by oracle/gen_golden.py:gen_sup_learn) against this repository without a GPU:
  * the file layout (data_procs/logger_iml.py:377-401) round-trips through data_log.read_sup_learn / write_sup_learn,
  * the recorded LIP prediction x_nex[0:2] (MPC_LIP_modi.get_next_states, :149-178) is reproduced to the bit,
  * the oracle and the host build of the CUDA lane code, re-solving every recorded tick from a cold start, land on the foot
    placement the reference's Ipopt run chose (the recorded run warm-started from its previous plan and stopped after at most 30
    L-BFGS iterations, so agreement is statistical: tolerances below)."""
import os

import numpy as np

import hostsim_binding as hs
from mujoco_lip_mpc_simulation_b200 import _lipmodel, data_log
from oracle import c_oracle

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sup_learn.npz"))
SAFE_DIS = 0.4    # main_sim_mpc.py:11-14


def recorded_ticks():
    f = data_log.sup_learn_fields(G["X"], G["y_mpc"])
    T = len(f["pos"])
    xn = np.zeros((T, 5))
    for i in range(T):
        t = float(f["rest_t"][i])
        A, B = _lipmodel.flow_matrices(t, t / 0.4)
        xn[i] = A @ np.array([*f["pos"][i], *f["vel"][i], f["heading"][i]]) + B @ np.array([*f["stance"][i], f["hd_input_pr"][i]])
    return f, xn


def foot_agreement(p_plan, status, f):
    d = np.linalg.norm(p_plan[:, 0, :2] - f["foot"], axis=1)
    ok = status == 0
    return d, ok


def test_layout_round_trip(tmp_path):
    data_log.write_sup_learn(str(tmp_path), G["X"], G["y_mpc"], G["y_act"])
    with open(tmp_path / "X_data.csv") as fh:
        assert fh.readline().strip() == str(G["first_line"])          # byte-identical to the reference's file
    back = data_log.read_sup_learn(str(tmp_path))
    assert np.array_equal(back["X"], G["X"]) and np.array_equal(back["y_mpc"], G["y_mpc"]) and np.array_equal(back["y_act"], G["y_act"])
    assert back["obs"].shape == (640, 6, 3) and set(np.unique(back["leg_ind"])) == {-1, 1}
    assert np.allclose(np.unique(np.round(back["rest_t"], 6)), np.arange(1, 9) * 0.05)
    f = back
    X, y = data_log.sup_learn_rows(f["obs"][0], f["pos"], f["vel"], f["heading"], f["stance"], f["goal"], f["leg_ind"], f["rest_t"],
                                   f["foot"], f["hd_input_pr"], f["x_nex_pos"], f["v_des"])
    assert np.array_equal(X, G["X"])
    np.testing.assert_allclose(y, G["y_mpc"], rtol=0, atol=1e-15)


def test_recorded_prediction_bit_exact():
    f, xn = recorded_ticks()
    assert np.array_equal(xn[:, :2], f["x_nex_pos"])


def test_solvers_land_on_the_recorded_foot_placement():
    f, xn = recorded_ticks()
    T = len(xn)
    cir = (f["obs"][0] + np.array([0.0, 0.0, SAFE_DIS]))[None]
    leg = (-f["leg_ind"]).astype(np.int32)                                  # logger_mpc.py:336 passes (-1) * leg_ind
    warm = np.tile(xn, (1, 3))
    r_host = hs.solve(hs.default_params("modi"), xn, f["goal"], leg, cir, None, warm)
    r_orc = c_oracle.solve_batch(c_oracle.params("modi"), xn, f["goal"], leg, cir, None, warm, field=np.zeros(T, np.int32), threads=4)
    for name, r in (("host build of the CUDA lanes", r_host), ("oracle", r_orc)):
        d, ok = foot_agreement(r["p_plan"], r["status"], f)
        assert ok.mean() > 0.8, name
        assert np.median(d[ok]) < 5e-4, (name, np.median(d[ok]))          # measured 2.1e-4 m
        assert (d[ok] < 1e-3).mean() > 0.65, (name, (d[ok] < 1e-3).mean())   # measured 0.74
        assert (d < 1e-2).mean() > 0.9, (name, (d < 1e-2).mean())          # measured 0.95 over all 640 ticks
        assert d.max() < 0.12, name
    # and the two agree with each other far more tightly than either does with the 30-iteration L-BFGS run
    both = (r_host["status"] == 0) & (r_orc["status"] == 0)
    gap = np.abs(r_host["p_plan"][both] - r_orc["p_plan"][both]).reshape(both.sum(), -1).max(axis=1)
    assert (gap < 1e-5).mean() > 0.97, (gap < 1e-5).mean()               # a handful of ticks have two local optima
