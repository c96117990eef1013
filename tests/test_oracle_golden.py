"""CPU tests: the oracle (numpy + C restatements) against golden vectors frozen from the REFERENCE'S OWN classes
(tests/golden/*.npz, generator oracle/gen_golden.py)."""
import os

import numpy as np
import pytest

from oracle import c_oracle, lip_np

G = os.path.join(os.path.dirname(__file__), "golden")
FORM = {"sig_step": lip_np.SIG_STEP, "modi": lip_np.MODI, "dd": lip_np.DD}


def _load(name):
    return np.load(os.path.join(G, name), allow_pickle=True)


@pytest.mark.parametrize("form", ["sig_step", "modi", "dd"])
def test_numpy_callbacks_match_reference(form):
    g = _load(f"callbacks_{form}.npz")
    F = FORM[form]
    for b in range(len(g["xk"])):
        xk, u, goal = g["xk"][b], g["u"][b], g["goal_eff"][b]
        cir = g["cir"][b][g["sel_c"][b].astype(bool)]
        elp = g["elp"][b][g["sel_e"][b].astype(bool)]
        if form == "dd":
            lu = g["last_u"][b]
            f, gr = lip_np.dd_objective(F, xk, goal, lu, u), lip_np.dd_gradient(F, xk, goal, lu, u)
            c, J = lip_np.dd_constraints(F, xk, cir, elp, u), lip_np.dd_jacobian(F, xk, cir, elp, u)
            _, _, cl, cu = lip_np.dd_bounds(F, len(cir), len(elp))
        else:
            f, gr = lip_np.lip_objective(F, xk, goal, u), lip_np.lip_gradient(F, xk, goal, u)
            c, J = lip_np.lip_constraints(F, xk, cir, elp, u), lip_np.lip_jacobian(F, xk, cir, elp, u)
            cl, cu = lip_np.lip_bounds(F, int(g["leg"][b]), len(cir), len(elp))
        assert abs(f - g["f"][b]) <= 1e-12 * max(1.0, abs(g["f"][b]))
        np.testing.assert_allclose(gr, g["grad"][b], rtol=0, atol=1e-11)
        np.testing.assert_allclose(c, np.asarray(g["c"][b], dtype=float), rtol=0, atol=1e-12)
        np.testing.assert_allclose(J, np.asarray(g["jac"][b], dtype=float), rtol=0, atol=1e-12)
        np.testing.assert_array_equal(cl, np.asarray(g["cl"][b], dtype=float))
        np.testing.assert_array_equal(cu, np.asarray(g["cu"][b], dtype=float))


@pytest.mark.parametrize("form", ["sig_step", "modi", "dd"])
def test_c_oracle_callbacks_match_reference(form):
    """C restatement incl. goal shift, obstacle selection and bounds, against the recorded reference problem."""
    g = _load(f"callbacks_{form}.npz")
    P = c_oracle.params(form)
    for b in range(len(g["xk"])):
        xk, u = g["xk"][b], g["u"][b]
        elp = g["elp"][b] if len(g["elp"][b]) else None
        info = c_oracle.setup_info(P, xk, g["goal"][b], int(g["leg"][b]), g["cir"][b], elp)
        np.testing.assert_allclose(info["goal"], g["goal_eff"][b], rtol=0, atol=1e-12)
        assert info["nc"] == int(g["sel_c"][b].sum()) and info["ne"] == int(g["sel_e"][b].sum())
        f, gr, c, J, cl, cu = c_oracle.evaluate(P, xk, g["goal"][b], int(g["leg"][b]), g["cir"][b], elp, u, g["last_u"][b])
        assert abs(f - g["f"][b]) <= 1e-12 * max(1.0, abs(g["f"][b]))
        np.testing.assert_allclose(gr, g["grad"][b], rtol=0, atol=1e-11)
        np.testing.assert_allclose(c, np.asarray(g["c"][b], dtype=float), rtol=0, atol=1e-12)
        np.testing.assert_allclose(J, np.asarray(g["jac"][b], dtype=float), rtol=0, atol=1e-12)
        np.testing.assert_array_equal(cl, np.asarray(g["cl"][b], dtype=float))
        np.testing.assert_array_equal(cu, np.asarray(g["cu"][b], dtype=float))


def test_sig_step_warm_start_rule():
    """u0 = [x,x,x] for None, else the shifted previous plan (MPC_LIP_sig_step.py:185-189)."""
    g = _load("callbacks_sig_step.npz")
    for b in range(len(g["xk"])):
        guess = None if b % 2 == 0 else list(g["warm_in"][b].reshape(3, 5))
        np.testing.assert_array_equal(lip_np.sig_step_warm_start(g["xk"][b], guess), g["u0"][b])


@pytest.mark.parametrize("form", ["sig_step", "modi", "dd"])
def test_c_oracle_solves_match_golden(form):
    """The golden optima are KKT points verified with the reference's own callbacks (kkt_ref, viol_ref)."""
    g = _load(f"solves_{form}.npz")
    ok = g["status"] == 0
    assert np.nanmax(g["kkt_ref"][ok]) <= 1e-5 and np.nanmax(g["viol_ref"][ok]) <= 1e-6
    P = c_oracle.params(form, max_iter=500)
    n = len(g["x0"])
    elp = g["elp"] if g["elp"].shape[1] else None
    r = c_oracle.solve_batch(P, g["x0"], g["goal"], g["leg"], g["cir"], elp, g["warm"], field=np.arange(n), last_u=g["last_u"], threads=4)
    np.testing.assert_array_equal(r["status"], g["status"])
    np.testing.assert_allclose(r["u"][ok], g["u"][ok], rtol=0, atol=1e-9)


def test_config1_known_answer():
    """Reference __main__ scenario (MPC_LIP_sig_step.py:553-575): first foot placement (0.0767879, -0.1791675, 0.1963495)
    as probed in SURVEY.md 8(c)."""
    g = _load("config1_closed_loop.npz")
    np.testing.assert_allclose(g["p0"][0], [0.0767879, -0.1791675, 0.1963495], atol=2e-7)
    assert np.all(g["status"] == 0) and g["kkt_ref"].max() <= 1e-6 and g["viol_ref"].max() <= 1e-6
    P = c_oracle.params("sig_step", max_iter=500)
    state, leg, guess = g["state"][0], 1, None
    for k in range(5):
        r = c_oracle.solve(P, state, [10, 10], leg, g["cir"], None, lip_np.sig_step_warm_start(state, guess))
        np.testing.assert_allclose(r["p_plan"][0], g["p0"][k], atol=1e-8)
        guess = list(r["x_plan"])
        state, leg = r["x_plan"][0], -leg


def test_reduced_space_map():
    """u_k := x_{k+1} reproduces p under the reference's p = W(u - A x) (W B = I)."""
    m = lip_np.model()
    np.testing.assert_allclose(m.W @ m.B, np.eye(3), atol=1e-12)
    rng = np.random.default_rng(0)
    xk, z = rng.normal(size=5), rng.normal(size=9) * 0.3
    _, p = lip_np.lip_rollout(xk, lip_np.u_from_p(xk, z))
    np.testing.assert_allclose(p.ravel(), z, atol=1e-12)
    assert np.linalg.matrix_rank(m.dx_du) == 9
