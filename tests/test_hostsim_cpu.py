"""CPU tests of the PRODUCT's per-lane solver code (csrc/dcbf_core.cuh, dcbf_lanes.cuh) compiled for the host by
tests/hostsim -- a debugging build used only here; the shipped library has no CPU path.  The same assertions run
against the real CUDA kernels in test_gpu_parity.py."""
import os

import numpy as np
import pytest

import hostsim_binding as H
from mujoco_lip_mpc_simulation_b200 import scenarios
from oracle import c_oracle, lip_np

G = os.path.join(os.path.dirname(__file__), "golden")


def _ref_eval_in_p(form, g, b, z, lam):
    """oracle callbacks mapped to the reduced space: grad_z = U^T grad_u, jac_z = jac_u U."""
    P = c_oracle.params(form, goal_shift=0, select_obs=0)
    elp = g["elp"][b] if len(g["elp"][b]) else None
    if form == "dd":
        return c_oracle.evaluate(P, g["xk"][b], g["goal"][b], 1, g["cir"][b], elp, z, g["last_u"][b])
    U = lip_np.p_map()
    f, gr, c, J, cl, cu = c_oracle.evaluate(P, g["xk"][b], g["goal"][b], int(g["leg"][b]), g["cir"][b], elp, lip_np.u_from_p(g["xk"][b], z))
    return f, U.T @ gr, c, J @ U, cl, cu


@pytest.mark.parametrize("form", ["sig_step", "modi", "dd"])
def test_eval_matches_oracle(form):
    g = np.load(os.path.join(G, f"callbacks_{form}.npz"), allow_pickle=True)
    rng = np.random.default_rng(5)
    P = H.default_params(form)
    for b in range(0, len(g["xk"]), 3):
        elp = g["elp"][b] if len(g["elp"][b]) else None
        if form == "dd":
            z = g["u"][b]
        else:
            _, p = lip_np.lip_rollout(g["xk"][b], g["u"][b])
            z = p.ravel()
        K = len(g["cir"][b]) + len(g["elp"][b])
        m = 3 * (K + 1) if form == "dd" else 3 * (4 + K + (1 if form == "modi" else 0))
        lam = rng.normal(size=(1, m))
        r = H.evaluate(P, g["xk"][b], g["goal"][b], int(g["leg"][b]), g["cir"][b], elp, z, lam=lam, last_u=g["last_u"][b][None])
        f, gr, c, J, cl, cu = _ref_eval_in_p(form, g, b, z, lam)
        assert abs(r["f"][0] - f) <= 1e-12 * max(1.0, abs(f))
        np.testing.assert_allclose(r["grad"][0], gr, rtol=0, atol=1e-10)
        np.testing.assert_allclose(r["c"][0], c, rtol=0, atol=1e-12)
        np.testing.assert_allclose(r["jac"][0], J, rtol=0, atol=1e-11)
        np.testing.assert_array_equal(r["cl"][0], cl)
        np.testing.assert_array_equal(r["cu"][0], cu)
        # Lagrangian Hessian against a central difference of the oracle's first derivatives
        n = len(z)
        Hfd = np.zeros((n, n))
        for j in range(n):
            e = np.zeros(n); e[j] = 1e-6
            _, gp, _, Jp, _, _ = _ref_eval_in_p(form, g, b, z + e, lam)
            _, gm, _, Jm, _, _ = _ref_eval_in_p(form, g, b, z - e, lam)
            Hfd[:, j] = ((gp + Jp.T @ lam[0]) - (gm + Jm.T @ lam[0])) / 2e-6
        np.testing.assert_allclose(r["hess"][0], Hfd, rtol=0, atol=2e-6 * max(1.0, np.abs(Hfd).max()))


@pytest.mark.parametrize("form", ["sig_step", "modi", "dd"])
def test_solve_matches_golden(form):
    g = np.load(os.path.join(G, f"solves_{form}.npz"))
    n = len(g["x0"])
    P = H.default_params(form)
    P.max_iter = 500
    elp = g["elp"] if g["elp"].shape[1] else None
    r = H.solve(P, g["x0"], g["goal"], g["leg"], g["cir"], elp, g["warm"], field=np.arange(n), last_u=g["last_u"])
    assert np.mean((r["status"] == 2) == (g["status"] == 2)) >= 0.98
    both = (r["status"] == 0) & (g["status"] == 0)
    if form == "dd":
        dp = np.abs(r["u"] - g["u"]).max(axis=1)
    else:
        dp = np.abs(r["p_plan"] - g["p_plan"]).reshape(n, -1).max(axis=1)
    rel = np.abs(r["f"] - g["f"]) / np.maximum(1.0, np.abs(g["f"]))
    # tolerances of BASELINE.json north_star: 1e-4 m / 1e-4 rad, objective 1e-6 relative
    assert np.mean(dp[both] <= 1e-4) >= 0.98
    assert np.mean(rel[both] <= 1e-6) >= 0.98


def test_config1_closed_loop():
    g = np.load(os.path.join(G, "config1_closed_loop.npz"))
    sc = scenarios.config1()
    r = H.rollout(H.default_params("sig_step"), 5, sc.x0, sc.goal, sc.leg, sc.cir, None)
    np.testing.assert_allclose(r["traj"][0][:, 5:7], g["p0"][:, :2], atol=1e-5)
    np.testing.assert_allclose(r["traj"][0][:, :5], g["x_plan"][:, 0, :], atol=1e-5)
    assert np.all(r["traj"][0][:, 7] == 0)


def test_edge_cases():
    """no obstacles at all; every obstacle out of detection range (modi); one-scenario batch."""
    x0 = np.array([[0.0, 0.0, 0.6, -0.3, 0.0]])
    P = H.default_params("sig_step")
    r = H.solve(P, x0, [10, 10], [1], np.zeros((1, 0, 3)), None, np.tile(x0, (1, 3)))
    o = c_oracle.solve(c_oracle.params("sig_step", max_iter=300), x0[0], [10, 10], 1, None, None, np.tile(x0[0], 3))
    assert r["status"][0] == 0 and o["status"] == 0
    np.testing.assert_allclose(r["p_plan"][0], o["p_plan"], atol=1e-5)
    Pm = H.default_params("modi")
    far = np.array([[[30.0, 30.0, 1.0]]])
    r2 = H.solve(Pm, x0, [10, 10], [1], far, np.array([[[40.0, 40.0, 1.0, 0.5, 0.3]]]), np.tile(x0, (1, 3)))
    o2 = c_oracle.solve(c_oracle.params("modi", max_iter=300), x0[0], [10, 10], 1, far[0], np.array([[40.0, 40.0, 1.0, 0.5, 0.3]]), np.tile(x0[0], 3))
    assert r2["status"][0] == o2["status"] == 0
    np.testing.assert_allclose(r2["p_plan"][0], o2["p_plan"], atol=1e-5)
