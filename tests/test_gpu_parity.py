"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI (libdcbf_mpc.so) via
mujoco_lip_mpc_simulation_b200.batch.DcbfSolver; the oracle (oracle/) is only the checker.

Tolerances are the ones BASELINE.json's north_star states: foot placement 1e-4 m, heading 1e-4 rad, objective 1e-6
relative, identical feasible/infeasible class (Ipopt status 2 vs not)."""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from mujoco_lip_mpc_simulation_b200 import _lib, scenarios  # noqa: E402
from oracle import c_oracle, lip_np  # noqa: E402

G = os.path.join(os.path.dirname(__file__), "golden")
POS_TOL, OBJ_TOL = 1e-4, 1e-6


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.fail("the gpu tests need a CUDA device; there is no CPU fallback")
    from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
    return DcbfSolver


def _solver(gpu, form, sc, **kw):
    s = gpu(form, device=0, **kw)
    s.set_fields(sc.cir, sc.elp if sc.elp.shape[1] else None)
    return s


def _agreement(form, res, ref, B):
    st = res.status.cpu().numpy()
    same_class = (st == 2) == (ref["status"] == 2)
    both = (st == 0) & (ref["status"] == 0)
    if form == "dd":
        dp = np.abs(res.u.cpu().numpy() - ref["u"]).max(axis=1)
    else:
        dp = np.abs(res.p_plan.cpu().numpy() - ref["p_plan"]).reshape(B, -1).max(axis=1)
    rel = np.abs(res.obj.cpu().numpy() - ref["f"]) / np.maximum(1.0, np.abs(ref["f"]))
    return same_class, both, dp, rel


def _distinct_optima(res, ref, both, dp, rel):
    """jointly converged pairs whose plans differ AND which the data prove to be two local optima of the non-convex NLP:
    both points are feasible to 1e-6 and their objectives differ (north_star: "any mismatch explained as a distinct local
    optimum, with the fraction reported")."""
    return both & (dp > POS_TOL) & (rel > OBJ_TOL) & (res.viol.cpu().numpy() <= 1e-6) & (ref["viol"] <= 1e-6)


@pytest.mark.parametrize("form", ["sig_step", "modi", "dd"])
def test_eval_kernel_matches_reference_callbacks(gpu, form):
    """K1 against golden vectors from the reference's own LIP_Prob classes (mapped to the reduced space)."""
    g = np.load(os.path.join(G, f"callbacks_{form}.npz"), allow_pickle=True)
    n_case = len(g["xk"])
    U = lip_np.p_map()
    ran = 0
    for b in range(n_case):
        # the reference builds its rows from the SELECTED obstacles (MPC_LIP_modi.py:325-338; all of them for the other two
        # formulations); dcbf_eval emits a row per obstacle of the field, so the field handed over is the selection
        cir = g["cir"][b][g["sel_c"][b].astype(bool)]
        elp = g["elp"][b][g["sel_e"][b].astype(bool)] if len(g["elp"][b]) else np.zeros((0, 5))
        s = gpu(form, device=0)
        s.set_fields(cir if len(cir) else None, elp if len(elp) else None)
        assert s.m == len(g["c"][b])
        ran += 1
        if form == "dd":
            z = g["u"][b]
        else:
            z = lip_np.lip_rollout(g["xk"][b], g["u"][b])[1].ravel()
        r = s.evaluate(g["xk"][b], g["goal_eff"][b], [int(g["leg"][b])], z, last_u=g["last_u"][b][None], want_hess=False)
        f, gr, c, J = (r[k][0].cpu().numpy() for k in ("f", "grad", "c", "jac"))
        gref = g["grad"][b] if form == "dd" else U.T @ g["grad"][b]
        Jref = np.asarray(g["jac"][b], float) if form == "dd" else np.asarray(g["jac"][b], float) @ U
        assert abs(f - g["f"][b]) <= 1e-12 * max(1.0, abs(g["f"][b]))
        np.testing.assert_allclose(gr, gref, rtol=0, atol=1e-10)
        np.testing.assert_allclose(c, np.asarray(g["c"][b], float), rtol=0, atol=1e-12)
        np.testing.assert_allclose(J, Jref, rtol=0, atol=1e-11)
        np.testing.assert_array_equal(r["cl"][0].cpu().numpy(), np.asarray(g["cl"][b], float))
        np.testing.assert_array_equal(r["cu"][0].cpu().numpy(), np.asarray(g["cu"][b], float))
    assert ran == n_case and ran >= 20


@pytest.mark.parametrize("mode", ["warp", "thread"])
def test_obstacle_selection_and_detour_goal_match_reference(gpu, monkeypatch, mode):
    """select_obs (MPC_LIP_modi.py:325-338) and the goal shift over the selected circles (:249-271) exactly as the solve kernels
    apply them (dcbf_setup_info runs their setup code), against the masks / effective goals recorded from the reference classes;
    sig_step: every obstacle is a row and the goal shift searches all circles (MPC_LIP_sig_step.py:229-253)."""
    monkeypatch.setenv("DCBF_KERNEL", mode)
    for form in ("modi", "sig_step", "dd"):
        g = np.load(os.path.join(G, f"callbacks_{form}.npz"), allow_pickle=True)
        n_sel = n_shift = 0
        for b in range(len(g["xk"])):
            cir, elp = g["cir"][b], g["elp"][b]
            s = gpu(form, device=0)
            s.set_fields(cir if len(cir) else None, elp if len(elp) else None)
            r = s.setup_info(g["xk"][b][None, :s.nx], g["goal"][b][None])
            want = sum(int(v) << j for j, v in enumerate(list(g["sel_c"][b]) + list(g["sel_e"][b])))
            assert int(r["mask"][0]) == want, (form, b)
            assert int(r["count"][0]) == int(g["sel_c"][b].sum() + g["sel_e"][b].sum())
            np.testing.assert_allclose(r["goal_eff"][0].cpu().numpy(), g["goal_eff"][b], rtol=0, atol=1e-12)
            n_sel += int(want != (1 << (len(cir) + len(elp))) - 1)
            n_shift += int(np.abs(g["goal_eff"][b] - g["goal"][b]).max() > 1e-9)
        if form == "modi":
            assert n_sel >= 20      # the fixture exercises the selection: nearly every case drops an obstacle
        if form != "dd":
            assert n_shift >= 1     # ... and the detour
    # a batch: the mask is the definition, evaluated in numpy on the same fields
    sc = scenarios.make_batch("modi", 4096, seed=5)
    s = _solver(gpu, "modi", sc)
    r = s.setup_info(sc.x0, sc.goal, field=sc.field)
    px, py = sc.x0[:, 0:1], sc.x0[:, 1:2]
    c, e = sc.cir[sc.field], sc.elp[sc.field]
    sel_c = (px - c[:, :, 0]) ** 2 + (py - c[:, :, 1]) ** 2 - c[:, :, 2] ** 2 <= 16.0
    sel_e = (px - e[:, :, 0]) ** 2 + (py - e[:, :, 1]) ** 2 - np.maximum(e[:, :, 2], e[:, :, 3]) ** 2 <= 16.0
    bits = np.concatenate([sel_c, sel_e], axis=1)
    want = (bits * (1 << np.arange(bits.shape[1]))[None]).sum(axis=1)
    mask = r["mask"].cpu().numpy()
    # a scenario whose distance sits within rounding of the threshold may differ (d^2 - r^2 is formed from r^2 prepared on the device)
    assert np.mean(mask == want) >= 0.9999
    assert np.array_equal(r["count"].cpu().numpy(), np.array([bin(int(v)).count("1") for v in mask]))


def test_invalid_field_index_is_reported_not_dereferenced(gpu):
    """a stale / out-of-range field index must not become an out-of-bounds read: that scenario returns status -13"""
    sc = scenarios.make_batch("sig_step", 512, seed=3, n_fields=8)
    s = _solver(gpu, "sig_step", sc)
    fld = sc.field.copy()
    fld[7] = 8          # one past the last field
    fld[100] = -1
    fld[300] = 1 << 30
    res = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=fld)
    ref = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field)
    st = res.status.cpu().numpy()
    assert list(st[[7, 100, 300]]) == [-13, -13, -13]
    keep = np.ones(512, bool); keep[[7, 100, 300]] = False
    assert np.array_equal(st[keep], ref.status.cpu().numpy()[keep])
    assert torch.equal(res.u[torch.as_tensor(keep)], ref.u[torch.as_tensor(keep)])
    with pytest.raises(ValueError):
        s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field[:100])


@pytest.mark.parametrize("form", ["sig_step", "modi", "dd"])
def test_eval_hessian_matches_finite_difference_of_oracle(gpu, form):
    sc = scenarios.make_batch(form, 64, seed=31, n_fields=8)
    s = _solver(gpu, form, sc)
    rng = np.random.default_rng(1)
    n, m = s.n, s.m
    if form == "dd":
        z = sc.warm + rng.normal(size=(64, 6)) * 0.02
    else:
        z = np.tile([0.1, -0.1, 0.02], (64, 3)) * 0 + rng.normal(size=(64, 9)) * 0.05
        z[:, 0::3] += sc.x0[:, 0:1]; z[:, 1::3] += sc.x0[:, 1:2]
    lam = rng.normal(size=(64, m))
    base = s.evaluate(sc.x0, sc.goal, sc.leg, z, lam=lam, field=sc.field, last_u=sc.last_u)
    Hk = base["hess"].cpu().numpy()
    Hfd = np.zeros_like(Hk)
    for j in range(n):
        e = np.zeros((1, n)); e[0, j] = 1e-6
        rp = s.evaluate(sc.x0, sc.goal, sc.leg, z + e, lam=lam, field=sc.field, last_u=sc.last_u, want_hess=False)
        rm = s.evaluate(sc.x0, sc.goal, sc.leg, z - e, lam=lam, field=sc.field, last_u=sc.last_u, want_hess=False)
        gp = rp["grad"].cpu().numpy() + np.einsum("bmn,bm->bn", rp["jac"].cpu().numpy(), lam)
        gm = rm["grad"].cpu().numpy() + np.einsum("bmn,bm->bn", rm["jac"].cpu().numpy(), lam)
        Hfd[:, :, j] = (gp - gm) / 2e-6
    scale = np.maximum(1.0, np.abs(Hfd).max(axis=(1, 2)))[:, None, None]
    assert np.max(np.abs(Hk - Hfd) / scale) <= 5e-6


@pytest.mark.parametrize("form", ["sig_step", "modi", "dd"])
def test_solve_matches_golden(gpu, form):
    """Golden optima are KKT points certified with the reference's callbacks (oracle/gen_golden.py)."""
    g = np.load(os.path.join(G, f"solves_{form}.npz"))
    n = len(g["x0"])
    s = gpu(form, device=0, max_iter=500)
    s.set_fields(g["cir"], g["elp"] if g["elp"].shape[1] else None)
    res = s.solve(g["x0"], g["goal"], g["leg"], g["warm"], field=np.arange(n, dtype=np.int32), last_u=g["last_u"])
    ref = dict(status=g["status"], u=g["u"], p_plan=g["p_plan"], f=g["f"])
    same_class, both, dp, rel = _agreement(form, res, ref, n)
    # the fixture holds a few hundred problems: one mismatch is 0.3-0.5 %, so the bound here is "at most one", the 99.9 % bound
    # of the north_star is asserted on the full-size batches below
    print(f"{form}: {n} golden problems, class mismatches {(~same_class).sum()}, plan mismatches {(dp[both] > POS_TOL).sum()}")
    assert (~same_class).sum() <= 1
    assert (dp[both] > POS_TOL).sum() <= 1 and (rel[both] > OBJ_TOL).sum() <= 1


@pytest.mark.parametrize("form,B,seed", [("sig_step", 4096, 0), ("modi", 65536, 1), ("dd", 65536, 2)])
def test_solve_matches_oracle_on_bench_distribution(gpu, form, B, seed):
    """configs 2, 3 and 4 at their FULL sizes (4096 / 65536 / 65536) against the C oracle on the same seeded inputs, with the
    north_star's numbers: the same feasible / infeasible class on >= 99.9 %, plans within 1e-4 and objectives within 1e-6
    relative on >= 99.9 % of the jointly converged solves once the pairs this test PROVES to be two distinct local optima
    (both feasible to 1e-6, objectives differ) are set aside; their fraction is printed and bounded."""
    sc = scenarios.make_batch(form, B, seed=seed)
    s = _solver(gpu, form, sc)
    res = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field, last_u=sc.last_u)
    P = c_oracle.params(form, max_iter=200)
    ref = c_oracle.solve_batch(P, sc.x0, sc.goal, sc.leg, sc.cir, sc.elp if sc.elp.shape[1] else None, sc.warm,
                               field=sc.field, last_u=sc.last_u, threads=os.cpu_count() or 4)
    same_class, both, dp, rel = _agreement(form, res, ref, B)
    distinct = _distinct_optima(res, ref, both, dp, rel)
    rest = both & ~distinct
    print(f"{form} B={B}: class agreement {same_class.mean():.5f}; jointly converged {both.sum()}; distinct local optima "
          f"{distinct.sum()} ({distinct.sum() / max(1, both.sum()):.5f}); plans within 1e-4 {np.mean(dp[both] <= POS_TOL):.5f} raw, "
          f"{np.mean(dp[rest] <= POS_TOL):.5f} without them; objectives within 1e-6 {np.mean(rel[both] <= OBJ_TOL):.5f} raw, "
          f"{np.mean(rel[rest] <= OBJ_TOL):.5f} without them")
    assert same_class.mean() >= 0.999
    assert both.mean() >= 0.5
    assert distinct.sum() <= 0.003 * both.sum()
    assert np.mean(dp[rest] <= POS_TOL) >= 0.999
    assert np.mean(rel[rest] <= OBJ_TOL) >= 0.999


@pytest.mark.parametrize("mode", ["thread", "warp"])
@pytest.mark.parametrize("form,B,seed", [("sig_step", 2048, 21), ("modi", 2048, 22)])
def test_both_kernel_families_match_oracle(gpu, monkeypatch, mode, form, B, seed):
    """the per-thread kernels (large batches) and the warp-cooperative kernels (small batches, single solves) are
    selected by batch size; force each one (DCBF_KERNEL is read by dcbf_create) and check it against the oracle."""
    monkeypatch.setenv("DCBF_KERNEL", mode)
    sc = scenarios.make_batch(form, B, seed=seed)
    s = _solver(gpu, form, sc, max_iter=300)
    res = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field)
    P = c_oracle.params(form, max_iter=300)
    ref = c_oracle.solve_batch(P, sc.x0, sc.goal, sc.leg, sc.cir, sc.elp if sc.elp.shape[1] else None, sc.warm,
                               field=sc.field, threads=os.cpu_count() or 4)
    same_class, both, dp, rel = _agreement(form, res, ref, B)
    assert same_class.mean() >= 0.995 and np.mean(dp[both] <= POS_TOL) >= 0.995 and np.mean(rel[both] <= OBJ_TOL) >= 0.995
    # closed loop through the same kernel family
    ro = s.rollout(4, sc.x0[:256], sc.goal[:256], sc.leg[:256], field=sc.field[:256])
    one = s.solve(sc.x0[:256], sc.goal[:256], sc.leg[:256], sc.warm[:256], field=sc.field[:256])
    np.testing.assert_allclose(ro["traj"][:, 0, :5].cpu().numpy(), one.x_plan[:, 0].cpu().numpy(), atol=1e-12)


@pytest.mark.parametrize("mode", ["thread", "warp"])
def test_dd_kernel_families_match_oracle(gpu, monkeypatch, mode):
    """differential-drive formulation: the per-thread kernel and the warp kernel (wp::DdW, 6 variables, node Jacobians per
    iterate) against the oracle, and against each other on the plan"""
    B = 2048
    sc = scenarios.make_batch("dd", B, seed=23)
    P = c_oracle.params("dd", max_iter=300)
    ref = c_oracle.solve_batch(P, sc.x0, sc.goal, sc.leg, sc.cir, sc.elp, sc.warm, field=sc.field, last_u=sc.last_u,
                               threads=os.cpu_count() or 4)
    monkeypatch.setenv("DCBF_KERNEL", mode)
    s = _solver(gpu, "dd", sc, max_iter=300)
    res = s.solve(sc.x0, sc.goal, None, sc.warm, field=sc.field, last_u=sc.last_u)
    same_class, both, dp, rel = _agreement("dd", res, ref, B)
    assert same_class.mean() >= 0.995 and np.mean(dp[both] <= POS_TOL) >= 0.995 and np.mean(rel[both] <= OBJ_TOL) >= 0.995
    monkeypatch.setenv("DCBF_KERNEL", "thread" if mode == "warp" else "warp")
    s2 = _solver(gpu, "dd", sc, max_iter=300)
    other = s2.solve(sc.x0, sc.goal, None, sc.warm, field=sc.field, last_u=sc.last_u)
    st, st2 = res.status.cpu().numpy(), other.status.cpu().numpy()
    assert np.mean(st == st2) >= 0.998
    ok = (st == 0) & (st2 == 0)
    assert np.max(np.abs(res.u.cpu().numpy()[ok] - other.u.cpu().numpy()[ok])) <= POS_TOL


def test_warm_started_resolves_match_oracle(gpu):
    """config 2: "half of the batch additionally re-solved warm from the shifted solution of a first solve"
    (warm start u0 = [x_2, x_3, x_3] at the advanced state x_1, flipped leg: MPC_LIP_sig_step.py:188-189,565-575)."""
    B = 2048
    sc = scenarios.make_batch("sig_step", B, seed=0)
    s = _solver(gpu, "sig_step", sc, max_iter=300)
    first = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field)
    xp = first.x_plan.cpu().numpy()
    keep = first.status.cpu().numpy() == 0
    x1, warm = xp[:, 0], np.concatenate([xp[:, 1], xp[:, 2], xp[:, 2]], axis=1)
    res = s.solve(x1, sc.goal, -sc.leg, warm, field=sc.field)
    P = c_oracle.params("sig_step", max_iter=300)
    ref = c_oracle.solve_batch(P, x1, sc.goal, -sc.leg, sc.cir, None, warm, field=sc.field, threads=os.cpu_count() or 4)
    same_class, both, dp, rel = _agreement("sig_step", res, ref, B)
    both &= keep
    assert same_class[keep].mean() >= 0.995 and np.mean(dp[both] <= POS_TOL) >= 0.995 and np.mean(rel[both] <= OBJ_TOL) >= 0.995
    it_cold, it_warm = first.iters.cpu().numpy()[both].mean(), res.iters.cpu().numpy()[both].mean()
    print(f"mean iterations cold {it_cold:.1f} -> warm {it_warm:.1f}")


def test_warm_started_resolve_is_idempotent(gpu):
    """size-independent property at config 3's full size: re-solving from the returned plan returns the same plan."""
    B = 65536
    sc = scenarios.make_batch("modi", B, seed=1)
    s = _solver(gpu, "modi", sc)
    r1 = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field)
    r2 = s.solve(sc.x0, sc.goal, sc.leg, r1.u.reshape(B, 15), field=sc.field)
    ok = (r1.status == 0) & (r2.status == 0)
    assert ok.float().mean() > 0.5
    d = (r1.p_plan - r2.p_plan).abs().reshape(B, -1).max(dim=1).values
    assert (d[ok] <= POS_TOL).float().mean() >= 0.999
    assert ((r1.status == 2) == (r2.status == 2)).float().mean() >= 0.99


def test_returned_plans_are_feasible_and_consistent(gpu):
    """config 4's full size: every status-0 plan satisfies all rows when re-evaluated by the K1 kernel, the reported
    objective equals the re-evaluated one, and x_plan is the rollout of u."""
    B = 65536
    sc = scenarios.make_batch("dd", B, seed=2)
    s = _solver(gpu, "dd", sc)
    r = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field, last_u=sc.last_u)
    ev = s.evaluate(sc.x0, sc.goal, sc.leg, r.u, field=sc.field, last_u=sc.last_u, want_hess=False)
    ok = r.status == 0
    viol = torch.maximum((ev["cl"] - ev["c"]).clamp(min=0).max(dim=1).values, (ev["c"] - ev["cu"]).clamp(min=0).max(dim=1).values)
    assert float(viol[ok].max()) <= 1e-6
    assert float((ev["f"][ok] - r.obj[ok]).abs().max()) <= 1e-9 * float(r.obj[ok].abs().max())
    u = r.u
    lo = torch.tensor([0.4, -np.pi / 16] * 3, device=u.device) - 1e-7
    hi = torch.tensor([0.8, np.pi / 16] * 3, device=u.device) + 1e-7
    assert bool(((u[ok] >= lo) & (u[ok] <= hi)).all())
    x = torch.as_tensor(sc.x0, device=u.device).clone()
    for i in range(3):
        x = torch.stack([x[:, 0] + 0.4 * torch.cos(x[:, 2]) * u[:, 2 * i], x[:, 1] + 0.4 * torch.sin(x[:, 2]) * u[:, 2 * i],
                         x[:, 2] + u[:, 2 * i + 1]], dim=1)
        assert float((x - r.x_plan[:, i]).abs().max()) <= 1e-12


def test_determinism_and_shard_invariance(gpu):
    """bitwise: two runs agree, and solving the two halves separately equals solving the whole batch."""
    B = 8192
    sc = scenarios.make_batch("sig_step", B, seed=4)
    s = _solver(gpu, "sig_step", sc)
    a = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field)
    b = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field)
    assert torch.equal(a.u, b.u) and torch.equal(a.status, b.status) and torch.equal(a.iters, b.iters)
    h = B // 2
    lo = s.solve(sc.x0[:h], sc.goal[:h], sc.leg[:h], sc.warm[:h], field=sc.field[:h])
    hi = s.solve(sc.x0[h:], sc.goal[h:], sc.leg[h:], sc.warm[h:], field=sc.field[h:])
    assert torch.equal(torch.cat([lo.u, hi.u]), a.u) and torch.equal(torch.cat([lo.status, hi.status]), a.status)


def test_rollout_matches_oracle_closed_loop(gpu):
    """K3 against the plan -> apply -> re-plan loop of MPC_LIP_sig_step.py:565-575 driven by the oracle."""
    g = np.load(os.path.join(G, "config1_closed_loop.npz"))
    sc = scenarios.config1()
    s = _solver(gpu, "sig_step", sc)
    r = s.rollout(5, sc.x0, sc.goal, sc.leg)
    traj = r["traj"][0].cpu().numpy()
    np.testing.assert_allclose(traj[:, 5:7], g["p0"][:, :2], atol=POS_TOL)
    np.testing.assert_allclose(traj[:, :5], g["x_plan"][:, 0, :], atol=POS_TOL)
    # random scenarios, 8 steps
    B, steps = 64, 8
    sc = scenarios.make_batch("sig_step", B, seed=9, n_fields=16)
    s = _solver(gpu, "sig_step", sc)
    r = s.rollout(steps, sc.x0, sc.goal, sc.leg, field=sc.field)
    traj = r["traj"].cpu().numpy()
    P = c_oracle.params("sig_step", max_iter=300)
    agree = total = 0
    for b in range(B):
        state, leg, guess = sc.x0[b].copy(), int(sc.leg[b]), None
        for k in range(steps):
            o = c_oracle.solve(P, state, sc.goal[b], leg, sc.cir[sc.field[b]], None, lip_np.sig_step_warm_start(state, guess))
            if o["status"] != 0 or traj[b, k, 7] != 0:
                break   # after a non-converged / infeasible re-plan the two chains may legitimately part
            total += 1
            same = np.abs(traj[b, k, :5] - o["x_plan"][0]).max() <= POS_TOL
            agree += int(same)
            if o["close2goal"] or not same:
                break   # a different local optimum at one step makes every later step a different problem
            guess = list(o["x_plan"])
            state, leg = o["x_plan"][0].copy(), -leg
    assert total > 100 and agree / total >= 0.99


def test_host_buffer_entry_point_equals_device_entry_point(gpu):
    sc = scenarios.make_batch("modi", 1000, seed=6)
    s = _solver(gpu, "modi", sc)
    d = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field)
    s.set_fields_host(sc.cir, sc.elp)
    h = s.solve_host(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field)
    np.testing.assert_array_equal(h.u, d.u.cpu().numpy())
    np.testing.assert_array_equal(h.status, d.status.cpu().numpy())
    np.testing.assert_array_equal(h.p_plan, d.p_plan.cpu().numpy())
    np.testing.assert_array_equal(h.close2goal.astype(bool), d.close2goal.cpu().numpy())


def test_size_class_split_is_transparent(gpu, monkeypatch):
    """modi batches from 5 120 scenarios on are split by selected-obstacle count (classify pre-pass; one-slot kernel, typed
    turn-row slot kernel wp::LipL and generic two-slot kernel on forked streams).  With two classes (DCBF_LIPL=0) every scenario
    gets the result of the unsplit launch bit for bit; the LipL class adds the turn rows to the condensed system in closed form
    instead of through the dot products, i.e. in a different summation order: same status everywhere, same plans to rounding"""
    B = 16384
    sc = scenarios.make_batch("modi", B, seed=41)
    monkeypatch.setenv("DCBF_SPLIT", "0")
    ref = _solver(gpu, "modi", sc).solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field)
    monkeypatch.setenv("DCBF_SPLIT", "16384")
    monkeypatch.setenv("DCBF_LIPL", "0")
    s = _solver(gpu, "modi", sc)
    l0 = s.launches
    res = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field)
    torch.cuda.synchronize()
    assert s.launches - l0 == 3                      # classify + two solve kernels
    assert torch.equal(res.status, ref.status) and torch.equal(res.iters, ref.iters)
    assert torch.equal(res.p_plan, ref.p_plan) and torch.equal(res.x_plan, ref.x_plan)
    monkeypatch.delenv("DCBF_LIPL")
    s = _solver(gpu, "modi", sc)
    l0 = s.launches
    res = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field)
    torch.cuda.synchronize()
    assert s.launches - l0 == 4                      # classify + three solve kernels
    assert torch.equal(res.status, ref.status)
    assert (res.iters == ref.iters).double().mean().item() >= 0.99
    both = (res.status == 0) & (ref.status == 0)
    dp = (res.p_plan - ref.p_plan).abs().reshape(B, -1).max(dim=1).values[both]
    assert dp.max().item() <= 1e-4 and (dp <= 1e-9).double().mean().item() >= 0.97, (dp.max().item(), (dp <= 1e-9).double().mean().item())
    # the classes the typed slot does not touch are still bit-identical
    other = s.setup_info(sc.x0, sc.goal, field=sc.field)["count"] != 5
    assert 0.1 < (~other).double().mean().item() < 0.4           # (a quarter of the batch selects five obstacles)
    assert torch.equal(res.p_plan[other], ref.p_plan[other]) and torch.equal(res.iters[other], ref.iters[other])


def test_size_class_split_with_empty_classes(gpu, monkeypatch):
    """the three-class split when a class is empty: every obstacle out of detection range (all scenarios in the one-slot class) and
    a batch made of five-obstacle scenarios only (typed turn-row slot class alone); same results as the unsplit launch"""
    B = 8192
    sc = scenarios.make_batch("modi", B, seed=44)
    far_c, far_e = sc.cir.copy(), sc.elp.copy()
    far_c[:, :, :2] += 100.0
    far_e[:, :, :2] += 100.0
    for cir, elp, idx in ((far_c, far_e, np.arange(B)), (sc.cir, sc.elp, None)):
        monkeypatch.setenv("DCBF_SPLIT", "0")
        s0 = gpu("modi", device=0)
        s0.set_fields(cir, elp)
        if idx is None:   # keep the scenarios that select exactly five obstacles
            cnt = s0.setup_info(sc.x0, sc.goal, field=sc.field)["count"].cpu().numpy()
            idx = np.nonzero(cnt == 5)[0]
            assert len(idx) > 1000
        a = (sc.x0[idx], sc.goal[idx], sc.leg[idx], sc.warm[idx])
        ref = s0.solve(*a, field=sc.field[idx])
        monkeypatch.setenv("DCBF_SPLIT", "1")
        s1 = gpu("modi", device=0)
        s1.set_fields(cir, elp)
        res = s1.solve(*a, field=sc.field[idx])
        torch.cuda.synchronize()
        assert torch.equal(res.status, ref.status)
        both = (res.status == 0) & (ref.status == 0)
        dp = (res.p_plan - ref.p_plan).abs().reshape(len(idx), -1).max(dim=1).values[both]
        assert both.any() and dp.max().item() <= 1e-4 and (dp <= 1e-9).double().mean().item() >= 0.97


@pytest.mark.parametrize("form", ["sig_step", "dd"])
def test_scheduling_order_is_transparent(gpu, monkeypatch, form):
    """batches of 2048+ scenarios are started in the order of their predicted clearance (hard problems first, so that none of them
    is left for the tail: sched_classify_kernel / sched_scatter_kernel); the order decides which warp solves which scenario and
    must not change any result"""
    B = 4096
    sc = scenarios.make_batch(form, B, seed=43)
    monkeypatch.setenv("DCBF_ORDER", "0")
    ref = _solver(gpu, form, sc).solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field, last_u=sc.last_u)
    monkeypatch.delenv("DCBF_ORDER")
    s = _solver(gpu, form, sc)
    l0 = s.launches
    res = s.solve(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field, last_u=sc.last_u)
    torch.cuda.synchronize()
    assert s.launches - l0 == 3                      # classify + scatter + solve
    assert torch.equal(res.status, ref.status) and torch.equal(res.iters, ref.iters)
    assert torch.equal(res.u, ref.u) and torch.equal(res.x_plan, ref.x_plan) and torch.equal(res.obj, ref.obj)
    small = s.solve(sc.x0[:512], sc.goal[:512], sc.leg[:512], sc.warm[:512], field=sc.field[:512], last_u=None if sc.last_u is None else sc.last_u[:512])
    assert torch.equal(small.u, ref.u[:512])          # below the threshold: natural order, same numbers


def test_host_entry_point_with_page_locked_buffers(gpu):
    """dcbf_solve_host copies straight from / to page-locked caller buffers (no staging); same results as the staged path"""
    sc = scenarios.make_batch("sig_step", 512, seed=31)
    s = _solver(gpu, "sig_step", sc)
    s.set_fields_host(sc.cir)
    ref = s.solve_host(sc.x0, sc.goal, sc.leg, sc.warm, field=sc.field)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    from mujoco_lip_mpc_simulation_b200.batch import SolveResult
    out = SolveResult(pin(np.empty((512, 15))), pin(np.empty((512, 3, 5))), pin(np.empty((512, 3, 3))), pin(np.empty(512, np.int32)),
                      pin(np.empty(512, np.int32)), pin(np.empty(512)), pin(np.empty(512)), pin(np.empty(512, np.uint8)))
    s.solve_host(pin(sc.x0), pin(sc.goal), pin(sc.leg.astype(np.int32)), pin(sc.warm), field=pin(sc.field.astype(np.int32)), out=out)
    assert np.array_equal(out.status, ref.status) and np.array_equal(out.iters, ref.iters)
    assert np.array_equal(out.p_plan, ref.p_plan) and np.array_equal(out.x_plan, ref.x_plan) and np.array_equal(out.u, ref.u)
    assert np.array_equal(out.close2goal, ref.close2goal)


def test_host_pipeline_delivers_the_results_of_the_synchronous_call(gpu):
    """dcbf_solve_host_async / dcbf_wait: consecutive host-buffer batches on three contexts used round-robin; every batch gets the
    result of the synchronous call bit for bit, pageable buffers are refused"""
    from mujoco_lip_mpc_simulation_b200.batch import HostPipeline
    B = 512
    batches = [scenarios.make_batch("sig_step", B, seed=60 + j) for j in range(5)]
    F = batches[0].cir.shape[0]
    cir_all = np.concatenate([b.cir for b in batches], axis=0)
    s = _solver(gpu, "sig_step", batches[0])
    s.set_fields_host(cir_all)
    refs = [s.solve_host(b.x0, b.goal, b.leg, b.warm, field=b.field + j * F) for j, b in enumerate(batches)]
    pipe = HostPipeline("sig_step", lanes=3, device=0)
    pipe.set_fields_host(cir_all)
    pin = HostPipeline.pin
    outs = []
    for j, b in enumerate(batches):
        outs.append(pipe.submit(pin(b.x0), pin(b.goal), pin(b.leg.astype(np.int32)), pin(b.warm), field=pin((b.field + j * F).astype(np.int32))))
    pipe.drain()
    for (out, _), ref in zip(outs, refs):
        assert np.array_equal(out.status, ref.status) and np.array_equal(out.iters, ref.iters)
        assert np.array_equal(out.p_plan, ref.p_plan) and np.array_equal(out.x_plan, ref.x_plan)
    b = batches[0]
    with pytest.raises(RuntimeError):
        pipe.solvers[0].solve_host(b.x0, b.goal, b.leg, b.warm, field=b.field, wait=False)     # pageable buffers
    small, _ = pipe.submit(pin(b.x0[:7]), pin(b.goal[:7]), pin(b.leg[:7].astype(np.int32)), pin(b.warm[:7]), field=pin(b.field[:7].astype(np.int32)))
    pipe.drain()                                                                                # a tiny batch takes the same path
    assert np.array_equal(small.status, refs[0].status[:7]) and np.array_equal(small.p_plan, refs[0].p_plan[:7])
    pipe.solvers[0].wait()                                                                      # nothing pending: returns at once
    # differential drive through the same entry points (previous control, no p_plan)
    sd = scenarios.make_batch("dd", 300, seed=66)
    ref = _solver(gpu, "dd", sd)
    ref.set_fields_host(sd.cir, sd.elp)
    rd = ref.solve_host(sd.x0, sd.goal, sd.leg, sd.warm, field=sd.field, last_u=sd.last_u)
    pd = HostPipeline("dd", lanes=2, device=0)
    pd.set_fields_host(sd.cir, sd.elp)
    od, _ = pd.submit(pin(sd.x0), pin(sd.goal), pin(sd.leg.astype(np.int32)), pin(sd.warm), field=pin(sd.field.astype(np.int32)), last_u=pin(sd.last_u))
    pd.drain()
    assert od.p_plan is None and np.array_equal(od.status, rd.status) and np.array_equal(od.u, rd.u) and np.array_equal(od.x_plan, rd.x_plan)


def test_error_codes(gpu):
    lib = _lib.load()
    P = _lib.DcbfParams()
    lib.dcbf_default_params(0, C.byref(P))
    ctx = C.c_void_p()
    assert lib.dcbf_create(C.byref(P), 0, C.byref(ctx)) == 0
    x = torch.zeros((1, 15), dtype=torch.float64, device="cuda")
    args = [x.data_ptr(), x.data_ptr(), None, None, x.data_ptr(), None] + [None] * 8   # x0, goal, leg, field, warm, last_u, outputs
    assert lib.dcbf_solve(ctx, 1, *args[:14], None) == -3          # DCBF_ERR_NO_FIELDS
    assert lib.dcbf_solve(ctx, 1, None, *args[1:14], None) == -1    # DCBF_ERR_ARG
    assert lib.dcbf_set_fields(ctx, 1, 99, x.data_ptr(), 0, None, None) == -1
    assert lib.dcbf_create(C.byref(P), 12345, C.byref(C.c_void_p())) == -2
    lib.dcbf_destroy(ctx)


@pytest.mark.parametrize("mode", ["warp", "thread"])
def test_cold_start_rule_without_a_start_vector(gpu, monkeypatch, mode):
    """warm = NULL is the reference's init_guess = None, u0 = [x_k, x_k, x_k] (MPC_LIP_sig_step.py:185-187), formed on the device:
    bit for bit the result of passing that vector, through the device entry point, the host-buffer entry point (pageable and
    page-locked buffers) and both kernel families; the differential drive has no such rule and rejects the call"""
    monkeypatch.setenv("DCBF_KERNEL", mode)
    sc = scenarios.make_batch("sig_step", 3000, seed=31)
    s = _solver(gpu, "sig_step", sc)
    ref = s.solve(sc.x0, sc.goal, sc.leg, np.tile(sc.x0, (1, 3)), field=sc.field)
    res = s.solve(sc.x0, sc.goal, sc.leg, None, field=sc.field)
    for a, b in ((ref.u, res.u), (ref.p_plan, res.p_plan), (ref.status, res.status), (ref.iters, res.iters)):
        assert torch.equal(a, b)
    small = s.solve(sc.x0[:100], sc.goal[:100], sc.leg[:100], None, field=sc.field[:100])   # (no start-order pass in front of the kernel)
    assert torch.equal(small.u, ref.u[:100]) and torch.equal(small.iters, ref.iters[:100])
    h = s.solve_host(sc.x0, sc.goal, sc.leg, None, field=sc.field)
    np.testing.assert_array_equal(h.u, ref.u.cpu().numpy())
    np.testing.assert_array_equal(h.iters, ref.iters.cpu().numpy())
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory().numpy()   # noqa: E731
    pe = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()           # noqa: E731
    B = 3000
    from mujoco_lip_mpc_simulation_b200.batch import SolveResult
    out = SolveResult(pe((B, 15), torch.float64), pe((B, 3, 5), torch.float64), pe((B, 3, 3), torch.float64), pe((B,), torch.int32),
                      pe((B,), torch.int32), pe((B,), torch.float64), pe((B,), torch.float64), pe((B,), torch.uint8))
    s.solve_host(pin(sc.x0), pin(sc.goal), pin(sc.leg), None, field=pin(sc.field), out=out)
    np.testing.assert_array_equal(out.u, ref.u.cpu().numpy())
    np.testing.assert_array_equal(out.status, ref.status.cpu().numpy())
    scd = scenarios.make_batch("dd", 64, seed=32)
    sd = _solver(gpu, "dd", scd)
    with pytest.raises(ValueError):
        sd.solve(scd.x0, scd.goal, None, None, field=scd.field, last_u=scd.last_u)
    lib = _lib.load()
    x = torch.zeros((64, 3), dtype=torch.float64, device="cuda")
    assert lib.dcbf_solve(sd._ctx, 64, x.data_ptr(), x.data_ptr(), None, None, None, None, *([None] * 8), None) == -1   # DCBF_ERR_ARG


def test_empty_and_single_batches(gpu):
    sc = scenarios.config1()
    s = _solver(gpu, "sig_step", sc)
    r0 = s.solve(np.zeros((0, 5)), np.zeros((0, 2)), np.zeros(0, np.int32), np.zeros((0, 15)))
    assert r0.u.shape == (0, 15)
    r1 = s.solve(sc.x0, sc.goal, sc.leg, sc.warm)
    np.testing.assert_allclose(r1.p_plan[0, 0].cpu().numpy(), [0.0767879, -0.1791675, 0.1963495], atol=1e-6)
    assert int(r1.status[0]) == 0
