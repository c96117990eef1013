"""GPU tests of the control-tick path (dcbf_tick): the prediction, the warm-start rule and the dense plan trajectory are
checked against the host restatements that tests/test_helpers_cpu.py pins to golden vectors of the reference classes
(_lipmodel.flow_matrices / track_det <- MPCCBF.get_next_states / xk_track_det), and the re-plan inside the tick must be the
very solve dcbf_solve performs on the same inputs."""
import math

import numpy as np
import pytest
import torch

from mujoco_lip_mpc_simulation_b200 import _lipmodel, scenarios
from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver

pytestmark = pytest.mark.gpu


def _state_before_step_end(sc, t_rest, rng):
    """current (pos, vel, hd, stance foot) such that the LIP flow over t_rest is well defined: take the scenario state as the
    present and a stance foot near the body"""
    B = len(sc.x0)
    pos, vel, hd = sc.x0[:, 0:2].copy(), sc.x0[:, 2:4].copy(), sc.x0[:, 4].copy()
    gp = np.concatenate([pos + rng.normal(scale=0.05, size=(B, 2)), rng.uniform(-0.1, 0.1, size=(B, 1))], axis=1)
    return pos, vel, hd, gp


@pytest.mark.parametrize("form", ["sig_step", "modi"])
def test_tick_matches_host_restatement_and_plain_solve(form):
    rng = np.random.default_rng(5)
    B = 192
    sc = scenarios.make_batch(form, B, seed=21)
    s = DcbfSolver(form, device=0)
    s.set_fields(sc.cir, sc.elp if sc.elp.shape[1] else None)
    t_rest = rng.uniform(0.0, 0.4, size=B)
    t_rest[:4] = [0.0, 0.4, 0.25, 0.01]
    pos, vel, hd, gp = _state_before_step_end(sc, t_rest, rng)
    prev = rng.normal(size=(B, 15)) * 0.1 + np.tile(sc.x0, (1, 3))
    mode = rng.integers(0, 3, size=B).astype(np.uint8)
    out = s.tick(pos, vel, hd, gp, t_rest, sc.goal, -sc.leg, prev_plan=prev, mode=mode, field=sc.field)
    torch.cuda.synchronize()
    xn = out["x_next"].cpu().numpy()
    warm = out["warm"].cpu().numpy()
    # prediction: A(t) x + B(t) p with the heading gain t / dt
    for b in range(B):
        A, Bm = _lipmodel.flow_matrices(float(t_rest[b]), float(t_rest[b]) * (1.0 / 0.4))
        ref = A @ np.concatenate([pos[b], vel[b], [hd[b]]]) + Bm @ gp[b]
        np.testing.assert_allclose(xn[b], ref, rtol=0, atol=2e-13)
    # warm-start rule of data_procs/logger_mpc.py:326-333
    for b in range(B):
        p3 = prev[b].reshape(3, 5)
        want = {0: p3, 1: np.stack([p3[1], p3[2], p3[2]]), 2: np.stack([xn[b]] * 3)}[int(mode[b])]
        np.testing.assert_array_equal(warm[b].reshape(3, 5), want)
    # the re-plan is the plain solve on (x_next, warm): bit for bit where the start is cold (mode 2); a warm-started scenario begins
    # at a lower barrier parameter (dcbf_params mu_warm / mu_shift), i.e. other iterates towards the same optimum
    ref = s.solve(xn, sc.goal, -sc.leg, warm, field=sc.field)
    torch.cuda.synchronize()
    cold = torch.as_tensor(mode == 2, device=ref.status.device)
    pl = out["plan"]
    assert torch.equal(pl.status[cold], ref.status[cold]) and torch.equal(pl.iters[cold], ref.iters[cold])
    assert torch.equal(pl.p_plan[cold], ref.p_plan[cold]) and torch.equal(pl.x_plan[cold], ref.x_plan[cold])
    both = (pl.status == 0) & (ref.status == 0) & ~cold
    assert float(((pl.status == 2) == (ref.status == 2)).float().mean()) >= 0.98 and int(both.sum()) >= 40
    d = (pl.p_plan - ref.p_plan).abs().reshape(B, -1).max(dim=1).values
    assert float((d[both] <= 1e-4).float().mean()) >= 0.95      # `prev` is noise around x0: a few land in another local optimum
    # dense plan trajectory: three segments of 1 + 41 samples (MPC_LIP_modi.py:117-122)
    pd = out["pos_det"].cpu().numpy()
    xp, pp = pl.x_plan.cpu().numpy(), pl.p_plan.cpu().numpy()
    for b in range(0, B, 7):
        starts = [xn[b], xp[b, 0], xp[b, 1]]
        want = np.concatenate([_lipmodel.track_det(starts[j], pp[b, j], 0.4) for j in range(3)])
        assert want.shape == (126, 2)
        np.testing.assert_allclose(pd[b], want, rtol=0, atol=5e-13)


def test_tick_without_previous_plan_is_a_cold_start():
    sc = scenarios.make_batch("modi", 64, seed=22)
    s = DcbfSolver("modi", device=0)
    s.set_fields(sc.cir, sc.elp)
    z2, z1 = np.zeros((64, 2)), np.zeros(64)
    # t_rest = 0: the prediction is the state itself and the cold start is [x, x, x] -> same as the plain cold solve
    out = s.tick(sc.x0[:, 0:2], sc.x0[:, 2:4], sc.x0[:, 4], np.zeros((64, 3)), z1, sc.goal, sc.leg, field=sc.field, want_pos_det=False)
    ref = s.solve(sc.x0, sc.goal, sc.leg, np.tile(sc.x0, (1, 3)), field=sc.field)
    torch.cuda.synchronize()
    assert out["pos_det"] is None
    np.testing.assert_allclose(out["x_next"].cpu().numpy(), sc.x0, rtol=0, atol=1e-15)
    assert torch.equal(out["plan"].status, ref.status)
    assert torch.allclose(out["plan"].p_plan, ref.p_plan, rtol=0, atol=1e-9)
    del z2


def test_alip_foot_placement_chained_behind_the_dd_solve():
    """dcbf_alip_foot against the host mirror of ALIP_plan/planner.py (pinned to the reference by tests/test_helpers_cpu.py),
    reading the forward speed straight from the DD plan (speed_stride = 6)"""
    from mujoco_lip_mpc_simulation_b200.ALIP_plan.planner import ALIP, ALIPParam
    B = 512
    rng = np.random.default_rng(9)
    sc = scenarios.make_batch("dd", B, seed=24)
    s = DcbfSolver("dd", device=0)
    s.set_fields(sc.cir, sc.elp)
    plan = s.solve(sc.x0, sc.goal, None, sc.warm, field=sc.field, last_u=sc.last_u)
    x_alip = np.stack([rng.uniform(-0.1, 0.2, B), rng.uniform(5.0, 40.0, B)], axis=1)
    y_alip = np.stack([rng.uniform(-0.2, 0.2, B), rng.uniform(-15.0, 15.0, B)], axis=1)
    time = rng.uniform(0.0, 0.45, B)          # beyond T the reference clamps
    sup = rng.choice([-1, 1], size=B).astype(np.int32)
    out = s.alip_foot(x_alip, y_alip, time, sup, plan.u, speed_stride=6)
    torch.cuda.synchronize()
    a = ALIP(ALIPParam(H=1.0, T=0.4, m=45.0))
    speed = plan.u.cpu().numpy()[:, 0]
    ux, uy, Ly, Lx = a.getFootPlacement(speed, sup, time, x_alip, y_alip)
    np.testing.assert_allclose(out["foot"].cpu().numpy(), np.stack([ux, uy], axis=1), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(out["am"].cpu().numpy(), np.stack([Ly, Lx], axis=1), rtol=1e-12, atol=1e-10)
    xt, yt = a.getTimedState(x_alip, y_alip, 0.4 - np.minimum(time, 0.4))
    np.testing.assert_allclose(out["next"].cpu().numpy(), np.concatenate([xt, yt], axis=1), rtol=1e-12, atol=1e-10)
    # the regulation clamps are hit on both sides
    uyg = out["foot"].cpu().numpy()[:, 1]
    assert np.all((np.abs(uyg) >= 0.1 - 1e-15) & (np.abs(uyg) <= 0.45 + 1e-15)) and np.all(np.sign(uyg) == sup)


def test_lean_elementary_functions_on_the_device():
    """csrc/dcbf_math.cuh as compiled for sm_100a (rcp / rsqrt seeds + Newton, Cody-Waite sincos, one-division atan2) against
    numpy: <= 1-2 ulp on the argument ranges of the solver"""
    from mujoco_lip_mpc_simulation_b200.batch import _ptr
    rng = np.random.default_rng(3)
    n = 200000
    a = np.concatenate([rng.uniform(-7, 7, n // 2), rng.uniform(-300, 300, n // 2)])
    b = np.concatenate([rng.uniform(-12, 12, n // 2), 10.0 ** rng.uniform(-12, 3, n // 2) * rng.choice([-1, 1], n // 2)])
    s = DcbfSolver("sig_step", device=0)
    ta, tb = torch.as_tensor(a, device="cuda"), torch.as_tensor(b, device="cuda")
    out = torch.empty((n, 7), dtype=torch.float64, device="cuda")
    assert s.lib.dcbf_math_probe(s._ctx, n, _ptr(ta), _ptr(tb), _ptr(out), s._stream()) == 0
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    assert np.max(np.abs(o[:, 0] - np.sin(a))) <= 3e-16 and np.max(np.abs(o[:, 1] - np.cos(a))) <= 3e-16
    assert np.max(np.abs(o[:, 2] - np.arctan2(a, b))) <= 5e-16
    assert np.max(np.abs(o[:, 3] * b - 1.0)) <= 3e-16
    assert np.max(np.abs(o[:, 4] - a / b) / np.abs(a / b)) <= 3e-16
    assert np.max(np.abs(o[:, 5] * np.sqrt(np.abs(b)) - 1.0)) <= 5e-16
    lg = np.log(np.abs(b))
    assert np.max(np.abs(o[:, 6] - lg) / np.spacing(np.abs(lg))) <= 1.0                        # flog: within one ulp of libm


def test_eval_kernel_reproduces_recorded_labels():
    """the D-CBF rows of the evaluation kernel (dcbf_eval, modi) at the plans the reference RECORDED (tests/golden/data_log_plans.npz,
    see tests/test_data_log_cpu.py): plans filed under pred_fail (Ipopt status 2) violate a D-CBF row, the others do not"""
    import os
    from mujoco_lip_mpc_simulation_b200 import data_log
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "data_log_plans.npz"))
    cir = G["cir"] + np.array([0.0, 0.0, 0.4])
    elp = G["elp"] + np.array([0.0, 0.0, 0.4, 0.4, 0.0])
    n = len(G["plan"])
    x0, z = np.zeros((n, 5)), np.zeros((n, 9))
    for b, a in enumerate(G["plan"]):
        x, v, p, _ = data_log.plan_from_pos_det(a)
        x0[b, 0:2], x0[b, 2:4] = x[0], v[0]
        z[b, 0::3], z[b, 1::3] = p[:, 0], p[:, 1]
    s = DcbfSolver("modi", device=0)
    s.set_fields(cir, elp)
    ev = s.evaluate(x0, np.tile([10.0, 10.0], (n, 1)), np.ones(n, np.int32), z, field=G["run"].astype(np.int32), want_hess=False)
    c = ev["c"].cpu().numpy().reshape(n, 3, 5 + 8)          # per step: v_bx, v_by, 4 circles, 4 ellipses, leg, turn, coupling
    worst = c[:, :, 2:10].reshape(n, -1).min(axis=1)
    fail, feasi = G["label"] == 2, G["label"] == 0
    assert np.mean(worst[fail] < -1e-4) >= 0.9 and np.mean(worst[feasi] >= -1e-4) >= 0.95
    # and the leg-length row holds to Ipopt's bound relaxation on every recorded plan
    assert np.max(c[:, :, 10]) <= 0.09 + 1e-6


def test_dd_eval_kernel_reproduces_recorded_labels():
    """same for the differential-drive runs (data_log/DD_me*.pkl): recorded states -> controls -> dcbf_eval (dd) rows"""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "data_log_plans.npz"))
    P = G["dd_plan"]
    n = len(P)
    d = np.diff(P, axis=1)
    v = (d[:, :, 0] * np.cos(P[:, :3, 2]) + d[:, :, 1] * np.sin(P[:, :3, 2])) / 0.4
    z = np.stack([v, d[:, :, 2]], axis=2).reshape(n, 6)
    s = DcbfSolver("dd", device=0)
    s.set_fields(G["dd_cir"] + np.array([0.0, 0.0, 0.4]), G["dd_elp"] + np.array([0.0, 0.0, 0.4, 0.4, 0.0]))
    ev = s.evaluate(P[:, 0], np.tile([10.0, 10.0], (n, 1)), None, z, field=G["dd_run"].astype(np.int32), last_u=np.zeros((n, 2)), want_hess=False)
    c = ev["c"].cpu().numpy().reshape(n, 3, 9)              # per step: 4 circles, 4 ellipses, coupling row
    worst = c[:, :, :8].reshape(n, -1).min(axis=1)
    fail, feasi = G["dd_label"] == 2, G["dd_label"] == 0
    assert np.mean(worst[fail] < -1e-4) >= 0.95 and np.mean(worst[feasi] >= -1e-4) >= 0.9


def test_dd_solver_reproduces_ipopt_feasibility_verdicts():
    """dcbf_solve (dd) from the start state of the 110 recorded differential-drive re-plans: status 2 exactly where the reference's
    Ipopt run gave up (pred_fail), as for the oracle in tests/test_data_log_cpu.py; both kernel families"""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "data_log_plans.npz"))
    P = G["dd_plan"]
    n = len(P)
    d = np.diff(P, axis=1)
    v = (d[:, :, 0] * np.cos(P[:, :3, 2]) + d[:, :, 1] * np.sin(P[:, :3, 2])) / 0.4
    u = np.stack([v, d[:, :, 2]], axis=2).reshape(n, 6)
    want = G["dd_label"] == 2
    for mode in ("warp", "thread"):
        os.environ["DCBF_KERNEL"] = mode
        try:
            s = DcbfSolver("dd", device=0)
        finally:
            del os.environ["DCBF_KERNEL"]
        s.set_fields(G["dd_cir"] + np.array([0.0, 0.0, 0.4]), G["dd_elp"] + np.array([0.0, 0.0, 0.4, 0.4, 0.0]))
        r = s.solve(P[:, 0], np.tile([10.0, 10.0], (n, 1)), None, u, field=G["dd_run"].astype(np.int32), last_u=u[:, :2].copy())
        got = r.status.cpu().numpy() == 2
        assert (got == want).mean() >= 0.98 and not (want & ~got).any(), mode


def test_heading_input_kernel_matches_reference_logger():
    """dcbf_heading_input against Logger.tube_func / avg_hd outputs frozen from the reference (tests/golden/helpers.npz), bit for
    bit, reading the plan headings in place from an x_plan buffer and writing the third column of the next tick's glo_p."""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "helpers.npz"))
    s = DcbfSolver("modi")
    dev = s.tdev
    B = len(G["hdin_cur"])
    x_plan = torch.zeros((B, 3, 5), device=dev, dtype=torch.float64)
    x_plan[:, :, 4] = torch.as_tensor(G["hdin_hds"], device=dev)
    nex_turn = torch.as_tensor(G["hdin_turn"], device=dev).clone()
    glo_p = torch.full((B, 3), 7.0, device=dev, dtype=torch.float64)
    out = s.heading_input(G["hdin_cur"], nex_turn, x_plan=x_plan, glo_p=glo_p)
    assert np.array_equal(nex_turn.cpu().numpy(), G["hdin_nex_turn"])
    assert np.array_equal(out.cpu().numpy(), G["hdin_pr"])
    assert torch.equal(glo_p[:, :2], torch.full((B, 2), 7.0, device=dev, dtype=torch.float64))
    # dense [B,3] headings and a separate output give the same numbers
    nt2 = torch.as_tensor(G["hdin_turn"], device=dev).clone()
    out2 = s.heading_input(G["hdin_cur"], nt2, mpc_hds=G["hdin_hds"])
    assert torch.equal(out2, glo_p[:, 2]) and torch.equal(nt2, nex_turn)
    # the decay between re-plans: ten ticks without a new plan shrink the pending turn geometrically
    for _ in range(10):
        s.heading_input(G["hdin_cur"], nt2, mpc_hds=G["hdin_hds"])
    assert float(nt2.abs().max()) < 1e-4     # 0.4 rad * 0.7^3 * 0.4^8


def test_tick_on_the_recorded_learning_set():
    """dcbf_tick on the 640 control ticks the reference recorded in sup_learn/*.csv (main_sim_mpc.py run, real cyipopt; fixture
    tests/golden/sup_learn.npz): the LIP prediction equals the recorded x_nex[0:2] and the re-plan, cold-started, lands on the
    recorded foot placement (the reference run warm-started and stopped after <= 30 L-BFGS iterations: statistical tolerances,
    the same as for the oracle in tests/test_sup_learn_cpu.py)."""
    import os
    from mujoco_lip_mpc_simulation_b200 import data_log
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sup_learn.npz"))
    f = data_log.sup_learn_fields(G["X"], G["y_mpc"])
    T = len(f["pos"])
    s = DcbfSolver("modi", device=0)
    s.set_fields((f["obs"][0] + np.array([0.0, 0.0, 0.4]))[None])
    glo_p = np.column_stack([f["stance"], f["hd_input_pr"]])
    out = s.tick(f["pos"], f["vel"], f["heading"], glo_p, f["rest_t"], f["goal"], -f["leg_ind"], want_pos_det=False)
    torch.cuda.synchronize()
    xn = out["x_next"].cpu().numpy()
    np.testing.assert_allclose(xn[:, :2], f["x_nex_pos"], rtol=0, atol=2e-13)      # device cosh / sinh and FMA contraction
    st = out["plan"].status.cpu().numpy()
    d = np.linalg.norm(out["plan"].p_plan.cpu().numpy()[:, 0, :2] - f["foot"], axis=1)
    ok = st == 0
    assert ok.mean() > 0.8
    assert np.median(d[ok]) < 5e-4 and (d[ok] < 1e-3).mean() > 0.65
    assert (d < 1e-2).mean() > 0.9 and d.max() < 0.12
    # the learning-set rows written from the tick's own inputs and outputs have the reference layout
    X, y = data_log.sup_learn_rows(f["obs"][0], f["pos"], f["vel"], f["heading"], f["stance"], f["goal"], f["leg_ind"], f["rest_t"],
                                   out["plan"].p_plan.cpu().numpy()[:, 0, :2], f["hd_input_pr"], xn[:, :2], f["v_des"])
    assert np.array_equal(X, G["X"]) and y.shape == G["y_mpc"].shape
    np.testing.assert_allclose(y[:, 2:], G["y_mpc"][:, 2:], rtol=0, atol=2e-13)


def test_tick_rate_closed_loop_stays_on_the_device():
    """SURVEY.md 8(f) row 1 end to end: robots walk to the goal re-planning EIGHT times per step from a noisy state, every piece on
    the GPU -- scenarios from dcbf_gen_fields / dcbf_gen_states, heading input from dcbf_heading_input, prediction + warm-start
    rule + re-plan from dcbf_tick; the plant is the LIP flow about the stance foot with velocity noise (torch ops on the same
    stream).  Nothing is copied to the host inside the loop."""
    torch.manual_seed(3)
    s = DcbfSolver("sig_step", device=0)
    dev, B, F, beta, dtk = s.tdev, 512, 64, _lipmodel.BETA, 0.05
    sc = scenarios.make_batch_device(s, B, seed=17, n_fields=F)
    assert int((sc["attempts"] < 0).sum()) == 0
    x0, goal, field = sc["x0"], sc["goal"], sc["field"]
    first = s.solve(x0, goal, sc["leg"], sc["warm"], field=field)                     # the step that is about to start
    pos, vel, hd = x0[:, 0:2].clone(), x0[:, 2:4].clone(), x0[:, 4].clone()
    stance, nex_turn = first.p_plan[:, 0, :2].clone(), first.p_plan[:, 0, 2].clone()
    leg_next = (-sc["leg"]).contiguous()
    prev = first.x_plan.reshape(B, 15).clone()
    glo_p = torch.zeros((B, 3), device=dev, dtype=torch.float64)
    cir = sc["cir"][field.long()]                                                    # [B, K, 3] inflated by 0.4
    d0 = torch.linalg.norm(pos - goal, dim=1)
    done = torch.zeros(B, dtype=torch.bool, device=dev)
    min_gap = torch.full((B,), float("inf"), device=dev, dtype=torch.float64)
    n_ok = torch.zeros((), device=dev, dtype=torch.float64)
    n_ticks, launches0 = 0, s.launches
    ch, sh = math.cosh(beta * dtk), math.sinh(beta * dtk)
    for step in range(40):
        for j in range(8):
            glo_p[:, :2] = stance
            s.heading_input(hd, nex_turn, x_plan=prev, glo_p=glo_p)
            out = s.tick(pos, vel, hd, glo_p, 0.4 - dtk * j, goal, leg_next, prev_plan=prev,
                         mode=torch.full((B,), 1 if j == 0 else 0, dtype=torch.uint8, device=dev), field=field, want_pos_det=False)
            plan = out["plan"]
            prev = plan.x_plan.reshape(B, 15)
            nex_turn = plan.p_plan[:, 0, 2].clone()
            n_ok += (plan.status == 0).double().mean()
            n_ticks += 1
            # plant: LIP flow about the stance foot over one tick, velocity noise, heading follows the commanded rate
            rel = pos - stance
            npos = stance + ch * rel + (sh / beta) * vel
            nvel = (sh * beta) * rel + ch * vel + 0.01 * torch.randn_like(vel)
            live = (~done)[:, None]
            pos, vel = torch.where(live, npos, pos), torch.where(live, nvel, vel)
            hd = torch.where(~done, hd + glo_p[:, 2] * (dtk / 0.4), hd)
            gap = (torch.linalg.norm(pos[:, None, :] - cir[:, :, :2], dim=2) - (cir[:, :, 2] - 0.4)).min(dim=1).values
            min_gap = torch.minimum(min_gap, gap)
        stance = torch.where((~done)[:, None], plan.p_plan[:, 0, :2], stance)          # touchdown on the planned foothold
        leg_next = (-leg_next).contiguous()
        done |= torch.linalg.norm(pos - goal, dim=1) < 0.3
    torch.cuda.synchronize()
    d1 = torch.linalg.norm(pos - goal, dim=1)
    print(f"closed loop: {n_ticks} ticks x {B} robots, converged re-plans {float(n_ok) / n_ticks:.3f}, arrived {float(done.double().mean()):.3f}, "
          f"never inside an obstacle {float((min_gap > 0.0).double().mean()):.3f}, median progress {float((d0 - d1).median()):.2f} m")
    assert s.launches - launches0 >= 3 * n_ticks                 # heading input + prepare + solve per tick, all ours
    assert float(n_ok) / n_ticks > 0.85                          # converged re-plans
    assert float((min_gap > 0.0).double().mean()) > 0.97         # the body of the obstacle is never entered
    assert float((min_gap > -0.05).double().mean()) > 0.99
    assert float(((d1 < d0 - 2.0) | done).double().mean()) > 0.9  # walked at least 2 m towards the goal, or arrived
    assert float(done.double().mean()) > 0.3


def test_veldes_foot_kernel_matches_reference_helpers():
    """dcbf_veldes_foot = MPCCBF.alip_des_vel + MPCCBF.cal_foot_with_veldes (MPC_LIP_sig_step.py:168-181) against the values the
    reference's own methods returned (tests/golden/helpers.npz), and against the host mirror on a batch"""
    import os
    H = np.load(os.path.join(os.path.dirname(__file__), "golden", "helpers.npz"))
    s = DcbfSolver("sig_step", device=0)
    a = s.veldes_foot(leg=[1], vx_max=0.7)["vel_des"].cpu().numpy()[0]
    b = s.veldes_foot(leg=[-1], vx_max=0.5)["vel_des"].cpu().numpy()[0]
    np.testing.assert_allclose([a, b], H["alip_des_vel"], rtol=0, atol=1e-14)
    r = s.veldes_foot(x_state=H["cfv_in"][None, :5], vel_des=H["cfv_in"][None, 5:])
    np.testing.assert_allclose(r["foot"].cpu().numpy()[0], H["cfv_out"], rtol=0, atol=1e-13)
    from mujoco_lip_mpc_simulation_b200.MPC_LIP_sig_step import MPCCBF
    rng = np.random.default_rng(3)
    xs = rng.normal(size=(257, 5)); leg = rng.choice([-1, 1], size=257).astype(np.int32)
    r = s.veldes_foot(x_state=xs, leg=leg, vx_max=0.6)
    pl = MPCCBF([[10.0, 10.0]], [[1, 1, 0.5]], [[1, 1, 0.9]], [-0.5, 10.5])
    for i in (0, 1, 100, 256):
        vd = pl.alip_des_vel(0.6, int(leg[i]))
        np.testing.assert_allclose(r["vel_des"][i].cpu().numpy(), vd, rtol=0, atol=1e-14)
        np.testing.assert_allclose(r["foot"][i].cpu().numpy(), pl.cal_foot_with_veldes(xs[i], vd), rtol=0, atol=1e-12)
