// dcbf_lanes.cuh -- per-lane bodies of the kernels (one problem per thread), shared by dcbf_kernels.cu and by the
// host-compiled debugging harness tests/hostsim/hostsim.cpp.  See dcbf_core.cuh for the model and the solver.
#pragma once
#include "dcbf_core.cuh"

namespace dcbf {

#ifndef DCBF_KT
#define DCBF_KT (2 * DCBF_MAX_OBS)
#endif

// obstacle lists -> quadratic-form records (MPC_LIP_modi.py:598-609 hoisted out of the solve)
DCBF_HD void prep_circle(const double *c, double *o) { o[0] = c[0]; o[1] = c[1]; o[2] = c[2] * c[2]; }
DCBF_HD void prep_ellipse(const double *e, double *o) {
    // a' = (b cos)^2 + (a sin)^2, b' = 2 cos sin (b^2 - a^2), c' = (b sin)^2 + (a cos)^2, rhs = (a b)^2
    const double cp = cos(e[4]), sp = sin(e[4]);
    o[0] = e[0]; o[1] = e[1];
    o[2] = (e[3] * cp) * (e[3] * cp) + (e[2] * sp) * (e[2] * sp);
    o[3] = 2.0 * cp * sp * (e[3] * e[3] - e[2] * e[2]);
    o[4] = (e[3] * sp) * (e[3] * sp) + (e[2] * cp) * (e[2] * cp);
    o[5] = (e[3] * e[2]) * (e[3] * e[2]);
    const double rm = e[2] > e[3] ? e[2] : e[3];
    o[6] = rm * rm; o[7] = 0.0;
}

struct BatchIn {
    const double *x0, *goal, *warm, *last_u;
    const int32_t *leg, *field;
    const double *cir_rec, *elp_rec;
    int Kc, Ke;
    int F;   // number of prepared fields (0: unknown, no range check)
    const uint8_t *mode;   // per scenario: how the start vector was made (0 previous plan verbatim, 1 shifted plan, 2 / NULL cold start)
};

// field index of scenario b.  An index outside [0, F) -- a stale or short index array -- must not become an out-of-bounds read:
// the scenario then reads field 0 and `bad` makes the caller poison its state, so that it comes back with status -13
// (invalid number) and NaN plans instead of silent garbage.
DCBF_HD int batch_field(const BatchIn &in, int b, bool &bad) {
    const int f = in.field ? in.field[b] : 0;
    bad = in.F > 0 && (unsigned)f >= (unsigned)in.F;
    return bad ? 0 : f;
}

template <bool DD>
DCBF_HD void load_problem(const dcbf_params &P, const BatchIn &in, int b, Problem &pb) {
    constexpr int NX = DD ? 3 : 5;
    DCBF_UNROLL
    for (int i = 0; i < NX; i++) pb.x0[i] = in.x0[(size_t)NX * b + i];
    pb.goal_raw[0] = in.goal[2 * (size_t)b]; pb.goal_raw[1] = in.goal[2 * (size_t)b + 1];
    pb.leg = in.leg ? in.leg[b] : 1;
    if (DD && in.last_u) { pb.last_u[0] = in.last_u[2 * (size_t)b]; pb.last_u[1] = in.last_u[2 * (size_t)b + 1]; }
    else { pb.last_u[0] = 0.0; pb.last_u[1] = 0.0; }
    bool bad;
    const int f = batch_field(in, b, bad);
    if (bad) { pb.x0[2] = nan(""); pb.x0[NX - 1] = nan(""); }
    pb.nc = in.Kc; pb.ne = in.Ke;
    pb.cir = in.cir_rec + (size_t)f * in.Kc * DCBF_CIR_REC;
    pb.elp = in.elp_rec + (size_t)f * in.Ke * DCBF_ELP_REC;
    setup_problem(P, pb);
}

struct SolveOut {
    double *u, *x_plan, *p_plan, *obj, *viol;
    int32_t *status, *iters;
    uint8_t *close;
};

DCBF_HD bool lip_close(const dcbf_params &P, const Problem &pb, const LipNodes &nd) {
    bool close = false;
    DCBF_UNROLL
    for (int i = 0; i < 3; i++) {
        const double dxg = nd.x[i + 1] - pb.goal_raw[0], dyg = nd.y[i + 1] - pb.goal_raw[1];
        if ((i == 0 || P.close_any) && sqrt(dxg * dxg + dyg * dyg) <= P.close_radius) close = true;
    }
    return close;
}

// plan re-roll of gen_control_test (MPC_LIP_sig_step.py:99-111, MPC_LIP_modi.py:102-115)
DCBF_HD void write_lip_plan(const dcbf_params &P, const Problem &pb, const LipNodes &nd, const double *z, int b, const SolveOut &o) {
    DCBF_UNROLL
    for (int i = 0; i < 3; i++) {
        const double st[5] = {nd.x[i + 1], nd.y[i + 1], nd.vx[i + 1], nd.vy[i + 1], nd.th[i + 1]};
        DCBF_UNROLL
        for (int j = 0; j < 5; j++) {
            if (o.u) o.u[15 * (size_t)b + 5 * i + j] = st[j];            // representative u_k := x_{k+1}
            if (o.x_plan) o.x_plan[15 * (size_t)b + 5 * i + j] = st[j];
        }
        if (o.p_plan) {
            o.p_plan[9 * (size_t)b + 3 * i + 0] = z[FXI(i)];
            o.p_plan[9 * (size_t)b + 3 * i + 1] = z[FYI(i)];
            o.p_plan[9 * (size_t)b + 3 * i + 2] = z[THI(i)];
        }
    }
    if (o.close) o.close[b] = lip_close(P, pb, nd) ? 1 : 0;
}

// ---- K1+K2: one solve, split into begin / iterate / finish so that a lane can pick up a new problem as soon as its
// current one has converged (persistent "refill" kernels) ----------------------------------------------------------------
DCBF_HD void lip_lane_begin(const dcbf_params &P, const Consts &K, const BatchIn &in, int b, LipModel<DCBF_KT> &M, IpmState<9> &S) {
    load_problem<false>(P, in, b, M.pb);
    ipm_init(P, S);
    double u0[15];
    DCBF_UNROLL
    for (int i = 0; i < 15; i++) u0[i] = in.warm[15 * (size_t)b + i];
    lip_z_from_u(K, M.pb.x0, u0, S.z);
    if (in.mode) S.mu = in.mode[b] == 0 ? P.mu_warm : (in.mode[b] == 1 ? P.mu_shift : P.mu_init);
}

DCBF_HD void lip_lane_finish(const dcbf_params &P, const LipModel<DCBF_KT> &M, const IpmState<9> &S, int b, const SolveOut &out) {
    if (out.status) out.status[b] = S.status;
    if (out.iters) out.iters[b] = S.iters;
    if (out.obj) out.obj[b] = S.obj;
    if (out.viol) out.viol[b] = S.viol;
    write_lip_plan(P, M.pb, M.nd, S.z, b, out);
}

DCBF_HD void solve_lip_lane(const dcbf_params &P, const Consts &K, const BatchIn &in, const SolveOut &out, int b) {
    LipModel<DCBF_KT> M;
    IpmState<9> S;
    lip_lane_begin(P, K, in, b, M, S);
    while (!ipm_iterate(K, P, M, S)) {}
    lip_lane_finish(P, M, S, b, out);
}

DCBF_HD void dd_lane_begin(const dcbf_params &P, const BatchIn &in, int b, DdModel<DCBF_KT> &M, IpmState<6> &S) {
    load_problem<true>(P, in, b, M.pb);
    ipm_init(P, S);
    DCBF_UNROLL
    for (int i = 0; i < 6; i++) S.z[i] = in.warm[6 * (size_t)b + i];
}

DCBF_HD void dd_lane_finish(const dcbf_params &P, const Consts &K, DdModel<DCBF_KT> &M, const IpmState<6> &S, int b, const SolveOut &out) {
    if (out.status) out.status[b] = S.status;
    if (out.iters) out.iters[b] = S.iters;
    if (out.obj) out.obj[b] = S.obj;
    if (out.viol) out.viol[b] = S.viol;
    // plan re-roll (MPC_DD_sig_step.py:83-99)
    dd_rollout(K, M.pb.x0, S.z, M.nd, false);
    DCBF_UNROLL
    for (int i = 0; i < 3; i++) {
        if (out.x_plan) {
            out.x_plan[9 * (size_t)b + 3 * i + 0] = M.nd.x[i + 1];
            out.x_plan[9 * (size_t)b + 3 * i + 1] = M.nd.y[i + 1];
            out.x_plan[9 * (size_t)b + 3 * i + 2] = M.nd.th[i + 1];
        }
        if (out.u) { out.u[6 * (size_t)b + 2 * i] = S.z[2 * i]; out.u[6 * (size_t)b + 2 * i + 1] = S.z[2 * i + 1]; }
    }
    if (out.close) {
        const double dxg = M.nd.x[1] - M.pb.goal_raw[0], dyg = M.nd.y[1] - M.pb.goal_raw[1];
        out.close[b] = sqrt(dxg * dxg + dyg * dyg) <= P.close_radius ? 1 : 0;
    }
}

DCBF_HD void solve_dd_lane(const dcbf_params &P, const Consts &K, const BatchIn &in, const SolveOut &out, int b) {
    DdModel<DCBF_KT> M;
    IpmState<6> S;
    dd_lane_begin(P, in, b, M, S);
    while (!ipm_iterate(K, P, M, S)) {}
    dd_lane_finish(P, K, M, S, b, out);
}

// ---- K1: evaluation at given points ------------------------------------------------------------------------------------
struct EvalPtrs {
    const double *z, *lambda;
    double *f, *grad, *c, *jac, *cl, *cu, *hess;
    int m;
};

DCBF_HD void eval_lip_lane(const dcbf_params &P, const Consts &K, const BatchIn &in, const EvalPtrs &ev, int b) {
    Problem pb;
    {
        dcbf_params Q = P; Q.goal_shift = 0; Q.select_obs = 0;
        load_problem<false>(Q, in, b, pb);
    }
    double z[9];
    DCBF_UNROLL
    for (int i = 0; i < 3; i++) {   // reference p order -> internal order
        z[FXI(i)] = ev.z[9 * (size_t)b + 3 * i];
        z[FYI(i)] = ev.z[9 * (size_t)b + 3 * i + 1];
        z[THI(i)] = ev.z[9 * (size_t)b + 3 * i + 2];
    }
    LipNodes nd;
    lip_rollout(K, pb.x0, z, nd);
    Acc<9> A;
    acc_reset(A);
    LogAcc LA; LA.sum = 0.0; LA.prod = 1.0; LA.cnt = 0;
    RowCtl ctl; ctl.mu = 0.0; ctl.alpha = 0.0; ctl.alpha_z = 0.0; ctl.phase = PH_MAIN; ctl.pending = false; ctl.reinit = false;
    EvalOut E;
    const size_t m = ev.m;
    E.c = ev.c ? ev.c + m * b : nullptr;
    E.jac = ev.jac ? ev.jac + 9 * m * b : nullptr;
    E.cl = ev.cl ? ev.cl + m * b : nullptr;
    E.cu = ev.cu ? ev.cu + m * b : nullptr;
    E.lambda = ev.lambda ? ev.lambda + m * b : nullptr;
    E.row = 0;
    LipRows<1> unused;
    lip_full_step<0, MODE_EVAL, 1>(K, P, pb, nd, z, 1.0, ctl, unused, A, LA, &E);
    lip_full_step<1, MODE_EVAL, 1>(K, P, pb, nd, z, 1.0, ctl, unused, A, LA, &E);
    lip_full_step<2, MODE_EVAL, 1>(K, P, pb, nd, z, 1.0, ctl, unused, A, LA, &E);
    if (ev.f) ev.f[b] = A.f;
    if (ev.grad)
        for (int a = 0; a < 9; a++) ev.grad[9 * (size_t)b + lip_ref_var(a)] = A.grad[a];
    if (ev.hess)
        for (int a = 0; a < 9; a++)
            for (int c = 0; c < 9; c++) ev.hess[81 * (size_t)b + 9 * lip_ref_var(a) + lip_ref_var(c)] = A.K[tri(a, c)];
}

DCBF_HD void eval_dd_lane(const dcbf_params &P, const Consts &K, const BatchIn &in, const EvalPtrs &ev, int b) {
    Problem pb;
    {
        dcbf_params Q = P; Q.goal_shift = 0; Q.select_obs = 0;
        load_problem<true>(Q, in, b, pb);
    }
    double z[6];
    DCBF_UNROLL
    for (int i = 0; i < 6; i++) z[i] = ev.z[6 * (size_t)b + i];
    DdNodes nd;
    dd_rollout(K, pb.x0, z, nd, true);
    Acc<6> A;
    acc_reset(A);
    LogAcc LA; LA.sum = 0.0; LA.prod = 1.0; LA.cnt = 0;
    RowCtl ctl; ctl.mu = 0.0; ctl.alpha = 0.0; ctl.alpha_z = 0.0; ctl.phase = PH_MAIN; ctl.pending = false; ctl.reinit = false;
    EvalOut E;
    const size_t m = ev.m;
    E.c = ev.c ? ev.c + m * b : nullptr;
    E.jac = ev.jac ? ev.jac + 6 * m * b : nullptr;
    E.cl = ev.cl ? ev.cl + m * b : nullptr;
    E.cu = ev.cu ? ev.cu + m * b : nullptr;
    E.lambda = ev.lambda ? ev.lambda + m * b : nullptr;
    E.row = 0;
    DdSecond H2;
    DCBF_UNROLL
    for (int i = 0; i < 4; i++) { H2.qxx[i] = H2.qxy[i] = H2.qyy[i] = H2.cx[i] = H2.cy[i] = 0.0; }
    DdRows<1> unused;
    dd_full_step<0, MODE_EVAL, 1>(K, P, pb, nd, z, 1.0, ctl, unused, A, LA, H2, &E);
    dd_full_step<1, MODE_EVAL, 1>(K, P, pb, nd, z, 1.0, ctl, unused, A, LA, H2, &E);
    dd_full_step<2, MODE_EVAL, 1>(K, P, pb, nd, z, 1.0, ctl, unused, A, LA, H2, &E);
    dd_add_second(K, nd, z, H2, A.K);
    if (ev.f) ev.f[b] = A.f;
    if (ev.grad)
        for (int a = 0; a < 6; a++) ev.grad[6 * (size_t)b + a] = A.grad[a];
    if (ev.hess)
        for (int a = 0; a < 6; a++)
            for (int c = 0; c < 6; c++) ev.hess[36 * (size_t)b + 6 * a + c] = A.K[tri(a, c)];
}

// ---- K3: closed-loop rollout (LIP formulations): plan -> apply -> re-plan, MPC_LIP_sig_step.py:565-575 -------------------
struct RolloutOut {
    double *x_final, *traj;
    int32_t *steps_done, *n_infeasible, *total_iters;
};

DCBF_HD void rollout_lip_lane(const dcbf_params &P, const Consts &K, const BatchIn &in, const RolloutOut &out, int steps, int b) {
    LipModel<DCBF_KT> M;
    IpmState<9> S;
    load_problem<false>(P, in, b, M.pb);
    double u0[15];
    DCBF_UNROLL
    for (int i = 0; i < 3; i++) {
        DCBF_UNROLL
        for (int j = 0; j < 5; j++) u0[5 * i + j] = M.pb.x0[j];          // cold start [x, x, x]
    }
    int done = 0, ninf = 0, tot = 0;
    for (int st = 0; st < steps; st++) {
        setup_problem(P, M.pb);                                          // goal shift / selection at the new state
        ipm_init(P, S);
        if (st > 0) S.mu = P.mu_shift;                                   // warm-started re-plan: lower first barrier parameter
        lip_z_from_u(K, M.pb.x0, u0, S.z);
        while (!ipm_iterate(K, P, M, S)) {}
        tot += S.iters;
        if (S.status == 2) ninf++;
        const LipNodes &nd = M.nd;
        const bool close = lip_close(P, M.pb, nd);
        if (out.traj) {
            double *t = out.traj + ((size_t)b * steps + st) * 8;
            t[0] = nd.x[1]; t[1] = nd.y[1]; t[2] = nd.vx[1]; t[3] = nd.vy[1]; t[4] = nd.th[1];
            t[5] = S.z[FXI(0)]; t[6] = S.z[FYI(0)]; t[7] = (double)S.status;
        }
        // shifted warm start [x_2, x_3, x_3] (MPC_LIP_sig_step.py:188-189 with init_guess = x_list)
        const double x2[5] = {nd.x[2], nd.y[2], nd.vx[2], nd.vy[2], nd.th[2]};
        const double x3[5] = {nd.x[3], nd.y[3], nd.vx[3], nd.vy[3], nd.th[3]};
        DCBF_UNROLL
        for (int j = 0; j < 5; j++) { u0[j] = x2[j]; u0[5 + j] = x3[j]; u0[10 + j] = x3[j]; }
        // apply the first step exactly (model = plant) and flip the stance leg
        M.pb.x0[0] = nd.x[1]; M.pb.x0[1] = nd.y[1]; M.pb.x0[2] = nd.vx[1]; M.pb.x0[3] = nd.vy[1]; M.pb.x0[4] = nd.th[1];
        M.pb.leg = -M.pb.leg;
        done = st + 1;
        if (close) break;
    }
    if (out.traj) {
        const double nanv = nan("");
        for (int st = done; st < steps; st++)
            for (int j = 0; j < 8; j++) out.traj[((size_t)b * steps + st) * 8 + j] = nanv;
    }
    if (out.x_final) {
        DCBF_UNROLL
        for (int j = 0; j < 5; j++) out.x_final[5 * (size_t)b + j] = M.pb.x0[j];
    }
    if (out.steps_done) out.steps_done[b] = done;
    if (out.n_infeasible) out.n_infeasible[b] = ninf;
    if (out.total_iters) out.total_iters[b] = tot;
}

}  // namespace dcbf
