/*
 * ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C FP64 restatement of the reference hot path (one NLP per call, CPU).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Restated reference code (paths relative to /root/reference):
 *   model matrices A,B,W,M_A,M_B,dx_du,dP_du ....... MPC_LIP_sig_step.py:47-86
 *   objective / gradient ........................... MPC_LIP_sig_step.py:372-407, MPC_LIP_modi.py:430-465,
 *                                                    MPC_DD_sig_step.py:351-397
 *   constraints / jacobian ......................... MPC_LIP_sig_step.py:410-496, MPC_LIP_modi.py:468-583,
 *                                                    MPC_DD_sig_step.py:399-477,534-566
 *   bounds cl/cu (leg parity), goal shift .......... MPC_LIP_sig_step.py:191-253, MPC_LIP_modi.py:201-271,
 *                                                    MPC_DD_sig_step.py:127-141
 *   obstacle selection ............................. MPC_LIP_modi.py:325-338
 *   plan re-roll, close_2_goal ..................... MPC_LIP_sig_step.py:99-111, MPC_LIP_modi.py:102-115,
 *                                                    MPC_DD_sig_step.py:83-99
 * The callbacks are evaluated in the reference's own decision space (u in R^15, DD: R^6) with the dense
 * dx_du / dP_du products the reference uses -- deliberately NOT the reduced foot/turn space of the CUDA
 * kernels, so the two implementations share no derivation.
 *
 * The solve itself lives in third-party code that is absent from /root/reference and from this image:
 * cyipopt (PyPI, unpinned) -> Ipopt (COIN-OR, unpinned, >= 3.14 because of the `hsllib` option) -> HSL MA57.
 * It is restated here from the published algorithm (Waechter & Biegler, Math. Prog. 106(1), 2006): slack
 * reformulation of the general rows, monotone Fiacco-McCormick barrier update (mu_init 0.1, kappa_mu 0.2,
 * theta_mu 1.5, kappa_eps 10), fraction-to-boundary tau = max(0.99, 1-mu), filter line search with
 * switching/Armijo conditions, inertia correction by delta*I, gradient-based objective scaling (max grad 100),
 * bound_relax_factor 1e-8, bound_push/frac 1e-2, tol 1e-8.  Differences, stated: the Lagrangian Hessian is
 * a central finite difference of the analytic first derivatives (the reference gives Ipopt no Hessian and
 * Ipopt falls back to L-BFGS; both reach the same KKT points), the equality multipliers are tied to the
 * slack-bound multipliers (y = z_U - z_L), there is no second-order correction, and the restoration phase is
 * a Levenberg-Marquardt minimisation of the squared row violation instead of Ipopt's l1 restoration NLP.
 * Status 2 (Ipopt "Infeasible_Problem_Detected") is returned when that restoration converges to a stationary
 * point of the violation that is still infeasible by more than constr_viol_tol = 1e-4.
 *
 * PARITY PINNED for the callbacks (tests/golden/callbacks_*.npz come from the reference's own LIP_Prob
 * classes) and for the returned optima as KKT points of the reference callbacks (tests/golden/solves_*.npz);
 * and the solver itself is checked against OUTPUT of the real pipeline: re-solving the 640 control ticks recorded in
 * the reference's sup_learn/*.csv (a main_sim_mpc.py run with cyipopt; tests/golden/sup_learn.npz) from a cold start
 * gives the recorded foot placement to a median of 0.2 mm (74 % within 1 mm, 95 % within 1 cm; the recorded run
 * warm-started and stopped after <= 30 L-BFGS iterations, hence statistical) -- tests/test_sup_learn_cpu.py.
 * PARITY UNPINNED for Ipopt's iteration-capped exit codes, which nothing in the reference records.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define NMAX 15
#define MMAX 160
#define KMAX 48

typedef struct {
    int form;            /* 0 sig_step, 1 modi, 2 dd */
    double p, q, r, gamma, s_turn, t_smooth;
    double bvx_min, bvx_max, bvy_min, bvy_max, leg_sq, ang_max;
    int has_fen;
    int max_iter;
    double tol;
    int select_obs;      /* modi: apply detection-range selection */
    int goal_shift;      /* sig_step, modi: detour heuristic */
    double close_radius; /* close_2_goal threshold */
    int split_abs;       /* solve with the smooth two-row form of s*|dtheta|+v (see prob_eval) */
} orc_params;

typedef struct {
    const orc_params *P;
    double xk[5];
    double goal[2];      /* goal actually used by the NLP (after shift) */
    int leg;
    int nc, ne;
    double cir[KMAX][3];
    double elp[KMAX][5];
    double last_u[2];
    int n, m;            /* solver rows: reference rows, DD variable-bound rows, then the 3 split rows */
    int mref;            /* reference rows only */
    int split;
} orc_problem;

/* ------------------------------------------------------------------------------------------------ */
/* model constants                                                                                   */
/* ------------------------------------------------------------------------------------------------ */
static double g_A[5][5], g_B[5][3], g_W[3][5], g_MA[5][5], g_MB[5][5], g_dx[20][15], g_dp[9][15];
static int g_model_ready = 0;
static const double ORC_BETA = 3.1320919526731650; /* sqrt(9.81) */
static const double ORC_DT = 0.4;

static void mat55(const double a[5][5], const double b[5][5], double o[5][5]) {
    double t[5][5];
    for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) {
        double s = 0; for (int k = 0; k < 5; k++) s += a[i][k] * b[k][j]; t[i][j] = s; }
    memcpy(o, t, sizeof(t));
}

static void build_model(void) {
    if (g_model_ready) return;
    double beta = sqrt(9.81 / 1.0);
    double ch = cosh(beta * ORC_DT), sh = sinh(beta * ORC_DT);
    memset(g_A, 0, sizeof(g_A)); memset(g_B, 0, sizeof(g_B)); memset(g_W, 0, sizeof(g_W));
    for (int i = 0; i < 4; i++) g_A[i][i] = ch;
    g_A[4][4] = 1.0;
    g_A[0][2] = g_A[1][3] = sh / beta;
    g_A[2][0] = g_A[3][1] = sh * beta;
    g_B[0][0] = g_B[1][1] = 1.0 - ch;
    g_B[2][0] = g_B[3][1] = -sh * beta;
    g_B[4][2] = 1.0;
    double wa = 5.0, wb = 1.0;
    double den = wa * (ch - 1.0) * (ch - 1.0) + wb * (sh * beta) * (sh * beta);
    double c_h = -wa * (ch - 1.0) / den, s_h = -wb * sh * beta / den;
    g_W[0][0] = g_W[1][1] = c_h; g_W[0][2] = g_W[1][3] = s_h; g_W[2][4] = 1.0;
    /* M_B = B W (5x5), M_A = A - B W A */
    double WA[3][5];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 5; j++) {
        double s = 0; for (int k = 0; k < 5; k++) s += g_W[i][k] * g_A[k][j]; WA[i][j] = s; }
    for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) {
        double s = 0, t = 0;
        for (int k = 0; k < 3; k++) { s += g_B[i][k] * g_W[k][j]; t += g_B[i][k] * WA[k][j]; }
        g_MB[i][j] = s; g_MA[i][j] = g_A[i][j] - t; }
    double pre[3][5][5];
    memcpy(pre[0], g_MB, sizeof(g_MB));
    mat55(g_MA, g_MB, pre[1]);
    mat55(g_MA, pre[1], pre[2]);
    memset(g_dx, 0, sizeof(g_dx)); memset(g_dp, 0, sizeof(g_dp));
    for (int row = 1; row < 4; row++) for (int col = 0; col < row; col++)
        for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++)
            g_dx[5 * row + i][5 * col + j] = pre[row - 1 - col][i][j];
    /* dP blocks: W, -W A M_B, -W A M_A M_B */
    double pl[3][3][5];
    memcpy(pl[0], g_W, sizeof(g_W));
    for (int b = 1; b < 3; b++)
        for (int i = 0; i < 3; i++) for (int j = 0; j < 5; j++) {
            double s = 0; for (int k = 0; k < 5; k++) s += WA[i][k] * pre[b - 1][k][j]; pl[b][i][j] = -s; }
    for (int row = 0; row < 3; row++) for (int col = 0; col <= row; col++)
        for (int i = 0; i < 3; i++) for (int j = 0; j < 5; j++)
            g_dp[3 * row + i][5 * col + j] = pl[row - col][i][j];
    g_model_ready = 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* level sets                                                                                        */
/* ------------------------------------------------------------------------------------------------ */
static double h_cir(const double *c, double x, double y) {
    return (x - c[0]) * (x - c[0]) + (y - c[1]) * (y - c[1]) - c[2] * c[2];
}
static void elp_coef(const double *e, double *a, double *b, double *c, double *rhs) {
    double cp = cos(e[4]), sp = sin(e[4]);
    *a = (e[3] * cp) * (e[3] * cp) + (e[2] * sp) * (e[2] * sp);
    *b = 2.0 * cp * sp * (e[3] * e[3] - e[2] * e[2]);
    *c = (e[3] * sp) * (e[3] * sp) + (e[2] * cp) * (e[2] * cp);
    *rhs = (e[3] * e[2]) * (e[3] * e[2]);
}
static double h_elp(const double *e, double x, double y) {
    double a, b, c, rhs; elp_coef(e, &a, &b, &c, &rhs);
    double dx = x - e[0], dy = y - e[1];
    return a * dx * dx + b * dx * dy + c * dy * dy - rhs;
}
static void dh_elp(const double *e, double x, double y, double *d1, double *d2) {
    double a, b, c, rhs; elp_coef(e, &a, &b, &c, &rhs);
    double dx = x - e[0], dy = y - e[1];
    *d1 = 2.0 * a * dx + b * dy; *d2 = 2.0 * c * dy + b * dx;
}

/* ------------------------------------------------------------------------------------------------ */
/* LIP callbacks in u-space                                                                          */
/* ------------------------------------------------------------------------------------------------ */
static void lip_roll(const orc_problem *pb, const double *u, double x[4][5], double p[3][3]) {
    memcpy(x[0], pb->xk, 5 * sizeof(double));
    for (int i = 0; i < 3; i++) {
        const double *ui = u + 5 * i;
        double ax[5];
        for (int a = 0; a < 5; a++) { double s = 0; for (int k = 0; k < 5; k++) s += g_A[a][k] * x[i][k]; ax[a] = s; }
        for (int a = 0; a < 3; a++) { double s = 0; for (int k = 0; k < 5; k++) s += g_W[a][k] * (ui[k] - ax[k]); p[i][a] = s; }
        for (int a = 0; a < 5; a++) {
            double s = 0; for (int k = 0; k < 5; k++) s += g_MA[a][k] * x[i][k] + g_MB[a][k] * ui[k];
            x[i + 1][a] = s; }
    }
}

static void lip_eval(const orc_problem *pb, const double *u, double *f, double *grad, double *c, double *jac) {
    const orc_params *P = pb->P;
    double x[4][5], p[3][3];
    lip_roll(pb, u, x, p);
    const int n = 15;
    if (f || grad) {
        double cost = 0.0;
        if (grad) for (int j = 0; j < n; j++) grad[j] = 0.0;
        for (int i = 1; i < 4; i++) {
            double w = P->q + (i == 1 ? P->p : 0.0);
            double ex = x[i][0] - pb->goal[0], ey = x[i][1] - pb->goal[1];
            double dx_ = -ex, dy_ = -ey;
            double tar = atan2(dy_, dx_);
            double phi = x[i][4] - tar;
            cost += w * (ex * ex + ey * ey) + P->r * phi * phi;
            if (grad) {
                double r2 = dx_ * dx_ + dy_ * dy_;
                for (int j = 0; j < n; j++) {
                    double dtar = (dx_ * (-g_dx[5 * i + 1][j]) - dy_ * (-g_dx[5 * i][j])) / r2;
                    grad[j] += 2.0 * w * (ex * g_dx[5 * i][j] + ey * g_dx[5 * i + 1][j])
                             + 2.0 * P->r * phi * (g_dx[5 * i + 4][j] - dtar);
                }
            }
        }
        if (f) *f = cost;
    }
    if (!c && !jac) return;
    int row = 0;
    for (int i = 0; i < 3; i++) {
        int k = i + 1;
        double th = x[k][4], cs = cos(th), sn = sin(th), vx = x[k][2], vy = x[k][3];
        double vbx = cs * vx + sn * vy, vby = -sn * vx + cs * vy;
        int r_vbx = row;
        if (c) { c[row] = vbx; c[row + 1] = vby; }
        if (jac) for (int j = 0; j < n; j++) {
            jac[row * n + j] = cs * g_dx[5 * k + 2][j] + sn * g_dx[5 * k + 3][j] + (-sn * vx + cs * vy) * g_dx[5 * k + 4][j];
            jac[(row + 1) * n + j] = -sn * g_dx[5 * k + 2][j] + cs * g_dx[5 * k + 3][j] + (-cs * vx - sn * vy) * g_dx[5 * k + 4][j];
        }
        row += 2;
        for (int o = 0; o < pb->nc; o++, row++) {
            const double *ci = pb->cir[o];
            if (c) c[row] = h_cir(ci, x[k][0], x[k][1]) + (P->gamma - 1.0) * h_cir(ci, x[i][0], x[i][1]);
            if (jac) {
                double a1 = 2.0 * (x[k][0] - ci[0]), a2 = 2.0 * (x[k][1] - ci[1]);
                double b1 = 2.0 * (x[i][0] - ci[0]), b2 = 2.0 * (x[i][1] - ci[1]);
                for (int j = 0; j < n; j++)
                    jac[row * n + j] = a1 * g_dx[5 * k][j] + a2 * g_dx[5 * k + 1][j]
                                     + (P->gamma - 1.0) * (b1 * g_dx[5 * i][j] + b2 * g_dx[5 * i + 1][j]);
            }
        }
        for (int o = 0; o < pb->ne; o++, row++) {
            const double *e = pb->elp[o];
            if (c) c[row] = h_elp(e, x[k][0], x[k][1]) + (P->gamma - 1.0) * h_elp(e, x[i][0], x[i][1]);
            if (jac) {
                double a1, a2, b1, b2;
                dh_elp(e, x[k][0], x[k][1], &a1, &a2);
                dh_elp(e, x[i][0], x[i][1], &b1, &b2);
                for (int j = 0; j < n; j++)
                    jac[row * n + j] = a1 * g_dx[5 * k][j] + a2 * g_dx[5 * k + 1][j]
                                     + (P->gamma - 1.0) * (b1 * g_dx[5 * i][j] + b2 * g_dx[5 * i + 1][j]);
            }
        }
        double lx = x[i][0] - p[i][0], ly = x[i][1] - p[i][1];
        if (c) { c[row] = lx * lx + ly * ly; c[row + 1] = p[i][2]; }
        if (jac) for (int j = 0; j < n; j++) {
            jac[row * n + j] = 2.0 * lx * (g_dx[5 * i][j] - g_dp[3 * i][j]) + 2.0 * ly * (g_dx[5 * i + 1][j] - g_dp[3 * i + 1][j]);
            jac[(row + 1) * n + j] = g_dp[3 * i + 2][j];
        }
        row += 2;
        if (P->has_fen) {
            double d = p[i][2];
            if (c) c[row] = P->s_turn * fabs(d) + vbx;
            if (jac) {
                double sg = d == 0.0 ? 0.0 : (d > 0 ? P->s_turn : -P->s_turn);
                for (int j = 0; j < n; j++) jac[row * n + j] = sg * g_dp[3 * i + 2][j] + jac[r_vbx * n + j];
            }
            row++;
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* DD callbacks in u-space (R^6)                                                                     */
/* ------------------------------------------------------------------------------------------------ */
static void dd_roll(const orc_problem *pb, const double *u, double x[4][3]) {
    x[0][0] = pb->xk[0]; x[0][1] = pb->xk[1]; x[0][2] = pb->xk[2];
    for (int i = 0; i < 3; i++) {
        x[i + 1][0] = x[i][0] + ORC_DT * cos(x[i][2]) * u[2 * i];
        x[i + 1][1] = x[i][1] + ORC_DT * sin(x[i][2]) * u[2 * i];
        x[i + 1][2] = x[i][2] + u[2 * i + 1];
    }
}
static void dd_sens(const double x[4][3], const double *u, double d[12][6]) {
    memset(d, 0, 12 * 6 * sizeof(double));
    for (int k = 1; k < 4; k++) for (int j = 0; j < k; j++) {
        d[3 * k][2 * j] = ORC_DT * cos(x[j][2]);
        d[3 * k + 1][2 * j] = ORC_DT * sin(x[j][2]);
        double sx = 0, sy = 0;
        for (int l = j + 1; l < k; l++) { sx += -u[2 * l] * ORC_DT * sin(x[l][2]); sy += u[2 * l] * ORC_DT * cos(x[l][2]); }
        d[3 * k][2 * j + 1] = sx; d[3 * k + 1][2 * j + 1] = sy; d[3 * k + 2][2 * j + 1] = 1.0;
    }
}
static void dd_eval(const orc_problem *pb, const double *u, double *f, double *grad, double *c, double *jac) {
    const orc_params *P = pb->P;
    const int n = 6;
    double x[4][3], d[12][6];
    dd_roll(pb, u, x);
    dd_sens(x, u, d);
    if (f || grad) {
        double cost = 0;
        if (grad) for (int j = 0; j < n; j++) grad[j] = 0;
        for (int k = 1; k < 4; k++) {
            int i = k - 1;
            double w = P->q + (k == 1 ? P->p : 0.0);
            double ex = x[k][0] - pb->goal[0], ey = x[k][1] - pb->goal[1], dx_ = -ex, dy_ = -ey;
            double phi = x[k][2] - atan2(dy_, dx_);
            double pv = i == 0 ? pb->last_u[0] : u[2 * (i - 1)], pw = i == 0 ? pb->last_u[1] : u[2 * (i - 1) + 1];
            double dv = u[2 * i] - pv, dw = u[2 * i + 1] - pw;
            cost += w * (ex * ex + ey * ey) + P->r * phi * phi + P->t_smooth * (dv * dv + dw * dw);
            if (grad) {
                double r2 = dx_ * dx_ + dy_ * dy_;
                for (int j = 0; j < n; j++) {
                    double dtar = (dx_ * (-d[3 * k + 1][j]) - dy_ * (-d[3 * k][j])) / r2;
                    grad[j] += 2.0 * w * (ex * d[3 * k][j] + ey * d[3 * k + 1][j]) + 2.0 * P->r * phi * (d[3 * k + 2][j] - dtar);
                }
                grad[2 * i] += 2.0 * P->t_smooth * dv; grad[2 * i + 1] += 2.0 * P->t_smooth * dw;
                if (i > 0) { grad[2 * (i - 1)] -= 2.0 * P->t_smooth * dv; grad[2 * (i - 1) + 1] -= 2.0 * P->t_smooth * dw; }
            }
        }
        if (f) *f = cost;
    }
    if (!c && !jac) return;
    int row = 0;
    for (int i = 0; i < 3; i++) {
        int k = i + 1;
        for (int o = 0; o < pb->nc + pb->ne; o++, row++) {
            double a1, a2, b1, b2, hv;
            if (o < pb->nc) {
                const double *ci = pb->cir[o];
                hv = h_cir(ci, x[k][0], x[k][1]) + (P->gamma - 1.0) * h_cir(ci, x[i][0], x[i][1]);
                a1 = 2.0 * (x[k][0] - ci[0]); a2 = 2.0 * (x[k][1] - ci[1]);
                b1 = 2.0 * (x[i][0] - ci[0]); b2 = 2.0 * (x[i][1] - ci[1]);
            } else {
                const double *e = pb->elp[o - pb->nc];
                hv = h_elp(e, x[k][0], x[k][1]) + (P->gamma - 1.0) * h_elp(e, x[i][0], x[i][1]);
                dh_elp(e, x[k][0], x[k][1], &a1, &a2); dh_elp(e, x[i][0], x[i][1], &b1, &b2);
            }
            if (c) c[row] = hv;
            if (jac) for (int j = 0; j < n; j++)
                jac[row * n + j] = a1 * d[3 * k][j] + a2 * d[3 * k + 1][j] + (P->gamma - 1.0) * (b1 * d[3 * i][j] + b2 * d[3 * i + 1][j]);
        }
        double w = u[2 * i + 1];
        if (c) c[row] = P->s_turn * fabs(w) + u[2 * i];
        if (jac) {
            for (int j = 0; j < n; j++) jac[row * n + j] = 0.0;
            jac[row * n + 2 * i] = 1.0;
            jac[row * n + 2 * i + 1] = w == 0.0 ? 0.0 : (w > 0 ? P->s_turn : -P->s_turn);
        }
        row++;
    }
    /* DD variable bounds lb<=u<=ub (MPC_DD_sig_step.py:131-140) appended as identity rows */
    for (int j = 0; j < n; j++, row++) {
        if (c) c[row] = u[j];
        if (jac) { for (int q = 0; q < n; q++) jac[row * n + q] = 0.0; jac[row * n + j] = 1.0; }
    }
}

static void ref_eval(const orc_problem *pb, const double *u, double *f, double *grad, double *c, double *jac) {
    if (pb->P->form == 2) dd_eval(pb, u, f, grad, c, jac); else lip_eval(pb, u, f, grad, c, jac);
}

/* Solver-side view of the rows.  With pb->split the non-smooth row  s*|dtheta| + v  in [v_min, v_max]
 * (MPC_LIP_modi.py:493, MPC_DD_sig_step.py:416) is replaced by the two smooth rows  v + s*dtheta <= v_max  and
 * v - s*dtheta <= v_max  (the second appended after all other rows).  The feasible set is identical: the lower
 * bound v_min is implied by the v_bx row (LIP) / the variable bound on v (DD), which carry the same limits. */
static void prob_eval(const orc_problem *pb, const double *u, double *f, double *grad, double *c, double *jac) {
    ref_eval(pb, u, f, grad, c, jac);
    if (!pb->split || (!c && !jac)) return;
    const int n = pb->n, K = pb->nc + pb->ne, mb = pb->m - 3;
    const double s = pb->P->s_turn;
    for (int i = 0; i < 3; i++) {
        if (pb->P->form == 2) {
            int rf = i * (K + 1) + K;
            if (c) { c[rf] = u[2 * i] + s * u[2 * i + 1]; c[mb + i] = u[2 * i] - s * u[2 * i + 1]; }
            if (jac) {
                for (int j = 0; j < n; j++) jac[rf * n + j] = jac[(mb + i) * n + j] = 0.0;
                jac[rf * n + 2 * i] = 1.0; jac[rf * n + 2 * i + 1] = s;
                jac[(mb + i) * n + 2 * i] = 1.0; jac[(mb + i) * n + 2 * i + 1] = -s;
            }
        } else {
            int rv = i * (5 + K), rd = rv + 3 + K, rf = rv + 4 + K;
            if (c) { double v = c[rv], d = c[rd]; c[rf] = v + s * d; c[mb + i] = v - s * d; }
            if (jac) for (int j = 0; j < n; j++) {
                double a = jac[rv * n + j], b = jac[rd * n + j];
                jac[rf * n + j] = a + s * b; jac[(mb + i) * n + j] = a - s * b;
            }
        }
    }
}

/* rows in reference order (+ DD variable-bound rows at the end) */
static int ref_bounds(const orc_problem *pb, double *cl, double *cu) {
    const orc_params *P = pb->P;
    int row = 0, K = pb->nc + pb->ne;
    if (P->form == 2) {
        for (int i = 0; i < 3; i++) {
            for (int o = 0; o < K; o++, row++) { cl[row] = 0.0; cu[row] = INFINITY; }
            cl[row] = P->bvx_min; cu[row] = P->bvx_max; row++;
        }
        for (int i = 0; i < 3; i++) {
            cl[row] = P->bvx_min; cu[row] = P->bvx_max; row++;
            cl[row] = -P->ang_max; cu[row] = P->ang_max; row++;
        }
        return row;
    }
    for (int i = 0; i < 3; i++) {
        int plus = (pb->leg > 0) == (i % 2 == 0);
        cl[row] = P->bvx_min; cu[row] = P->bvx_max; row++;
        cl[row] = plus ? P->bvy_min : -P->bvy_max; cu[row] = plus ? P->bvy_max : -P->bvy_min; row++;
        for (int o = 0; o < K; o++, row++) { cl[row] = 0.0; cu[row] = INFINITY; }
        cl[row] = 0.0; cu[row] = P->leg_sq; row++;
        cl[row] = -P->ang_max; cu[row] = P->ang_max; row++;
        if (P->has_fen) { cl[row] = P->bvx_min; cu[row] = P->bvx_max; row++; }
    }
    return row;
}

static int prob_bounds(const orc_problem *pb, double *cl, double *cu) {
    int rows = ref_bounds(pb, cl, cu);
    if (!pb->split) return rows;
    int K = pb->nc + pb->ne, mb = pb->m - 3;
    for (int i = 0; i < 3; i++) {
        int rf = pb->P->form == 2 ? i * (K + 1) + K : i * (5 + K) + 4 + K;
        cl[rf] = -INFINITY; cl[mb + i] = -INFINITY; cu[mb + i] = cu[rf];
    }
    return pb->m;
}

/* goal shift + selection: fills pb from raw inputs */
static void prob_setup(orc_problem *pb, const orc_params *P, const double *xk, const double *goal, int leg,
                       int nc, const double *cir, int ne, const double *elp, const double *last_u) {
    build_model();
    memset(pb, 0, sizeof(*pb));
    pb->P = P;
    int nx = P->form == 2 ? 3 : 5;
    for (int i = 0; i < nx; i++) pb->xk[i] = xk[i];
    pb->leg = leg;
    if (last_u) { pb->last_u[0] = last_u[0]; pb->last_u[1] = last_u[1]; }
    double px = xk[0], py = xk[1];
    for (int o = 0; o < nc && pb->nc < KMAX; o++) {
        const double *c = cir + 3 * o;
        if (P->select_obs && !((px - c[0]) * (px - c[0]) + (py - c[1]) * (py - c[1]) - c[2] * c[2] <= 16.0)) continue;
        memcpy(pb->cir[pb->nc++], c, 3 * sizeof(double));
    }
    for (int o = 0; o < ne && pb->ne < KMAX; o++) {
        const double *e = elp + 5 * o;
        double r = e[2] > e[3] ? e[2] : e[3];
        if (P->select_obs && !((px - e[0]) * (px - e[0]) + (py - e[1]) * (py - e[1]) - r * r <= 16.0)) continue;
        memcpy(pb->elp[pb->ne++], e, 5 * sizeof(double));
    }
    pb->goal[0] = goal[0]; pb->goal[1] = goal[1];
    if (P->goal_shift) {
        double dg = (px - goal[0]) * (px - goal[0]) + (py - goal[1]) * (py - goal[1]);
        for (int o = 0; o < pb->nc; o++) {
            const double *c = pb->cir[o];
            double dc = (px - c[0]) * (px - c[0]) + (py - c[1]) * (py - c[1]);
            if (dc < dg && dc < 9.0 * c[2] * c[2]) {
                double th = atan2(goal[1] - py, goal[0] - px), al = atan2(c[1] - py, c[0] - px);
                double d = th - al;
                if (d < 0 && fabs(d) > M_PI) d += 2.0 * M_PI;
                else if (d > 0 && fabs(d) > M_PI) d -= 2.0 * M_PI;
                if (fabs(d) < M_PI / 12) {
                    double na = d < 0 ? th - M_PI / 12 : th + M_PI / 12;
                    double rad = sqrt(dg);
                    pb->goal[0] = px + rad * cos(na); pb->goal[1] = py + rad * sin(na);
                    break;
                }
            }
        }
    }
    pb->n = P->form == 2 ? 6 : 15;
    int K = pb->nc + pb->ne;
    pb->mref = P->form == 2 ? 3 * (K + 1) : 3 * (4 + K + (P->has_fen ? 1 : 0));
    pb->split = P->has_fen && P->split_abs;
    pb->m = pb->mref + (P->form == 2 ? 6 : 0) + (pb->split ? 3 : 0);
}

/* ------------------------------------------------------------------------------------------------ */
/* dense helpers                                                                                     */
/* ------------------------------------------------------------------------------------------------ */
static int chol(int n, double *a /* n x n, lower used, overwritten */) {
    for (int j = 0; j < n; j++) {
        double d = a[j * n + j];
        for (int k = 0; k < j; k++) d -= a[j * n + k] * a[j * n + k];
        if (!(d > 1e-14)) return 0;
        d = sqrt(d); a[j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = a[i * n + j];
            for (int k = 0; k < j; k++) s -= a[i * n + k] * a[j * n + k];
            a[i * n + j] = s / d;
        }
    }
    return 1;
}
static void chol_solve(int n, const double *l, double *b) {
    for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= l[i * n + k] * b[k]; b[i] = s / l[i * n + i]; }
    for (int i = n - 1; i >= 0; i--) { double s = b[i]; for (int k = i + 1; k < n; k++) s -= l[k * n + i] * b[k]; b[i] = s / l[i * n + i]; }
}

/* ------------------------------------------------------------------------------------------------ */
/* interior-point solve                                                                              */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { double theta, phi; } filt_t;
#define FILT_MAX 64

typedef struct {
    int n, m;
    double sf;                 /* objective scaling */
    double dl[MMAX], du[MMAX]; /* relaxed bounds */
    int hasl[MMAX], hasu[MMAX];
    double x[NMAX], s[MMAX], zl[MMAX], zu[MMAX];
    double mu;
    filt_t filt[FILT_MAX]; int nf;
    double theta_max, theta_min;
} ipm_t;

static void lag_grad(const orc_problem *pb, const ipm_t *S, const double *u, const double *y, double *out) {
    double g[NMAX], jac[MMAX * NMAX];
    prob_eval(pb, u, NULL, g, NULL, jac);
    /* need jac only: prob_eval computes both when asked */
    for (int j = 0; j < S->n; j++) { double s = S->sf * g[j]; for (int r = 0; r < S->m; r++) s += jac[r * S->n + j] * y[r]; out[j] = s; }
}

static void fd_hessian(const orc_problem *pb, const ipm_t *S, const double *y, double *H) {
    int n = S->n;
    double up[NMAX], gp[NMAX], gm[NMAX];
    for (int j = 0; j < n; j++) {
        memcpy(up, S->x, n * sizeof(double));
        double h = 1e-6 * fmax(1.0, fabs(S->x[j]));
        up[j] = S->x[j] + h; lag_grad(pb, S, up, y, gp);
        up[j] = S->x[j] - h; lag_grad(pb, S, up, y, gm);
        for (int i = 0; i < n; i++) H[i * n + j] = (gp[i] - gm[i]) / (2.0 * h);
    }
    for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) { double a = 0.5 * (H[i * n + j] + H[j * n + i]); H[i * n + j] = H[j * n + i] = a; }
}

static double barrier_phi(const ipm_t *S, double f, const double *s) {
    double v = S->sf * f;
    for (int r = 0; r < S->m; r++) {
        if (S->hasl[r]) v -= S->mu * log(s[r] - S->dl[r]);
        if (S->hasu[r]) v -= S->mu * log(S->du[r] - s[r]);
    }
    return v;
}

static double row_violation(const ipm_t *S, const double *c, const double *cl, const double *cu) {
    double v = 0;
    for (int r = 0; r < S->m; r++) { double a = cl[r] - c[r], b = c[r] - cu[r]; if (a > v) v = a; if (b > v) v = b; }
    return v;
}

static void init_slacks(ipm_t *S, const double *c) {
    for (int r = 0; r < S->m; r++) {
        double lo = S->dl[r], hi = S->du[r], v = c[r];
        if (S->hasl[r] && S->hasu[r]) {
            double pl = fmin(1e-2 * fmax(1.0, fabs(lo)), 1e-2 * (hi - lo));
            double pu = fmin(1e-2 * fmax(1.0, fabs(hi)), 1e-2 * (hi - lo));
            if (v < lo + pl) v = lo + pl; if (v > hi - pu) v = hi - pu;
        } else if (S->hasl[r]) { double pl = 1e-2 * fmax(1.0, fabs(lo)); if (v < lo + pl) v = lo + pl; }
        else if (S->hasu[r]) { double pu = 1e-2 * fmax(1.0, fabs(hi)); if (v > hi - pu) v = hi - pu; }
        S->s[r] = v;
    }
}

/* Levenberg-Marquardt on 0.5*sum viol^2; returns 1 if violation reduced below target, 0 if stationary */
static int restore(const orc_problem *pb, ipm_t *S, const double *cl, const double *cu, double target, int *iters, int max_iter) {
    int n = S->n, m = S->m;
    double c[MMAX], jac[MMAX * NMAX], K[NMAX * NMAX], rhs[NMAX], xt[NMAX], ct[MMAX];
    double lam = 1e-4;
    int stall = 0;
    prob_eval(pb, S->x, NULL, NULL, c, jac);
    for (int it = 0; it < 200 && *iters < max_iter; it++) {
        double v2 = 0, vmax = 0;
        for (int j = 0; j < n; j++) rhs[j] = 0; memset(K, 0, sizeof(double) * n * n);
        for (int r = 0; r < m; r++) {
            double v = 0; if (cl[r] - c[r] > 0) v = c[r] - cl[r]; else if (c[r] - cu[r] > 0) v = c[r] - cu[r];
            if (v == 0) continue;
            v2 += v * v; if (fabs(v) > vmax) vmax = fabs(v);
            for (int i = 0; i < n; i++) { rhs[i] -= jac[r * n + i] * v; for (int j = 0; j <= i; j++) K[i * n + j] += jac[r * n + i] * jac[r * n + j]; }
        }
        if (vmax <= target) return 1;
        {   /* curvature term sum_r v_r Hess(c_r): Newton instead of Gauss-Newton on 0.5*sum v^2 (nonzero-residual problem) */
            double yv[MMAX], H2[NMAX * NMAX], sf_keep = S->sf;
            for (int r = 0; r < m; r++) { double v = 0; if (cl[r] - c[r] > 0) v = c[r] - cl[r]; else if (c[r] - cu[r] > 0) v = c[r] - cu[r]; yv[r] = v; }
            S->sf = 0.0; fd_hessian(pb, S, yv, H2); S->sf = sf_keep;
            for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) K[i * n + j] += H2[i * n + j];
        }
        double gn = 0; for (int j = 0; j < n; j++) if (fabs(rhs[j]) > gn) gn = fabs(rhs[j]);
        if (gn <= 1e-10 * fmax(1.0, vmax)) return 0;
        int ok = 0;
        for (int tr = 0; tr < 30; tr++) {
            double L[NMAX * NMAX], d[NMAX];
            memcpy(L, K, sizeof(double) * n * n);
            for (int j = 0; j < n; j++) L[j * n + j] += lam;
            if (!chol(n, L)) { lam *= 10; continue; }
            memcpy(d, rhs, sizeof(double) * n); chol_solve(n, L, d);
            for (int j = 0; j < n; j++) xt[j] = S->x[j] + d[j];
            prob_eval(pb, xt, NULL, NULL, ct, NULL);
            double w2 = 0;
            for (int r = 0; r < m; r++) { double v = 0; if (cl[r] - ct[r] > 0) v = ct[r] - cl[r]; else if (ct[r] - cu[r] > 0) v = ct[r] - cu[r]; w2 += v * v; }
            if (w2 < v2 * (1.0 - 1e-12)) {
                double dn = 0; for (int j = 0; j < n; j++) if (fabs(d[j]) > dn) dn = fabs(d[j]);
                memcpy(S->x, xt, sizeof(double) * n);
                lam = fmax(lam * 0.2, 1e-12); ok = 1;
                (*iters)++;
                prob_eval(pb, S->x, NULL, NULL, c, jac);
                if (dn < 1e-12 && sqrt(w2) > target) return 0;
                /* stagnation: two consecutive accepted steps with a relative decrease below 1e-4 */
                if (v2 - w2 <= 1e-4 * v2) stall++; else stall = 0;
                if (stall >= 2 && sqrt(w2) > target) return 0;
                break;
            }
            lam *= 10;
            if (lam > 1e12) break;
        }
        if (!ok) return 0;
    }
    return 0;
}

typedef struct {
    double u[NMAX]; double f; int status; int iters; double viol; double kkt;
} orc_result;

static void ipm_solve(const orc_problem *pb, const double *u0, orc_result *R) {
    const orc_params *P = pb->P;
    ipm_t Sv; ipm_t *S = &Sv;
    memset(S, 0, sizeof(*S));
    int n = pb->n, m = pb->m;
    S->n = n; S->m = m;
    double cl[MMAX], cu[MMAX];
    prob_bounds(pb, cl, cu);
    for (int r = 0; r < m; r++) {
        S->hasl[r] = isfinite(cl[r]); S->hasu[r] = isfinite(cu[r]);
        S->dl[r] = S->hasl[r] ? cl[r] - 1e-8 * fmax(1.0, fabs(cl[r])) : -INFINITY;
        S->du[r] = S->hasu[r] ? cu[r] + 1e-8 * fmax(1.0, fabs(cu[r])) : INFINITY;
    }
    memcpy(S->x, u0, n * sizeof(double));
    double f, g[NMAX], c[MMAX], jac[MMAX * NMAX];
    prob_eval(pb, S->x, &f, g, c, jac);
    double gmax = 0; for (int j = 0; j < n; j++) if (fabs(g[j]) > gmax) gmax = fabs(g[j]);
    S->sf = gmax > 100.0 ? 100.0 / gmax : 1.0;
    init_slacks(S, c);
    for (int r = 0; r < m; r++) { S->zl[r] = S->hasl[r] ? 1.0 : 0.0; S->zu[r] = S->hasu[r] ? 1.0 : 0.0; }
    S->mu = 0.1;
    double theta0 = 0; for (int r = 0; r < m; r++) theta0 += fabs(c[r] - S->s[r]);
    S->theta_max = 1e4 * fmax(1.0, theta0); S->theta_min = 1e-4 * fmax(1.0, theta0);
    const double tol = P->tol > 0 ? P->tol : 1e-8;
    int iters = 0, status = -1;
    double delta_last = 0.0;
    int acceptable_cnt = 0, nstall = 0, tiny = 0;

    for (;;) {
        /* --- optimality error ----------------------------------------------------------------- */
        double y[MMAX];
        for (int r = 0; r < m; r++) y[r] = S->zu[r] - S->zl[r];
        double dinf = 0, pinf = 0, zsum = 0; int nz = 0;
        for (int j = 0; j < n; j++) { double s = S->sf * g[j]; for (int r = 0; r < m; r++) s += jac[r * n + j] * y[r]; if (fabs(s) > dinf) dinf = fabs(s); }
        for (int r = 0; r < m; r++) { double v = fabs(c[r] - S->s[r]); if (v > pinf) pinf = v; }
        for (int r = 0; r < m; r++) { if (S->hasl[r]) { zsum += S->zl[r]; nz++; } if (S->hasu[r]) { zsum += S->zu[r]; nz++; } }
        double sd = fmax(100.0, 2.0 * zsum / fmax(1, m + nz)) / 100.0;   /* y tied to z: ||y||_1 <= zsum */
        double sc = fmax(100.0, zsum / fmax(1, nz)) / 100.0;
        double E0, Emu;
        for (;;) {
            double comp0 = 0, compm = 0;
            for (int r = 0; r < m; r++) {
                if (S->hasl[r]) { double v = (S->s[r] - S->dl[r]) * S->zl[r]; if (fabs(v) > comp0) comp0 = fabs(v); if (fabs(v - S->mu) > compm) compm = fabs(v - S->mu); }
                if (S->hasu[r]) { double v = (S->du[r] - S->s[r]) * S->zu[r]; if (fabs(v) > comp0) comp0 = fabs(v); if (fabs(v - S->mu) > compm) compm = fabs(v - S->mu); }
            }
            E0 = fmax(fmax(dinf / sd, pinf), comp0 / sc);
            Emu = fmax(fmax(dinf / sd, pinf), compm / sc);
            if (E0 <= tol) break;
            if (Emu <= 10.0 * S->mu && S->mu > tol / 10.0 * (1 + 1e-12)) {
                S->mu = fmax(tol / 10.0, fmin(0.2 * S->mu, pow(S->mu, 1.5)));
                S->nf = 0;
                continue;
            }
            break;
        }
        double viol = row_violation(S, c, cl, cu);
        if (E0 <= tol) { status = 0; break; }
        if (E0 <= 1e-6 && viol <= 1e-4) { if (++acceptable_cnt >= 15) { status = 1; break; } } else acceptable_cnt = 0;
        if (iters >= P->max_iter) { status = -1; break; }

        /* --- Newton system --------------------------------------------------------------------- */
        double H[NMAX * NMAX], K[NMAX * NMAX], L[NMAX * NMAX], rhs[NMAX], sig[MMAX], bb[MMAX];
        fd_hessian(pb, S, y, H);
        for (int r = 0; r < m; r++) {
            double a = 0, b = 0;
            if (S->hasl[r]) { a += S->zl[r] / (S->s[r] - S->dl[r]); b += S->mu / (S->s[r] - S->dl[r]); }
            if (S->hasu[r]) { a += S->zu[r] / (S->du[r] - S->s[r]); b -= S->mu / (S->du[r] - S->s[r]); }
            sig[r] = a; bb[r] = b;
        }
        memcpy(K, H, sizeof(double) * n * n);
        for (int j = 0; j < n; j++) rhs[j] = -S->sf * g[j];
        for (int r = 0; r < m; r++) {
            double w = sig[r] * (c[r] - S->s[r]) - bb[r];
            for (int i = 0; i < n; i++) {
                rhs[i] -= jac[r * n + i] * w;
                double t = sig[r] * jac[r * n + i];
                for (int j = 0; j < n; j++) K[i * n + j] += t * jac[r * n + j];
            }
        }
        double delta = 0.0; int fact = 0;
        for (int tr = 0; tr < 40; tr++) {
            memcpy(L, K, sizeof(double) * n * n);
            for (int j = 0; j < n; j++) L[j * n + j] += delta;
            if (chol(n, L)) { fact = 1; break; }
            if (delta == 0.0) delta = delta_last == 0.0 ? 1e-4 : fmax(1e-20, delta_last / 3.0);
            else delta *= (delta_last == 0.0 ? 100.0 : 8.0);
            if (delta > 1e40) break;
        }
        if (!fact) { status = -3; break; }
        if (delta > 0) delta_last = delta;
        double dx[NMAX], ds[MMAX], dzl[MMAX], dzu[MMAX];
        memcpy(dx, rhs, sizeof(double) * n); chol_solve(n, L, dx);
        for (int r = 0; r < m; r++) {
            double jd = 0; for (int j = 0; j < n; j++) jd += jac[r * n + j] * dx[j];
            ds[r] = jd + (c[r] - S->s[r]);
            dzl[r] = S->hasl[r] ? (S->mu / (S->s[r] - S->dl[r]) - S->zl[r]) - S->zl[r] / (S->s[r] - S->dl[r]) * ds[r] : 0.0;
            dzu[r] = S->hasu[r] ? (S->mu / (S->du[r] - S->s[r]) - S->zu[r]) + S->zu[r] / (S->du[r] - S->s[r]) * ds[r] : 0.0;
        }
        /* --- step sizes ------------------------------------------------------------------------ */
        double tau = fmax(0.99, 1.0 - S->mu), amax = 1.0, az = 1.0;
        for (int r = 0; r < m; r++) {
            if (S->hasl[r] && ds[r] < 0) { double a = -tau * (S->s[r] - S->dl[r]) / ds[r]; if (a < amax) amax = a; }
            if (S->hasu[r] && ds[r] > 0) { double a = tau * (S->du[r] - S->s[r]) / ds[r]; if (a < amax) amax = a; }
            if (S->hasl[r] && dzl[r] < 0) { double a = -tau * S->zl[r] / dzl[r]; if (a < az) az = a; }
            if (S->hasu[r] && dzu[r] < 0) { double a = -tau * S->zu[r] / dzu[r]; if (a < az) az = a; }
        }
        /* --- filter line search ---------------------------------------------------------------- */
        double theta = 0; for (int r = 0; r < m; r++) theta += fabs(c[r] - S->s[r]);
        double phi = barrier_phi(S, f, S->s);
        double dphi = 0; for (int j = 0; j < n; j++) dphi += S->sf * g[j] * dx[j];
        for (int r = 0; r < m; r++) {
            if (S->hasl[r]) dphi -= S->mu * ds[r] / (S->s[r] - S->dl[r]);
            if (S->hasu[r]) dphi += S->mu * ds[r] / (S->du[r] - S->s[r]);
        }
        double alpha = amax; int accepted = 0;
        double xt[NMAX], st[MMAX], ct[MMAX], ft;
        for (int ls = 0; ls < 25; ls++, alpha *= 0.5) {
            for (int j = 0; j < n; j++) xt[j] = S->x[j] + alpha * dx[j];
            for (int r = 0; r < m; r++) st[r] = S->s[r] + alpha * ds[r];
            prob_eval(pb, xt, &ft, NULL, ct, NULL);
            double th_t = 0; for (int r = 0; r < m; r++) th_t += fabs(ct[r] - st[r]);
            double ph_t = barrier_phi(S, ft, st);
            if (!isfinite(ph_t) || !isfinite(th_t) || th_t > S->theta_max) continue;
            int in_filter = 0;
            for (int k = 0; k < S->nf; k++) if (th_t >= S->filt[k].theta && ph_t >= S->filt[k].phi) { in_filter = 1; break; }
            if (in_filter) continue;
            int sw = dphi < 0 && theta <= S->theta_min && alpha * pow(-dphi, 2.3) > pow(theta, 1.1);
            if (sw) {
                if (ph_t <= phi + 1e-8 * alpha * dphi + 10 * 2.2e-16 * fabs(phi)) { accepted = 1; }
            } else if (th_t <= (1 - 1e-5) * theta || ph_t <= phi - 1e-5 * theta + 10 * 2.2e-16 * fabs(phi)) {
                accepted = 2;
            }
            if (accepted) break;
        }
        if (accepted && tiny >= 3) { accepted = 0; }   /* pinned by the fraction-to-boundary rule while infeasible: restoration now */
        if (!accepted) {
            tiny = 0;
            /* restoration: reduce the violation of the original rows from the current x */
            double c_now = row_violation(S, c, cl, cu);
            double entry = 0; for (int r = 0; r < m; r++) entry += fabs(c[r] - S->s[r]);
            if (S->nf < FILT_MAX) { S->filt[S->nf].theta = (1 - 1e-5) * theta; S->filt[S->nf].phi = phi - 1e-5 * theta; S->nf++; }
            iters++;   /* entering restoration counts as an iteration (guards against cycling) */
            int ok = restore(pb, S, cl, cu, fmax(0.1 * c_now, 1e-9), &iters, P->max_iter);
            prob_eval(pb, S->x, &f, g, c, jac);
            double vnow = row_violation(S, c, cl, cu);
            if (!ok && vnow > 1e-4) { status = 2; break; }
            /* feasible but the line search is stuck, or a second stall at a marginally infeasible stationary point */
            if (!ok && (c_now <= 1e-9 || nstall++ >= 1)) { status = -2; break; }
            init_slacks(S, c);
            for (int r = 0; r < m; r++) { S->zl[r] = S->hasl[r] ? 1.0 : 0.0; S->zu[r] = S->hasu[r] ? 1.0 : 0.0; }
            (void)entry;
            continue;
        }
        if (accepted == 2 && S->nf < FILT_MAX) { S->filt[S->nf].theta = (1 - 1e-5) * theta; S->filt[S->nf].phi = phi - 1e-5 * theta; S->nf++; }
        memcpy(S->x, xt, sizeof(double) * n);
        memcpy(S->s, st, sizeof(double) * m);
        for (int r = 0; r < m; r++) {
            if (S->hasl[r]) { double z = S->zl[r] + az * dzl[r], gap = S->s[r] - S->dl[r]; z = fmax(fmin(z, 1e10 * S->mu / gap), S->mu / (1e10 * gap)); S->zl[r] = z; }
            if (S->hasu[r]) { double z = S->zu[r] + az * dzu[r], gap = S->du[r] - S->s[r]; z = fmax(fmin(z, 1e10 * S->mu / gap), S->mu / (1e10 * gap)); S->zu[r] = z; }
        }
        if (alpha < 1e-2 && viol > 1e-4) tiny++; else tiny = 0;
        prob_eval(pb, S->x, &f, g, c, jac);
        iters++;
    }
    memcpy(R->u, S->x, n * sizeof(double));
    prob_eval(pb, S->x, &f, g, c, jac);
    R->f = f; R->status = status; R->iters = iters; R->viol = row_violation(S, c, cl, cu);
    double y[MMAX]; for (int r = 0; r < m; r++) y[r] = S->zu[r] - S->zl[r];
    double dinf = 0; for (int j = 0; j < n; j++) { double s = S->sf * g[j]; for (int r = 0; r < m; r++) s += jac[r * n + j] * y[r]; if (fabs(s) > dinf) dinf = fabs(s); }
    R->kkt = dinf / S->sf;
}

/* ------------------------------------------------------------------------------------------------ */
/* exported C entry points (ctypes)                                                                  */
/* ------------------------------------------------------------------------------------------------ */
void orc_default_params(int form, orc_params *P) {
    memset(P, 0, sizeof(*P));
    P->form = form;
    P->q = 1.0; P->bvx_min = 0.4; P->bvx_max = 0.8; P->bvy_min = 0.15; P->leg_sq = 0.09; P->ang_max = M_PI / 16;
    P->tol = 1e-8; P->split_abs = 1;
    if (form == 0) { P->p = 2.0; P->r = 15.0; P->gamma = 0.4; P->s_turn = 0.014 * 180 / M_PI; P->bvy_max = 0.3; P->max_iter = 20; P->goal_shift = 1; P->close_radius = 0.35; }
    else if (form == 1) { P->p = 0.0; P->r = 50.0; P->gamma = 0.2; P->s_turn = 0.024 * 180 / M_PI; P->bvy_max = 0.35; P->has_fen = 1; P->max_iter = 30; P->select_obs = 1; P->goal_shift = 1; P->close_radius = 0.15; }
    else { P->p = 0.0; P->r = 50.0; P->gamma = 0.2; P->s_turn = 0.024 * 180 / M_PI; P->bvy_max = 0.35; P->has_fen = 1; P->t_smooth = 2.0; P->max_iter = 40; P->close_radius = 0.35; }
}

/* dims after selection: out[0]=n, out[1]=m_reference_rows, out[2]=nc_sel, out[3]=ne_sel; goal_eff[2] */
void orc_setup_info(const orc_params *P, const double *xk, const double *goal, int leg, int nc, const double *cir,
                    int ne, const double *elp, int *out, double *goal_eff) {
    orc_problem pb; prob_setup(&pb, P, xk, goal, leg, nc, cir, ne, elp, NULL);
    out[0] = pb.n; out[1] = pb.mref; out[2] = pb.nc; out[3] = pb.ne;
    goal_eff[0] = pb.goal[0]; goal_eff[1] = pb.goal[1];
}

/* callbacks at u, reference row order; cl/cu too. */
void orc_eval(const orc_params *P, const double *xk, const double *goal, int leg, int nc, const double *cir, int ne,
              const double *elp, const double *last_u, const double *u, double *f, double *grad, double *c, double *jac,
              double *cl, double *cu) {
    orc_problem pb; prob_setup(&pb, P, xk, goal, leg, nc, cir, ne, elp, last_u);
    double cc[MMAX], jj[MMAX * NMAX], l[MMAX], h[MMAX];
    ref_eval(&pb, u, f, grad, cc, jj);
    ref_bounds(&pb, l, h);
    int mref = pb.mref;
    if (c) memcpy(c, cc, mref * sizeof(double));
    if (jac) memcpy(jac, jj, mref * pb.n * sizeof(double));
    if (cl) memcpy(cl, l, mref * sizeof(double));
    if (cu) memcpy(cu, h, mref * sizeof(double));
}

/* one solve.  u0: reference warm start (R^15 / R^6).  outputs: u (R^15/R^6), plan x[3][5|3], p[3][3] (LIP),
 * f, status, iters, viol, close2goal */
void orc_solve(const orc_params *P, const double *xk, const double *goal, int leg, int nc, const double *cir, int ne,
               const double *elp, const double *last_u, const double *u0, double *u_out, double *x_plan, double *p_plan,
               double *f, int *status, int *iters, double *viol, int *close2goal) {
    orc_problem pb; prob_setup(&pb, P, xk, goal, leg, nc, cir, ne, elp, last_u);
    orc_result R; memset(&R, 0, sizeof(R));
    ipm_solve(&pb, u0, &R);
    memcpy(u_out, R.u, pb.n * sizeof(double));
    if (f) *f = R.f; if (status) *status = R.status; if (iters) *iters = R.iters; if (viol) *viol = R.viol;
    int close = 0;
    if (P->form == 2) {
        double x[4][3]; dd_roll(&pb, R.u, x);
        if (x_plan) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) x_plan[3 * i + j] = x[i + 1][j];
        double d = hypot(x[1][0] - goal[0], x[1][1] - goal[1]);
        close = d <= P->close_radius;
    } else {
        double x[4][5], p[3][3]; lip_roll(&pb, R.u, x, p);
        if (x_plan) for (int i = 0; i < 3; i++) for (int j = 0; j < 5; j++) x_plan[5 * i + j] = x[i + 1][j];
        if (p_plan) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) p_plan[3 * i + j] = p[i][j];
        if (P->form == 0) { for (int i = 1; i < 4; i++) if (hypot(x[i][0] - goal[0], x[i][1] - goal[1]) <= P->close_radius) close = 1; }
        else close = hypot(x[1][0] - goal[0], x[1][1] - goal[1]) <= P->close_radius;
    }
    if (close2goal) *close2goal = close;
}

/* batch of independent solves on `threads` host threads (cpu baseline).  Layouts: xk[B][5|3], goal[B][2],
 * leg[B], cir[F][nc][3], elp[F][ne][5], field[B] (index into F, NULL => b), last_u[B][2]|NULL,
 * u0[B][15|6] -> u_out[B][15|6], p_plan[B][9]|NULL, f[B], status[B], iters[B], viol[B] */
typedef struct {
    const orc_params *P; int B, t, nt; const double *xk, *goal; const int *leg; int nc; const double *cir; int ne;
    const double *elp; const int *field; const double *last_u, *u0; double *u_out, *x_plan, *p_plan, *f; int *status, *iters;
    double *viol; int *close;
} batch_arg;

static void *batch_worker(void *vp) {
    batch_arg *a = (batch_arg *)vp;
    int nx = a->P->form == 2 ? 3 : 5, n = a->P->form == 2 ? 6 : 15;
    for (int b = a->t; b < a->B; b += a->nt) {
        int fi = a->field ? a->field[b] : b;
        orc_solve(a->P, a->xk + nx * b, a->goal + 2 * b, a->leg ? a->leg[b] : 1, a->nc, a->cir + (size_t)fi * a->nc * 3, a->ne,
                  a->elp ? a->elp + (size_t)fi * a->ne * 5 : NULL, a->last_u ? a->last_u + 2 * b : NULL, a->u0 + n * b,
                  a->u_out + n * b, a->x_plan ? a->x_plan + 3 * nx * b : NULL, a->p_plan ? a->p_plan + 9 * b : NULL,
                  a->f + b, a->status + b, a->iters + b, a->viol + b, a->close ? a->close + b : NULL);
    }
    return NULL;
}

void orc_solve_batch(const orc_params *P, int B, int threads, const double *xk, const double *goal, const int *leg, int nc,
                     const double *cir, int ne, const double *elp, const int *field, const double *last_u, const double *u0,
                     double *u_out, double *x_plan, double *p_plan, double *f, int *status, int *iters, double *viol, int *close) {
    build_model();
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256]; batch_arg args[256];
    for (int t = 0; t < threads; t++) {
        batch_arg a = { P, B, t, threads, xk, goal, leg, nc, cir, ne, elp, field, last_u, u0, u_out, x_plan, p_plan, f, status, iters, viol, close };
        args[t] = a;
        if (threads == 1) batch_worker(&args[t]); else pthread_create(&th[t], NULL, batch_worker, &args[t]);
    }
    if (threads > 1) for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}

int orc_sizeof_params(void) { return (int)sizeof(orc_params); }
