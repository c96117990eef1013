// dcbf_core.cuh -- per-lane (one problem per thread) D-CBF LIP/DD MPC: problem assembly + interior point method.
//
// Everything here is FP64 and works in the REDUCED decision space: for the LIP formulations z = (p0, p1, p2) with
// p_k = (foot_x, foot_y, dtheta) -- the only combination of the reference's 15 variables that reaches the dynamics,
// cost and constraints (p_k = W(u_k - A x_k), /root/reference/MPC_LIP_sig_step.py:302-306, rank(dx_du) = 9); for DD
// z = u = (v0, w0, v1, w1, v2, w2).  In z the LIP rollout is affine, so every row has a closed-form gradient and
// Hessian with compile-time sparsity.  Internal variable order (LIP): (fx0, fy0, fx1, fy1, fx2, fy2, t0, t1, t2).
//
// What is restated from the reference (file:line in /root/reference):
//   cost          MPC_LIP_sig_step.py:372-386 / MPC_LIP_modi.py:430-444 / MPC_DD_sig_step.py:351-369
//   rows          MPC_LIP_sig_step.py:410-437 / MPC_LIP_modi.py:468-500 / MPC_DD_sig_step.py:399-421
//   bounds        MPC_LIP_sig_step.py:193-227 / MPC_LIP_modi.py:203-245 / MPC_DD_sig_step.py:127-141
//   goal shift    MPC_LIP_sig_step.py:229-253;  obstacle selection MPC_LIP_modi.py:325-338
// The solve (cyipopt -> Ipopt, absent third-party code) is a from-scratch primal-dual interior point method in the
// style of Waechter & Biegler (2006): slack form of the rows, monotone barrier update, fraction-to-boundary rule,
// filter line search, inertia correction by delta*I on the condensed matrix  W + J^T Sigma J, Levenberg-Marquardt
// feasibility restoration that doubles as the infeasibility detector (status 2).
//
// The file compiles for the device (nvcc) and, for CPU-side debugging of the algorithm in tests/hostsim only, for
// the host (g++).  No product entry point runs the host instantiation.
#pragma once
#include <math.h>
#include <stdint.h>
#ifdef DCBF_TRACE
#include <stdio.h>
#endif

#include "../../include/dcbf_mpc.h"
#include "dcbf_math.cuh"

#if defined(__CUDACC__)
#define DCBF_HD __host__ __device__ __forceinline__
#define DCBF_UNROLL _Pragma("unroll")
#define DCBF_CE __host__ __device__ constexpr
#define DCBF_MATH __host__ __device__ __noinline__
#else
#define DCBF_CE constexpr
#define DCBF_MATH inline
#define DCBF_HD inline
#define DCBF_UNROLL
#endif

namespace dcbf {

// ---------------------------------------------------------------------------------------------------------------
// constants
// ---------------------------------------------------------------------------------------------------------------
struct Consts {
    double C, Sb, bS;      // cosh(beta dt), sinh(beta dt)/beta, beta sinh(beta dt)      (MPC_LIP_sig_step.py:47-56)
    double gx[3], gv[3];   // d pos_{l+1+d} / d foot_l  and  d vel_{l+1+d} / d foot_l,  d = 0,1,2
    double dt;
};

inline Consts make_consts() {
    Consts k;
    const double beta = sqrt(9.81 / 1.0), dt = 0.4;
    k.C = cosh(beta * dt);
    const double S = sinh(beta * dt);
    k.Sb = S / beta;
    k.bS = S * beta;
    k.gx[0] = 1.0 - k.C;
    k.gv[0] = -k.bS;
    for (int d = 1; d < 3; d++) {
        k.gx[d] = k.C * k.gx[d - 1] + k.Sb * k.gv[d - 1];
        k.gv[d] = k.bS * k.gx[d - 1] + k.C * k.gv[d - 1];
    }
    k.dt = dt;
    return k;
}

// prepared obstacle records (written by the field-preparation kernel)
#define DCBF_CIR_REC 3  // cx, cy, r^2
#define DCBF_ELP_REC 8  // cx, cy, a', b', c', rhs, rmax^2, pad

DCBF_CE int tri(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

DCBF_CE int FXI(int l) { return 2 * l; }
DCBF_CE int FYI(int l) { return 2 * l + 1; }
DCBF_CE int THI(int l) { return 6 + l; }

#define DCBF_KAPPA_SIGMA 1e10
#ifndef DCBF_KAPPA_MU
#define DCBF_KAPPA_MU 0.2            /* Ipopt mu_linear_decrease_factor */
#endif
#ifndef DCBF_MU_POW
#define DCBF_MU_POW(mu) ((mu) * sqrt(mu))   /* Ipopt mu_superlinear_decrease_power = 1.5 */
#endif
#ifndef DCBF_KAPPA_EPS
#define DCBF_KAPPA_EPS 10.0          /* Ipopt barrier_tol_factor (default of dcbf_params::kappa_eps for the differential drive) */
#endif
#ifndef DCBF_BOUND_PUSH
#define DCBF_BOUND_PUSH 1e-2         /* Ipopt bound_push: initial slacks at least this far (times max(1, |bound|)) inside their bounds ... */
#endif
#ifndef DCBF_BOUND_FRAC
#define DCBF_BOUND_FRAC 1e-2         /* ... Ipopt bound_frac: and at most this fraction of the range for two-sided rows */
#endif
#ifndef DCBF_LM_UP
#define DCBF_LM_UP 10.0              /* restoration: Levenberg-Marquardt parameter after a rejected trial ... */
#endif
#ifndef DCBF_LM_DOWN
#define DCBF_LM_DOWN 0.2             /* ... and after an accepted one */
#endif
#ifndef DCBF_RESTO_WINDOW
#define DCBF_RESTO_WINDOW 1e-2       /* restoration: relative decrease of the squared violation over three steps (round-1 value; now dcbf_params::resto_window) */
#endif
#ifdef DCBF_COUNT
static long g_trials = 0;
#endif
#define DCBF_FILT 8
#define DCBF_LS_MAX 22

enum { MODE_SOLVE = 0, MODE_EVAL = 1 };
enum { PH_MAIN = 0, PH_RESTO = 1 };

DCBF_HD double dmax(double a, double b) { return a > b ? a : b; }
DCBF_HD double dmin(double a, double b) { return a < b ? a : b; }
// Elementary functions.  The warp kernels (dcbf_warp.cuh) inline the lean versions of dcbf_math.cuh.  The per-thread kernels keep
// one out-of-line copy of each routine (they call them from dozens of sites and their register allocation is fragile: the lean
// versions measured 5-10 % slower there, profiles/r01_summary.md), built on the library functions unless DCBF_LEAN_MATH is set.
#if defined(DCBF_LEAN_MATH) || !defined(__CUDA_ARCH__)
DCBF_MATH void dsincos(double a, double *s, double *c) { fsincos(a, s, c); }
DCBF_MATH double datan2(double y, double x) { return fatan2(y, x); }
DCBF_MATH double drcp(double x) { return frcp(x); }
DCBF_MATH double ddiv(double a, double b) { return fdiv(a, b); }
#else
DCBF_MATH void dsincos(double a, double *s, double *c) { sincos(a, s, c); }
DCBF_MATH double datan2(double y, double x) { return atan2(y, x); }
DCBF_MATH double drcp(double x) { return 1.0 / x; }
DCBF_MATH double ddiv(double a, double b) { return a / b; }
#endif
DCBF_MATH double dlog(double x) { return log(x); }
DCBF_HD double drsqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);   // MUFU.RSQ64H + Newton steps, <= 1 ulp; ~4x shorter than sqrt followed by a division
#else
    return 1.0 / sqrt(x);
#endif
}
// alpha * a^2.3 > t^1.1 for a > 0, t >= 0 (switching condition of the filter line search) without pow()
DCBF_HD bool switch_cond(double alpha, double a, double t) {
    if (!(t > 0.0)) return alpha > 0.0;
    return dlog(alpha) + 2.3 * dlog(a) > 1.1 * dlog(t);
}

// ---------------------------------------------------------------------------------------------------------------
// per-problem data
// ---------------------------------------------------------------------------------------------------------------
struct Problem {
    double x0[5];         // LIP: px,py,vx,vy,theta   DD: x,y,theta
    double goal[2];       // goal used by the NLP (after the detour heuristic)
    double goal_raw[2];   // goal as given (close_2_goal uses it)
    double last_u[2];     // DD
    int leg;
    int nc, ne;
    uint32_t mc, me;      // selection masks (bit j = obstacle j is a row of the NLP)
    const double *cir;    // prepared records of this scenario's field
    const double *elp;
};

// ---------------------------------------------------------------------------------------------------------------
// solver state shared by both models
// ---------------------------------------------------------------------------------------------------------------
template <int N>
struct IpmState {
    double z[N], dz[N];
    double mu, sf;
    double alpha, alpha_z;
    double delta_last, lm_lambda, resto_target, resto_entry;
    double theta_max, theta_min;
    double v2_h1, v2_h2;   // squared violation one and two accepted restoration steps ago (windowed stagnation test)
    double filt_th[DCBF_FILT], filt_ph[DCBF_FILT];
    double obj, viol;
    int nf, iters, acc_cnt, status;
    int phase, nstall, tiny, nresto;
    bool pending, reinit, first, done;
};

// accumulators of one full pass
template <int N>
struct Acc {
    double K[N * (N + 1) / 2];
    double q1[N], q2[N], q3[N], grad[N];
    double f;
    double theta, pinf;       // sum |c - s|, max |c - s|
    double cmin, cmax;        // extremes of gap*z (complementarity)
    double zsum;              // sum of bound multipliers
    double logsum, logprod;   // sum log(gaps) = logsum + log(logprod)
    double v2, vmax;          // violation of the original bounds
    int nz, nrows, logcnt;
};

template <int N>
DCBF_HD void acc_reset(Acc<N> &A) {
    DCBF_UNROLL
    for (int i = 0; i < N * (N + 1) / 2; i++) A.K[i] = 0.0;
    DCBF_UNROLL
    for (int i = 0; i < N; i++) { A.q1[i] = 0.0; A.q2[i] = 0.0; A.q3[i] = 0.0; A.grad[i] = 0.0; }
    A.f = 0.0; A.theta = 0.0; A.pinf = 0.0; A.cmin = 1e300; A.cmax = 0.0; A.zsum = 0.0;
    A.logsum = 0.0; A.logprod = 1.0; A.v2 = 0.0; A.vmax = 0.0; A.nz = 0; A.nrows = 0; A.logcnt = 0;
}

struct RowW { double sig, w1, binv, y; };

struct LogAcc { double sum, prod; int cnt; };
DCBF_HD void log_push(LogAcc &L, double gap) {
    L.prod *= gap;
    if (++L.cnt == 6) { L.sum += dlog(L.prod); L.prod = 1.0; L.cnt = 0; }
}
DCBF_HD double log_total(const LogAcc &L) { return L.cnt ? L.sum + dlog(L.prod) : L.sum; }

// control block handed to the row helpers
struct RowCtl {
    double mu, alpha, alpha_z;
    int phase;
    bool pending, reinit;
};

DCBF_HD double relax_lo(double lo) { return lo - 1e-8 * dmax(1.0, fabs(lo)); }
DCBF_HD double relax_hi(double hi) { return hi + 1e-8 * dmax(1.0, fabs(hi)); }

// One row in a full pass (solve mode).  Applies the pending step / (re)initialisation to the row state, then returns
// the weights for the condensed system and updates the statistics.  ds/el/eu slots: on entry the pending step
// (ds, dzl, dzu); on exit r_c = c - s and the reciprocal gaps, consumed by row_dir().
template <bool LO, bool HI, int N>
DCBF_HD RowW row_full(const RowCtl &ctl, Acc<N> &A, LogAcc &LA, double c, double lo, double hi, double &s, double &zl,
                      double &zu, double &ds, double &el, double &eu) {
    RowW o;
    // violation of the original bounds
    double v = 0.0;
    if (LO && c < lo) v = c - lo;
    if (HI && c > hi) v = c - hi;
    A.v2 += v * v;
    A.vmax = dmax(A.vmax, fabs(v));
    A.nrows++;
    if (ctl.phase == PH_RESTO) {
        // Newton on 0.5*sum v^2: Gauss-Newton term (sig) plus the curvature term v * Hess(c_r) (carried by y)
        o.sig = v != 0.0 ? 1.0 : 0.0; o.w1 = v; o.binv = 0.0; o.y = v;
        return o;
    }
    const double lr = LO ? relax_lo(lo) : 0.0, hr = HI ? relax_hi(hi) : 0.0;
    if (ctl.reinit) {
        double sv = c;
        if (LO && HI) {
            const double pl = dmin(DCBF_BOUND_PUSH * dmax(1.0, fabs(lr)), DCBF_BOUND_FRAC * (hr - lr));
            const double pu = dmin(DCBF_BOUND_PUSH * dmax(1.0, fabs(hr)), DCBF_BOUND_FRAC * (hr - lr));
            sv = dmin(dmax(sv, lr + pl), hr - pu);
        } else if (LO) {
            sv = dmax(sv, lr + DCBF_BOUND_PUSH * dmax(1.0, fabs(lr)));
        } else if (HI) {
            sv = dmin(sv, hr - DCBF_BOUND_PUSH * dmax(1.0, fabs(hr)));
        }
        s = sv;
        if (LO) zl = 1.0;
        if (HI) zu = 1.0;
    } else if (ctl.pending) {
        s += ctl.alpha * ds;
        if (LO) {
            const double gap = s - lr;
            double z = zl + ctl.alpha_z * el;
            z = dmax(dmin(z, DCBF_KAPPA_SIGMA * ctl.mu / gap), ctl.mu / (DCBF_KAPPA_SIGMA * gap));
            zl = z;
        }
        if (HI) {
            const double gap = hr - s;
            double z = zu + ctl.alpha_z * eu;
            z = dmax(dmin(z, DCBF_KAPPA_SIGMA * ctl.mu / gap), ctl.mu / (DCBF_KAPPA_SIGMA * gap));
            zu = z;
        }
    }
    const double rc = c - s;
    double sig = 0.0, binv = 0.0, y = 0.0;
    if (LO) {
        const double gap = s - lr, inv = 1.0 / gap;
        sig += zl * inv; binv += inv; y -= zl;
        const double cz = gap * zl;
        A.cmin = dmin(A.cmin, cz); A.cmax = dmax(A.cmax, cz); A.zsum += zl; A.nz++;
        log_push(LA, gap);
        el = inv;
    }
    if (HI) {
        const double gap = hr - s, inv = 1.0 / gap;
        sig += zu * inv; binv -= inv; y += zu;
        const double cz = gap * zu;
        A.cmin = dmin(A.cmin, cz); A.cmax = dmax(A.cmax, cz); A.zsum += zu; A.nz++;
        log_push(LA, gap);
        eu = inv;
    }
    ds = rc;
    A.theta += fabs(rc);
    A.pinf = dmax(A.pinf, fabs(rc));
    o.sig = sig; o.w1 = sig * rc; o.binv = binv; o.y = y;
    return o;
}

struct DirStat { double amax, az, dphi; };

// direction pass for one row: jd = grad_r . dz.  Consumes r_c / reciprocal gaps, leaves (ds, dzl, dzu).
template <bool LO, bool HI>
DCBF_HD void row_dir(double mu, double tau, DirStat &D, double jd, double s, double lo, double hi, double zl, double zu,
                     double &ds, double &el, double &eu) {
    const double d = jd + ds;   // ds slot holds r_c
    ds = d;
    if (LO) {
        const double inv = el, gap = s - relax_lo(lo);
        const double dzl = mu * inv - zl - zl * inv * d;
        el = dzl;
        D.dphi -= mu * d * inv;
        if (d < 0.0) D.amax = dmin(D.amax, -tau * gap / d);
        if (dzl < 0.0) D.az = dmin(D.az, -tau * zl / dzl);
    }
    if (HI) {
        const double inv = eu, gap = relax_hi(hi) - s;
        const double dzu = mu * inv - zu + zu * inv * d;
        eu = dzu;
        D.dphi += mu * d * inv;
        if (d > 0.0) D.amax = dmin(D.amax, tau * gap / d);
        if (dzu < 0.0) D.az = dmin(D.az, -tau * zu / dzu);
    }
}

struct ValStat { double f, theta, v2, vmax; LogAcc la; bool ok; };

template <bool LO, bool HI>
DCBF_HD void row_val(int phase, double alpha, ValStat &V, double c, double lo, double hi, double s, double ds) {
    double v = 0.0;
    if (LO && c < lo) v = c - lo;
    if (HI && c > hi) v = c - hi;
    V.v2 += v * v;
    V.vmax = dmax(V.vmax, fabs(v));
    if (phase == PH_RESTO) return;
    const double st = s + alpha * ds;
    V.theta += fabs(c - st);
    if (LO) { const double gap = st - relax_lo(lo); if (!(gap > 0.0)) V.ok = false; log_push(V.la, gap); }
    if (HI) { const double gap = relax_hi(hi) - st; if (!(gap > 0.0)) V.ok = false; log_push(V.la, gap); }
}

// ---------------------------------------------------------------------------------------------------------------
// dense SPD solve on the packed lower triangle (fully unrolled; N = 9 or 6)
// ---------------------------------------------------------------------------------------------------------------
template <int N>
DCBF_HD bool chol_packed(const double *K, double delta, double *L) {
    bool ok = true;
    DCBF_UNROLL
    for (int j = 0; j < N; j++) {
        double d = K[tri(j, j)] + delta;
        DCBF_UNROLL
        for (int k = 0; k < j; k++) d -= L[tri(j, k)] * L[tri(j, k)];
        if (!(d > 1e-14)) { ok = false; d = 1.0; }
        const double r = drsqrt(d);
        L[tri(j, j)] = r;   // store the reciprocal of the diagonal
        DCBF_UNROLL
        for (int i = j + 1; i < N; i++) {
            double s = K[tri(i, j)];
            DCBF_UNROLL
            for (int k = 0; k < j; k++) s -= L[tri(i, k)] * L[tri(j, k)];
            L[tri(i, j)] = s * r;
        }
    }
    return ok;
}

template <int N>
DCBF_HD void chol_solve_packed(const double *L, double *b) {
    DCBF_UNROLL
    for (int i = 0; i < N; i++) {
        double s = b[i];
        DCBF_UNROLL
        for (int k = 0; k < i; k++) s -= L[tri(i, k)] * b[k];
        b[i] = s * L[tri(i, i)];
    }
    DCBF_UNROLL
    for (int i = N - 1; i >= 0; i--) {
        double s = b[i];
        DCBF_UNROLL
        for (int k = i + 1; k < N; k++) s -= L[tri(k, i)] * b[k];
        b[i] = s * L[tri(i, i)];
    }
}

// ===============================================================================================================
// LIP model
// ===============================================================================================================
struct LipNodes {
    double x[4], y[4], vx[4], vy[4], th[4], cs[4], sn[4];
};

DCBF_HD void lip_rollout(const Consts &k, const double *x0, const double *z, LipNodes &nd) {
    nd.x[0] = x0[0]; nd.y[0] = x0[1]; nd.vx[0] = x0[2]; nd.vy[0] = x0[3]; nd.th[0] = x0[4];
    DCBF_UNROLL
    for (int i = 0; i < 3; i++) {
        const double fx = z[FXI(i)], fy = z[FYI(i)];
        nd.x[i + 1] = k.C * nd.x[i] + k.Sb * nd.vx[i] + k.gx[0] * fx;
        nd.y[i + 1] = k.C * nd.y[i] + k.Sb * nd.vy[i] + k.gx[0] * fy;
        nd.vx[i + 1] = k.bS * nd.x[i] + k.C * nd.vx[i] + k.gv[0] * fx;
        nd.vy[i + 1] = k.bS * nd.y[i] + k.C * nd.vy[i] + k.gv[0] * fy;
        nd.th[i + 1] = nd.th[i] + z[THI(i)];
        dsincos(nd.th[i + 1], &nd.sn[i + 1], &nd.cs[i + 1]);
    }
}

// reference warm start u0 (R^15) -> z0 : p_k = W(u_k - A x_k), x_{k+1} = M_A x_k + M_B u_k = A x_k + B p_k
DCBF_HD void lip_z_from_u(const Consts &k, const double *x0, const double *u, double *z) {
    const double wa = 5.0, wb = 1.0;
    const double den = wa * (k.C - 1.0) * (k.C - 1.0) + wb * k.bS * k.bS;
    const double Ch = -wa * (k.C - 1.0) / den, Sh = -wb * k.bS / den;
    double x = x0[0], y = x0[1], vx = x0[2], vy = x0[3], th = x0[4];
    DCBF_UNROLL
    for (int i = 0; i < 3; i++) {
        const double ax = k.C * x + k.Sb * vx, ay = k.C * y + k.Sb * vy;
        const double avx = k.bS * x + k.C * vx, avy = k.bS * y + k.C * vy;
        const double fx = Ch * (u[5 * i + 0] - ax) + Sh * (u[5 * i + 2] - avx);
        const double fy = Ch * (u[5 * i + 1] - ay) + Sh * (u[5 * i + 3] - avy);
        const double dth = u[5 * i + 4] - th;
        z[FXI(i)] = fx; z[FYI(i)] = fy; z[THI(i)] = dth;
        x = ax + k.gx[0] * fx; y = ay + k.gx[0] * fy;
        vx = avx + k.gv[0] * fx; vy = avy + k.gv[0] * fy;
        th += dth;
    }
}

template <int KT>
struct LipRows {
    // two-sided rows, index 3*i + {0: v_bx, 1: v_by, 2: dtheta}
    double s2[9], zl2[9], zu2[9], d2[9], el2[9], eu2[9];
    // upper-only rows, index 3*i + {0: leg, 1: fen+, 2: fen-}
    double su[9], zu1[9], du[9], eu1[9];
    // lower-only rows (D-CBF), index i*KT + j
    double sc[3 * KT], zc[3 * KT], dc[3 * KT], ec[3 * KT];
};

// accumulate one row into K and the q vectors.  Pattern: foot entries [0, NF) and theta entries THI(0..NT-1).
template <int NF, int NT, int MODE>
DCBF_HD void lip_acc_row(Acc<9> &A, const double *g, const RowW &w) {
    if (MODE == MODE_SOLVE) {
        DCBF_UNROLL
        for (int a = 0; a < NF; a++) {
            const double sa = w.sig * g[a];
            DCBF_UNROLL
            for (int b = 0; b <= a; b++) A.K[tri(a, b)] = fma(sa, g[b], A.K[tri(a, b)]);
            A.q1[a] = fma(g[a], w.w1, A.q1[a]);
            A.q2[a] = fma(g[a], w.binv, A.q2[a]);
            A.q3[a] = fma(g[a], w.y, A.q3[a]);
        }
        DCBF_UNROLL
        for (int a = 0; a < NT; a++) {
            const int ia = THI(a);
            const double sa = w.sig * g[ia];
            DCBF_UNROLL
            for (int b = 0; b < NF; b++) A.K[tri(ia, b)] = fma(sa, g[b], A.K[tri(ia, b)]);
            DCBF_UNROLL
            for (int b = 0; b <= a; b++) A.K[tri(ia, THI(b))] = fma(sa, g[THI(b)], A.K[tri(ia, THI(b))]);
            A.q1[ia] = fma(g[ia], w.w1, A.q1[ia]);
            A.q2[ia] = fma(g[ia], w.binv, A.q2[ia]);
            A.q3[ia] = fma(g[ia], w.y, A.q3[ia]);
        }
    }
}

// K += pull-back of a node Hessian over (x_k, y_k, theta_k), node k = I+1 (depends on foot/turn 0..I)
template <int I>
DCBF_HD void lip_add_node_hess(const Consts &k, double *K, double hxx, double hxy, double hyy, double hxt, double hyt,
                               double htt) {
    DCBF_UNROLL
    for (int l = 0; l <= I; l++) {
        const double gl = k.gx[I - l];
        DCBF_UNROLL
        for (int m = 0; m <= I; m++) {
            const double gg = gl * k.gx[I - m];
            if (m <= l) {
                K[tri(FXI(l), FXI(m))] = fma(gg, hxx, K[tri(FXI(l), FXI(m))]);
                K[tri(FYI(l), FYI(m))] = fma(gg, hyy, K[tri(FYI(l), FYI(m))]);
            }
            K[tri(FYI(l), FXI(m))] = fma(gg, hxy, K[tri(FYI(l), FXI(m))]);   // each (y_l, x_m) pair once
        }
        DCBF_UNROLL
        for (int a = 0; a <= I; a++) {
            K[tri(THI(a), FXI(l))] = fma(gl, hxt, K[tri(THI(a), FXI(l))]);
            K[tri(THI(a), FYI(l))] = fma(gl, hyt, K[tri(THI(a), FYI(l))]);
        }
    }
    DCBF_UNROLL
    for (int a = 0; a <= I; a++) {
        DCBF_UNROLL
        for (int b = 0; b <= a; b++) K[tri(THI(a), THI(b))] += htt;
    }
}

// position-only pull-back for node NODE in {1,2,3} (node 0 is constant)
template <int NODE>
DCBF_HD void lip_add_pos_hess(const Consts &k, double *K, double hxx, double hxy, double hyy) {
    DCBF_UNROLL
    for (int l = 0; l < NODE; l++) {
        const double gl = k.gx[NODE - 1 - l];
        DCBF_UNROLL
        for (int m = 0; m < NODE; m++) {
            const double gg = gl * k.gx[NODE - 1 - m];
            if (m <= l) {
                K[tri(FXI(l), FXI(m))] = fma(gg, hxx, K[tri(FXI(l), FXI(m))]);
                K[tri(FYI(l), FYI(m))] = fma(gg, hyy, K[tri(FYI(l), FYI(m))]);
            }
            K[tri(FYI(l), FXI(m))] = fma(gg, hxy, K[tri(FYI(l), FXI(m))]);
        }
    }
}
template <>
DCBF_HD void lip_add_pos_hess<0>(const Consts &, double *, double, double, double) {}

// outputs of the evaluation kernel (reference row order, reference p ordering of the variables)
struct EvalOut {
    double *c, *jac, *cl, *cu;   // row-major [m], [m][9]; may be NULL
    const double *lambda;        // [m] or NULL
    int row;                     // running reference row index
};

DCBF_HD int lip_ref_var(int internal) {   // internal index -> reference p index
    return internal < 6 ? 3 * (internal >> 1) + (internal & 1) : 3 * (internal - 6) + 2;
}

template <int NF, int NT>
DCBF_HD void lip_eval_emit(EvalOut &E, double c, const double *g, double lo, double hi) {
    const int r = E.row++;
    if (E.c) E.c[r] = c;
    if (E.cl) E.cl[r] = lo;
    if (E.cu) E.cu[r] = hi;
    if (E.jac) {
        for (int j = 0; j < 9; j++) E.jac[9 * r + j] = 0.0;
        DCBF_UNROLL
        for (int a = 0; a < NF; a++) E.jac[9 * r + lip_ref_var(a)] = g[a];
        DCBF_UNROLL
        for (int a = 0; a < NT; a++) E.jac[9 * r + lip_ref_var(THI(a))] = g[THI(a)];
    }
}
DCBF_HD double eval_lambda(const EvalOut &E) { return E.lambda ? E.lambda[E.row] : 0.0; }


DCBF_HD double gxs(const Consts &k, int d) { return d >= 0 ? k.gx[d] : 0.0; }

// gradient of a D-CBF row of step I from the level-set gradients at both ends (foot entries only)
template <int I>
DCBF_HD void lip_cbf_grad(const Consts &k, double h1x, double h1y, double h0x, double h0y, double *g) {
    DCBF_UNROLL
    for (int l = 0; l <= I; l++) {
        g[FXI(l)] = k.gx[I - l] * h1x + gxs(k, I - 1 - l) * h0x;
        g[FYI(l)] = k.gx[I - l] * h1y + gxs(k, I - 1 - l) * h0y;
    }
}

// one step of the horizon in a full pass
template <int I, int MODE, int KT>
DCBF_HD void lip_full_step(const Consts &k, const dcbf_params &P, const Problem &pb, const LipNodes &nd, const double *z,
                           double sf, const RowCtl &ctl, LipRows<KT> &R, Acc<9> &A, LogAcc &LA, EvalOut *E) {
    constexpr int kn = I + 1;
    constexpr int NF = 2 * (I + 1), NT = I + 1;
    const double INF = 1e300;
    const bool hess = true;   // restoration uses the curvature of the violated rows too (objective scaled by sf = 0)
    // ---- objective at node kn -------------------------------------------------------------------------------
    {
        const double w = P.w_q + (I == 0 ? P.w_p : 0.0);
        const double ex = nd.x[kn] - pb.goal[0], ey = nd.y[kn] - pb.goal[1];
        const double dx = -ex, dy = -ey;
        const double r2 = dx * dx + dy * dy, ir2 = 1.0 / r2;
        const double phi = nd.th[kn] - datan2(dy, dx);
        A.f += w * (ex * ex + ey * ey) + P.w_r * phi * phi;
        const double px = -dy * ir2, py = dx * ir2;           // d phi / d(x_k, y_k)
        const double nx = 2.0 * w * ex + 2.0 * P.w_r * phi * px;
        const double ny = 2.0 * w * ey + 2.0 * P.w_r * phi * py;
        const double nt = 2.0 * P.w_r * phi;
        DCBF_UNROLL
        for (int l = 0; l <= I; l++) {
            A.grad[FXI(l)] = fma(k.gx[I - l], nx, A.grad[FXI(l)]);
            A.grad[FYI(l)] = fma(k.gx[I - l], ny, A.grad[FYI(l)]);
            A.grad[THI(l)] += nt;
        }
        if (hess) {
            const double ir4 = ir2 * ir2;
            const double pxx = -2.0 * dx * dy * ir4, pyy = -pxx, pxy = (dx * dx - dy * dy) * ir4;
            const double r2w = 2.0 * P.w_r * sf;
            lip_add_node_hess<I>(k, A.K, sf * 2.0 * w + r2w * (px * px + phi * pxx), r2w * (px * py + phi * pxy),
                                 sf * 2.0 * w + r2w * (py * py + phi * pyy), r2w * px, r2w * py, r2w);
        }
    }
    // ---- body-velocity rows -----------------------------------------------------------------------------------
    const double cs = nd.cs[kn], sn = nd.sn[kn];
    const double vbx = cs * nd.vx[kn] + sn * nd.vy[kn], vby = -sn * nd.vx[kn] + cs * nd.vy[kn];
    const bool plus = (pb.leg > 0) == ((I & 1) == 0);
    const double vy_lo = plus ? P.bvy_min : -P.bvy_max, vy_hi = plus ? P.bvy_max : -P.bvy_min;
    double gvx[9], gvy[9];
    DCBF_UNROLL
    for (int l = 0; l <= I; l++) {
        gvx[FXI(l)] = cs * k.gv[I - l]; gvx[FYI(l)] = sn * k.gv[I - l]; gvx[THI(l)] = vby;
        gvy[FXI(l)] = -sn * k.gv[I - l]; gvy[FYI(l)] = cs * k.gv[I - l]; gvy[THI(l)] = -vbx;
    }
    double Yx, Yy;   // multipliers weighting the Hessians of v_bx and v_by
    if (MODE == MODE_SOLVE) {
        RowW wx = row_full<true, true>(ctl, A, LA, vbx, P.bvx_min, P.bvx_max, R.s2[3 * I], R.zl2[3 * I], R.zu2[3 * I],
                                       R.d2[3 * I], R.el2[3 * I], R.eu2[3 * I]);
        lip_acc_row<NF, NT, MODE>(A, gvx, wx);
        RowW wy = row_full<true, true>(ctl, A, LA, vby, vy_lo, vy_hi, R.s2[3 * I + 1], R.zl2[3 * I + 1], R.zu2[3 * I + 1],
                                       R.d2[3 * I + 1], R.el2[3 * I + 1], R.eu2[3 * I + 1]);
        lip_acc_row<NF, NT, MODE>(A, gvy, wy);
        Yx = wx.y; Yy = wy.y;
    } else {
        Yx = eval_lambda(*E); lip_eval_emit<NF, NT>(*E, vbx, gvx, P.bvx_min, P.bvx_max);
        Yy = eval_lambda(*E); lip_eval_emit<NF, NT>(*E, vby, gvy, vy_lo, vy_hi);
    }
    // ---- D-CBF rows -------------------------------------------------------------------------------------------
    double qxx = 0.0, qxy = 0.0, qyy = 0.0;   // sum_r y_r Q_r
    const double gm1 = P.gamma - 1.0;
    for (int j = 0; j < pb.nc; j++) {
        if (MODE == MODE_SOLVE && !((pb.mc >> j) & 1u)) continue;
        const double *o = pb.cir + DCBF_CIR_REC * j;
        const double ax = nd.x[kn] - o[0], ay = nd.y[kn] - o[1], bx = nd.x[I] - o[0], by = nd.y[I] - o[1];
        const double c = (ax * ax + ay * ay - o[2]) + gm1 * (bx * bx + by * by - o[2]);
        double g[9];
        lip_cbf_grad<I>(k, 2.0 * ax, 2.0 * ay, 2.0 * gm1 * bx, 2.0 * gm1 * by, g);
        double y;
        if (MODE == MODE_SOLVE) {
            const int ri = I * KT + j;
            RowW w = row_full<true, false>(ctl, A, LA, c, 0.0, INF, R.sc[ri], R.zc[ri], R.zc[ri], R.dc[ri], R.ec[ri], R.ec[ri]);
            lip_acc_row<NF, 0, MODE>(A, g, w);
            y = w.y;
        } else {
            y = eval_lambda(*E);
            lip_eval_emit<NF, 0>(*E, c, g, 0.0, INFINITY);
        }
        qxx += 2.0 * y; qyy += 2.0 * y;
    }
    for (int j = 0; j < pb.ne; j++) {
        if (MODE == MODE_SOLVE && !((pb.me >> j) & 1u)) continue;
        const double *o = pb.elp + DCBF_ELP_REC * j;
        const double ax = nd.x[kn] - o[0], ay = nd.y[kn] - o[1], bx = nd.x[I] - o[0], by = nd.y[I] - o[1];
        const double ea = o[2], eb = o[3], ec = o[4];
        const double c = (ea * ax * ax + eb * ax * ay + ec * ay * ay - o[5])
                       + gm1 * (ea * bx * bx + eb * bx * by + ec * by * by - o[5]);
        double g[9];
        lip_cbf_grad<I>(k, 2.0 * ea * ax + eb * ay, 2.0 * ec * ay + eb * ax, gm1 * (2.0 * ea * bx + eb * by),
                        gm1 * (2.0 * ec * by + eb * bx), g);
        double y;
        if (MODE == MODE_SOLVE) {
            const int ri = I * KT + pb.nc + j;
            RowW w = row_full<true, false>(ctl, A, LA, c, 0.0, INF, R.sc[ri], R.zc[ri], R.zc[ri], R.dc[ri], R.ec[ri], R.ec[ri]);
            lip_acc_row<NF, 0, MODE>(A, g, w);
            y = w.y;
        } else {
            y = eval_lambda(*E);
            lip_eval_emit<NF, 0>(*E, c, g, 0.0, INFINITY);
        }
        qxx += 2.0 * ea * y; qxy += eb * y; qyy += 2.0 * ec * y;
    }
    if (hess) {
        lip_add_pos_hess<I + 1>(k, A.K, qxx, qxy, qyy);
        lip_add_pos_hess<I>(k, A.K, gm1 * qxx, gm1 * qxy, gm1 * qyy);
    }
    // ---- leg length row: |pos_I - foot_I|^2 <= leg_sq -------------------------------------------------------------
    {
        const double lx = nd.x[I] - z[FXI(I)], ly = nd.y[I] - z[FYI(I)];
        const double c = lx * lx + ly * ly;
        double e[I + 1], g[9];   // d(pos_I - foot_I)/d foot_l (same for x and y)
        DCBF_UNROLL
        for (int l = 0; l < I; l++) e[l] = k.gx[I - 1 - l];
        e[I] = -1.0;
        DCBF_UNROLL
        for (int l = 0; l <= I; l++) { g[FXI(l)] = 2.0 * lx * e[l]; g[FYI(l)] = 2.0 * ly * e[l]; }
        double y;
        if (MODE == MODE_SOLVE) {
            RowW w = row_full<false, true>(ctl, A, LA, c, -INF, P.leg_sq, R.su[3 * I], R.zu1[3 * I], R.zu1[3 * I], R.du[3 * I],
                                           R.eu1[3 * I], R.eu1[3 * I]);
            lip_acc_row<NF, 0, MODE>(A, g, w);
            y = w.y;
        } else {
            y = eval_lambda(*E);
            lip_eval_emit<NF, 0>(*E, c, g, 0.0, P.leg_sq);
        }
        if (hess) {
            DCBF_UNROLL
            for (int l = 0; l <= I; l++) {
                DCBF_UNROLL
                for (int m = 0; m <= l; m++) {
                    const double v = 2.0 * y * e[l] * e[m];
                    A.K[tri(FXI(l), FXI(m))] += v;
                    A.K[tri(FYI(l), FYI(m))] += v;
                }
            }
        }
    }
    // ---- turn row and speed/turn coupling ---------------------------------------------------------------------------
    {
        const double dth = z[THI(I)];
        double g[9];
        DCBF_UNROLL
        for (int l = 0; l <= I; l++) { g[FXI(l)] = 0.0; g[FYI(l)] = 0.0; g[THI(l)] = 0.0; }
        g[THI(I)] = 1.0;
        if (MODE == MODE_SOLVE) {
            RowW w = row_full<true, true>(ctl, A, LA, dth, -P.ang_max, P.ang_max, R.s2[3 * I + 2], R.zl2[3 * I + 2],
                                          R.zu2[3 * I + 2], R.d2[3 * I + 2], R.el2[3 * I + 2], R.eu2[3 * I + 2]);
            // single-entry row: accumulate by hand
            A.K[tri(THI(I), THI(I))] += w.sig;
            A.q1[THI(I)] += w.w1; A.q2[THI(I)] += w.binv; A.q3[THI(I)] += w.y;
        } else {
            lip_eval_emit<NF, NT>(*E, dth, g, -P.ang_max, P.ang_max);
        }
        if (P.has_fen) {
            if (MODE == MODE_SOLVE) {
                // smooth split of  s|dth| + v_bx <= v_max :  v_bx + s dth <= v_max  and  v_bx - s dth <= v_max
                double gp[9], gm[9];
                DCBF_UNROLL
                for (int l = 0; l <= I; l++) {
                    gp[FXI(l)] = gvx[FXI(l)]; gp[FYI(l)] = gvx[FYI(l)]; gp[THI(l)] = gvx[THI(l)];
                    gm[FXI(l)] = gvx[FXI(l)]; gm[FYI(l)] = gvx[FYI(l)]; gm[THI(l)] = gvx[THI(l)];
                }
                gp[THI(I)] += P.s_turn; gm[THI(I)] -= P.s_turn;
                RowW wp = row_full<false, true>(ctl, A, LA, vbx + P.s_turn * dth, -INF, P.bvx_max, R.su[3 * I + 1], R.zu1[3 * I + 1],
                                                R.zu1[3 * I + 1], R.du[3 * I + 1], R.eu1[3 * I + 1], R.eu1[3 * I + 1]);
                lip_acc_row<NF, NT, MODE>(A, gp, wp);
                RowW wm = row_full<false, true>(ctl, A, LA, vbx - P.s_turn * dth, -INF, P.bvx_max, R.su[3 * I + 2], R.zu1[3 * I + 2],
                                                R.zu1[3 * I + 2], R.du[3 * I + 2], R.eu1[3 * I + 2], R.eu1[3 * I + 2]);
                lip_acc_row<NF, NT, MODE>(A, gm, wm);
                Yx += wp.y + wm.y;
            } else {
                // literal reference row (MPC_LIP_modi.py:493, 637-643)
                const double sg = dth == 0.0 ? 0.0 : (dth > 0.0 ? P.s_turn : -P.s_turn);
                double gf[9];
                DCBF_UNROLL
                for (int l = 0; l <= I; l++) { gf[FXI(l)] = gvx[FXI(l)]; gf[FYI(l)] = gvx[FYI(l)]; gf[THI(l)] = gvx[THI(l)]; }
                gf[THI(I)] += sg;
                Yx += eval_lambda(*E);
                lip_eval_emit<NF, NT>(*E, P.s_turn * fabs(dth) + vbx, gf, P.bvx_min, P.bvx_max);
            }
        }
    }
    // ---- Hessians of the rotated-velocity rows (v_bx carries Yx, v_by carries Yy) ------------------------------------
    if (hess) {
        const double hfx = -sn * Yx - cs * Yy, hfy = cs * Yx - sn * Yy, htt = -(Yx * vbx + Yy * vby);
        DCBF_UNROLL
        for (int a = 0; a <= I; a++) {
            DCBF_UNROLL
            for (int l = 0; l <= I; l++) {
                A.K[tri(THI(a), FXI(l))] = fma(k.gv[I - l], hfx, A.K[tri(THI(a), FXI(l))]);
                A.K[tri(THI(a), FYI(l))] = fma(k.gv[I - l], hfy, A.K[tri(THI(a), FYI(l))]);
            }
            DCBF_UNROLL
            for (int b = 0; b <= a; b++) A.K[tri(THI(a), THI(b))] += htt;
        }
    }
}

// node directions J_node dz for the direction pass
struct LipDir { double dx[4], dy[4], dvx[4], dvy[4], dth[4]; };

DCBF_HD void lip_dir_nodes(const Consts &k, const double *dz, LipDir &d) {
    d.dx[0] = d.dy[0] = d.dvx[0] = d.dvy[0] = d.dth[0] = 0.0;
    DCBF_UNROLL
    for (int i = 0; i < 3; i++) {
        d.dx[i + 1] = k.C * d.dx[i] + k.Sb * d.dvx[i] + k.gx[0] * dz[FXI(i)];
        d.dy[i + 1] = k.C * d.dy[i] + k.Sb * d.dvy[i] + k.gx[0] * dz[FYI(i)];
        d.dvx[i + 1] = k.bS * d.dx[i] + k.C * d.dvx[i] + k.gv[0] * dz[FXI(i)];
        d.dvy[i + 1] = k.bS * d.dy[i] + k.C * d.dvy[i] + k.gv[0] * dz[FYI(i)];
        d.dth[i + 1] = d.dth[i] + dz[THI(i)];
    }
}

template <int I, int KT>
DCBF_HD void lip_dir_step(const Consts &k, const dcbf_params &P, const Problem &pb, const LipNodes &nd, const double *z,
                          const double *dz, const LipDir &d, double mu, double tau, LipRows<KT> &R, DirStat &D) {
    constexpr int kn = I + 1;
    const double INF = 1e300;
    const double cs = nd.cs[kn], sn = nd.sn[kn];
    const double vbx = cs * nd.vx[kn] + sn * nd.vy[kn], vby = -sn * nd.vx[kn] + cs * nd.vy[kn];
    const bool plus = (pb.leg > 0) == ((I & 1) == 0);
    const double vy_lo = plus ? P.bvy_min : -P.bvy_max, vy_hi = plus ? P.bvy_max : -P.bvy_min;
    const double jvx = cs * d.dvx[kn] + sn * d.dvy[kn] + vby * d.dth[kn];
    const double jvy = -sn * d.dvx[kn] + cs * d.dvy[kn] - vbx * d.dth[kn];
    row_dir<true, true>(mu, tau, D, jvx, R.s2[3 * I], P.bvx_min, P.bvx_max, R.zl2[3 * I], R.zu2[3 * I], R.d2[3 * I],
                        R.el2[3 * I], R.eu2[3 * I]);
    row_dir<true, true>(mu, tau, D, jvy, R.s2[3 * I + 1], vy_lo, vy_hi, R.zl2[3 * I + 1], R.zu2[3 * I + 1], R.d2[3 * I + 1],
                        R.el2[3 * I + 1], R.eu2[3 * I + 1]);
    const double gm1 = P.gamma - 1.0;
    for (int j = 0; j < pb.nc; j++) {
        if (!((pb.mc >> j) & 1u)) continue;
        const double *o = pb.cir + DCBF_CIR_REC * j;
        const double ax = nd.x[kn] - o[0], ay = nd.y[kn] - o[1], bx = nd.x[I] - o[0], by = nd.y[I] - o[1];
        const double jd = 2.0 * (ax * d.dx[kn] + ay * d.dy[kn]) + 2.0 * gm1 * (bx * d.dx[I] + by * d.dy[I]);
        const int ri = I * KT + j;
        row_dir<true, false>(mu, tau, D, jd, R.sc[ri], 0.0, INF, R.zc[ri], 0.0, R.dc[ri], R.ec[ri], R.ec[ri]);
    }
    for (int j = 0; j < pb.ne; j++) {
        if (!((pb.me >> j) & 1u)) continue;
        const double *o = pb.elp + DCBF_ELP_REC * j;
        const double ax = nd.x[kn] - o[0], ay = nd.y[kn] - o[1], bx = nd.x[I] - o[0], by = nd.y[I] - o[1];
        const double ea = o[2], eb = o[3], ec = o[4];
        const double jd = (2.0 * ea * ax + eb * ay) * d.dx[kn] + (2.0 * ec * ay + eb * ax) * d.dy[kn]
                        + gm1 * ((2.0 * ea * bx + eb * by) * d.dx[I] + (2.0 * ec * by + eb * bx) * d.dy[I]);
        const int ri = I * KT + pb.nc + j;
        row_dir<true, false>(mu, tau, D, jd, R.sc[ri], 0.0, INF, R.zc[ri], 0.0, R.dc[ri], R.ec[ri], R.ec[ri]);
    }
    {
        const double lx = nd.x[I] - z[FXI(I)], ly = nd.y[I] - z[FYI(I)];
        const double jd = 2.0 * lx * (d.dx[I] - dz[FXI(I)]) + 2.0 * ly * (d.dy[I] - dz[FYI(I)]);
        row_dir<false, true>(mu, tau, D, jd, R.su[3 * I], -INF, P.leg_sq, 0.0, R.zu1[3 * I], R.du[3 * I], R.eu1[3 * I], R.eu1[3 * I]);
    }
    row_dir<true, true>(mu, tau, D, dz[THI(I)], R.s2[3 * I + 2], -P.ang_max, P.ang_max, R.zl2[3 * I + 2], R.zu2[3 * I + 2],
                        R.d2[3 * I + 2], R.el2[3 * I + 2], R.eu2[3 * I + 2]);
    if (P.has_fen) {
        row_dir<false, true>(mu, tau, D, jvx + P.s_turn * dz[THI(I)], R.su[3 * I + 1], -INF, P.bvx_max, 0.0, R.zu1[3 * I + 1],
                             R.du[3 * I + 1], R.eu1[3 * I + 1], R.eu1[3 * I + 1]);
        row_dir<false, true>(mu, tau, D, jvx - P.s_turn * dz[THI(I)], R.su[3 * I + 2], -INF, P.bvx_max, 0.0, R.zu1[3 * I + 2],
                             R.du[3 * I + 2], R.eu1[3 * I + 2], R.eu1[3 * I + 2]);
    }
}

template <int I, int KT>
DCBF_HD void lip_val_step(const dcbf_params &P, const Problem &pb, const LipNodes &nd, const double *z, int phase,
                          double alpha, const LipRows<KT> &R, ValStat &V) {
    constexpr int kn = I + 1;
    const double INF = 1e300;
    {
        const double w = P.w_q + (I == 0 ? P.w_p : 0.0);
        const double ex = nd.x[kn] - pb.goal[0], ey = nd.y[kn] - pb.goal[1];
        const double phi = nd.th[kn] - datan2(-ey, -ex);
        V.f += w * (ex * ex + ey * ey) + P.w_r * phi * phi;
    }
    const double cs = nd.cs[kn], sn = nd.sn[kn];
    const double vbx = cs * nd.vx[kn] + sn * nd.vy[kn], vby = -sn * nd.vx[kn] + cs * nd.vy[kn];
    const bool plus = (pb.leg > 0) == ((I & 1) == 0);
    const double vy_lo = plus ? P.bvy_min : -P.bvy_max, vy_hi = plus ? P.bvy_max : -P.bvy_min;
    row_val<true, true>(phase, alpha, V, vbx, P.bvx_min, P.bvx_max, R.s2[3 * I], R.d2[3 * I]);
    row_val<true, true>(phase, alpha, V, vby, vy_lo, vy_hi, R.s2[3 * I + 1], R.d2[3 * I + 1]);
    const double gm1 = P.gamma - 1.0;
    for (int j = 0; j < pb.nc; j++) {
        if (!((pb.mc >> j) & 1u)) continue;
        const double *o = pb.cir + DCBF_CIR_REC * j;
        const double ax = nd.x[kn] - o[0], ay = nd.y[kn] - o[1], bx = nd.x[I] - o[0], by = nd.y[I] - o[1];
        const double c = (ax * ax + ay * ay - o[2]) + gm1 * (bx * bx + by * by - o[2]);
        row_val<true, false>(phase, alpha, V, c, 0.0, INF, R.sc[I * KT + j], R.dc[I * KT + j]);
    }
    for (int j = 0; j < pb.ne; j++) {
        if (!((pb.me >> j) & 1u)) continue;
        const double *o = pb.elp + DCBF_ELP_REC * j;
        const double ax = nd.x[kn] - o[0], ay = nd.y[kn] - o[1], bx = nd.x[I] - o[0], by = nd.y[I] - o[1];
        const double ea = o[2], eb = o[3], ec = o[4];
        const double c = (ea * ax * ax + eb * ax * ay + ec * ay * ay - o[5])
                       + gm1 * (ea * bx * bx + eb * bx * by + ec * by * by - o[5]);
        row_val<true, false>(phase, alpha, V, c, 0.0, INF, R.sc[I * KT + pb.nc + j], R.dc[I * KT + pb.nc + j]);
    }
    {
        const double lx = nd.x[I] - z[FXI(I)], ly = nd.y[I] - z[FYI(I)];
        row_val<false, true>(phase, alpha, V, lx * lx + ly * ly, -INF, P.leg_sq, R.su[3 * I], R.du[3 * I]);
    }
    const double dth = z[THI(I)];
    row_val<true, true>(phase, alpha, V, dth, -P.ang_max, P.ang_max, R.s2[3 * I + 2], R.d2[3 * I + 2]);
    if (P.has_fen) {
        row_val<false, true>(phase, alpha, V, vbx + P.s_turn * dth, -INF, P.bvx_max, R.su[3 * I + 1], R.du[3 * I + 1]);
        row_val<false, true>(phase, alpha, V, vbx - P.s_turn * dth, -INF, P.bvx_max, R.su[3 * I + 2], R.du[3 * I + 2]);
    }
}

// The LIP model: problem + row state + the three passes the solver needs.
template <int KT>
struct LipModel {
    static constexpr int N = 9;
    Problem pb;
    LipRows<KT> R;
    LipNodes nd;     // nodes of the current iterate (valid after pass_full)

    DCBF_HD void pass_full(const Consts &k, const dcbf_params &P, const IpmState<9> &S, Acc<9> &A, LogAcc &LA) {
        RowCtl ctl;
        ctl.mu = S.mu; ctl.alpha = S.alpha; ctl.alpha_z = S.alpha_z; ctl.phase = S.phase; ctl.pending = S.pending;
        ctl.reinit = S.reinit;
        acc_reset(A);
        LA.sum = 0.0; LA.prod = 1.0; LA.cnt = 0;
        lip_rollout(k, pb.x0, S.z, nd);
        const double sf = S.phase == PH_RESTO ? 0.0 : S.sf;
        lip_full_step<0, MODE_SOLVE, KT>(k, P, pb, nd, S.z, sf, ctl, R, A, LA, nullptr);
        lip_full_step<1, MODE_SOLVE, KT>(k, P, pb, nd, S.z, sf, ctl, R, A, LA, nullptr);
        lip_full_step<2, MODE_SOLVE, KT>(k, P, pb, nd, S.z, sf, ctl, R, A, LA, nullptr);
    }
    // max |d f / d z| at the start point (objective scaling); no rows, no trigonometry of the headings
    DCBF_HD double grad_inf(const Consts &k, const dcbf_params &P, const IpmState<9> &S) const {
        double x = pb.x0[0], y = pb.x0[1], vx = pb.x0[2], vy = pb.x0[3], th = pb.x0[4];
        double g[9];
        DCBF_UNROLL
        for (int i = 0; i < 9; i++) g[i] = 0.0;
        DCBF_UNROLL
        for (int i = 0; i < 3; i++) {
            const double fx = S.z[FXI(i)], fy = S.z[FYI(i)];
            const double xn = k.C * x + k.Sb * vx + k.gx[0] * fx, yn = k.C * y + k.Sb * vy + k.gx[0] * fy;
            vx = k.bS * x + k.C * vx + k.gv[0] * fx; vy = k.bS * y + k.C * vy + k.gv[0] * fy;
            x = xn; y = yn; th += S.z[THI(i)];
            const double w = P.w_q + (i == 0 ? P.w_p : 0.0);
            const double ex = x - pb.goal[0], ey = y - pb.goal[1];
            const double ir2 = 1.0 / (ex * ex + ey * ey);
            const double phi = th - datan2(-ey, -ex);
            const double nx = 2.0 * w * ex + 2.0 * P.w_r * phi * (ey * ir2), ny = 2.0 * w * ey + 2.0 * P.w_r * phi * (-ex * ir2);
            const double nt = 2.0 * P.w_r * phi;
            DCBF_UNROLL
            for (int l = 0; l <= i; l++) {
                g[FXI(l)] = fma(k.gx[i - l], nx, g[FXI(l)]);
                g[FYI(l)] = fma(k.gx[i - l], ny, g[FYI(l)]);
                g[THI(l)] += nt;
            }
        }
        double m = 0.0;
        DCBF_UNROLL
        for (int i = 0; i < 9; i++) m = dmax(m, fabs(g[i]));
        return m;
    }
    DCBF_HD void pass_dir(const Consts &k, const dcbf_params &P, const IpmState<9> &S, double tau, DirStat &D) {
        LipDir d;
        lip_dir_nodes(k, S.dz, d);
        lip_dir_step<0, KT>(k, P, pb, nd, S.z, S.dz, d, S.mu, tau, R, D);
        lip_dir_step<1, KT>(k, P, pb, nd, S.z, S.dz, d, S.mu, tau, R, D);
        lip_dir_step<2, KT>(k, P, pb, nd, S.z, S.dz, d, S.mu, tau, R, D);
    }
    DCBF_HD void pass_value(const Consts &k, const dcbf_params &P, const IpmState<9> &S, double alpha, ValStat &V) const {
        double zt[9];
        DCBF_UNROLL
        for (int i = 0; i < 9; i++) zt[i] = fma(alpha, S.dz[i], S.z[i]);
        LipNodes nt;
        lip_rollout(k, pb.x0, zt, nt);
        lip_val_step<0, KT>(P, pb, nt, zt, S.phase, alpha, R, V);
        lip_val_step<1, KT>(P, pb, nt, zt, S.phase, alpha, R, V);
        lip_val_step<2, KT>(P, pb, nt, zt, S.phase, alpha, R, V);
    }
};

// ===============================================================================================================
// DD (differential drive / unicycle) model:  z = (v0, w0, v1, w1, v2, w2),  state (x, y, theta)
//   x+ = x + dt v cos(theta), y+ = y + dt v sin(theta), theta+ = theta + w      (MPC_DD_sig_step.py:356-363)
// ===============================================================================================================
struct DdNodes {
    double x[4], y[4], th[4], cs[3], sn[3];
    double Jx[4][6], Jy[4][6];   // d pos_k / dz  (node k depends on z[0 .. 2k))
};

DCBF_HD void dd_rollout(const Consts &k, const double *x0, const double *z, DdNodes &nd, bool jac) {
    nd.x[0] = x0[0]; nd.y[0] = x0[1]; nd.th[0] = x0[2];
    if (jac) {
        DCBF_UNROLL
        for (int j = 0; j < 6; j++) { nd.Jx[0][j] = 0.0; nd.Jy[0][j] = 0.0; }
    }
    DCBF_UNROLL
    for (int i = 0; i < 3; i++) {
        dsincos(nd.th[i], &nd.sn[i], &nd.cs[i]);
        const double v = z[2 * i], w = z[2 * i + 1];
        nd.x[i + 1] = fma(k.dt * nd.cs[i], v, nd.x[i]);
        nd.y[i + 1] = fma(k.dt * nd.sn[i], v, nd.y[i]);
        nd.th[i + 1] = nd.th[i] + w;
        if (jac) {
            DCBF_UNROLL
            for (int j = 0; j < 6; j++) { nd.Jx[i + 1][j] = nd.Jx[i][j]; nd.Jy[i + 1][j] = nd.Jy[i][j]; }
            nd.Jx[i + 1][2 * i] += k.dt * nd.cs[i];
            nd.Jy[i + 1][2 * i] += k.dt * nd.sn[i];
            DCBF_UNROLL
            for (int l = 0; l < i; l++) {   // theta_i depends on w_l, l < i
                nd.Jx[i + 1][2 * l + 1] -= k.dt * v * nd.sn[i];
                nd.Jy[i + 1][2 * l + 1] += k.dt * v * nd.cs[i];
            }
        }
    }
}

template <int KT>
struct DdRows {
    double s2[6], zl2[6], zu2[6], d2[6], el2[6], eu2[6];   // variable bounds v_i, w_i (index 2i, 2i+1)
    double su[6], zu1[6], du[6], eu1[6];                   // fen+ / fen-  (index 2i, 2i+1)
    double sc[3 * KT], zc[3 * KT], dc[3 * KT], ec[3 * KT]; // D-CBF rows
};

template <int NV>
DCBF_HD void dd_acc_row(Acc<6> &A, const double *g, const RowW &w) {
    DCBF_UNROLL
    for (int a = 0; a < NV; a++) {
        const double sa = w.sig * g[a];
        DCBF_UNROLL
        for (int b = 0; b <= a; b++) A.K[tri(a, b)] = fma(sa, g[b], A.K[tri(a, b)]);
        A.q1[a] = fma(g[a], w.w1, A.q1[a]);
        A.q2[a] = fma(g[a], w.binv, A.q2[a]);
        A.q3[a] = fma(g[a], w.y, A.q3[a]);
    }
}

DCBF_HD void dd_eval_emit(EvalOut &E, double c, const double *g, int nv, double lo, double hi) {
    const int r = E.row++;
    if (E.c) E.c[r] = c;
    if (E.cl) E.cl[r] = lo;
    if (E.cu) E.cu[r] = hi;
    if (E.jac) for (int j = 0; j < 6; j++) E.jac[6 * r + j] = j < nv ? g[j] : 0.0;
}

// per-node second-order bookkeeping: K += sum_k [Jx_k;Jy_k]^T Q_k [Jx_k;Jy_k] + cx_k Hx_k + cy_k Hy_k
struct DdSecond { double qxx[4], qxy[4], qyy[4], cx[4], cy[4]; };

template <int I, int MODE, int KT>
DCBF_HD void dd_full_step(const Consts &k, const dcbf_params &P, const Problem &pb, const DdNodes &nd, const double *z,
                          double sf, const RowCtl &ctl, DdRows<KT> &R, Acc<6> &A, LogAcc &LA, DdSecond &H2, EvalOut *E) {
    constexpr int kn = I + 1;
    constexpr int NV = 2 * (I + 1);
    const double INF = 1e300;
    const bool hess = true;   // restoration uses the curvature of the violated rows too (objective scaled by sf = 0)
    // ---- objective: node kn and the smoothness term of step I --------------------------------------------------------
    {
        const double w = P.w_q + (I == 0 ? P.w_p : 0.0);
        const double ex = nd.x[kn] - pb.goal[0], ey = nd.y[kn] - pb.goal[1];
        const double dx = -ex, dy = -ey;
        const double r2 = dx * dx + dy * dy, ir2 = 1.0 / r2;
        const double phi = nd.th[kn] - datan2(dy, dx);
        const double pv = I == 0 ? pb.last_u[0] : z[2 * (I > 0 ? I - 1 : 0)];
        const double pw = I == 0 ? pb.last_u[1] : z[2 * (I > 0 ? I - 1 : 0) + 1];
        const double dv = z[2 * I] - pv, dw = z[2 * I + 1] - pw;
        A.f += w * (ex * ex + ey * ey) + P.w_r * phi * phi + P.w_t * (dv * dv + dw * dw);
        const double px = -dy * ir2, py = dx * ir2;
        const double nx = 2.0 * w * ex + 2.0 * P.w_r * phi * px;
        const double ny = 2.0 * w * ey + 2.0 * P.w_r * phi * py;
        const double nt = 2.0 * P.w_r * phi;
        DCBF_UNROLL
        for (int a = 0; a < NV; a++) A.grad[a] = fma(nd.Jx[kn][a], nx, fma(nd.Jy[kn][a], ny, A.grad[a]));
        DCBF_UNROLL
        for (int l = 0; l <= I; l++) A.grad[2 * l + 1] += nt;
        A.grad[2 * I] += 2.0 * P.w_t * dv; A.grad[2 * I + 1] += 2.0 * P.w_t * dw;
        if (I > 0) { A.grad[2 * (I > 0 ? I - 1 : 0)] -= 2.0 * P.w_t * dv; A.grad[2 * (I > 0 ? I - 1 : 0) + 1] -= 2.0 * P.w_t * dw; }
        if (hess) {
            const double ir4 = ir2 * ir2;
            const double pxx = -2.0 * dx * dy * ir4, pyy = -pxx, pxy = (dx * dx - dy * dy) * ir4;
            const double r2w = 2.0 * P.w_r * sf;
            H2.qxx[kn] += sf * 2.0 * w + r2w * (px * px + phi * pxx);
            H2.qxy[kn] += r2w * (px * py + phi * pxy);
            H2.qyy[kn] += sf * 2.0 * w + r2w * (py * py + phi * pyy);
            H2.cx[kn] += sf * nx; H2.cy[kn] += sf * ny;
            // cross terms with theta_k (d theta_k / d w_l = 1, l <= I) and theta-theta
            const double hxt = r2w * px, hyt = r2w * py;
            DCBF_UNROLL
            for (int l = 0; l <= I; l++) {
                const int tw = 2 * l + 1;
                DCBF_UNROLL
                for (int a = 0; a < NV; a++) {
                    const double v = hxt * nd.Jx[kn][a] + hyt * nd.Jy[kn][a];
                    if (a == tw) A.K[tri(tw, tw)] += 2.0 * v; else A.K[tri(tw, a)] += v;
                }
                DCBF_UNROLL
                for (int m = 0; m <= l; m++) A.K[tri(tw, 2 * m + 1)] += r2w;
            }
            const double t2 = 2.0 * P.w_t * sf;
            A.K[tri(2 * I, 2 * I)] += t2; A.K[tri(2 * I + 1, 2 * I + 1)] += t2;
            if (I > 0) {
                constexpr int Im = I > 0 ? I - 1 : 0;
                A.K[tri(2 * Im, 2 * Im)] += t2; A.K[tri(2 * Im + 1, 2 * Im + 1)] += t2;
                A.K[tri(2 * I, 2 * Im)] -= t2; A.K[tri(2 * I + 1, 2 * Im + 1)] -= t2;
            }
        }
    }
    // ---- D-CBF rows -----------------------------------------------------------------------------------------------------
    const double gm1 = P.gamma - 1.0;
    for (int j = 0; j < pb.nc + pb.ne; j++) {
        const bool is_c = j < pb.nc;
        if (MODE == MODE_SOLVE && !(((is_c ? pb.mc : pb.me) >> (is_c ? j : j - pb.nc)) & 1u)) continue;
        double c, h1x, h1y, h0x, h0y, ea, eb, ec;
        if (is_c) {
            const double *o = pb.cir + DCBF_CIR_REC * j;
            const double ax = nd.x[kn] - o[0], ay = nd.y[kn] - o[1], bx = nd.x[I] - o[0], by = nd.y[I] - o[1];
            c = (ax * ax + ay * ay - o[2]) + gm1 * (bx * bx + by * by - o[2]);
            h1x = 2.0 * ax; h1y = 2.0 * ay; h0x = 2.0 * gm1 * bx; h0y = 2.0 * gm1 * by;
            ea = 1.0; eb = 0.0; ec = 1.0;
        } else {
            const double *o = pb.elp + DCBF_ELP_REC * (j - pb.nc);
            const double ax = nd.x[kn] - o[0], ay = nd.y[kn] - o[1], bx = nd.x[I] - o[0], by = nd.y[I] - o[1];
            ea = o[2]; eb = o[3]; ec = o[4];
            c = (ea * ax * ax + eb * ax * ay + ec * ay * ay - o[5]) + gm1 * (ea * bx * bx + eb * bx * by + ec * by * by - o[5]);
            h1x = 2.0 * ea * ax + eb * ay; h1y = 2.0 * ec * ay + eb * ax;
            h0x = gm1 * (2.0 * ea * bx + eb * by); h0y = gm1 * (2.0 * ec * by + eb * bx);
        }
        double g[6];
        DCBF_UNROLL
        for (int a = 0; a < NV; a++) g[a] = h1x * nd.Jx[kn][a] + h1y * nd.Jy[kn][a] + h0x * nd.Jx[I][a] + h0y * nd.Jy[I][a];
        double y;
        if (MODE == MODE_SOLVE) {
            const int ri = I * KT + j;
            RowW w = row_full<true, false>(ctl, A, LA, c, 0.0, INF, R.sc[ri], R.zc[ri], R.zc[ri], R.dc[ri], R.ec[ri], R.ec[ri]);
            dd_acc_row<NV>(A, g, w);
            y = w.y;
        } else {
            y = eval_lambda(*E);
            dd_eval_emit(*E, c, g, NV, 0.0, INFINITY);
        }
        H2.qxx[kn] += 2.0 * ea * y; H2.qxy[kn] += eb * y; H2.qyy[kn] += 2.0 * ec * y;
        H2.qxx[I] += gm1 * 2.0 * ea * y; H2.qxy[I] += gm1 * eb * y; H2.qyy[I] += gm1 * 2.0 * ec * y;
        H2.cx[kn] += y * h1x; H2.cy[kn] += y * h1y; H2.cx[I] += y * h0x; H2.cy[I] += y * h0y;
    }
    // ---- speed/turn coupling and the variable bounds -----------------------------------------------------------------------
    const double v = z[2 * I], w = z[2 * I + 1];
    if (MODE == MODE_SOLVE) {
        RowW wp = row_full<false, true>(ctl, A, LA, v + P.s_turn * w, -INF, P.bvx_max, R.su[2 * I], R.zu1[2 * I], R.zu1[2 * I],
                                        R.du[2 * I], R.eu1[2 * I], R.eu1[2 * I]);
        RowW wm = row_full<false, true>(ctl, A, LA, v - P.s_turn * w, -INF, P.bvx_max, R.su[2 * I + 1], R.zu1[2 * I + 1],
                                        R.zu1[2 * I + 1], R.du[2 * I + 1], R.eu1[2 * I + 1], R.eu1[2 * I + 1]);
        RowW bv = row_full<true, true>(ctl, A, LA, v, P.bvx_min, P.bvx_max, R.s2[2 * I], R.zl2[2 * I], R.zu2[2 * I], R.d2[2 * I],
                                       R.el2[2 * I], R.eu2[2 * I]);
        RowW bw = row_full<true, true>(ctl, A, LA, w, -P.ang_max, P.ang_max, R.s2[2 * I + 1], R.zl2[2 * I + 1], R.zu2[2 * I + 1],
                                       R.d2[2 * I + 1], R.el2[2 * I + 1], R.eu2[2 * I + 1]);
        const int iv = 2 * I, iw = 2 * I + 1;
        const double st = P.s_turn;
        A.K[tri(iv, iv)] += wp.sig + wm.sig + bv.sig;
        A.K[tri(iw, iv)] += st * (wp.sig - wm.sig);
        A.K[tri(iw, iw)] += st * st * (wp.sig + wm.sig) + bw.sig;
        A.q1[iv] += wp.w1 + wm.w1 + bv.w1;       A.q1[iw] += st * (wp.w1 - wm.w1) + bw.w1;
        A.q2[iv] += wp.binv + wm.binv + bv.binv; A.q2[iw] += st * (wp.binv - wm.binv) + bw.binv;
        A.q3[iv] += wp.y + wm.y + bv.y;          A.q3[iw] += st * (wp.y - wm.y) + bw.y;
    } else {
        double g[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        g[2 * I] = 1.0;
        g[2 * I + 1] = w == 0.0 ? 0.0 : (w > 0.0 ? P.s_turn : -P.s_turn);
        dd_eval_emit(*E, P.s_turn * fabs(w) + v, g, 6, P.bvx_min, P.bvx_max);
    }
}

// add the accumulated second-order terms of the nodes to K
DCBF_HD void dd_add_second(const Consts &k, const DdNodes &nd, const double *z, const DdSecond &H2, double *K) {
    // Gauss-Newton-like part: [Jx;Jy]^T Q [Jx;Jy] per node
    DCBF_UNROLL
    for (int kn = 1; kn < 4; kn++) {
        DCBF_UNROLL
        for (int a = 0; a < 2 * kn; a++) {
            const double ta = H2.qxx[kn] * nd.Jx[kn][a] + H2.qxy[kn] * nd.Jy[kn][a];
            const double tb = H2.qxy[kn] * nd.Jx[kn][a] + H2.qyy[kn] * nd.Jy[kn][a];
            DCBF_UNROLL
            for (int b = 0; b <= a; b++) K[tri(a, b)] += ta * nd.Jx[kn][b] + tb * nd.Jy[kn][b];
        }
    }
    // curvature of the positions: H x_k = sum_{l<k} [-dt sin(th_l)(e_vl Jth_l^T + sym) - dt v_l cos(th_l) Jth_l Jth_l^T]
    DCBF_UNROLL
    for (int l = 1; l < 3; l++) {
        double CX = 0.0, CY = 0.0;
        DCBF_UNROLL
        for (int kn = l + 1; kn < 4; kn++) { CX += H2.cx[kn]; CY += H2.cy[kn]; }
        const double v = z[2 * l];
        const double cvw = k.dt * (-CX * nd.sn[l] + CY * nd.cs[l]);          // coefficient of e_vl Jth^T + sym
        const double cww = -k.dt * v * (CX * nd.cs[l] + CY * nd.sn[l]);      // coefficient of Jth Jth^T
        DCBF_UNROLL
        for (int a = 0; a < l; a++) {
            K[tri(2 * l, 2 * a + 1)] += cvw;
            DCBF_UNROLL
            for (int b = 0; b <= a; b++) K[tri(2 * a + 1, 2 * b + 1)] += cww;
        }
    }
}

template <int I, int KT>
DCBF_HD void dd_dir_step(const dcbf_params &P, const Problem &pb, const DdNodes &nd, const double *dz, const double *dxn,
                         const double *dyn, double mu, double tau, DdRows<KT> &R, DirStat &D) {
    constexpr int kn = I + 1;
    const double INF = 1e300;
    const double gm1 = P.gamma - 1.0;
    for (int j = 0; j < pb.nc + pb.ne; j++) {
        const bool is_c = j < pb.nc;
        if (!(((is_c ? pb.mc : pb.me) >> (is_c ? j : j - pb.nc)) & 1u)) continue;
        double h1x, h1y, h0x, h0y;
        if (is_c) {
            const double *o = pb.cir + DCBF_CIR_REC * j;
            h1x = 2.0 * (nd.x[kn] - o[0]); h1y = 2.0 * (nd.y[kn] - o[1]);
            h0x = 2.0 * gm1 * (nd.x[I] - o[0]); h0y = 2.0 * gm1 * (nd.y[I] - o[1]);
        } else {
            const double *o = pb.elp + DCBF_ELP_REC * (j - pb.nc);
            const double ax = nd.x[kn] - o[0], ay = nd.y[kn] - o[1], bx = nd.x[I] - o[0], by = nd.y[I] - o[1];
            h1x = 2.0 * o[2] * ax + o[3] * ay; h1y = 2.0 * o[4] * ay + o[3] * ax;
            h0x = gm1 * (2.0 * o[2] * bx + o[3] * by); h0y = gm1 * (2.0 * o[4] * by + o[3] * bx);
        }
        const double jd = h1x * dxn[kn] + h1y * dyn[kn] + h0x * dxn[I] + h0y * dyn[I];
        const int ri = I * KT + j;
        row_dir<true, false>(mu, tau, D, jd, R.sc[ri], 0.0, INF, R.zc[ri], 0.0, R.dc[ri], R.ec[ri], R.ec[ri]);
    }
    const double dv = dz[2 * I], dw = dz[2 * I + 1];
    row_dir<false, true>(mu, tau, D, dv + P.s_turn * dw, R.su[2 * I], -INF, P.bvx_max, 0.0, R.zu1[2 * I], R.du[2 * I], R.eu1[2 * I], R.eu1[2 * I]);
    row_dir<false, true>(mu, tau, D, dv - P.s_turn * dw, R.su[2 * I + 1], -INF, P.bvx_max, 0.0, R.zu1[2 * I + 1], R.du[2 * I + 1],
                         R.eu1[2 * I + 1], R.eu1[2 * I + 1]);
    row_dir<true, true>(mu, tau, D, dv, R.s2[2 * I], P.bvx_min, P.bvx_max, R.zl2[2 * I], R.zu2[2 * I], R.d2[2 * I], R.el2[2 * I], R.eu2[2 * I]);
    row_dir<true, true>(mu, tau, D, dw, R.s2[2 * I + 1], -P.ang_max, P.ang_max, R.zl2[2 * I + 1], R.zu2[2 * I + 1], R.d2[2 * I + 1],
                        R.el2[2 * I + 1], R.eu2[2 * I + 1]);
}

template <int I, int KT>
DCBF_HD void dd_val_step(const dcbf_params &P, const Problem &pb, const DdNodes &nd, const double *z, int phase, double alpha,
                         const DdRows<KT> &R, ValStat &V) {
    constexpr int kn = I + 1;
    const double INF = 1e300;
    {
        const double w = P.w_q + (I == 0 ? P.w_p : 0.0);
        const double ex = nd.x[kn] - pb.goal[0], ey = nd.y[kn] - pb.goal[1];
        const double phi = nd.th[kn] - datan2(-ey, -ex);
        const double pv = I == 0 ? pb.last_u[0] : z[2 * (I > 0 ? I - 1 : 0)];
        const double pw = I == 0 ? pb.last_u[1] : z[2 * (I > 0 ? I - 1 : 0) + 1];
        const double dv = z[2 * I] - pv, dw = z[2 * I + 1] - pw;
        V.f += w * (ex * ex + ey * ey) + P.w_r * phi * phi + P.w_t * (dv * dv + dw * dw);
    }
    const double gm1 = P.gamma - 1.0;
    for (int j = 0; j < pb.nc + pb.ne; j++) {
        const bool is_c = j < pb.nc;
        if (!(((is_c ? pb.mc : pb.me) >> (is_c ? j : j - pb.nc)) & 1u)) continue;
        double c;
        if (is_c) {
            const double *o = pb.cir + DCBF_CIR_REC * j;
            const double ax = nd.x[kn] - o[0], ay = nd.y[kn] - o[1], bx = nd.x[I] - o[0], by = nd.y[I] - o[1];
            c = (ax * ax + ay * ay - o[2]) + gm1 * (bx * bx + by * by - o[2]);
        } else {
            const double *o = pb.elp + DCBF_ELP_REC * (j - pb.nc);
            const double ax = nd.x[kn] - o[0], ay = nd.y[kn] - o[1], bx = nd.x[I] - o[0], by = nd.y[I] - o[1];
            c = (o[2] * ax * ax + o[3] * ax * ay + o[4] * ay * ay - o[5]) + gm1 * (o[2] * bx * bx + o[3] * bx * by + o[4] * by * by - o[5]);
        }
        row_val<true, false>(phase, alpha, V, c, 0.0, INF, R.sc[I * KT + j], R.dc[I * KT + j]);
    }
    const double v = z[2 * I], w = z[2 * I + 1];
    row_val<false, true>(phase, alpha, V, v + P.s_turn * w, -INF, P.bvx_max, R.su[2 * I], R.du[2 * I]);
    row_val<false, true>(phase, alpha, V, v - P.s_turn * w, -INF, P.bvx_max, R.su[2 * I + 1], R.du[2 * I + 1]);
    row_val<true, true>(phase, alpha, V, v, P.bvx_min, P.bvx_max, R.s2[2 * I], R.d2[2 * I]);
    row_val<true, true>(phase, alpha, V, w, -P.ang_max, P.ang_max, R.s2[2 * I + 1], R.d2[2 * I + 1]);
}

template <int KT>
struct DdModel {
    static constexpr int N = 6;
    Problem pb;
    DdRows<KT> R;
    DdNodes nd;

    DCBF_HD void pass_full(const Consts &k, const dcbf_params &P, const IpmState<6> &S, Acc<6> &A, LogAcc &LA) {
        RowCtl ctl;
        ctl.mu = S.mu; ctl.alpha = S.alpha; ctl.alpha_z = S.alpha_z; ctl.phase = S.phase; ctl.pending = S.pending;
        ctl.reinit = S.reinit;
        acc_reset(A);
        LA.sum = 0.0; LA.prod = 1.0; LA.cnt = 0;
        dd_rollout(k, pb.x0, S.z, nd, true);
        DdSecond H2;
        DCBF_UNROLL
        for (int i = 0; i < 4; i++) { H2.qxx[i] = H2.qxy[i] = H2.qyy[i] = H2.cx[i] = H2.cy[i] = 0.0; }
        const double sf = S.phase == PH_RESTO ? 0.0 : S.sf;
        dd_full_step<0, MODE_SOLVE, KT>(k, P, pb, nd, S.z, sf, ctl, R, A, LA, H2, nullptr);
        dd_full_step<1, MODE_SOLVE, KT>(k, P, pb, nd, S.z, sf, ctl, R, A, LA, H2, nullptr);
        dd_full_step<2, MODE_SOLVE, KT>(k, P, pb, nd, S.z, sf, ctl, R, A, LA, H2, nullptr);
        dd_add_second(k, nd, S.z, H2, A.K);
    }
    DCBF_HD double grad_inf(const Consts &k, const dcbf_params &P, const IpmState<6> &S) {
        // small problem: reuse the full pass in restoration mode (touches no row state)
        Acc<6> A; LogAcc LA;
        IpmState<6> T = S;
        T.phase = PH_RESTO;
        pass_full(k, P, T, A, LA);
        double m = 0.0;
        DCBF_UNROLL
        for (int i = 0; i < 6; i++) m = dmax(m, fabs(A.grad[i]));
        return m;
    }
    DCBF_HD void pass_dir(const Consts &k, const dcbf_params &P, const IpmState<6> &S, double tau, DirStat &D) {
        double dxn[4], dyn[4];
        DCBF_UNROLL
        for (int kn = 0; kn < 4; kn++) {
            double sx = 0.0, sy = 0.0;
            DCBF_UNROLL
            for (int a = 0; a < 2 * kn; a++) { sx = fma(nd.Jx[kn][a], S.dz[a], sx); sy = fma(nd.Jy[kn][a], S.dz[a], sy); }
            dxn[kn] = sx; dyn[kn] = sy;
        }
        dd_dir_step<0, KT>(P, pb, nd, S.dz, dxn, dyn, S.mu, tau, R, D);
        dd_dir_step<1, KT>(P, pb, nd, S.dz, dxn, dyn, S.mu, tau, R, D);
        dd_dir_step<2, KT>(P, pb, nd, S.dz, dxn, dyn, S.mu, tau, R, D);
    }
    DCBF_HD void pass_value(const Consts &k, const dcbf_params &P, const IpmState<6> &S, double alpha, ValStat &V) const {
        double zt[6];
        DCBF_UNROLL
        for (int i = 0; i < 6; i++) zt[i] = fma(alpha, S.dz[i], S.z[i]);
        DdNodes nt;
        dd_rollout(k, pb.x0, zt, nt, false);
        dd_val_step<0, KT>(P, pb, nt, zt, S.phase, alpha, R, V);
        dd_val_step<1, KT>(P, pb, nt, zt, S.phase, alpha, R, V);
        dd_val_step<2, KT>(P, pb, nt, zt, S.phase, alpha, R, V);
    }
};

// ===============================================================================================================
// interior-point driver (one call = one iteration of one lane)
// ===============================================================================================================
template <int N>
DCBF_HD void ipm_init(const dcbf_params &P, IpmState<N> &S) {
    S.mu = P.mu_init; S.sf = 1.0; S.alpha = 0.0; S.alpha_z = 0.0; S.delta_last = 0.0; S.lm_lambda = 1e-4;
    S.resto_target = 0.0; S.resto_entry = 0.0; S.theta_max = 1e300; S.theta_min = 0.0; S.nf = 0; S.iters = 0; S.acc_cnt = 0;
    S.status = -1; S.nstall = 0; S.tiny = 0; S.nresto = 0; S.v2_h1 = 0.0; S.v2_h2 = 0.0; S.phase = PH_MAIN; S.pending = false; S.reinit = true; S.first = true; S.done = false;
    S.obj = 0.0; S.viol = 0.0;
}

template <int N>
DCBF_HD void filter_add(IpmState<N> &S, double th, double ph) {
    const int slot = S.nf < DCBF_FILT ? S.nf : (S.iters % DCBF_FILT);
    S.filt_th[slot] = th; S.filt_ph[slot] = ph;
    if (S.nf < DCBF_FILT) S.nf++;
}

// returns true when the lane has finished (S.status, S.obj, S.viol are final and S.z is the answer)
template <class Model>
DCBF_HD bool ipm_iterate(const Consts &k, const dcbf_params &P, Model &M, IpmState<Model::N> &S) {
    constexpr int N = Model::N;
    Acc<N> A;
    LogAcc LA;
    if (S.first) {
        const double gmax = M.grad_inf(k, P, S);   // gradient-based objective scaling (Ipopt nlp_scaling_max_gradient = 100)
        S.sf = gmax > 100.0 ? 100.0 / gmax : 1.0;
    }
    M.pass_full(k, P, S, A, LA);
#ifdef DCBF_TRACE
    printf("it %3d ph %d mu %.2e f %.6f theta %.3e vmax %.3e alpha %.3e delta %.1e lam %.1e nf %d\n", S.iters, S.phase, S.mu, A.f, A.theta, A.vmax, S.alpha, S.delta_last, S.lm_lambda, S.nf);
#endif
    S.pending = false;
    S.reinit = false;
    S.obj = A.f;
    S.viol = A.vmax;
    if (!(A.f == A.f) || !(A.theta == A.theta)) { S.status = -13; S.done = true; return true; }
    double rhs[N];
    const double tol = P.tol;
    if (S.phase == PH_MAIN) {
        if (S.first) {
            S.theta_max = 1e4 * dmax(1.0, A.theta);
            S.theta_min = 1e-4 * dmax(1.0, A.theta);
            S.first = false;
        }
        double dinf = 0.0;
        DCBF_UNROLL
        for (int i = 0; i < N; i++) dinf = dmax(dinf, fabs(fma(S.sf, A.grad[i], A.q3[i])));
        const double sd = dmax(100.0, 2.0 * A.zsum / (double)(A.nrows + A.nz)) * 0.01;
        const double sc = dmax(100.0, A.zsum / (double)(A.nz > 0 ? A.nz : 1)) * 0.01;
        double E0;
        for (;;) {
            const double compm = dmax(fabs(A.cmax - S.mu), fabs(A.cmin - S.mu));
            E0 = dmax(dmax(dinf / sd, A.pinf), A.cmax / sc);
            const double Emu = dmax(dmax(dinf / sd, A.pinf), compm / sc);
            if (E0 <= tol) break;
            if (Emu <= P.kappa_eps * S.mu && S.mu > tol * 0.1 * (1.0 + 1e-12)) {
                S.mu = dmax(tol * 0.1, dmin(DCBF_KAPPA_MU * S.mu, DCBF_MU_POW(S.mu)));
                S.nf = 0;
                continue;
            }
            break;
        }
        if (E0 <= tol) { S.status = 0; S.done = true; return true; }
        if (E0 <= 1e-6 && A.vmax <= P.constr_viol_tol) {
            if (++S.acc_cnt >= 15) { S.status = 1; S.done = true; return true; }
        } else {
            S.acc_cnt = 0;
        }
        if (S.iters >= P.max_iter) { S.status = -1; S.done = true; return true; }
        if (S.tiny >= P.tiny_count) {
            // three consecutive accepted steps shorter than 1e-2 while the rows are still violated: the fraction-to-boundary
            // rule is pinning the iterate (typical for infeasible problems) -> go to restoration now instead of crawling
            S.tiny = 0;
            filter_add(S, (1.0 - 1e-5) * A.theta, (S.sf * A.f - S.mu * log_total(LA)) - 1e-5 * A.theta);
            S.phase = PH_RESTO; S.resto_entry = A.vmax; S.resto_target = dmax(0.1 * A.vmax, 1e-9); S.lm_lambda = 1e-4; S.acc_cnt = 0; S.nresto = 0;
            S.iters++;
            return false;
        }
        DCBF_UNROLL
        for (int i = 0; i < N; i++) rhs[i] = -S.sf * A.grad[i] - A.q1[i] + S.mu * A.q2[i];
    } else {
        if (A.vmax <= S.resto_target) {   // restoration succeeded: back to the main phase with fresh slacks
            S.phase = PH_MAIN; S.reinit = true;
            return false;
        }
        double gn = 0.0;
        DCBF_UNROLL
        for (int i = 0; i < N; i++) gn = dmax(gn, fabs(A.q1[i]));
#ifdef DCBF_TRACE
        printf("      resto: v2 %.12e gn %.3e\n", A.v2, gn);
#endif
        bool stationary = gn <= 1e-10 * dmax(1.0, A.vmax) || S.lm_lambda > 1e12;
        if (stationary) {
            if (A.vmax > P.constr_viol_tol) { S.status = 2; S.done = true; return true; }
            // stationary, violation within constr_viol_tol: give the main phase one more chance; a second stall at such a
            // point (marginally infeasible problem, filter blocks every step) ends as Restoration_Failed instead of cycling
            if (S.resto_entry <= 1e-9 || S.nstall++ >= 1) { S.status = -2; S.done = true; return true; }
            S.phase = PH_MAIN; S.reinit = true;
            return false;
        }
        if (S.iters >= P.max_iter) { S.status = -1; S.done = true; return true; }
        // Levenberg-Marquardt trials: the assembled K and right-hand side are reused while lambda is escalated
        double L[N * (N + 1) / 2];
        for (int rt = 0; rt < 20 && S.lm_lambda <= 1e12; rt++) {
            if (!chol_packed<N>(A.K, S.lm_lambda, L)) { S.lm_lambda *= DCBF_LM_UP; continue; }
            DCBF_UNROLL
            for (int i = 0; i < N; i++) S.dz[i] = -A.q1[i];
            chol_solve_packed<N>(L, S.dz);
            ValStat V;
            V.f = 0.0; V.theta = 0.0; V.v2 = 0.0; V.vmax = 0.0; V.ok = true; V.la.sum = 0.0; V.la.prod = 1.0; V.la.cnt = 0;
            M.pass_value(k, P, S, 1.0, V);
#ifdef DCBF_COUNT
            g_trials++;
#endif
            if (V.v2 < A.v2 * (1.0 - 1e-12)) {
                double dn = 0.0;
                DCBF_UNROLL
                for (int i = 0; i < N; i++) { dn = dmax(dn, fabs(S.dz[i])); S.z[i] += S.dz[i]; }
                S.iters++;
                S.lm_lambda = dmax(S.lm_lambda * DCBF_LM_DOWN, 1e-12);
                // stagnation: the squared violation has stopped decreasing (two consecutive accepted steps with a relative
                // decrease below 1e-4) -> the iterate is (numerically) a stationary point of the violation
                if (A.v2 - V.v2 <= 1e-4 * A.v2) S.acc_cnt++; else S.acc_cnt = 0;
                // ... or three consecutive accepted steps that together gained less than DCBF_RESTO_WINDOW (the iterate
                // crawls along a kink of the violation: rows entering and leaving the violated set)
                const bool crawl = S.nresto >= 2 && S.v2_h2 - V.v2 <= P.resto_window * S.v2_h2;
                S.v2_h2 = S.v2_h1; S.v2_h1 = A.v2; S.nresto++;
                if ((dn < 1e-12 || S.acc_cnt >= 2 || crawl) && V.vmax > S.resto_target) S.lm_lambda = 1e13;
                return false;
            }
            S.lm_lambda *= DCBF_LM_UP;
        }
        return false;   // lambda exhausted: the next call classifies the point as stationary
    }
    // ---- factor + solve ---------------------------------------------------------------------------------------------
    double L[N * (N + 1) / 2];
    double delta = 0.0;
    bool ok = false;
    for (int tr = 0; tr < 48; tr++) {
        ok = chol_packed<N>(A.K, delta, L);
        if (ok) break;
        if (delta == 0.0) delta = S.delta_last == 0.0 ? 1e-4 : dmax(1e-20, S.delta_last * (1.0 / 3.0));
        else delta *= (S.delta_last == 0.0 ? 100.0 : 8.0);
    }
    if (!ok) { S.status = -3; S.done = true; return true; }
    if (delta > 0.0) S.delta_last = delta;
    DCBF_UNROLL
    for (int i = 0; i < N; i++) S.dz[i] = rhs[i];
    chol_solve_packed<N>(L, S.dz);

    // ---- main phase: step sizes and filter line search -----------------------------------------------------------------
    const double tau = dmax(0.99, 1.0 - S.mu);
    DirStat D;
    D.amax = 1.0; D.az = 1.0; D.dphi = 0.0;
    M.pass_dir(k, P, S, tau, D);
    double dphi = D.dphi;
    DCBF_UNROLL
    for (int i = 0; i < N; i++) dphi = fma(S.sf * A.grad[i], S.dz[i], dphi);
    const double theta = A.theta;
    const double phi = S.sf * A.f - S.mu * log_total(LA);
    double alpha = D.amax;
    int accepted = 0;
    const double eps_phi = 10.0 * 2.2e-16 * fabs(phi);
    for (int ls = 0; ls < DCBF_LS_MAX; ls++, alpha *= 0.5) {
        ValStat V;
        V.f = 0.0; V.theta = 0.0; V.v2 = 0.0; V.vmax = 0.0; V.ok = true; V.la.sum = 0.0; V.la.prod = 1.0; V.la.cnt = 0;
        M.pass_value(k, P, S, alpha, V);
        const double ph_t = S.sf * V.f - S.mu * log_total(V.la);
        const double th_t = V.theta;
        if (!V.ok || !(ph_t == ph_t) || !(th_t <= S.theta_max)) continue;
        bool in_filter = false;
        for (int q = 0; q < S.nf; q++)
            if (th_t >= S.filt_th[q] && ph_t >= S.filt_ph[q]) in_filter = true;
        if (in_filter) continue;
        const bool sw = dphi < 0.0 && theta <= S.theta_min && switch_cond(alpha, -dphi, theta);
        if (sw) {
            if (ph_t <= phi + 1e-8 * alpha * dphi + eps_phi) accepted = 1;
        } else if (th_t <= (1.0 - 1e-5) * theta || ph_t <= phi - 1e-5 * theta + eps_phi) {
            accepted = 2;
        }
        if (accepted) break;
    }
#ifdef DCBF_TRACE
    printf("      ls: accepted %d alpha %.3e amax %.3e dphi %.3e\n", accepted, alpha, D.amax, dphi);
#endif
    if (!accepted) {
        filter_add(S, (1.0 - 1e-5) * theta, phi - 1e-5 * theta);
        S.phase = PH_RESTO;
        S.resto_entry = A.vmax;
        S.resto_target = dmax(0.1 * A.vmax, 1e-9);
        S.lm_lambda = 1e-4;
        S.acc_cnt = 0;
        S.nresto = 0;
        S.iters++;
        return false;
    }
    if (accepted == 2) filter_add(S, (1.0 - 1e-5) * theta, phi - 1e-5 * theta);
    DCBF_UNROLL
    for (int i = 0; i < N; i++) S.z[i] = fma(alpha, S.dz[i], S.z[i]);
    S.alpha = alpha; S.alpha_z = D.az; S.pending = true;
    if (alpha < P.tiny_alpha && A.vmax > P.constr_viol_tol) S.tiny++; else S.tiny = 0;
    S.iters++;
    return false;
}

// ---------------------------------------------------------------------------------------------------------------
// problem setup shared by the kernels: detour heuristic and obstacle selection on the prepared records
// ---------------------------------------------------------------------------------------------------------------
DCBF_HD void setup_problem(const dcbf_params &P, Problem &pb) {
    const double px = pb.x0[0], py = pb.x0[1];
    pb.mc = 0u; pb.me = 0u;
    for (int j = 0; j < pb.nc; j++) {
        const double *o = pb.cir + DCBF_CIR_REC * j;
        const double d = (px - o[0]) * (px - o[0]) + (py - o[1]) * (py - o[1]) - o[2];
        if (!P.select_obs || d <= P.detect_sq) pb.mc |= 1u << j;
    }
    for (int j = 0; j < pb.ne; j++) {
        const double *o = pb.elp + DCBF_ELP_REC * j;
        const double d = (px - o[0]) * (px - o[0]) + (py - o[1]) * (py - o[1]) - o[6];
        if (!P.select_obs || d <= P.detect_sq) pb.me |= 1u << j;
    }
    pb.goal[0] = pb.goal_raw[0]; pb.goal[1] = pb.goal_raw[1];
    if (P.goal_shift) {
        const double gx = pb.goal_raw[0], gy = pb.goal_raw[1];
        const double dg = (px - gx) * (px - gx) + (py - gy) * (py - gy);
        const double PI = 3.14159265358979323846;
        for (int j = 0; j < pb.nc; j++) {
            if (!((pb.mc >> j) & 1u)) continue;
            const double *o = pb.cir + DCBF_CIR_REC * j;
            const double dc = (px - o[0]) * (px - o[0]) + (py - o[1]) * (py - o[1]);
            if (dc < dg && dc < 9.0 * o[2]) {
                const double th = datan2(gy - py, gx - px), al = datan2(o[1] - py, o[0] - px);
                double d = th - al;
                if (d < 0.0 && fabs(d) > PI) d += 2.0 * PI;
                else if (d > 0.0 && fabs(d) > PI) d -= 2.0 * PI;
                if (fabs(d) < PI / 12.0) {
                    const double na = d < 0.0 ? th - PI / 12.0 : th + PI / 12.0;
                    const double rad = sqrt(dg);
                    double sn, cs;
                    dsincos(na, &sn, &cs);
                    pb.goal[0] = px + rad * cs; pb.goal[1] = py + rad * sn;
                    break;
                }
            }
        }
    }
}

}  // namespace dcbf
