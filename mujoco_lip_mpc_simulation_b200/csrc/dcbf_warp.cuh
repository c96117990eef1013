// dcbf_warp.cuh -- warp-cooperative variant of the LIP solver: ONE PROBLEM PER WARP (device only).
//
// The per-thread kernels of dcbf_lanes.cuh have the lowest instruction count per problem but a long serial chain per
// interior-point iteration (~7 k dependent instructions) and, at 255 registers, only two warps per scheduler: a
// 4096-scenario batch is bound by single-warp latency times the slowest problem's iteration count (profiles/
// r01_summary.md).  Here the 32 lanes of a warp share one problem:
//   * rows are distributed over lanes (row r -> lane r % 32, slot r / 32); every lane evaluates its rows, applies
//     the slack/multiplier updates and publishes the row gradient (9 doubles) and weights to shared memory;
//   * the condensed matrix  W + J^T Sigma J  and the three J^T-vectors are accumulated entry-parallel (72 entries over
//     32 lanes) from the staged rows; the Lagrangian Hessian is assembled from per-node Hessians (5x5 with 8 distinct
//     entries) through the constant feature map T (24 features x 9 variables);
//   * norms, merit values and step sizes are warp-shuffle reductions; the 9x9 Cholesky is row-owned (lane i owns row i)
//     with the factor staged in shared memory; all scalar control flow is replicated and therefore uniform.
// ~2.5 k warp instructions per iteration at ~100 registers (4-5 warps per scheduler), ~3x shorter dependent chain.
// The algorithm (barrier rule, filter, restoration, status codes) is the one of ipm_iterate() in dcbf_core.cuh.
#pragma once
#include "dcbf_lanes.cuh"

namespace dcbf {
namespace wp {

constexpr int NFEAT = 24;   // 3 nodes x (x, y, vx, vy, th) | 3 x (lx, ly) | 3 x dth
enum { RT_NONE = 0, RT_CBF, RT_VBX, RT_VBY, RT_LEG, RT_DTH, RT_FENP, RT_FENM };
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int FN(int k, int c) { return 5 * (k - 1) + c; }   // node k in 1..3, c: 0 x,1 y,2 vx,3 vy,4 th
__device__ __forceinline__ int FLX(int i) { return 15 + 2 * i; }
__device__ __forceinline__ int FLY(int i) { return 16 + 2 * i; }
__device__ __forceinline__ int FDT(int i) { return 21 + i; }

template <int NS>
struct WarpShared {
    double G[32 * NS][9];    // row gradients (internal variable order)
    double SG[32 * NS][9];   // sigma * row gradient
    double RW[32 * NS][4];   // sigma, w1, binv, y
    double HQ[32 * NS][3];   // y * (2a', b', 2c') of the D-CBF rows
    double obs[DCBF_KT][6];  // selected obstacles: cx, cy, a', b', c', rhs
    double Kf[45], Lf[45], q[27], dz[9], zc[9], zt[9], red[8];
    double x0[5], goal[2], graw[2];
    double nodes[4][5];      // x, y, vx, vy, th of nodes 0..3
    double trig[4][3];       // sin, cos, atan2 target of nodes 1..3
    double nobj[4][10];      // f_k, nx, ny, nt, hxx, hxy, hyy, hxt, hyt, htt of the objective at node k
    double NH[4][8];         // node Hessians xx, xy, yy, xt, yt, tt, vxt, vyt
    double legy[3];
    double filt_th[DCBF_FILT], filt_ph[DCBF_FILT];
};

// per-CTA constants staged in shared memory: parameters, model constants, feature map, (row, col) of the 72 entries
struct CtaShared {
    dcbf_params P;
    Consts K;
    double T[NFEAT][9];
    int ea[72], eb[72];
};

__device__ __noinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __noinline__ double wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __noinline__ double wmin(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int wsumi(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// constant feature map T[f][a] = d feature_f / d z_a  (built once per CTA)
__device__ __forceinline__ void build_T(const Consts &k, double (*T)[9]) {
    for (int t = threadIdx.x; t < NFEAT * 9; t += blockDim.x) {
        const int f = t / 9, a = t % 9;
        double v = 0.0;
        const bool ax = a < 6 && (a & 1) == 0, ay = a < 6 && (a & 1) == 1;
        const int l = a < 6 ? a >> 1 : a - 6;
        if (f < 15) {
            const int kn = f / 5 + 1, c = f % 5;
            if (l < kn) {
                if (c == 0 && ax) v = k.gx[kn - 1 - l];
                if (c == 1 && ay) v = k.gx[kn - 1 - l];
                if (c == 2 && ax) v = k.gv[kn - 1 - l];
                if (c == 3 && ay) v = k.gv[kn - 1 - l];
                if (c == 4 && a >= 6) v = 1.0;
            }
        } else if (f < 21) {
            const int i = (f - 15) >> 1, c = (f - 15) & 1;
            if ((c == 0 && ax) || (c == 1 && ay)) {
                if (l < i) v = k.gx[i - 1 - l];
                else if (l == i) v = -1.0;
            }
        } else {
            if (a >= 6 && l == f - 21) v = 1.0;
        }
        T[f][a] = v;
    }
}

struct RowDesc { int type, step, obs; };

__device__ __forceinline__ RowDesc row_desc(int r, int Ks, bool has_fen) {
    RowDesc d;
    d.type = RT_NONE; d.step = 0; d.obs = 0;
    if (r < 3 * Ks) { d.type = RT_CBF; d.step = r / Ks; d.obs = r - d.step * Ks; return d; }
    const int q = r - 3 * Ks, t = q / 3;
    if (t < (has_fen ? 6 : 4)) { d.type = RT_VBX + t; d.step = q - 3 * t; }
    return d;
}

// z is replicated in registers; dynamic element access goes through a switch-free select
__device__ __forceinline__ double zsel(const double *z, int a) {
    double v = z[0];
#pragma unroll
    for (int j = 1; j < 9; j++) v = (a == j) ? z[j] : v;
    return v;
}

// value, feature ids and feature derivatives of one row
struct RowEval { double c, d[4]; int f[4]; int nf; double hq0, hq1, hq2; };
struct RowBnd { double lo, hi; bool has_lo, has_hi; };

// static bounds of a row (MPC_LIP_sig_step.py:193-227, MPC_LIP_modi.py:203-245; split form of the coupling row)
__device__ __forceinline__ RowBnd row_bounds(const dcbf_params &P, const RowDesc &rd, int leg) {
    RowBnd b;
    b.lo = -1e300; b.hi = 1e300; b.has_lo = false; b.has_hi = false;
    if (rd.type == RT_CBF) { b.lo = 0.0; b.has_lo = true; }
    else if (rd.type == RT_VBX) { b.lo = P.bvx_min; b.hi = P.bvx_max; b.has_lo = b.has_hi = true; }
    else if (rd.type == RT_VBY) {
        const bool plus = (leg > 0) == ((rd.step & 1) == 0);
        b.lo = plus ? P.bvy_min : -P.bvy_max; b.hi = plus ? P.bvy_max : -P.bvy_min; b.has_lo = b.has_hi = true;
    }
    else if (rd.type == RT_LEG) { b.hi = P.leg_sq; b.has_hi = true; }
    else if (rd.type == RT_DTH) { b.lo = -P.ang_max; b.hi = P.ang_max; b.has_lo = b.has_hi = true; }
    else if (rd.type == RT_FENP || rd.type == RT_FENM) { b.hi = P.bvx_max; b.has_hi = true; }
    return b;
}

template <int NS, bool GRAD>
__device__ __forceinline__ void eval_row(const dcbf_params &P, const WarpShared<NS> &sm, const RowDesc &rd, const double *z, RowEval &e) {
    e.nf = 0; e.c = 0.0; e.hq0 = e.hq1 = e.hq2 = 0.0;
    if (GRAD) {
#pragma unroll
        for (int p = 0; p < 4; p++) { e.d[p] = 0.0; e.f[p] = 0; }
    }
    const int i = rd.step, kn = i + 1;
    if (rd.type == RT_CBF) {
        const double *o = sm.obs[rd.obs];
        const double gm1 = P.gamma - 1.0;
        const double ax = sm.nodes[kn][0] - o[0], ay = sm.nodes[kn][1] - o[1], bx = sm.nodes[i][0] - o[0], by = sm.nodes[i][1] - o[1];
        const double ea = o[2], eb = o[3], ec = o[4];
        e.c = (ea * ax * ax + eb * ax * ay + ec * ay * ay - o[5]) + gm1 * (ea * bx * bx + eb * bx * by + ec * by * by - o[5]);
        if (GRAD) {
            e.f[0] = FN(kn, 0); e.f[1] = FN(kn, 1);
            e.d[0] = 2.0 * ea * ax + eb * ay; e.d[1] = 2.0 * ec * ay + eb * ax;
            e.nf = 2;
            if (i > 0) {
                e.f[2] = FN(i, 0); e.f[3] = FN(i, 1);
                e.d[2] = gm1 * (2.0 * ea * bx + eb * by); e.d[3] = gm1 * (2.0 * ec * by + eb * bx);
                e.nf = 4;
            }
            e.hq0 = 2.0 * ea; e.hq1 = eb; e.hq2 = 2.0 * ec;
        }
    } else if (rd.type == RT_VBX || rd.type == RT_VBY || rd.type == RT_FENP || rd.type == RT_FENM) {
        const double sn = sm.trig[kn][0], cs = sm.trig[kn][1];
        const double vx = sm.nodes[kn][2], vy = sm.nodes[kn][3];
        const double vbx = cs * vx + sn * vy, vby = -sn * vx + cs * vy;
        if (GRAD) { e.f[0] = FN(kn, 2); e.f[1] = FN(kn, 3); e.f[2] = FN(kn, 4); e.nf = 3; }
        if (rd.type == RT_VBY) {
            e.c = vby;
            if (GRAD) { e.d[0] = -sn; e.d[1] = cs; e.d[2] = -vbx; }
        } else {
            if (GRAD) { e.d[0] = cs; e.d[1] = sn; e.d[2] = vby; }
            if (rd.type == RT_VBX) e.c = vbx;
            else {
                const double sg = rd.type == RT_FENP ? P.s_turn : -P.s_turn;
                e.c = vbx + sg * z[6 + i];
                if (GRAD) { e.f[3] = FDT(i); e.d[3] = sg; e.nf = 4; }
            }
        }
    } else if (rd.type == RT_LEG) {
        const double lx = sm.nodes[i][0] - z[2 * i], ly = sm.nodes[i][1] - z[2 * i + 1];
        e.c = lx * lx + ly * ly;
        if (GRAD) { e.f[0] = FLX(i); e.f[1] = FLY(i); e.d[0] = 2.0 * lx; e.d[1] = 2.0 * ly; e.nf = 2; }
    } else if (rd.type == RT_DTH) {
        e.c = z[6 + i];
        if (GRAD) { e.f[0] = FDT(i); e.d[0] = 1.0; e.nf = 1; }
    }
}

struct WState {   // replicated scalars of one problem
    double mu, sf, alpha, alpha_z, delta_last, lm_lambda, resto_target, resto_entry, theta_max, theta_min, obj, viol;
    int nf, iters, acc_cnt, status, phase, nstall, tiny;
    bool pending, reinit, first;
};

// ---------------------------------------------------------------------------------------------------------------
// rollout + per-node trigonometry and objective terms at the point z (shared memory); collective, out of line
// ---------------------------------------------------------------------------------------------------------------
template <int NS>
__device__ __noinline__ void w_nodes(const CtaShared &cs_, WarpShared<NS> &sm, const double *z, int lane, double sf, bool want_hess) {
    const Consts &k = cs_.K;
    const dcbf_params &P = cs_.P;
    if (lane == 0) {
        double x = sm.x0[0], y = sm.x0[1], vx = sm.x0[2], vy = sm.x0[3], th = sm.x0[4];
        sm.nodes[0][0] = x; sm.nodes[0][1] = y; sm.nodes[0][2] = vx; sm.nodes[0][3] = vy; sm.nodes[0][4] = th;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double fx = z[2 * i], fy = z[2 * i + 1];
            const double xn = k.C * x + k.Sb * vx + k.gx[0] * fx, yn = k.C * y + k.Sb * vy + k.gx[0] * fy;
            vx = k.bS * x + k.C * vx + k.gv[0] * fx; vy = k.bS * y + k.C * vy + k.gv[0] * fy;
            x = xn; y = yn; th += z[6 + i];
            sm.nodes[i + 1][0] = x; sm.nodes[i + 1][1] = y; sm.nodes[i + 1][2] = vx; sm.nodes[i + 1][3] = vy; sm.nodes[i + 1][4] = th;
        }
    }
    __syncwarp();
    if (lane < 3) {
        const int kn = lane + 1;
        const double th = sm.nodes[kn][4];
        double sn, cs;
        dsincos(th, &sn, &cs);
        const double w = P.w_q + (kn == 1 ? P.w_p : 0.0);
        const double ex = sm.nodes[kn][0] - sm.goal[0], ey = sm.nodes[kn][1] - sm.goal[1];
        const double dx = -ex, dy = -ey;
        const double r2 = dx * dx + dy * dy, ir2 = drcp(r2);
        const double tar = datan2(dy, dx);
        const double phi = th - tar;
        sm.trig[kn][0] = sn; sm.trig[kn][1] = cs; sm.trig[kn][2] = tar;
        const double px = -dy * ir2, py = dx * ir2;
        sm.nobj[kn][0] = w * (ex * ex + ey * ey) + P.w_r * phi * phi;
        sm.nobj[kn][1] = 2.0 * w * ex + 2.0 * P.w_r * phi * px;
        sm.nobj[kn][2] = 2.0 * w * ey + 2.0 * P.w_r * phi * py;
        sm.nobj[kn][3] = 2.0 * P.w_r * phi;
        if (want_hess) {
            const double ir4 = ir2 * ir2;
            const double pxx = -2.0 * dx * dy * ir4, pyy = -pxx, pxy = (dx * dx - dy * dy) * ir4;
            const double r2w = 2.0 * P.w_r * sf;
            sm.nobj[kn][4] = sf * 2.0 * w + r2w * (px * px + phi * pxx);
            sm.nobj[kn][5] = r2w * (px * py + phi * pxy);
            sm.nobj[kn][6] = sf * 2.0 * w + r2w * (py * py + phi * pyy);
            sm.nobj[kn][7] = r2w * px; sm.nobj[kn][8] = r2w * py; sm.nobj[kn][9] = r2w;
        }
    }
    __syncwarp();
}

// eight statistics reduced together (interleaved butterflies): four sums and four maxima
struct Stat8 { double s0, s1, s2, s3, m0, m1, m2, m3; };
__device__ __forceinline__ void reduce8_inline(Stat8 &t) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t.s0 += __shfl_xor_sync(FULL, t.s0, o); t.s1 += __shfl_xor_sync(FULL, t.s1, o);
        t.s2 += __shfl_xor_sync(FULL, t.s2, o); t.s3 += __shfl_xor_sync(FULL, t.s3, o);
        t.m0 = fmax(t.m0, __shfl_xor_sync(FULL, t.m0, o)); t.m1 = fmax(t.m1, __shfl_xor_sync(FULL, t.m1, o));
        t.m2 = fmax(t.m2, __shfl_xor_sync(FULL, t.m2, o)); t.m3 = fmax(t.m3, __shfl_xor_sync(FULL, t.m3, o));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the solver for one problem (all 32 lanes call it with identical arguments).  Inputs in sm: x0, graw, zc.
// ---------------------------------------------------------------------------------------------------------------
template <int NS>
__device__ void solve_lip_warp(const CtaShared &cs_, WarpShared<NS> &sm, const BatchIn &in, int b, int lane, int leg, WState &S) {
    const dcbf_params &P = cs_.P;
    const double (*T)[9] = cs_.T;
    // ---- problem setup -------------------------------------------------------------------------------------------
    const int fld = in.field ? in.field[b] : 0;
    const double *cir = in.cir_rec + (size_t)fld * in.Kc * DCBF_CIR_REC;
    const double *elp = in.elp_rec + (size_t)fld * in.Ke * DCBF_ELP_REC;
    const double px = sm.x0[0], py = sm.x0[1];
    int Ks;
    {
        // one lane per obstacle: selection (MPC_LIP_modi.py:325-338), compaction, detour heuristic (MPC_LIP_sig_step.py:229-253)
        const int j = lane;
        const bool is_c = j < in.Kc, is_e = !is_c && j < in.Kc + in.Ke;
        double rec[6] = {0, 0, 1, 0, 1, 0};
        double dsel = 1e300;
        if (is_c) {
            const double *o = cir + DCBF_CIR_REC * j;
            rec[0] = o[0]; rec[1] = o[1]; rec[5] = o[2];
            dsel = (px - o[0]) * (px - o[0]) + (py - o[1]) * (py - o[1]) - o[2];
        } else if (is_e) {
            const double *o = elp + DCBF_ELP_REC * (j - in.Kc);
            rec[0] = o[0]; rec[1] = o[1]; rec[2] = o[2]; rec[3] = o[3]; rec[4] = o[4]; rec[5] = o[5];
            dsel = (px - o[0]) * (px - o[0]) + (py - o[1]) * (py - o[1]) - o[6];
        }
        const bool sel = (is_c || is_e) && (!P.select_obs || dsel <= P.detect_sq);
        const unsigned mask = __ballot_sync(FULL, sel);
        Ks = __popc(mask);
        if (sel) {
            const int pos = __popc(mask & ((1u << lane) - 1u));
#pragma unroll
            for (int c = 0; c < 6; c++) sm.obs[pos][c] = rec[c];
        }
        bool hit = false;
        double ngx = 0.0, ngy = 0.0;
        const double gx = sm.graw[0], gy = sm.graw[1];
        if (P.goal_shift && sel && is_c) {
            const double PI = 3.14159265358979323846;
            const double dg = (px - gx) * (px - gx) + (py - gy) * (py - gy);
            const double dc = (px - rec[0]) * (px - rec[0]) + (py - rec[1]) * (py - rec[1]);
            if (dc < dg && dc < 9.0 * rec[5]) {
                const double th = datan2(gy - py, gx - px), al = datan2(rec[1] - py, rec[0] - px);
                double d = th - al;
                if (d < 0.0 && fabs(d) > PI) d += 2.0 * PI;
                else if (d > 0.0 && fabs(d) > PI) d -= 2.0 * PI;
                if (fabs(d) < PI / 12.0) {
                    const double na = d < 0.0 ? th - PI / 12.0 : th + PI / 12.0;
                    const double rad = sqrt(dg);
                    double sn, cs;
                    dsincos(na, &sn, &cs);
                    ngx = px + rad * cs; ngy = py + rad * sn; hit = true;
                }
            }
        }
        const unsigned hm = __ballot_sync(FULL, hit);
        double g0 = gx, g1 = gy;
        if (hm) {
            const int src = __ffs(hm) - 1;
            g0 = __shfl_sync(FULL, ngx, src); g1 = __shfl_sync(FULL, ngy, src);
        }
        if (lane == 0) { sm.goal[0] = g0; sm.goal[1] = g1; }
        __syncwarp();
    }
    const bool has_fen = P.has_fen != 0;
    const int m = 3 * (Ks + (has_fen ? 6 : 4));
    RowDesc rd[NS];
    RowBnd rb[NS];
    double rs[NS], rzl[NS], rzu[NS], rds[NS], rel[NS], reu[NS];
    int nz_l = 0;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        rd[s] = row_desc(s * 32 + lane, Ks, has_fen);
        if (s * 32 + lane >= m) rd[s].type = RT_NONE;
        rb[s] = row_bounds(P, rd[s], leg);
        nz_l += (rb[s].has_lo ? 1 : 0) + (rb[s].has_hi ? 1 : 0);
        rs[s] = rzl[s] = rzu[s] = rds[s] = rel[s] = reu[s] = 0.0;
    }
    const int nz = wsumi(nz_l), nrows = m;
    // ---- solver state ------------------------------------------------------------------------------------------------
    S.mu = P.mu_init; S.sf = 1.0; S.alpha = 0.0; S.alpha_z = 0.0; S.delta_last = 0.0; S.lm_lambda = 1e-4; S.resto_target = 0.0;
    S.resto_entry = 0.0; S.theta_max = 1e300; S.theta_min = 0.0; S.nf = 0; S.iters = 0; S.acc_cnt = 0; S.status = -1; S.nstall = 0; S.tiny = 0;
    S.phase = PH_MAIN; S.pending = false; S.reinit = true; S.first = true; S.obj = 0.0; S.viol = 0.0;
    const double tol = P.tol;

    bool nodes_valid = false;      // sm.nodes / trig / nobj (with Hessian terms) already describe zc (staged by the accepted trial)
    double carry_log = 0.0;        // sum of log(gaps) at the accepted trial point = barrier term of the next full pass
    bool carry_ok = false;
    for (;;) {
        const bool resto = S.phase == PH_RESTO;
        if (!nodes_valid) w_nodes<NS>(cs_, sm, sm.zc, lane, S.first ? 1.0 : (resto ? 0.0 : S.sf), true);
        nodes_valid = false;
        // objective value and gradient (lane a < 9 owns grad[a])
        const double fobj = sm.nobj[1][0] + sm.nobj[2][0] + sm.nobj[3][0];
        double grad_a = 0.0;
        if (lane < 9) {
#pragma unroll
            for (int kn = 1; kn <= 3; kn++)
                grad_a += sm.nobj[kn][1] * T[FN(kn, 0)][lane] + sm.nobj[kn][2] * T[FN(kn, 1)][lane] + sm.nobj[kn][3] * T[FN(kn, 4)][lane];
        }
        if (S.first) {
            const double gmax = wmax(fabs(grad_a));
            S.sf = gmax > 100.0 ? ddiv(100.0, gmax) : 1.0;
            if (lane < 3) {   // the objective Hessian staged above used sf = 1: rescale
#pragma unroll
                for (int c = 4; c < 10; c++) sm.nobj[lane + 1][c] *= S.sf;
            }
            __syncwarp();
        }
        // ---- rows: evaluate, update row state, stage gradients and weights; statistics stay in registers ----------------------
        Stat8 st8;
        st8.s0 = st8.s1 = st8.s2 = st8.s3 = 0.0; st8.m0 = 0.0; st8.m1 = -1e300; st8.m2 = 0.0; st8.m3 = 0.0;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int r = s * 32 + lane;
            RowEval e;
            eval_row<NS, true>(P, sm, rd[s], sm.zc, e);
            const RowBnd &bb = rb[s];
            double sig = 0.0, w1 = 0.0, binv = 0.0, y = 0.0;
            double t_rc = 0.0, t_cmin = 1e300, t_cmax = 0.0, t_z = 0.0, t_log = 0.0, t_v2 = 0.0, t_v = 0.0;
            if (rd[s].type != RT_NONE) {
                double v = 0.0;
                if (bb.has_lo && e.c < bb.lo) v = e.c - bb.lo;
                if (bb.has_hi && e.c > bb.hi) v = e.c - bb.hi;
                t_v2 = v * v; t_v = fabs(v);
                if (resto) {
                    sig = v != 0.0 ? 1.0 : 0.0; w1 = v; y = v;
                } else {
                    const double lr = bb.has_lo ? relax_lo(bb.lo) : 0.0, hr = bb.has_hi ? relax_hi(bb.hi) : 0.0;
                    if (S.reinit) {
                        double sv = e.c;
                        if (bb.has_lo && bb.has_hi) {
                            const double pl = fmin(1e-2 * fmax(1.0, fabs(lr)), 1e-2 * (hr - lr));
                            const double pu = fmin(1e-2 * fmax(1.0, fabs(hr)), 1e-2 * (hr - lr));
                            sv = fmin(fmax(sv, lr + pl), hr - pu);
                        } else if (bb.has_lo) sv = fmax(sv, lr + 1e-2 * fmax(1.0, fabs(lr)));
                        else if (bb.has_hi) sv = fmin(sv, hr - 1e-2 * fmax(1.0, fabs(hr)));
                        rs[s] = sv; rzl[s] = bb.has_lo ? 1.0 : 0.0; rzu[s] = bb.has_hi ? 1.0 : 0.0;
                    } else if (S.pending) {
                        rs[s] += S.alpha * rds[s];
                        if (bb.has_lo) {
                            const double gap = rs[s] - lr;
                            const double mg = ddiv(S.mu, gap);
                            rzl[s] = fmax(fmin(rzl[s] + S.alpha_z * rel[s], DCBF_KAPPA_SIGMA * mg), mg * (1.0 / DCBF_KAPPA_SIGMA));
                        }
                        if (bb.has_hi) {
                            const double gap = hr - rs[s];
                            const double mg = ddiv(S.mu, gap);
                            rzu[s] = fmax(fmin(rzu[s] + S.alpha_z * reu[s], DCBF_KAPPA_SIGMA * mg), mg * (1.0 / DCBF_KAPPA_SIGMA));
                        }
                    }
                    const double rc = e.c - rs[s];
                    double lp = 1.0;
                    if (bb.has_lo) {
                        const double gap = rs[s] - lr, inv = drcp(gap);
                        sig += rzl[s] * inv; binv += inv; y -= rzl[s];
                        const double cz = gap * rzl[s];
                        t_cmin = fmin(t_cmin, cz); t_cmax = fmax(t_cmax, cz); t_z += rzl[s];
                        lp *= gap; rel[s] = inv;
                    }
                    if (bb.has_hi) {
                        const double gap = hr - rs[s], inv = drcp(gap);
                        sig += rzu[s] * inv; binv -= inv; y += rzu[s];
                        const double cz = gap * rzu[s];
                        t_cmin = fmin(t_cmin, cz); t_cmax = fmax(t_cmax, cz); t_z += rzu[s];
                        lp *= gap; reu[s] = inv;
                    }
                    if (!carry_ok) t_log = dlog(lp);
                    rds[s] = rc;
                    t_rc = fabs(rc);
                    w1 = sig * rc;
                }
            }
            // stage (rows beyond m stage neutral values so that the entry-parallel loops need no guards)
            double g[9];
#pragma unroll
            for (int a = 0; a < 9; a++) g[a] = 0.0;
#pragma unroll
            for (int p = 0; p < 4; p++) {
                if (p < e.nf) {
                    const double dp = e.d[p];
                    const double *Tf = T[e.f[p]];
#pragma unroll
                    for (int a = 0; a < 9; a++) g[a] = fma(dp, Tf[a], g[a]);
                }
            }
#pragma unroll
            for (int a = 0; a < 9; a++) { sm.G[r][a] = g[a]; sm.SG[r][a] = sig * g[a]; }
            sm.RW[r][0] = sig; sm.RW[r][1] = w1; sm.RW[r][2] = binv; sm.RW[r][3] = y;
            sm.HQ[r][0] = y * e.hq0; sm.HQ[r][1] = y * e.hq1; sm.HQ[r][2] = y * e.hq2;
            st8.s0 += t_rc; st8.s1 += t_z; st8.s2 += t_log; st8.s3 += t_v2;
            st8.m0 = fmax(st8.m0, t_rc); st8.m1 = fmax(st8.m1, -t_cmin); st8.m2 = fmax(st8.m2, t_cmax); st8.m3 = fmax(st8.m3, t_v);
        }
        reduce8_inline(st8);
        __syncwarp();
        const double st_theta = st8.s0, st_zsum = st8.s1, st_logsum = carry_ok ? carry_log : st8.s2, st_v2 = st8.s3, st_pinf = st8.m0,
                     st_cmin = -st8.m1, st_cmax = st8.m2, st_vmax = st8.m3;
        carry_ok = false;
        // ---- node Hessians ---------------------------------------------------------------------------------------------
        {
            const double gm1 = P.gamma - 1.0;
            if (lane < 9) {            // position block: node kn = lane/3 + 1, component lane % 3
                const int kn = lane / 3 + 1, c = lane % 3;
                double acc = sm.nobj[kn][4 + c];
                for (int j = 0; j < Ks; j++) acc += sm.HQ[(kn - 1) * Ks + j][c];
                if (kn < 3) for (int j = 0; j < Ks; j++) acc = fma(gm1, sm.HQ[kn * Ks + j][c], acc);
                sm.NH[kn][c] = acc;
            } else if (lane < 12) {    // heading / velocity entries of node kn = lane - 8 (rows of step kn - 1)
                const int kn = lane - 8, i = kn - 1, base = 3 * Ks;
                double Yx = sm.RW[base + i][3], Yy = sm.RW[base + 3 + i][3];
                if (has_fen) Yx += sm.RW[base + 12 + i][3] + sm.RW[base + 15 + i][3];
                const double sn = sm.trig[kn][0], cs = sm.trig[kn][1];
                const double vx = sm.nodes[kn][2], vy = sm.nodes[kn][3];
                const double vbx = cs * vx + sn * vy, vby = -sn * vx + cs * vy;
                sm.NH[kn][3] = sm.nobj[kn][7]; sm.NH[kn][4] = sm.nobj[kn][8];
                sm.NH[kn][5] = sm.nobj[kn][9] - (Yx * vbx + Yy * vby);
                sm.NH[kn][6] = -sn * Yx - cs * Yy; sm.NH[kn][7] = cs * Yx - sn * Yy;
                sm.legy[i] = 2.0 * sm.RW[base + 6 + i][3];
            }
        }
        __syncwarp();
        // ---- condensed matrix and J^T vectors, entry-parallel (72 entries over 32 lanes) ---------------------------------------
        for (int e = lane; e < 72; e += 32) {
            const int a = cs_.ea[e], bcol = cs_.eb[e];
            double acc = 0.0;
            if (e < 45) {
                const double *pa = &sm.SG[0][a], *pb = &sm.G[0][bcol];
                double acc1 = 0.0;
                int r = 0;
                for (; r + 1 < m; r += 2) { acc = fma(pa[9 * r], pb[9 * r], acc); acc1 = fma(pa[9 * r + 9], pb[9 * r + 9], acc1); }
                if (r < m) acc = fma(pa[9 * r], pb[9 * r], acc);
                acc += acc1;
                // Lagrangian Hessian through the feature map
#pragma unroll 1
                for (int kn = 1; kn <= 3; kn++) {
                    const double Xa = T[FN(kn, 0)][a], Ya = T[FN(kn, 1)][a], VXa = T[FN(kn, 2)][a], VYa = T[FN(kn, 3)][a], Ta = T[FN(kn, 4)][a];
                    const double Xb = T[FN(kn, 0)][bcol], Yb = T[FN(kn, 1)][bcol], VXb = T[FN(kn, 2)][bcol], VYb = T[FN(kn, 3)][bcol], Tb = T[FN(kn, 4)][bcol];
                    const double *H = sm.NH[kn];
                    const double ux = H[0] * Xb + H[1] * Yb + H[3] * Tb;
                    const double uy = H[1] * Xb + H[2] * Yb + H[4] * Tb;
                    const double ut = H[3] * Xb + H[4] * Yb + H[5] * Tb + H[6] * VXb + H[7] * VYb;
                    acc += Xa * ux + Ya * uy + Ta * ut + (VXa * H[6] + VYa * H[7]) * Tb;
                }
#pragma unroll 1
                for (int i = 0; i < 3; i++)
                    acc += sm.legy[i] * (T[FLX(i)][a] * T[FLX(i)][bcol] + T[FLY(i)][a] * T[FLY(i)][bcol]);
                sm.Kf[e] = acc;
            } else {
                const double *pw = &sm.RW[0][bcol], *pg = &sm.G[0][a];   // bcol = 1 + vector index
                double acc1 = 0.0;
                int r = 0;
                for (; r + 1 < m; r += 2) { acc = fma(pw[4 * r], pg[9 * r], acc); acc1 = fma(pw[4 * r + 4], pg[9 * r + 9], acc1); }
                if (r < m) acc = fma(pw[4 * r], pg[9 * r], acc);
                acc += acc1;
                sm.q[e - 45] = acc;
            }
        }
        __syncwarp();
        S.pending = false; S.reinit = false;
        S.obj = fobj; S.viol = st_vmax;
        if (!(fobj == fobj) || !(st_theta == st_theta)) { S.status = -13; break; }
        // ---- convergence / barrier update / right-hand side ---------------------------------------------------------------
        double rhs_a = 0.0;
        if (!resto) {
            if (S.first) { S.theta_max = 1e4 * fmax(1.0, st_theta); S.theta_min = 1e-4 * fmax(1.0, st_theta); S.first = false; }
            const double dinf = wmax(lane < 9 ? fabs(fma(S.sf, grad_a, sm.q[18 + lane])) : 0.0);
            const double sd = fmax(100.0, ddiv(2.0 * st_zsum, (double)(nrows + nz))) * 0.01;
            const double sc = fmax(100.0, ddiv(st_zsum, (double)(nz > 0 ? nz : 1))) * 0.01;
            const double isd = drcp(sd), isc = drcp(sc);
            double E0;
            for (;;) {
                const double compm = fmax(fabs(st_cmax - S.mu), fabs(st_cmin - S.mu));
                E0 = fmax(fmax(dinf * isd, st_pinf), st_cmax * isc);
                const double Emu = fmax(fmax(dinf * isd, st_pinf), compm * isc);
                if (E0 <= tol) break;
                if (Emu <= 10.0 * S.mu && S.mu > tol * 0.1 * (1.0 + 1e-12)) {
                    S.mu = fmax(tol * 0.1, fmin(0.2 * S.mu, S.mu * sqrt(S.mu)));
                    S.nf = 0;
                    continue;
                }
                break;
            }
            if (E0 <= tol) { S.status = 0; break; }
            if (E0 <= 1e-6 && st_vmax <= P.constr_viol_tol) { if (++S.acc_cnt >= 15) { S.status = 1; break; } } else S.acc_cnt = 0;
            if (S.iters >= P.max_iter) { S.status = -1; break; }
            if (S.tiny >= 3) {   // pinned by the fraction-to-boundary rule while still infeasible: restoration now (see ipm_iterate())
                S.tiny = 0;
                const int slot = S.nf < DCBF_FILT ? S.nf : (S.iters % DCBF_FILT);
                if (lane == 0) { sm.filt_th[slot] = (1.0 - 1e-5) * st_theta; sm.filt_ph[slot] = (S.sf * fobj - S.mu * st_logsum) - 1e-5 * st_theta; }
                if (S.nf < DCBF_FILT) S.nf++;
                __syncwarp();
                S.phase = PH_RESTO; S.resto_entry = st_vmax; S.resto_target = fmax(0.1 * st_vmax, 1e-9); S.lm_lambda = 1e-4; S.acc_cnt = 0;
                S.iters++;
                continue;
            }
            if (lane < 9) rhs_a = -S.sf * grad_a - sm.q[lane] + S.mu * sm.q[9 + lane];
        } else {
            if (st_vmax <= S.resto_target) { S.phase = PH_MAIN; S.reinit = true; continue; }
            const double gn = wmax(lane < 9 ? fabs(sm.q[lane]) : 0.0);
            const bool stationary = gn <= 1e-10 * fmax(1.0, st_vmax) || S.lm_lambda > 1e12;
            if (stationary) {
                if (st_vmax > P.constr_viol_tol) { S.status = 2; break; }
                if (S.resto_entry <= 1e-9 || S.nstall++ >= 1) { S.status = -2; break; }   // see ipm_iterate()
                S.phase = PH_MAIN; S.reinit = true; continue;
            }
            if (S.iters >= P.max_iter) { S.status = -1; break; }
            if (lane < 9) rhs_a = -sm.q[lane];
        }
        // ---- restoration: Levenberg-Marquardt trials reuse the assembled K while lambda is escalated ---------------------------
        // ---- main phase: one factorisation with inertia correction by delta ----------------------------------------------------
        bool lm_accept = false;
        double v2t = 0.0, vmt = 0.0;
        for (int rt = 0; rt < 20; rt++) {
            double shift = resto ? S.lm_lambda : 0.0;
            bool ok = false;
            for (int tr = 0; tr < 48; tr++) {
                double Lrow[9];
                ok = true;
#pragma unroll
                for (int j = 0; j < 9; j++) {
                    double s_ = 0.0;
                    if (lane >= j && lane < 9) {
                        s_ = sm.Kf[tri(lane, j)] + (lane == j ? shift : 0.0);
#pragma unroll
                        for (int c = 0; c < j; c++) s_ = fma(-Lrow[c], sm.Lf[tri(j, c)], s_);
                    }
                    const double d = __shfl_sync(FULL, s_, j);
                    if (!(d > 1e-14)) { ok = false; break; }
                    const double rinv = drsqrt(d);
                    Lrow[j] = lane == j ? rinv : s_ * rinv;   // diagonal stored as its reciprocal
                    if (lane >= j && lane < 9) sm.Lf[tri(lane, j)] = Lrow[j];
                    __syncwarp();
                }
                if (ok || resto) break;
                if (shift == 0.0) shift = S.delta_last == 0.0 ? 1e-4 : fmax(1e-20, S.delta_last * (1.0 / 3.0));
                else shift *= (S.delta_last == 0.0 ? 100.0 : 8.0);
            }
            if (!ok) {
                if (!resto) { S.status = -3; break; }
                S.lm_lambda *= 10.0;
                if (S.lm_lambda > 1e12) break;
                continue;
            }
            if (!resto && shift > 0.0) S.delta_last = shift;
            // triangular solves: lane i holds component i
            {
                double bi = rhs_a;
#pragma unroll
                for (int c = 0; c < 9; c++) {
                    const double yc = __shfl_sync(FULL, bi * sm.Lf[tri(c, c)], c);   // y_c = b_c / L_cc
                    if (lane == c) bi = yc;
                    else if (lane > c && lane < 9) bi = fma(-sm.Lf[tri(lane, c)], yc, bi);
                }
#pragma unroll
                for (int c = 8; c >= 0; c--) {
                    const double xc = __shfl_sync(FULL, bi * sm.Lf[tri(c, c)], c);
                    if (lane == c) bi = xc;
                    else if (lane < c) bi = fma(-sm.Lf[tri(c, lane)], xc, bi);
                }
                if (lane < 9) sm.dz[lane] = bi;
            }
            __syncwarp();
            if (!resto) break;
            // Levenberg-Marquardt trial at full step (violation only)
            if (lane < 9) sm.zt[lane] = sm.zc[lane] + sm.dz[lane];
            __syncwarp();
            w_nodes<NS>(cs_, sm, sm.zt, lane, 0.0, false);
            v2t = 0.0; vmt = 0.0;
#pragma unroll
            for (int s = 0; s < NS; s++) {
                if (rd[s].type == RT_NONE) continue;
                RowEval e;
                eval_row<NS, false>(P, sm, rd[s], sm.zt, e);
                double v = 0.0;
                if (rb[s].has_lo && e.c < rb[s].lo) v = e.c - rb[s].lo;
                if (rb[s].has_hi && e.c > rb[s].hi) v = e.c - rb[s].hi;
                v2t += v * v; vmt = fmax(vmt, fabs(v));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                v2t += __shfl_xor_sync(FULL, v2t, o);
                vmt = fmax(vmt, __shfl_xor_sync(FULL, vmt, o));
            }
            if (v2t < st_v2 * (1.0 - 1e-12)) { lm_accept = true; break; }
            S.lm_lambda *= 10.0;
            if (S.lm_lambda > 1e12) break;
        }
        if (S.status == -3) break;
        if (resto) {
            if (lm_accept) {
                double dn = lane < 9 ? fabs(sm.dz[lane]) : 0.0;
                dn = wmax(dn);
                if (lane < 9) sm.zc[lane] = sm.zt[lane];
                S.iters++;
                S.lm_lambda = fmax(S.lm_lambda * 0.2, 1e-12);
                if (st_v2 - v2t <= 1e-4 * st_v2) S.acc_cnt++; else S.acc_cnt = 0;
                if ((dn < 1e-12 || S.acc_cnt >= 2) && vmt > S.resto_target) S.lm_lambda = 1e13;
            }
            __syncwarp();
            continue;
        }
        // ---- direction pass (main phase): ds, dz_L, dz_U, step sizes -----------------------------------------------------------
        double amax = 1.0, az = 1.0, dphi = 0.0;
        {
            const double tau = fmax(0.99, 1.0 - S.mu);
            if (lane < 9) dphi = S.sf * grad_a * sm.dz[lane];
#pragma unroll
            for (int s = 0; s < NS; s++) {
                if (rd[s].type == RT_NONE) continue;
                const int r = s * 32 + lane;
                double jd = 0.0;
#pragma unroll
                for (int a = 0; a < 9; a++) jd = fma(sm.G[r][a], sm.dz[a], jd);
                const RowBnd &e = rb[s];
                const double d = jd + rds[s];
                rds[s] = d;
                if (e.has_lo) {
                    const double inv = rel[s], gap = rs[s] - relax_lo(e.lo);
                    const double dzl = S.mu * inv - rzl[s] - rzl[s] * inv * d;
                    rel[s] = dzl;
                    dphi -= S.mu * d * inv;
                    if (d < 0.0) amax = fmin(amax, ddiv(-tau * gap, d));
                    if (dzl < 0.0) az = fmin(az, ddiv(-tau * rzl[s], dzl));
                }
                if (e.has_hi) {
                    const double inv = reu[s], gap = relax_hi(e.hi) - rs[s];
                    const double dzu = S.mu * inv - rzu[s] + rzu[s] * inv * d;
                    reu[s] = dzu;
                    dphi += S.mu * d * inv;
                    if (d > 0.0) amax = fmin(amax, ddiv(tau * gap, d));
                    if (dzu < 0.0) az = fmin(az, ddiv(-tau * rzu[s], dzu));
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {   // three reductions interleaved
                amax = fmin(amax, __shfl_xor_sync(FULL, amax, o));
                az = fmin(az, __shfl_xor_sync(FULL, az, o));
                dphi += __shfl_xor_sync(FULL, dphi, o);
            }
        }
        // ---- filter line search -------------------------------------------------------------------------------------------------
        const double theta = st_theta;
        const double phi = S.sf * fobj - S.mu * st_logsum;
        const double eps_phi = 10.0 * 2.2e-16 * fabs(phi);
        double alpha = amax;
        int accepted = 0;
        for (int ls = 0; ls < DCBF_LS_MAX; ls++, alpha *= 0.5) {
            if (lane < 9) sm.zt[lane] = fma(alpha, sm.dz[lane], sm.zc[lane]);
            __syncwarp();
            w_nodes<NS>(cs_, sm, sm.zt, lane, S.sf, true);
            const double ft = sm.nobj[1][0] + sm.nobj[2][0] + sm.nobj[3][0];
            double th_t = 0.0, lg_t = 0.0;
            bool okv = true;
#pragma unroll
            for (int s = 0; s < NS; s++) {
                if (rd[s].type == RT_NONE) continue;
                RowEval e;
                eval_row<NS, false>(P, sm, rd[s], sm.zt, e);
                const double stv = rs[s] + alpha * rds[s];
                th_t += fabs(e.c - stv);
                double lp = 1.0;
                if (rb[s].has_lo) { const double gap = stv - relax_lo(rb[s].lo); if (!(gap > 0.0)) okv = false; lp *= gap; }
                if (rb[s].has_hi) { const double gap = relax_hi(rb[s].hi) - stv; if (!(gap > 0.0)) okv = false; lp *= gap; }
                lg_t += dlog(lp);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                th_t += __shfl_xor_sync(FULL, th_t, o);
                lg_t += __shfl_xor_sync(FULL, lg_t, o);
            }
            okv = __all_sync(FULL, okv);
            const double ph_t = S.sf * ft - S.mu * lg_t;
            if (!okv || !(ph_t == ph_t) || !(th_t <= S.theta_max)) continue;
            bool in_filter = false;
            for (int q = 0; q < S.nf; q++)
                if (th_t >= sm.filt_th[q] && ph_t >= sm.filt_ph[q]) in_filter = true;
            if (in_filter) continue;
            const bool sw = dphi < 0.0 && theta <= S.theta_min && switch_cond(alpha, -dphi, theta);
            if (sw) { if (ph_t <= phi + 1e-8 * alpha * dphi + eps_phi) accepted = 1; }
            else if (th_t <= (1.0 - 1e-5) * theta || ph_t <= phi - 1e-5 * theta + eps_phi) accepted = 2;
            if (accepted) { carry_log = lg_t; break; }
        }
        if (accepted != 1) {   // filter augmentation (also before entering restoration)
            const int slot = S.nf < DCBF_FILT ? S.nf : (S.iters % DCBF_FILT);
            if (lane == 0) { sm.filt_th[slot] = (1.0 - 1e-5) * theta; sm.filt_ph[slot] = phi - 1e-5 * theta; }
            if (S.nf < DCBF_FILT) S.nf++;
            __syncwarp();
        }
        if (!accepted) {
            S.phase = PH_RESTO; S.resto_entry = st_vmax; S.resto_target = fmax(0.1 * st_vmax, 1e-9); S.lm_lambda = 1e-4; S.acc_cnt = 0;
            S.iters++;
            continue;
        }
        if (lane < 9) sm.zc[lane] = sm.zt[lane];
        __syncwarp();
        S.alpha = alpha; S.alpha_z = az; S.pending = true;
        nodes_valid = true; carry_ok = true;   // the accepted trial staged the nodes (with Hessian terms) and the barrier sum of the new point
        if (alpha < 1e-2 && st_vmax > P.constr_viol_tol) S.tiny++; else S.tiny = 0;
        S.iters++;
    }
    // every exit leaves the loop right after a full pass (or before any trial), so sm.nodes is the rollout of the final iterate
}

template <int NS>
__device__ __forceinline__ bool w_close(const dcbf_params &P, const WarpShared<NS> &sm) {
    bool close = false;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double dxg = sm.nodes[i + 1][0] - sm.graw[0], dyg = sm.nodes[i + 1][1] - sm.graw[1];
        if ((i == 0 || P.close_any) && sqrt(dxg * dxg + dyg * dyg) <= P.close_radius) close = true;
    }
    return close;
}

// per-CTA staging of the constants
__device__ __forceinline__ void stage_cta(const dcbf_params &P, const Consts &K, CtaShared &cs_) {
    if (threadIdx.x == 0) { cs_.P = P; cs_.K = K; }
    for (int e = threadIdx.x; e < 72; e += blockDim.x) {
        if (e < 45) {
            int a = 0;
            while ((a + 1) * (a + 2) / 2 <= e) a++;
            cs_.ea[e] = a; cs_.eb[e] = e - a * (a + 1) / 2;
        } else {
            const int t = e - 45;
            cs_.ea[e] = t % 9; cs_.eb[e] = 1 + t / 9;
        }
    }
    build_T(K, cs_.T);
    __syncthreads();
}

}  // namespace wp
}  // namespace dcbf
