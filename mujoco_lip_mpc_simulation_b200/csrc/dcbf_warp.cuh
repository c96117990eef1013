// dcbf_warp.cuh -- warp-cooperative variant of the LIP solver: ONE PROBLEM PER WARP (device only).
//
// The per-thread kernels of dcbf_lanes.cuh have the lowest instruction count per problem but a long serial chain per
// interior-point iteration and, at 255 registers, only two warps per scheduler: a 4096-scenario batch is bound by
// single-warp latency times the slowest problem's iteration count (profiles/r01_summary.md).  Here the 32 lanes of a
// warp share one problem:
//   * rows are ordered step-major (all rows of step 0, then step 1, then step 2) and distributed over lanes (row r ->
//     lane r % 32, slot r / 32); every lane evaluates its rows, applies the slack/multiplier updates in registers and
//     stages the row gradient, sigma * gradient and the weights TRANSPOSED in shared memory (ST[a][r]: consecutive lanes
//     write consecutive words, and an entry-parallel reader fetches two rows per 128-bit load);
//   * row gradients come from six scalars per row and two constant coefficient vectors per (row class, step): in
//     z = (foot, turn) space the rollout is affine, so d row / d foot_l = cA[l] * (d row / d node_{i+1}) + cB[l] * (d row /
//     d node_i);
//   * the condensed matrix  W + J^T Sigma J  and the three J^T-vectors are 72 dot products over the staged rows,
//     distributed over the lanes in three rounds sorted by the first step whose rows can contribute (an entry that
//     involves a variable of step s only sees rows of steps >= s); the Lagrangian Hessian is added from 27 node
//     quantities through a host-built sparse table (at most 6 terms per entry);
//   * norms, merit values and step sizes are warp-shuffle reductions; the 9x9 Cholesky is row-owned (lane i owns row
//     i, lane 9 carries the right-hand side as a tenth row, which is the forward substitution) with the factor staged
//     in shared memory; all scalar control flow is replicated and therefore uniform.
// The algorithm (barrier rule, filter, restoration, status codes) is the one of ipm_iterate() in dcbf_core.cuh.
#pragma once
#include "dcbf_lanes.cuh"

namespace dcbf {
namespace wp {

enum { RT_NONE = 0, RT_CBF, RT_VBX, RT_VBY, RT_LEG, RT_DTH, RT_FENP, RT_FENM };
constexpr unsigned FULL = 0xffffffffu;
constexpr int NSRC = 27;     // node quantities feeding the Lagrangian Hessian: 3 nodes x 8 entries, 3 leg multipliers
constexpr int NHT = 6;       // terms per matrix entry in the Hessian table
// flat layout of the assembled system: q1 q2 q3 (27) | pad | K (45, packed rows 0..8) | right-hand side (= packed row 9) | tail.
// The tail lets all 32 lanes store "their" right-hand-side component without a lane test (only 0..8 are meaningful).
constexpr int KQ_Q = 0;
constexpr int KQ_K = 28;
constexpr int KQ_RHS = KQ_K + 45;
constexpr int KQ_LEN = KQ_RHS + 32;
constexpr int LF_LEN = 56 + 32;   // factor rows 0..9, then one dump slot per lane for the branch-free stores

// Host-built constant tables, one copy per context in device memory (build_warp_tables()).
struct WarpTables {
    int desc[96];             // 3 rounds x 32 lanes: rowP | rowQ << 8 | class << 16 | out << 20   (-1: idle)
    double hc[NHT][48];       // K[e] += hc[t][e] * src[hs[t][e]]
    unsigned char hs[NHT][48];
    double cab[10][6];        // (cA[3], cB[3]) per row class: CBF step 0..2, velocity rows step 0..2, leg step 0..2, turn
};

// variable order (fx0, fy0, fx1, fy1, fx2, fy2, t0, t1, t2): step of a variable
inline int var_step(int a) { return a < 6 ? a >> 1 : a - 6; }

// d feature / d z for the 24 node features (3 nodes x (x, y, vx, vy, th) | 3 x (lx, ly) | 3 x dth); host only
inline void build_feature_map(const Consts &k, double (*T)[9]) {
    for (int f = 0; f < 24; f++) for (int a = 0; a < 9; a++) {
        double v = 0.0;
        const bool ax = a < 6 && (a & 1) == 0, ay = a < 6 && (a & 1) == 1;
        const int l = var_step(a);
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
        } else if (a >= 6 && l == f - 21) v = 1.0;
        T[f][a] = v;
    }
}

inline bool build_warp_tables(const Consts &k, WarpTables &W) {
    // ---- the 72 dot products, sorted by class = first step whose rows can contribute ---------------------------------
    struct Ent { int rowP, rowQ, cls, out; };
    Ent ent[72];
    int n = 0;
    for (int cls = 0; cls < 3; cls++) {
        for (int a = 0; a < 9; a++) for (int b = 0; b <= a; b++) {
            const int c = var_step(a) > var_step(b) ? var_step(a) : var_step(b);
            if (c == cls) ent[n++] = {9 + a, b, cls, KQ_K + tri(a, b)};            // sum_r ST[9 + a][r] * ST[b][r]
        }
        for (int v = 0; v < 3; v++) for (int a = 0; a < 9; a++)
            if (var_step(a) == cls) ent[n++] = {18 + 1 + v, a, cls, KQ_Q + 9 * v + a};   // sum_r ST[19 + v][r] * ST[a][r]
    }
    if (n != 72) return false;
    for (int t = 0; t < 96; t++) W.desc[t] = t < 72 ? (ent[t].rowP | ent[t].rowQ << 8 | ent[t].cls << 16 | ent[t].out << 20) : -1;
    // ---- Lagrangian Hessian map: sources = NH[kn][0..7] (xx, xy, yy, xt, yt, tt, vxt, vyt) for kn = 1..3, then legy[0..2] --
    double T[24][9];
    build_feature_map(k, T);
    for (int t = 0; t < NHT; t++) for (int e = 0; e < 48; e++) { W.hc[t][e] = 0.0; W.hs[t][e] = (unsigned char)NSRC; }
    static const int HP[8] = {0, 0, 1, 0, 1, 4, 2, 3}, HQ[8] = {0, 1, 1, 4, 4, 4, 4, 4};   // (p, q) of the 5x5 node Hessian
    for (int a = 0; a < 9; a++) for (int b = 0; b <= a; b++) {
        const int e = tri(a, b);
        int cnt = 0;
        for (int s = 0; s < NSRC; s++) {
            double c = 0.0;
            if (s < 24) {
                const int kn = s / 8 + 1, h = s % 8, p = HP[h], q = HQ[h], fp = 5 * (kn - 1) + p, fq = 5 * (kn - 1) + q;
                c = T[fp][a] * T[fq][b];
                if (p != q) c += T[fq][a] * T[fp][b];
            } else {
                const int i = s - 24;
                c = T[15 + 2 * i][a] * T[15 + 2 * i][b] + T[16 + 2 * i][a] * T[16 + 2 * i][b];
            }
            if (c != 0.0) {
                if (cnt >= NHT) return false;
                W.hc[cnt][e] = c; W.hs[cnt][e] = (unsigned char)s; cnt++;
            }
        }
    }
    // ---- gradient coefficient vectors ------------------------------------------------------------------------------------
    for (int c = 0; c < 10; c++) for (int j = 0; j < 6; j++) W.cab[c][j] = 0.0;
    for (int i = 0; i < 3; i++) for (int l = 0; l < 3; l++) {
        if (l <= i) { W.cab[i][l] = k.gx[i - l]; W.cab[3 + i][l] = k.gv[i - l]; }       // cA of D-CBF / velocity rows
        if (l < i) { W.cab[i][3 + l] = k.gx[i - 1 - l]; W.cab[6 + i][3 + l] = k.gx[i - 1 - l]; }   // cB of D-CBF / leg rows
        if (l == i) W.cab[6 + i][3 + l] = -1.0;
    }
    return true;
}

#if defined(__CUDACC__)

template <int NS> struct KsMax { static constexpr int v = NS == 1 ? 6 : (NS == 2 ? 17 : 2 * DCBF_MAX_OBS); };

template <int NS>
struct alignas(16) WarpShared {
    static constexpr int RP = 32 * NS + 2;   // padded row length: class starts / ends are rounded to even rows
    double ST[22][RP];       // staged rows, transposed: 0..8 gradient, 9..17 sigma * gradient, 18..21 sigma, w1, binv, y
    double HQ[32 * NS][3];   // y * (2a', b', 2c') of the D-CBF rows
    double obs[KsMax<NS>::v][6];   // selected obstacles: cx, cy, a', b', c', rhs
    double KQ[KQ_LEN];       // assembled system (see KQ_*)
    double Lf[LF_LEN];       // Cholesky factor (packed rows 0..8), the forward-substituted right-hand side (row 9), dump slots
    double dz[32], zc[32], zt[32];   // 9 meaningful entries each; all lanes store
    double x0[5], goal[2], graw[2];
    double fr[4][4];         // free response (zero foot placements) x, y, vx, vy at nodes 0..3
    double nodes[4][5];      // x, y, vx, vy, th of nodes 0..3
    double trig[4][3];       // sin, cos, atan2 target of nodes 1..3
    double nobj[4][10];      // f_k, nx, ny, nt, hxx, hxy, hyy, hxt, hyt, htt of the objective at node k
    double NHf[NSRC + 1];    // Hessian sources (+ one zero for the padding terms of the table)
    double cold[12];         // replicated scalars that are written once per event and read much later (see C_*): kept out of registers
    double filt_th[DCBF_FILT], filt_ph[DCBF_FILT];
};

// per-CTA constants staged in shared memory
struct CtaShared {
    dcbf_params P;
    Consts K;
    double cab[10][6];
    double hc[NHT][48];          // copy of the Hessian table (global loads in the assembly loop cost a long-scoreboard stall each)
    unsigned char hs[NHT][48];
    const WarpTables *tab;
};

// one warp per CTA: the per-problem scratch and the constants are static shared-memory objects, so every function sees
// them as shared-space symbols (LDS/STS with immediate offsets, no generic pointers through the out-of-line calls)
template <int NS> __shared__ WarpShared<NS> g_sm;
__shared__ CtaShared g_cs;

// %laneid through a volatile asm: the value stays in a register (the compiler otherwise re-reads SR_TID.X at every use)
__device__ __forceinline__ int lane_id() {
    int l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

__device__ __noinline__ double wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int wsumi(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

struct RowDesc { int type, step, obs, cls; };

// step-major row order: step i holds  [D-CBF x Ks, v_bx, v_by, leg, turn, (fen+, fen-)]
__device__ __forceinline__ RowDesc row_desc(int r, int Ks, int ms, int m) {
    RowDesc d;
    d.type = RT_NONE; d.step = 0; d.obs = 0; d.cls = 9;
    if (r >= m) return d;
    d.step = r / ms;
    const int w = r - d.step * ms;
    if (w < Ks) { d.type = RT_CBF; d.obs = w; d.cls = d.step; }
    else {
        d.type = RT_VBX + (w - Ks);
        d.cls = d.type == RT_LEG ? 6 + d.step : (d.type == RT_DTH ? 9 : 3 + d.step);
    }
    return d;
}

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

// value of one row and, with GRAD, the six scalars its gradient is made of:
//   d row / d foot_l = cA[l] * (p0, p1) + cB[l] * (q0, q1),   d row / d turn_l = (l <= step) * t_all + (l == step) * t_own
struct RowEval { double c, p0, p1, q0, q1, t_all, t_own, hq0, hq1, hq2; };

template <int NS, bool GRAD>
__device__ __forceinline__ void eval_row(const dcbf_params &P, const WarpShared<NS> &sm, const RowDesc &rd, const double *z, RowEval &e) {
    e.c = 0.0;
    if (GRAD) { e.p0 = e.p1 = e.q0 = e.q1 = e.t_all = e.t_own = 0.0; e.hq0 = e.hq1 = e.hq2 = 0.0; }
    const int i = rd.step, kn = i + 1;
    if (rd.type == RT_CBF) {
        const double *o = sm.obs[rd.obs];
        const double gm1 = P.gamma - 1.0;
        const double ax = sm.nodes[kn][0] - o[0], ay = sm.nodes[kn][1] - o[1], bx = sm.nodes[i][0] - o[0], by = sm.nodes[i][1] - o[1];
        const double ea = o[2], eb = o[3], ec = o[4];
        e.c = (ea * ax * ax + eb * ax * ay + ec * ay * ay - o[5]) + gm1 * (ea * bx * bx + eb * bx * by + ec * by * by - o[5]);
        if (GRAD) {
            e.p0 = 2.0 * ea * ax + eb * ay; e.p1 = 2.0 * ec * ay + eb * ax;
            e.q0 = gm1 * (2.0 * ea * bx + eb * by); e.q1 = gm1 * (2.0 * ec * by + eb * bx);
            e.hq0 = 2.0 * ea; e.hq1 = eb; e.hq2 = 2.0 * ec;
        }
    } else if (rd.type == RT_LEG) {
        const double lx = sm.nodes[i][0] - z[2 * i], ly = sm.nodes[i][1] - z[2 * i + 1];
        e.c = lx * lx + ly * ly;
        if (GRAD) { e.q0 = 2.0 * lx; e.q1 = 2.0 * ly; }
    } else if (rd.type == RT_DTH) {
        e.c = z[6 + i];
        if (GRAD) e.t_own = 1.0;
    } else if (rd.type != RT_NONE) {   // v_bx, v_by, fen+, fen-
        const double sn = sm.trig[kn][0], cs = sm.trig[kn][1];
        const double vx = sm.nodes[kn][2], vy = sm.nodes[kn][3];
        const double vbx = cs * vx + sn * vy, vby = -sn * vx + cs * vy;
        if (rd.type == RT_VBY) {
            e.c = vby;
            if (GRAD) { e.p0 = -sn; e.p1 = cs; e.t_all = -vbx; }
        } else {
            if (GRAD) { e.p0 = cs; e.p1 = sn; e.t_all = vby; }
            if (rd.type == RT_VBX) e.c = vbx;
            else {
                const double sg = rd.type == RT_FENP ? P.s_turn : -P.s_turn;
                e.c = vbx + sg * z[6 + i];
                if (GRAD) e.t_own = sg;
            }
        }
    }
}

// slots of WarpShared::cold.  Every lane stores the same value and nobody reads before the next __syncwarp(); none of these is
// updated by read-modify-write (that would not be safe if the lanes of a warp drifted apart).
enum { C_RESTO_TARGET = 0, C_RESTO_ENTRY, C_THETA_MAX, C_THETA_MIN, C_OBJ, C_VIOL, C_ST_THETA, C_ST_LOGSUM, C_ST_V2, C_ST_VMAX };

struct WState {   // replicated scalars of one problem (registers)
    double mu, sf, alpha, alpha_z, delta_last, lm_lambda;
    double v2_h1, v2_h2;   // windowed stagnation test of the restoration (see ipm_iterate())
    int nf, iters, acc_cnt, status, phase, nstall, tiny, nresto;
    bool pending, reinit, first;
#ifdef DCBF_DBG
    int n_fact, n_fail, n_trial, n_pass;
#endif
};

// ---------------------------------------------------------------------------------------------------------------
// nodes 1..3 at the point z (lanes 0..2, directly from the free response and the constant influence coefficients),
// per-node trigonometry and objective terms; collective, out of line
// ---------------------------------------------------------------------------------------------------------------
template <int NS>
__device__ __noinline__ void w_nodes(const double *z, int lane, double sf, bool want_hess) {
    WarpShared<NS> &sm = g_sm<NS>;
    const CtaShared &cs_ = g_cs;
    const dcbf_params &P = cs_.P;
    if (lane < 3) {
        const int kn = lane + 1;
        double x = sm.fr[kn][0], y = sm.fr[kn][1], vx = sm.fr[kn][2], vy = sm.fr[kn][3], th = sm.x0[4];
        const double *cx = cs_.cab[lane], *cv = cs_.cab[3 + lane];   // gx[kn-1-l], gv[kn-1-l] for l < kn, else 0
#pragma unroll
        for (int l = 0; l < 3; l++) {
            const double fx = z[2 * l], fy = z[2 * l + 1];
            x = fma(cx[l], fx, x); y = fma(cx[l], fy, y); vx = fma(cv[l], fx, vx); vy = fma(cv[l], fy, vy);
            th += l < kn ? z[6 + l] : 0.0;
        }
        sm.nodes[kn][0] = x; sm.nodes[kn][1] = y; sm.nodes[kn][2] = vx; sm.nodes[kn][3] = vy; sm.nodes[kn][4] = th;
        double sn, cs;
        fsincos(th, &sn, &cs);   // inline: interleaves with the atan2 chain below
        const double w = P.w_q + (kn == 1 ? P.w_p : 0.0);
        const double ex = x - sm.goal[0], ey = y - sm.goal[1];
        const double dx = -ex, dy = -ey;
        const double r2 = dx * dx + dy * dy, ir2 = frcp(r2);
        const double tar = fatan2(dy, dx);
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
// the solver for one problem (all 32 lanes call it with identical arguments).  Inputs in sm: x0, graw, zc, fr.
// ---------------------------------------------------------------------------------------------------------------
template <int NS>
__device__ void solve_lip_warp(const BatchIn &in, int b, int lane, int leg, WState &S) {
    constexpr int RP = WarpShared<NS>::RP;
    WarpShared<NS> &sm = g_sm<NS>;
    const CtaShared &cs_ = g_cs;
    const dcbf_params &P = cs_.P;
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
            if (pos < KsMax<NS>::v) {
#pragma unroll
                for (int c = 0; c < 6; c++) sm.obs[pos][c] = rec[c];
            }
        }
        bool hit = false;
        double ngx = 0.0, ngy = 0.0;
        const double gx = sm.graw[0], gy = sm.graw[1];
        if (P.goal_shift && sel && is_c) {
            const double PI = 3.14159265358979323846;
            const double dg = (px - gx) * (px - gx) + (py - gy) * (py - gy);
            const double dc = (px - rec[0]) * (px - rec[0]) + (py - rec[1]) * (py - rec[1]);
            if (dc < dg && dc < 9.0 * rec[5]) {
                const double th = fatan2(gy - py, gx - px), al = fatan2(rec[1] - py, rec[0] - px);
                double d = th - al;
                if (d < 0.0 && fabs(d) > PI) d += 2.0 * PI;
                else if (d > 0.0 && fabs(d) > PI) d -= 2.0 * PI;
                if (fabs(d) < PI / 12.0) {
                    const double na = d < 0.0 ? th - PI / 12.0 : th + PI / 12.0;
                    const double rad = sqrt(dg);
                    double sn, cs;
                    fsincos(na, &sn, &cs);
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
        if (lane < 5) sm.nodes[0][lane] = sm.x0[lane];
        __syncwarp();
    }
    const bool has_fen = P.has_fen != 0;
    const int ms = Ks + (has_fen ? 6 : 4);   // rows per step
    const int m = 3 * ms;
    const int mp = (m + 1) & ~1;             // dot products run over row pairs
    RowDesc rd[NS];
    RowBnd rb[NS];
    double rs[NS], rzl[NS], rzu[NS], rds[NS], rel[NS], reu[NS];
    int nz_l = 0;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        rd[s] = row_desc(s * 32 + lane, Ks, ms, m);
        rb[s] = row_bounds(P, rd[s], leg);
        nz_l += (rb[s].has_lo ? 1 : 0) + (rb[s].has_hi ? 1 : 0);
        rs[s] = rzl[s] = rzu[s] = rds[s] = rel[s] = reu[s] = 0.0;
    }
    const int nz = wsumi(nz_l), nrows = m;
    // the lane's dot products of the three assembly rounds (packed: operand rows, first contributing step, output slot)
    int dsc[3];
#pragma unroll
    for (int t = 0; t < 3; t++) dsc[t] = __ldg(&cs_.tab->desc[32 * t + lane]);
    const int l8 = lane < 9 ? lane : 8;
    const int rowbase = lane < 10 ? lane * (lane + 1) / 2 : 0;
    // ---- solver state ------------------------------------------------------------------------------------------------
    S.mu = P.mu_init; S.sf = 1.0; S.alpha = 0.0; S.alpha_z = 0.0; S.delta_last = 0.0; S.lm_lambda = 1e-4; 
    sm.cold[C_RESTO_TARGET] = 0.0; sm.cold[C_RESTO_ENTRY] = 0.0; sm.cold[C_THETA_MAX] = 1e300; sm.cold[C_THETA_MIN] = 0.0; S.nf = 0; S.iters = 0; S.acc_cnt = 0; S.status = -1; S.nstall = 0; S.tiny = 0;
    S.nresto = 0; S.v2_h1 = 0.0; S.v2_h2 = 0.0;
    S.phase = PH_MAIN; S.pending = false; S.reinit = true; S.first = true; sm.cold[C_OBJ] = 0.0; sm.cold[C_VIOL] = 0.0;
    const double tol = P.tol;
    const double *stf = &sm.ST[0][0];
#ifdef DCBF_DBG
    S.n_fact = S.n_fail = S.n_trial = S.n_pass = 0;
#endif

    bool nodes_valid = false;      // sm.nodes / trig / nobj (with Hessian terms) already describe zc (staged by the accepted trial)
    double carry_log = 0.0;        // sum of log(gaps) at the accepted trial point = barrier term of the next full pass
    bool carry_ok = false;
    for (;;) {
        const bool resto = S.phase == PH_RESTO;
#ifdef DCBF_DBG
        S.n_pass++;
#endif
        if (!nodes_valid) w_nodes<NS>(sm.zc, lane, S.first ? 1.0 : (resto ? 0.0 : S.sf), true);
        nodes_valid = false;
        // objective value and gradient (lane a < 9 owns grad[a]):  d node_kn / d foot_l = gx[kn-1-l],  d th_kn / d turn_l = 1
        const double fobj = sm.nobj[1][0] + sm.nobj[2][0] + sm.nobj[3][0];
        double grad_a = 0.0;
        {   // branch-free (see the note at the factorisation): lanes >= 9 compute on clamped indices and select 0
            const int l = l8 < 6 ? l8 >> 1 : l8 - 6;
            const int c = l8 < 6 ? 1 + (l8 & 1) : 3;
#pragma unroll
            for (int kn = 1; kn <= 3; kn++) {
                const double wf = cs_.cab[kn - 1][l];
                const double w = l8 < 6 ? wf : (l < kn ? 1.0 : 0.0);
                grad_a = fma(sm.nobj[kn][c], w, grad_a);
            }
            grad_a = lane < 9 ? grad_a : 0.0;
        }
        if (S.first) {
            const double gmax = wmax(fabs(grad_a));
            S.sf = gmax > 100.0 ? fdiv(100.0, gmax) : 1.0;
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
                            const double mg = fdiv(S.mu, gap);
                            rzl[s] = fmax(fmin(rzl[s] + S.alpha_z * rel[s], DCBF_KAPPA_SIGMA * mg), mg * (1.0 / DCBF_KAPPA_SIGMA));
                        }
                        if (bb.has_hi) {
                            const double gap = hr - rs[s];
                            const double mg = fdiv(S.mu, gap);
                            rzu[s] = fmax(fmin(rzu[s] + S.alpha_z * reu[s], DCBF_KAPPA_SIGMA * mg), mg * (1.0 / DCBF_KAPPA_SIGMA));
                        }
                    }
                    const double rc = e.c - rs[s];
                    double lp = 1.0;
                    if (bb.has_lo) {
                        const double gap = rs[s] - lr, inv = frcp(gap);
                        sig += rzl[s] * inv; binv += inv; y -= rzl[s];
                        const double cz = gap * rzl[s];
                        t_cmin = fmin(t_cmin, cz); t_cmax = fmax(t_cmax, cz); t_z += rzl[s];
                        lp *= gap; rel[s] = inv;
                    }
                    if (bb.has_hi) {
                        const double gap = hr - rs[s], inv = frcp(gap);
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
            // stage the transposed row (rows beyond m stage zeros so that the dot products need no guards)
            {
                const double *ab = cs_.cab[rd[s].cls];
                const int i = rd[s].step;
                double *col = &sm.ST[0][r];
#pragma unroll
                for (int l = 0; l < 3; l++) {
                    const double ca = ab[l], cb = ab[3 + l];
                    const double gxv = fma(ca, e.p0, cb * e.q0), gyv = fma(ca, e.p1, cb * e.q1);
                    const double gtv = (l <= i ? e.t_all : 0.0) + (l == i ? e.t_own : 0.0);
                    col[(2 * l) * RP] = gxv; col[(2 * l + 1) * RP] = gyv; col[(6 + l) * RP] = gtv;
                    col[(9 + 2 * l) * RP] = sig * gxv; col[(9 + 2 * l + 1) * RP] = sig * gyv; col[(9 + 6 + l) * RP] = sig * gtv;
                }
                col[18 * RP] = sig; col[19 * RP] = w1; col[20 * RP] = binv; col[21 * RP] = y;
            }
            sm.HQ[r][0] = y * e.hq0; sm.HQ[r][1] = y * e.hq1; sm.HQ[r][2] = y * e.hq2;
            st8.s0 += t_rc; st8.s1 += t_z; st8.s2 += t_log; st8.s3 += t_v2;
            st8.m0 = fmax(st8.m0, t_rc); st8.m1 = fmax(st8.m1, -t_cmin); st8.m2 = fmax(st8.m2, t_cmax); st8.m3 = fmax(st8.m3, t_v);
        }
        __syncwarp();   // the row branches reconverge here, before the shuffles
        reduce8_inline(st8);
        const double st_theta = st8.s0, st_zsum = st8.s1, st_logsum = carry_ok ? carry_log : st8.s2, st_v2 = st8.s3, st_pinf = st8.m0,
                     st_cmin = -st8.m1, st_cmax = st8.m2, st_vmax = st8.m3;
        carry_ok = false;
        sm.cold[C_ST_THETA] = st_theta; sm.cold[C_ST_LOGSUM] = st_logsum; sm.cold[C_ST_V2] = st_v2; sm.cold[C_ST_VMAX] = st_vmax;
        // ---- node Hessians -> the 27 sources of the Hessian table ---------------------------------------------------------------
        {
            const double gm1 = P.gamma - 1.0;
            const double *yv = sm.ST[21];
            const int lp = lane - 16;
            if ((unsigned)lp < 9u) {   // position block on lanes 16..24: node kn = lp/3 + 1, component lp % 3
                const int kn = lp / 3 + 1, c = lp % 3;
                double acc = sm.nobj[kn][4 + c];
                for (int j = 0; j < Ks; j++) acc += sm.HQ[(kn - 1) * ms + j][c];
                if (kn < 3) for (int j = 0; j < Ks; j++) acc = fma(gm1, sm.HQ[kn * ms + j][c], acc);
                sm.NHf[8 * (kn - 1) + c] = acc;
            } else if (lane < 3) {     // heading / velocity entries of node kn = lane + 1 (rows of step kn - 1)
                const int kn = lane + 1, i = kn - 1, base = i * ms + Ks;
                double Yx = yv[base], Yy = yv[base + 1];
                if (has_fen) Yx += yv[base + 4] + yv[base + 5];
                const double sn = sm.trig[kn][0], cs = sm.trig[kn][1];
                const double vx = sm.nodes[kn][2], vy = sm.nodes[kn][3];
                const double vbx = cs * vx + sn * vy, vby = -sn * vx + cs * vy;
                double *H = &sm.NHf[8 * i];
                H[3] = sm.nobj[kn][7]; H[4] = sm.nobj[kn][8];
                H[5] = sm.nobj[kn][9] - (Yx * vbx + Yy * vby);
                H[6] = -sn * Yx - cs * Yy; H[7] = cs * Yx - sn * Yy;
                sm.NHf[24 + i] = 2.0 * yv[base + 2];
            }
        }
        __syncwarp();
        // ---- condensed matrix and J^T vectors: three rounds of dot products over row pairs ---------------------------------------
#pragma unroll
        for (int t = 0; t < 3; t++) {
            const int d = dsc[t];
            if (d >= 0) {
                const double *pp = stf + (d & 0xff) * RP, *pq = stf + ((d >> 8) & 0xff) * RP;
                const int lo = (((d >> 16) & 0xf) * ms) & ~1, eo = d >> 20;
                double acc0 = 0.0, acc1 = 0.0;
                for (int r = mp - 2; r >= lo; r -= 2) {
                    const double2 u = *reinterpret_cast<const double2 *>(pp + r), v = *reinterpret_cast<const double2 *>(pq + r);
                    acc0 = fma(u.x, v.x, acc0); acc1 = fma(u.y, v.y, acc1);
                }
                double acc = acc0 + acc1;
                if (eo >= KQ_K) {   // Lagrangian Hessian through the table
                    const int e = eo - KQ_K;
#pragma unroll
                    for (int h = 0; h < NHT; h++) acc = fma(cs_.hc[h][e], sm.NHf[cs_.hs[h][e]], acc);
                }
                sm.KQ[eo] = acc;
            }
        }
        __syncwarp();
        S.pending = false; S.reinit = false;
        sm.cold[C_OBJ] = fobj; sm.cold[C_VIOL] = st_vmax;
        if (!(fobj == fobj) || !(st_theta == st_theta)) { S.status = -13; break; }
        // ---- convergence / barrier update / right-hand side ---------------------------------------------------------------
        const double *q = &sm.KQ[KQ_Q];
        double rhs_a = 0.0;
        if (!resto) {
            if (S.first) { sm.cold[C_THETA_MAX] = 1e4 * fmax(1.0, st_theta); sm.cold[C_THETA_MIN] = 1e-4 * fmax(1.0, st_theta); S.first = false; }
            const double dinf = wmax(lane < 9 ? fabs(fma(S.sf, grad_a, q[18 + l8])) : 0.0);
            const double sd = fmax(100.0, fdiv(2.0 * st_zsum, (double)(nrows + nz))) * 0.01;
            const double sc = fmax(100.0, fdiv(st_zsum, (double)(nz > 0 ? nz : 1))) * 0.01;
            const double isd = frcp(sd), isc = frcp(sc);
            double E0;
            for (;;) {
                const double compm = fmax(fabs(st_cmax - S.mu), fabs(st_cmin - S.mu));
                E0 = fmax(fmax(dinf * isd, st_pinf), st_cmax * isc);
                const double Emu = fmax(fmax(dinf * isd, st_pinf), compm * isc);
                if (E0 <= tol) break;
                if (Emu <= DCBF_KAPPA_EPS * S.mu && S.mu > tol * 0.1 * (1.0 + 1e-12)) {
                    S.mu = fmax(tol * 0.1, fmin(DCBF_KAPPA_MU * S.mu, DCBF_MU_POW(S.mu)));
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
                S.phase = PH_RESTO; sm.cold[C_RESTO_ENTRY] = st_vmax; sm.cold[C_RESTO_TARGET] = fmax(0.1 * st_vmax, 1e-9); S.lm_lambda = 1e-4; S.acc_cnt = 0; S.nresto = 0;
                S.iters++;
                continue;
            }
            rhs_a = -S.sf * grad_a - q[l8] + S.mu * q[9 + l8];
        } else {
            if (st_vmax <= sm.cold[C_RESTO_TARGET]) { S.phase = PH_MAIN; S.reinit = true; continue; }
            const double gn = wmax(lane < 9 ? fabs(q[l8]) : 0.0);
            const bool stationary = gn <= 1e-10 * fmax(1.0, st_vmax) || S.lm_lambda > 1e12;
            if (stationary) {
                if (st_vmax > P.constr_viol_tol) { S.status = 2; break; }
                if (sm.cold[C_RESTO_ENTRY] <= 1e-9 || S.nstall++ >= 1) { S.status = -2; break; }   // see ipm_iterate()
                S.phase = PH_MAIN; S.reinit = true; continue;
            }
            if (S.iters >= P.max_iter) { S.status = -1; break; }
            rhs_a = -q[l8];
        }
        sm.KQ[KQ_RHS + lane] = rhs_a;   // = packed row 9 of the system: tri(9, j) = 45 + j (lanes >= 9 write the tail)
        __syncwarp();
        // ---- restoration: Levenberg-Marquardt trials reuse the assembled K while lambda is escalated ---------------------------
        // ---- main phase: one factorisation with inertia correction by delta ----------------------------------------------------
        bool lm_accept = false;
        double v2t = 0.0, vmt = 0.0;
        for (int rt = 0; rt < 20; rt++) {
            double shift = resto ? S.lm_lambda : 0.0;
            bool ok = false;
            for (int tr = 0; tr < 48; tr++) {
                // lanes 0..8 own the rows of K, lane 9 the right-hand side (its "row" of the factor is L^-1 rhs).  Branch-free:
                // every lane runs the column arithmetic (idle lanes on harmless operands), only the stores are predicated,
                // so the warp reaches each shuffle converged.
                double Lrow[9];
                ok = true;
#ifdef DCBF_DBG
                S.n_fact++;
#endif
#pragma unroll
                for (int j = 0; j < 9; j++) {
                    const bool act = lane >= j && lane < 10;
                    double s_ = sm.KQ[KQ_K + rowbase + j] + (lane == j ? shift : 0.0);
#pragma unroll
                    for (int c = 0; c < j; c++) s_ = fma(-Lrow[c], sm.Lf[tri(j, c)], s_);
                    const double d = __shfl_sync(FULL, s_, j);
                    if (!(d > 1e-14)) {
                        ok = false;
#ifdef DCBF_DBG
                        S.n_fail++;
#endif
                        break;
                    }
                    const double rinv = drsqrt(d);
                    Lrow[j] = lane == j ? rinv : s_ * rinv;   // diagonal stored as its reciprocal
                    sm.Lf[act ? rowbase + j : 56 + lane] = Lrow[j];
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
            // backward substitution: lane i holds component i, starting from y = L^-1 rhs (row 9 of the factor); branch-free
            {
                const int li = l8;
                double bi = sm.Lf[45 + li];
#pragma unroll
                for (int c = 8; c >= 0; c--) {
                    const double xc = __shfl_sync(FULL, bi * sm.Lf[tri(c, c)], c);
                    const double lc = sm.Lf[tri(c, 0) + (li < c ? li : 0)];
                    bi = lane == c ? xc : (lane < c ? fma(-lc, xc, bi) : bi);
                }
                sm.dz[lane] = bi;
            }
            __syncwarp();
            if (!resto) break;
            // Levenberg-Marquardt trial at full step (violation only)
            sm.zt[lane] = sm.zc[l8] + sm.dz[l8];
            __syncwarp();
            w_nodes<NS>(sm.zt, lane, 0.0, false);
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
            __syncwarp();
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                v2t += __shfl_xor_sync(FULL, v2t, o);
                vmt = fmax(vmt, __shfl_xor_sync(FULL, vmt, o));
            }
            if (v2t < sm.cold[C_ST_V2] * (1.0 - 1e-12)) { lm_accept = true; break; }
            S.lm_lambda *= 10.0;
            if (S.lm_lambda > 1e12) break;
        }
        if (S.status == -3) break;
        if (resto) {
            if (lm_accept) {
                const double dn = wmax(fabs(sm.dz[l8]));
                sm.zc[lane] = sm.zt[lane];
                S.iters++;
                S.lm_lambda = fmax(S.lm_lambda * 0.2, 1e-12);
                const double v2c = sm.cold[C_ST_V2];
                if (v2c - v2t <= 1e-4 * v2c) S.acc_cnt++; else S.acc_cnt = 0;
                const bool crawl = S.nresto >= 2 && S.v2_h2 - v2t <= DCBF_RESTO_WINDOW * S.v2_h2;
                S.v2_h2 = S.v2_h1; S.v2_h1 = v2c; S.nresto++;
                if ((dn < 1e-12 || S.acc_cnt >= 2 || crawl) && vmt > sm.cold[C_RESTO_TARGET]) S.lm_lambda = 1e13;
            }
            __syncwarp();
            continue;
        }
        // ---- direction pass (main phase): ds, dz_L, dz_U, step sizes -----------------------------------------------------------
        double amax = 1.0, az = 1.0, dphi = 0.0;
        {
            const double tau = fmax(0.99, 1.0 - S.mu);
            dphi = S.sf * grad_a * sm.dz[l8];
#pragma unroll
            for (int s = 0; s < NS; s++) {
                if (rd[s].type == RT_NONE) continue;
                const int r = s * 32 + lane;
                double jd = 0.0;
#pragma unroll
                for (int a = 0; a < 9; a++) jd = fma(sm.ST[a][r], sm.dz[a], jd);
                const RowBnd &e = rb[s];
                const double d = jd + rds[s];
                rds[s] = d;
                if (e.has_lo) {
                    const double inv = rel[s], gap = rs[s] - relax_lo(e.lo);
                    const double dzl = S.mu * inv - rzl[s] - rzl[s] * inv * d;
                    rel[s] = dzl;
                    dphi -= S.mu * d * inv;
                    if (d < 0.0) amax = fmin(amax, fdiv(-tau * gap, d));
                    if (dzl < 0.0) az = fmin(az, fdiv(-tau * rzl[s], dzl));
                }
                if (e.has_hi) {
                    const double inv = reu[s], gap = relax_hi(e.hi) - rs[s];
                    const double dzu = S.mu * inv - rzu[s] + rzu[s] * inv * d;
                    reu[s] = dzu;
                    dphi += S.mu * d * inv;
                    if (d > 0.0) amax = fmin(amax, fdiv(tau * gap, d));
                    if (dzu < 0.0) az = fmin(az, fdiv(-tau * rzu[s], dzu));
                }
            }
            __syncwarp();
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {   // three reductions interleaved
                amax = fmin(amax, __shfl_xor_sync(FULL, amax, o));
                az = fmin(az, __shfl_xor_sync(FULL, az, o));
                dphi += __shfl_xor_sync(FULL, dphi, o);
            }
        }
        // ---- filter line search -------------------------------------------------------------------------------------------------
        const double theta = sm.cold[C_ST_THETA];
        const double phi = S.sf * sm.cold[C_OBJ] - S.mu * sm.cold[C_ST_LOGSUM];
        const double eps_phi = 10.0 * 2.2e-16 * fabs(phi);
        double alpha = amax;
        int accepted = 0;
        for (int ls = 0; ls < DCBF_LS_MAX; ls++, alpha *= 0.5) {
            sm.zt[lane] = fma(alpha, sm.dz[l8], sm.zc[l8]);
            __syncwarp();
#ifdef DCBF_DBG
            S.n_trial++;
#endif
            w_nodes<NS>(sm.zt, lane, S.sf, true);
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
            __syncwarp();
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                th_t += __shfl_xor_sync(FULL, th_t, o);
                lg_t += __shfl_xor_sync(FULL, lg_t, o);
            }
            okv = __all_sync(FULL, okv);
            const double ph_t = S.sf * ft - S.mu * lg_t;
            if (!okv || !(ph_t == ph_t) || !(th_t <= sm.cold[C_THETA_MAX])) continue;
            bool in_filter = false;
            for (int qf = 0; qf < S.nf; qf++)
                if (th_t >= sm.filt_th[qf] && ph_t >= sm.filt_ph[qf]) in_filter = true;
            if (in_filter) continue;
            const bool sw = dphi < 0.0 && theta <= sm.cold[C_THETA_MIN] && switch_cond(alpha, -dphi, theta);
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
            const double vm = sm.cold[C_ST_VMAX];
            S.phase = PH_RESTO; sm.cold[C_RESTO_ENTRY] = vm; sm.cold[C_RESTO_TARGET] = fmax(0.1 * vm, 1e-9); S.lm_lambda = 1e-4; S.acc_cnt = 0; S.nresto = 0;
            S.iters++;
            continue;
        }
        sm.zc[lane] = sm.zt[lane];
        __syncwarp();
        S.alpha = alpha; S.alpha_z = az; S.pending = true;
        nodes_valid = true; carry_ok = true;   // the accepted trial staged the nodes (with Hessian terms) and the barrier sum of the new point
        if (alpha < 1e-2 && sm.cold[C_ST_VMAX] > P.constr_viol_tol) S.tiny++; else S.tiny = 0;
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

// per-CTA staging of the constants; zeroes the pad columns of the staged rows
template <int NS>
__device__ __forceinline__ void stage_cta(const dcbf_params &P, const Consts &K, const WarpTables *tab, int lane) {
    WarpShared<NS> &sm = g_sm<NS>;
    if (lane == 0) { g_cs.P = P; g_cs.K = K; g_cs.tab = tab; sm.NHf[NSRC] = 0.0; }
    for (int t = lane; t < 60; t += 32) (&g_cs.cab[0][0])[t] = __ldg(&tab->cab[0][0] + t);
    for (int t = lane; t < NHT * 48; t += 32) (&g_cs.hc[0][0])[t] = __ldg(&tab->hc[0][0] + t);
    for (int t = lane; t < NHT * 48 / 4; t += 32)
        reinterpret_cast<unsigned *>(&g_cs.hs[0][0])[t] = __ldg(reinterpret_cast<const unsigned *>(&tab->hs[0][0]) + t);
    for (int t = lane; t < 44; t += 32) sm.ST[t >> 1][32 * NS + (t & 1)] = 0.0;
    __syncwarp();
}

#endif  // __CUDACC__

}  // namespace wp
}  // namespace dcbf
