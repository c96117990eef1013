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
//     in shared memory at a fixed row stride (see LF_DIAG); all scalar control flow is replicated and therefore uniform.
// The interior-point driver (solve_warp) is written once against a small model interface -- rows, nodes, gradient staging,
// Hessian sources -- and instantiated for the LIP formulations (LipW generic row slots, LipL with the three turn rows in a typed
// linear slot; 9 variables) and the differential-drive formulation (DdW / DdL, 6 variables; node Jacobians are recomputed per
// iterate because the unicycle rollout is not affine).
// The one-slot LIP kernel and DdL are built for 128 registers, i.e. 16 warps per SM (LaneRefresh: per-lane invariants are
// recomputed per iteration instead of being carried in registers; cold solver state lives in WarpShared::cold).
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
// Slot N of the right-hand side is a dump: all 32 lanes store "their" component without a lane test (index min(lane, N)).
constexpr int KQ_Q = 0;
constexpr int KQ_K = 28;
constexpr int KQ_RHS = KQ_K + 45;
constexpr int KQ_LEN = KQ_RHS + 12;   // right-hand side (N entries) + one dump slot (index N) for the lanes beyond N
// Cholesky factor of the (N + 1) x (N + 1) bordered system, one row per lane at a fixed stride of N entries: rows 0..N-1 of L, row N
// = L^-1 rhs, row N + 1 = dump row of the lanes beyond N; only the strictly lower triangle is ever written (diagonal and upper
// triangle stay zero from stage_cta on, which is what lets the backward substitution run without lane tests), the reciprocal
// diagonal follows at LF_DIAG
constexpr int LF_STRIDE_MAX = 9;
constexpr int LF_DIAG = (LF_STRIDE_MAX + 2) * LF_STRIDE_MAX;   // 99
constexpr int LF_LEN = LF_DIAG + LF_STRIDE_MAX + 2;            // 110

// Host-built constant tables, one copy per context in device memory (build_warp_tables()).
struct WarpTables {
    int desc[96];             // 3 rounds x 32 lanes: rowP | rowQ << 8 | class << 16 | out << 20   (-1: idle)
    double hc[NHT][48];       // K[e] += hc[t][e] * src[hs[t][e]]
    unsigned char hs[NHT][48];
    double cab[10][6];        // (cA[3], cB[3]) per row class: CBF step 0..2, velocity rows step 0..2, leg step 0..2, turn
    int desc_dd[64];          // DD: 2 rounds x 32 lanes, same packing (21 matrix entries + 18 vector entries)
    double sm_dd[24];         // DD: constant Hessian pattern of the control-smoothness cost (per unit 2 w_t)
    int desc_lin[64];         // DD, linear rows in their own slot (DdL): closed-form contribution of the four linear rows of a step to the
                              // lane's entry: flags (bits 0..4, see DdWT::lin_term) | source row << 8 | first column << 16
    int desc_lin_lip[96];     // LIP, turn rows in their own slot (LipL): the entry gets staged weight [source row][column] (bit 0 set) or nothing
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
    // LipL: the turn row of step i (column 32 + i of the staged weights) has the gradient e_{6+i}: sigma goes to the diagonal entry of
    // variable 6 + i, the weight rows w1 / binv / y (staged rows 19 + v) to component 6 + i of the three J^T vectors
    for (int t = 0; t < 96; t++) {
        int d = 0;
        if (t < 72) {
            if (ent[t].rowP < 18) {
                const int a = ent[t].rowP - 9, b = ent[t].rowQ;
                if (a == b && a >= 6) d = 1 | 18 << 8 | (32 + a - 6) << 16;
            } else if (ent[t].rowQ >= 6) d = 1 | ent[t].rowP << 8 | (32 + ent[t].rowQ - 6) << 16;
        }
        W.desc_lin_lip[t] = d;
    }
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
    // the kernels unroll the table per assembly round: round 0 up to 6 terms, round 1 up to 2, round 2 no matrix entries (LipW::round_terms)
    for (int t = 0; t < 72; t++) {
        const int out = W.desc[t] >> 20, rnd = t / 32;
        if (out < KQ_K) continue;
        const int lim = rnd == 0 ? 6 : (rnd == 1 ? 2 : 0);
        for (int h = lim; h < NHT; h++) if (W.hc[h][out - KQ_K] != 0.0) return false;
    }
    // ---- gradient coefficient vectors ------------------------------------------------------------------------------------
    for (int c = 0; c < 10; c++) for (int j = 0; j < 6; j++) W.cab[c][j] = 0.0;
    for (int i = 0; i < 3; i++) for (int l = 0; l < 3; l++) {
        if (l <= i) { W.cab[i][l] = k.gx[i - l]; W.cab[3 + i][l] = k.gv[i - l]; }       // cA of D-CBF / velocity rows
        if (l < i) { W.cab[i][3 + l] = k.gx[i - 1 - l]; W.cab[6 + i][3 + l] = k.gx[i - 1 - l]; }   // cB of D-CBF / leg rows
        if (l == i) W.cab[6 + i][3 + l] = -1.0;
    }
    // ---- DD: variables (v0, w0, v1, w1, v2, w2), step of variable a = a / 2; staged rows 0..5 gradient, 6..11 sigma * gradient,
    //      12..15 weights ------------------------------------------------------------------------------------------------------
    {
        struct Ent { int rowP, rowQ, cls, out; };
        Ent ent[39];
        int n = 0;
        for (int cls = 0; cls < 3; cls++) {
            for (int a = 0; a < 6; a++) for (int b = 0; b <= a; b++)
                if ((a >> 1) == cls) ent[n++] = {6 + a, b, cls, KQ_K + tri(a, b)};      // max(step a, step b) = step a for b <= a
            for (int v = 0; v < 3; v++) for (int a = 0; a < 6; a++)
                if ((a >> 1) == cls) ent[n++] = {12 + 1 + v, a, cls, KQ_Q + 6 * v + a};
        }
        if (n != 39) return false;
        for (int t = 0; t < 64; t++) W.desc_dd[t] = t < 39 ? (ent[t].rowP | ent[t].rowQ << 8 | ent[t].cls << 16 | ent[t].out << 20) : -1;
        // linear rows of step i (fen+: v + s w, fen-: v - s w, bound v, bound w) sit in columns 32 + 4 i .. 32 + 4 i + 3 of the staged
        // weights; with x_j the weight of row j:  (v, v): x1 + x2 + x3,  (w, v): s (x1 - x2),  (w, w): s^2 (x1 + x2) + x4  -- matrix entries
        // take sigma (staged row 2 N), the vector entries of variable v_i / w_i take  x1 + x2 + x3  /  s (x1 - x2) + x4  of their weight row
        for (int t = 0; t < 64; t++) {
            int flags = 0, src = 0, col = 32;
            if (t < 39) {
                if (ent[t].rowP < 12) {       // matrix entry (a, b)
                    const int a = ent[t].rowP - 6, b = ent[t].rowQ;
                    if ((a >> 1) == (b >> 1)) {
                        flags = (a & 1) == 0 ? (1 | 4) : ((b & 1) == 0 ? 2 : (16 | 8));
                        src = 12; col = 32 + 4 * (a >> 1);
                    }
                } else {                      // vector entry: weight row 13 + v, variable a
                    const int a = ent[t].rowQ;
                    flags = (a & 1) == 0 ? (1 | 4) : (2 | 8);
                    src = ent[t].rowP; col = 32 + 4 * (a >> 1);
                }
            }
            W.desc_lin[t] = flags | src << 8 | col << 16;
        }
        // t sum_i |u_i - u_{i-1}|^2  (MPC_DD_sig_step.py:351-369): Hessian 2 t * [[2,-1,0],[-1,2,-1],[0,-1,1]] on v and on w
        for (int e = 0; e < 24; e++) W.sm_dd[e] = 0.0;
        static const double D3[3][3] = {{2, -1, 0}, {-1, 2, -1}, {0, -1, 1}};
        for (int a = 0; a < 6; a++) for (int b = 0; b <= a; b++)
            if ((a & 1) == (b & 1)) W.sm_dd[tri(a, b)] = D3[a >> 1][b >> 1];
    }
    return true;
}

#if defined(__CUDACC__)

// Bounds-checked build (-DDCBF_BOUNDS=1, tools/sanitize_probe.py): every data-dependent index into the per-problem scratch is
// asserted on the device (a violated assert traps the kernel and the entry point returns a CUDA error).  compute-sanitizer is
// closed on the GPU pool (profiles/r03_sanitizer_refused.log), so this build is the memory check of the warp kernels.
#if DCBF_BOUNDS
#include <assert.h>
#define DCBF_ASSERT(c) assert(c)
#else
#define DCBF_ASSERT(c) ((void)0)
#endif

template <int NS> struct KsMax { static constexpr int v = NS == 1 ? 6 : (NS == 2 ? 17 : 2 * DCBF_MAX_OBS); };

// slots of WarpShared::cold.  Every lane stores the same value and nobody reads before the next __syncwarp(); none of these is
// updated by read-modify-write (that would not be safe if the lanes of a warp drifted apart).
enum { C_RESTO_TARGET = 0, C_RESTO_ENTRY, C_THETA_MAX, C_THETA_MIN, C_OBJ, C_VIOL, C_ST_THETA, C_ST_LOGSUM, C_ST_V2, C_ST_VMAX,
       // solver state that is touched once per iteration or less (was WState, i.e. registers carried across the whole solve).  alpha,
       // alpha_z, carry_log, delta_last: every lane overwrites them with the same value and reads its own write.  lm_lambda, v2_h1,
       // v2_h2 depend on their previous value: they are read into locals, the warp synchronises, then they are written.
       C_ALPHA, C_ALPHA_Z, C_CARRY_LOG, C_DELTA_LAST, C_LM_LAMBDA, C_V2_H1, C_V2_H2, C_NROWS_NZ, C_NZ, C_COUNT };

// model-specific node data -----------------------------------------------------------------------------------------------------
struct LipNodeData {
    double fr[4][4];         // free response (zero foot placements) x, y, vx, vy at nodes 0..3
    double nodes[4][5];      // x, y, vx, vy, th of nodes 0..3
    double trig[4][3];       // sin, cos, atan2 target of nodes 1..3
    double nobj[4][10];      // f_k, nx, ny, nt, hxx, hxy, hyy, hxt, hyt, htt of the objective at node k
    double NHf[NSRC + 1];    // Hessian sources (+ one zero for the padding terms of the table)
};
struct DdNodeData {
    double nodes[4][3];      // x, y, th of nodes 0..3
    double trig[3][2];       // sin, cos of th_0..2
    double Jx[4][6], Jy[4][6];   // d pos_k / dz (zero beyond the variables of steps < k)
    double nobj[4][10];      // as above (node k = 1..3); slot 0: the smoothness part of the objective
    double gsm[6];           // gradient of the smoothness cost
    double Q[4][3], Cc[4][2];    // sum of y * Q_r and y * grad h over the rows that touch node k
    double cvw[3], cww[3];   // curvature coefficients of the positions (see dd_add_second() in dcbf_core.cuh)
    double last_u[2];
    double sm_dd[24];        // constant Hessian pattern of the control-smoothness cost (copy of WarpTables::sm_dd)
};

template <class M, int NS>
struct alignas(16) WarpShared {
    static constexpr int N = M::N;
    // padded row length (a linear slot only stages its first NLINP columns)
    static constexpr int RP = (M::LIN2 ? 32 + M::NLINP : 32 * NS) + 2;
    static constexpr int NST = 2 * N + 4;
    double ST[NST][RP];      // staged rows, transposed: [0, N) gradient, [N, 2N) sigma * gradient, then sigma, w1, binv, y
    double HQ[M::LIN2 ? 32 : 32 * NS][M::NHQ];   // y * (second-order / first-order row coefficients) of the D-CBF rows
    double obs[KsMax<NS>::v][6];  // selected obstacles: cx, cy, a', b', c', rhs
    double KQ[KQ_LEN];       // assembled system (see KQ_*)
    double Lf[LF_LEN];       // Cholesky factor (see LF_DIAG)
    double dz[16];           // N meaningful entries; lanes >= 15 store into slot 15
    double zc[10], zt[10];   // N meaningful entries; lanes >= N store lane N-1's value into slot N-1 (index ln)
    double x0[5], goal[2], graw[2];
    // row state (slack, multipliers, step) of the kernels with more than one row per lane: the slot loop stays rolled there
    // (half the code, a third fewer registers); the single-slot kernel keeps it in registers
    static constexpr bool ROLLED = NS > 2 || (NS == 2 && M::ROLL2);
    double RS[ROLLED ? 6 : 1][ROLLED ? 32 * NS : 1];
    typename M::NodeData nd;
    double cold[C_COUNT];    // replicated scalars that are written once per event and read much later (see C_*): kept out of registers
    double filt_th[DCBF_FILT], filt_ph[DCBF_FILT];
};

// per-CTA constants staged in shared memory
struct CtaShared {
    dcbf_params P;
    Consts K;
    const WarpTables *tab;
};
// constant tables of the LIP models (a shared-memory object of their kernels only: the differential-drive kernels, which do not
// reference it, are 3 KB lighter per CTA)
struct LipTabShared {
    double cab[10][6];
    double hc[NHT][48];          // copy of the Hessian table (global loads in the assembly loop cost a long-scoreboard stall each)
    unsigned char hs[NHT][48];
};

// Warps per CTA.  The kernel body is ~130 KB, several times the SM's instruction cache: with independent warps every warp streams
// the whole iteration through the GPC-level instruction cache and its request rate saturates (ncu: gcc instruction requests 74 % of
// peak, sm__icc hit rate 71 %, 1.05 instructions per clock per SM whatever the occupancy).  The warps of a CTA therefore meet at a
// barrier at the top of every interior-point iteration (cta_tick): they walk through the iteration code together and share the
// fetched lines.  Static shared memory bounds the count (48 KB per CTA).
#ifndef DCBF_WPC_LIP
#define DCBF_WPC_LIP(NS) 1
#endif
#ifndef DCBF_WPC_DD
#define DCBF_WPC_DD(NS) ((NS) == 2 ? 2 : 1)
#endif
template <bool LIN> struct LipWT;
using LipW = LipWT<false>;   // LIP formulations, generic row slots
using LipL = LipWT<true>;    // LIP formulations, at most 32 non-linear rows: they fill slot 0, the three turn rows (linear) get slot 1
template <bool LIN> struct DdWT;
using DdW = DdWT<false>;   // differential drive, generic row slots
using DdL = DdWT<true>;    // differential drive, at most 10 obstacles: D-CBF rows in slot 0, the twelve linear rows in slot 1
template <class M, int NS> struct Wpc { static constexpr int v = DCBF_WPC_LIP(NS); };
template <int NS> struct Wpc<DdW, NS> { static constexpr int v = DCBF_WPC_DD(NS); };
#ifndef DCBF_WPC_DDL
#define DCBF_WPC_DDL 1
#endif
template <int NS> struct Wpc<DdL, NS> { static constexpr int v = DCBF_WPC_DDL; };

// kernels built for 128 registers (see solve_warp): the one-slot LIP kernel
#ifndef DCBF_LANE_REFRESH
#define DCBF_LANE_REFRESH 1
#endif
#ifndef DCBF_REFRESH_BOUNDS
#define DCBF_REFRESH_BOUNDS(NS) ((NS) > 1)
#endif
// Cholesky with the next pivot formed ahead of the shared-memory round trip of the column (solve_warp): the one-slot kernels, whose
// iteration is a chain of dependent latencies (4096-scenario step -1.0 %, identical bits; the multi-slot kernels lose 1-2 % with it).
#ifndef DCBF_PIVOT_AHEAD
#define DCBF_PIVOT_AHEAD(NS) ((NS) == 1)
#endif
template <class M, int NS> struct PivotAhead { static constexpr bool v = DCBF_PIVOT_AHEAD(NS); };
template <class M, int NS> struct LaneRefresh { static constexpr bool v = false; };
template <> struct LaneRefresh<LipW, 1> { static constexpr bool v = DCBF_LANE_REFRESH != 0; };
#ifndef DCBF_LANE_REFRESH_LIPL
#define DCBF_LANE_REFRESH_LIPL 0
#endif
template <> struct LaneRefresh<LipL, 2> { static constexpr bool v = DCBF_LANE_REFRESH_LIPL != 0; };
#ifndef DCBF_LANE_REFRESH_LIP2
#define DCBF_LANE_REFRESH_LIP2 0
#endif
template <> struct LaneRefresh<LipW, 2> { static constexpr bool v = DCBF_LANE_REFRESH_LIP2 != 0; };
#ifndef DCBF_LANE_REFRESH_DDL
#define DCBF_LANE_REFRESH_DDL 1
#endif
template <> struct LaneRefresh<DdL, 2> { static constexpr bool v = DCBF_LANE_REFRESH_DDL != 0; };

// per-problem scratch (one per warp) and the constants are static shared-memory objects, so every function sees them as
// shared-space symbols (no generic pointers through the out-of-line calls)
template <class M, int NS> __shared__ WarpShared<M, NS> g_sm[Wpc<M, NS>::v];
__shared__ CtaShared g_cs;
__shared__ LipTabShared g_lt;

// iteration barrier of a CTA: returns the number of threads that still have work (a warp without work keeps arriving until the
// count is zero).  One out-of-line copy so that every arrival is the same instruction.
__device__ __noinline__ int cta_tick(int working) { return __syncthreads_count(working); }
__device__ __forceinline__ int warp_in_cta() {
    int t;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
    return t >> 5;
}

// %laneid through a volatile asm: the value stays in a register (the compiler otherwise re-reads SR_TID.X at every use)
__device__ __forceinline__ int lane_id() {
    int l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

// alpha * a^2.3 > t^1.1 (switching condition of the filter) on the MUFU log2: the condition only selects which acceptance test a
// trial has to pass, single precision is plenty
__device__ __forceinline__ bool switch_cond_fast(double alpha, double a, double t) {
    if (!(t > 0.0)) return alpha > 0.0;
    return __log2f((float)alpha) + 2.3f * __log2f((float)a) > 1.1f * __log2f((float)t);
}

// max / min over the warp of NON-NEGATIVE doubles: their bit patterns order like unsigned integers, so two 32-bit hardware
// reductions (redux.sync) replace five shuffle rounds.  A NaN operand wins the max (and is then caught by the caller's checks).
__device__ __forceinline__ double wmax(double v) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(FULL, hi);
    const unsigned ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}
__device__ __forceinline__ double wmin(double v) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_min_sync(FULL, hi);
    const unsigned ml = __reduce_min_sync(FULL, hi == mh ? lo : 0xffffffffu);
    return __hiloint2double((int)mh, (int)ml);
}
__device__ __forceinline__ int wsumi(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

struct RowDesc { int type, step, obs, cls; };
struct RowBnd { double lo, hi; bool has_lo, has_hi; };

// value of one row and, with GRAD, the scalars its gradient is made of.
//   LIP: d row / d foot_l = cA[l] * (p0, p1) + cB[l] * (q0, q1),   d row / d turn_l = (l <= step) * t_all + (l == step) * t_own
//   DD : grad = p0 Jx[i+1] + p1 Jy[i+1] + q0 Jx[i] + q1 Jy[i] + t_all e_{v_i} + t_own e_{w_i}
struct RowEval { double c, p0, p1, q0, q1, t_all, t_own, hq0, hq1, hq2; };

struct WState {   // replicated scalars of one problem (registers); the small counters share one register, the rarely touched doubles
                  // live in WarpShared::cold (C_ALPHA ...)
    double mu, sf;
    int iters, status;
    unsigned phase : 1, pending : 1, reinit : 1, first : 1;
    unsigned nf : 4;        // filter entries (<= DCBF_FILT)
    unsigned acc_cnt : 5;   // consecutive acceptable iterates (15 ends the solve) / stagnating restoration steps (2 do)
    unsigned nstall : 2, tiny : 4;
    unsigned nresto : 2;    // restoration steps of this restoration phase, saturating (the windowed stagnation test needs >= 2)
};
static_assert(DCBF_FILT <= 15, "WState::nf is a 4-bit field");

// D-CBF row on the staged obstacle record (circle: a' = c' = 1, b' = 0); shared by both models
template <bool GRAD>
__device__ __forceinline__ void eval_cbf(const double *o, double gm1, double x1, double y1, double x0, double y0, RowEval &e) {
    const double ax = x1 - o[0], ay = y1 - o[1], bx = x0 - o[0], by = y0 - o[1];
    const double ea = o[2], eb = o[3], ec = o[4];
    e.c = (ea * ax * ax + eb * ax * ay + ec * ay * ay - o[5]) + gm1 * (ea * bx * bx + eb * bx * by + ec * by * by - o[5]);
    if (GRAD) {
        e.p0 = 2.0 * ea * ax + eb * ay; e.p1 = 2.0 * ec * ay + eb * ax;
        e.q0 = gm1 * (2.0 * ea * bx + eb * by); e.q1 = gm1 * (2.0 * ec * by + eb * bx);
        e.hq0 = 2.0 * ea; e.hq1 = eb; e.hq2 = 2.0 * ec;
    }
}

// one lane per obstacle: selection (MPC_LIP_modi.py:325-338), compaction into sm.obs; returns the number of selected obstacles.
// `sel_out` / `rec` let the LIP model run the detour heuristic on the same lanes.
template <class M, int NS>
__device__ __forceinline__ int stage_obstacles(const dcbf_params &P, WarpShared<M, NS> &sm, const BatchIn &in, int b, int lane, double px, double py,
                                               double *rec, bool &sel_out, bool &is_circle, unsigned *mask_out = nullptr) {
    bool bad;
    const int fld = batch_field(in, b, bad);
    if (bad && lane == 0) { sm.x0[2] = nan(""); sm.x0[4] = nan(""); }   // invalid field index: the solve returns -13 (the caller syncs the warp before it reads x0 again)
    const double *cir = in.cir_rec + (size_t)fld * in.Kc * DCBF_CIR_REC;
    const double *elp = in.elp_rec + (size_t)fld * in.Ke * DCBF_ELP_REC;
    const int j = lane;
    const bool is_c = j < in.Kc, is_e = !is_c && j < in.Kc + in.Ke;
    rec[0] = 0; rec[1] = 0; rec[2] = 1; rec[3] = 0; rec[4] = 1; rec[5] = 0;
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
    if (sel) {
        const int pos = __popc(mask & ((1u << lane) - 1u));
        DCBF_ASSERT(pos >= 0 && (size_t)fld * in.Kc + in.Kc <= (size_t)(in.F > 0 ? in.F : 1 << 30) * in.Kc);
        if (pos < KsMax<NS>::v) {
#pragma unroll
            for (int c = 0; c < 6; c++) sm.obs[pos][c] = rec[c];
        }
    }
    sel_out = sel; is_circle = is_c;
    if (mask_out) *mask_out = mask;
    return __popc(mask);
}

// ===============================================================================================================
// LIP model (sig_step / modi): z = (fx0, fy0, fx1, fy1, fx2, fy2, t0, t1, t2)
// ===============================================================================================================
template <bool LIN>
struct LipWT {
    using Self = LipWT<LIN>;
    static constexpr int N = 9;
    static constexpr int NHQ = 3;
    static constexpr int NROUND = 3;
    static constexpr bool HAS_CURV = false;
    static constexpr bool ROLL2 = false;   // two-slot kernel: unrolled slot loop, row state in registers (measured faster for modi)
    static constexpr bool LIN2 = LIN;      // slot 1 holds the three turn rows only (closed-form contributions, 32-column dot products)
    static constexpr int NLINP = 4;        // columns of the linear slot that are staged (rows 32..34, padded to an even count)
    using NodeData = LipNodeData;
    enum { RT_NONE = 0, RT_CBF, RT_VBX, RT_VBY, RT_LEG, RT_DTH, RT_FENP, RT_FENM };

    static __device__ __forceinline__ const int *desc(const WarpTables *tab) { return tab->desc; }
    static __device__ __forceinline__ const int *desc_lin(const WarpTables *tab) { return tab->desc_lin_lip; }
    static __device__ __forceinline__ int rows_per_step(const dcbf_params &P, int Ks) { return Ks + (P.has_fen ? 6 : 4); }
    static __device__ __forceinline__ int class_start(int cls, int, int ms) { return cls * ms; }   // first row that can touch a variable of step cls
    // rows of a step that sit in the generic slots (LipL: all but the turn row) and the offset of the fen rows behind the D-CBF rows
    static __device__ __forceinline__ int slot_rows_per_step(int ms) { return LIN ? ms - 1 : ms; }

    // problem setup: obstacle selection + compaction, detour heuristic (MPC_LIP_sig_step.py:229-253), node 0
    template <int NS>
    static __device__ __forceinline__ int setup(WarpShared<Self, NS> &sm, const dcbf_params &P, const BatchIn &in, int b, int lane, unsigned *mask_out = nullptr) {
        const double px = sm.x0[0], py = sm.x0[1];
        double rec[6];
        bool sel, is_c;
        const int Ks = stage_obstacles<Self, NS>(P, sm, in, b, lane, px, py, rec, sel, is_c, mask_out);
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
        __syncwarp();
        const unsigned hm = __ballot_sync(FULL, hit);
        double g0 = gx, g1 = gy;
        if (hm) {
            const int src = __ffs(hm) - 1;
            g0 = __shfl_sync(FULL, ngx, src); g1 = __shfl_sync(FULL, ngy, src);
        }
        if (lane == 0) { sm.goal[0] = g0; sm.goal[1] = g1; }
        if (lane < 5) sm.nd.nodes[0][lane] = sm.x0[lane];
        if (lane == 5) sm.nd.NHf[NSRC] = 0.0;   // the padding terms of the Hessian table read this slot
        __syncwarp();
        return Ks;
    }

    // step-major row order: step i holds  [D-CBF x Ks, v_bx, v_by, leg, turn, (fen+, fen-)]
    // LipL: slot 0 holds  [D-CBF x Ks, v_bx, v_by, leg, (fen+, fen-)]  per step (at most 32 rows), the turn row of step i is row 32 + i
    static __device__ __forceinline__ RowDesc row_desc(int r, int Ks, int ms, int m) {
        RowDesc d;
        d.type = RT_NONE; d.step = 0; d.obs = 0; d.cls = 9;
        if (LIN) {
            const int ms0 = ms - 1;
            if (r < 32) {
                if (r < 3 * ms0) {
                    d.step = r / ms0;
                    const int w = r - d.step * ms0;
                    if (w < Ks) { d.type = RT_CBF; d.obs = w; d.cls = d.step; }
                    else {
                        const int k = w - Ks;
                        d.type = k < 3 ? RT_VBX + k : RT_FENP + (k - 3);
                        d.cls = d.type == RT_LEG ? 6 + d.step : 3 + d.step;
                    }
                }
            } else if (r < 35) { d.type = RT_DTH; d.step = r - 32; }
            return d;
        }
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

    // static bounds of a row (MPC_LIP_sig_step.py:193-227, MPC_LIP_modi.py:203-245; split form of the coupling row)
    static __device__ __forceinline__ RowBnd row_bounds(const dcbf_params &P, const RowDesc &rd, int leg) {
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

    // a turn row: the value of the linear form on z (the direction pass calls it on dz)
    static __device__ __forceinline__ double eval_lin(const dcbf_params &, const RowDesc &rd, const double *z) {
        return rd.type == RT_DTH ? z[6 + rd.step] : 0.0;
    }
    // closed-form contribution of the turn rows to the lane's entry of the assembled system (descriptor: WarpTables::desc_lin_lip)
    template <int NS>
    static __device__ __forceinline__ double lin_term(const WarpShared<Self, NS> &sm, const dcbf_params &, int dl) {
        constexpr int RP = WarpShared<Self, NS>::RP;
        DCBF_ASSERT(((dl >> 8) & 0xff) < (WarpShared<Self, NS>::NST) && (dl >> 16) < RP);
        const double v = (&sm.ST[0][0])[((dl >> 8) & 0xff) * RP + (dl >> 16)];   // dl == 0 reads ST[0][0]: always initialised
        return (dl & 1) ? v : 0.0;
    }

    template <int NS, bool GRAD>
    static __device__ __forceinline__ void eval_row(const dcbf_params &P, const WarpShared<Self, NS> &sm, const RowDesc &rd, const double *z, RowEval &e) {
        e.c = 0.0;
        if (GRAD) { e.p0 = e.p1 = e.q0 = e.q1 = e.t_all = e.t_own = 0.0; e.hq0 = e.hq1 = e.hq2 = 0.0; }
        const int i = rd.step, kn = i + 1;
        const double (*nodes)[5] = sm.nd.nodes;
        if (rd.type == RT_CBF) {
            DCBF_ASSERT(rd.obs >= 0 && rd.obs < KsMax<NS>::v && i >= 0 && i < 3);
            eval_cbf<GRAD>(sm.obs[rd.obs], P.gamma - 1.0, nodes[kn][0], nodes[kn][1], nodes[i][0], nodes[i][1], e);
        } else if (rd.type != RT_NONE) {
            // every other row is  al * v_bx + be * v_by + ga * turn_i + de * |pos_i - foot_i|^2  with coefficients of the row type
            // (one code path for the twelve non-D-CBF lanes instead of three)
            const int t = rd.type;
            const double al = (t == RT_VBX || t == RT_FENP || t == RT_FENM) ? 1.0 : 0.0, be = t == RT_VBY ? 1.0 : 0.0;
            const double ga = t == RT_DTH ? 1.0 : (t == RT_FENP ? P.s_turn : (t == RT_FENM ? -P.s_turn : 0.0)), de = t == RT_LEG ? 1.0 : 0.0;
            const double sn = sm.nd.trig[kn][0], cs = sm.nd.trig[kn][1];
            const double vx = nodes[kn][2], vy = nodes[kn][3];
            const double vbx = cs * vx + sn * vy, vby = -sn * vx + cs * vy;
            const double lx = nodes[i][0] - z[2 * i], ly = nodes[i][1] - z[2 * i + 1];
            e.c = (al * vbx + be * vby) + (ga * z[6 + i] + de * (lx * lx + ly * ly));
            if (GRAD) {
                e.p0 = al * cs - be * sn; e.p1 = al * sn + be * cs; e.t_all = al * vby - be * vbx; e.t_own = ga;
                e.q0 = de * 2.0 * lx; e.q1 = de * 2.0 * ly;
            }
        }
    }

    // transposed staging of one row: gradient, sigma * gradient (rows beyond m stage zeros), D-CBF curvature weights
    template <int NS>
    static __device__ __forceinline__ void stage_row(WarpShared<Self, NS> &sm, const CtaShared &cs_, const RowDesc &rd, const RowEval &e, double sig, double y, int r) {
        constexpr int RP = WarpShared<Self, NS>::RP;
        const double *ab = g_lt.cab[rd.cls];
        const int i = rd.step;
        double *col = &sm.ST[0][r];
#pragma unroll
        for (int l = 0; l < 3; l++) {
            const double ca = ab[l], cb = ab[3 + l];
            const double gxv = fma(ca, e.p0, cb * e.q0), gyv = fma(ca, e.p1, cb * e.q1);
            const double gtv = (l <= i ? e.t_all : 0.0) + (l == i ? e.t_own : 0.0);
            col[(2 * l) * RP] = gxv; col[(2 * l + 1) * RP] = gyv; col[(6 + l) * RP] = gtv;
            col[(9 + 2 * l) * RP] = sig * gxv; col[(9 + 2 * l + 1) * RP] = sig * gyv; col[(9 + 6 + l) * RP] = sig * gtv;
        }
        DCBF_ASSERT(r >= 0 && r < (int)(sizeof(sm.HQ) / sizeof(sm.HQ[0])));
        sm.HQ[r][0] = y * e.hq0; sm.HQ[r][1] = y * e.hq1; sm.HQ[r][2] = y * e.hq2;
    }

    // nodes 1..3 at the point z (lanes 0..2, directly from the free response and the constant influence coefficients), per-node
    // trigonometry and objective terms; collective, out of line
    template <int NS>
    static __device__ __noinline__ void nodes(const double *z, int lane, int wid, double sf, bool want_hess) {
        WarpShared<Self, NS> &sm = g_sm<Self, NS>[wid];
        const CtaShared &cs_ = g_cs;
        const dcbf_params &P = cs_.P;
        if (lane < 3) {
            const int kn = lane + 1;
            double x = sm.nd.fr[kn][0], y = sm.nd.fr[kn][1], vx = sm.nd.fr[kn][2], vy = sm.nd.fr[kn][3], th = sm.x0[4];
            const double *cx = g_lt.cab[lane], *cv = g_lt.cab[3 + lane];   // gx[kn-1-l], gv[kn-1-l] for l < kn, else 0
#pragma unroll
            for (int l = 0; l < 3; l++) {
                const double fx = z[2 * l], fy = z[2 * l + 1];
                x = fma(cx[l], fx, x); y = fma(cx[l], fy, y); vx = fma(cv[l], fx, vx); vy = fma(cv[l], fy, vy);
                th += l < kn ? z[6 + l] : 0.0;
            }
            double *nk = sm.nd.nodes[kn];
            nk[0] = x; nk[1] = y; nk[2] = vx; nk[3] = vy; nk[4] = th;
            double sn, cs;
            fsincos(th, &sn, &cs);   // inline: interleaves with the atan2 chain below
            sm.nd.trig[kn][0] = sn; sm.nd.trig[kn][1] = cs;
            node_objective(P, sm.goal, P.w_q + (kn == 1 ? P.w_p : 0.0), x, y, th, sf, want_hess, sm.nd.nobj[kn]);
        }
        __syncwarp();
    }

    // objective value at the staged nodes and its gradient component for lane a < 9 (0 on the other lanes)
    template <int NS>
    static __device__ __forceinline__ double objective(const WarpShared<Self, NS> &sm) {
        return sm.nd.nobj[1][0] + sm.nd.nobj[2][0] + sm.nd.nobj[3][0];
    }
    template <int NS>
    static __device__ __forceinline__ double grad(const WarpShared<Self, NS> &sm, const CtaShared &cs_, int lane, int ln) {
        // d node_kn / d foot_l = gx[kn-1-l],  d th_kn / d turn_l = 1; branch-free: lanes >= 9 compute on clamped indices, select 0
        const int l = ln < 6 ? ln >> 1 : ln - 6;
        const int c = ln < 6 ? 1 + (ln & 1) : 3;
        double g = 0.0;
#pragma unroll
        for (int kn = 1; kn <= 3; kn++) {
            const double wf = g_lt.cab[kn - 1][l];
            const double w = ln < 6 ? wf : (l < kn ? 1.0 : 0.0);
            g = fma(sm.nd.nobj[kn][c], w, g);
        }
        return lane < 9 ? g : 0.0;
    }
    template <int NS>
    static __device__ __forceinline__ void rescale_objective_hessian(WarpShared<Self, NS> &sm, int lane, double sf) {
        if (lane < 3) {
#pragma unroll
            for (int c = 4; c < 10; c++) sm.nd.nobj[lane + 1][c] *= sf;
        }
    }

    // node Hessians -> the 27 sources of the Hessian table
    template <int NS>
    static __device__ __forceinline__ void hess_sources(WarpShared<Self, NS> &sm, const dcbf_params &P, int lane, int Ks, int ms_all) {
        const double gm1 = P.gamma - 1.0;
        const double *yv = sm.ST[2 * N + 3];
        const bool has_fen = P.has_fen != 0;
        const int ms = slot_rows_per_step(ms_all);   // rows of a step in the generic slots
        constexpr int FEN = LIN ? 3 : 4;             // fen+ / fen- behind the D-CBF rows of their step (LipL has no turn row in between)
        const int lp = lane - 16;
        if ((unsigned)lp < 9u) {   // position block on lanes 16..24: node kn = lp/3 + 1, component lp % 3
            const int kn = lp / 3 + 1, c = lp % 3;
            double acc = sm.nd.nobj[kn][4 + c];
            const int r0 = (kn - 1) * ms, r1 = kn < 3 ? kn * ms : r0;
            const double g1 = kn < 3 ? gm1 : 0.0;
            if (NS == 1 || LIN) {   // at most six D-CBF rows per step: fixed trip count, rows beyond Ks masked (no loop bookkeeping)
#pragma unroll
                for (int j = 0; j < KsMax<1>::v; j++) {
                    const int jj = j < Ks ? j : 0;
                    const double a0 = sm.HQ[r0 + jj][c], a1 = sm.HQ[r1 + jj][c];
                    acc += j < Ks ? fma(g1, a1, a0) : 0.0;
                }
            } else {         // same arithmetic in the same order (the size-class split must not change a result)
                for (int j = 0; j < Ks; j++) acc += fma(g1, sm.HQ[r1 + j][c], sm.HQ[r0 + j][c]);
            }
            sm.nd.NHf[8 * (kn - 1) + c] = acc;
        } else if (lane < 3) {     // heading / velocity entries of node kn = lane + 1 (rows of step kn - 1)
            const int kn = lane + 1, i = kn - 1, base = i * ms + Ks;
            double Yx = yv[base], Yy = yv[base + 1];
            if (has_fen) Yx += yv[base + FEN] + yv[base + FEN + 1];
            const double sn = sm.nd.trig[kn][0], cs = sm.nd.trig[kn][1];
            const double vx = sm.nd.nodes[kn][2], vy = sm.nd.nodes[kn][3];
            const double vbx = cs * vx + sn * vy, vby = -sn * vx + cs * vy;
            double *H = &sm.nd.NHf[8 * i];
            H[3] = sm.nd.nobj[kn][7]; H[4] = sm.nd.nobj[kn][8];
            H[5] = sm.nd.nobj[kn][9] - (Yx * vbx + Yy * vby);
            H[6] = -sn * Yx - cs * Yy; H[7] = cs * Yx - sn * Yy;
            sm.nd.NHf[24 + i] = 2.0 * yv[base + 2];
        }
    }
    template <int NS>
    static __device__ __forceinline__ void hess_curvature(WarpShared<Self, NS> &, int, double, const double *) {}
    // Lagrangian Hessian of matrix entry e (packed index) through the table.  TERMS = the largest number of table terms an entry of
    // this assembly round has (round 0 holds the entries among the variables of step 0: up to 6; round 1: 2; round 2 has vector
    // entries only) -- build_warp_tables() verifies the bound
    template <int NS>
    static __device__ __forceinline__ double hess_entry(const WarpShared<Self, NS> &sm, const CtaShared &cs_, int e, double sf, int terms, int) {
        double acc = 0.0;
#pragma unroll
        for (int h = 0; h < NHT; h++) if (h < terms) { DCBF_ASSERT(e >= 0 && e < 48 && g_lt.hs[h][e] <= NSRC); acc = fma(g_lt.hc[h][e], sm.nd.NHf[g_lt.hs[h][e]], acc); }   // `terms` is a constant after unrolling
        return acc;
    }
    static __host__ __device__ constexpr int round_terms(int t) { return t == 0 ? 6 : (t == 1 ? 2 : 0); }

    // objective terms at one node: f, (nx, ny, nt), Hessian (xx, xy, yy, xt, yt, tt) scaled by sf; shared with the DD model
    static __device__ __forceinline__ void node_objective(const dcbf_params &P, const double *goal, double w, double x, double y, double th, double sf,
                                                          bool want_hess, double *o) {
        const double ex = x - goal[0], ey = y - goal[1];
        const double dx = -ex, dy = -ey;
        const double r2 = dx * dx + dy * dy, ir2 = frcp(r2);
        const double phi = th - fatan2(dy, dx);
        const double px = -dy * ir2, py = dx * ir2;
        o[0] = w * (ex * ex + ey * ey) + P.w_r * phi * phi;
        o[1] = 2.0 * w * ex + 2.0 * P.w_r * phi * px;
        o[2] = 2.0 * w * ey + 2.0 * P.w_r * phi * py;
        o[3] = 2.0 * P.w_r * phi;
        if (want_hess) {
            const double ir4 = ir2 * ir2;
            const double pxx = -2.0 * dx * dy * ir4, pyy = -pxx, pxy = (dx * dx - dy * dy) * ir4;
            const double r2w = 2.0 * P.w_r * sf;
            o[4] = sf * 2.0 * w + r2w * (px * px + phi * pxx);
            o[5] = r2w * (px * py + phi * pxy);
            o[6] = sf * 2.0 * w + r2w * (py * py + phi * pyy);
            o[7] = r2w * px; o[8] = r2w * py; o[9] = r2w;
        }
    }
};

// ===============================================================================================================
// DD model (MPC_DD_sig_step.py): z = (v0, w0, v1, w1, v2, w2), x+ = x + dt v cos th, y+ = y + dt v sin th, th+ = th + w
// ===============================================================================================================
template <bool LIN>
struct DdWT {
    using Self = DdWT<LIN>;
    static constexpr int N = 6;
    static constexpr int NHQ = 7;   // y * (2a', b', 2c', h1x, h1y, h0x, h0y)
    static constexpr int NROUND = 2;
    static constexpr bool HAS_CURV = true;
    static constexpr bool ROLL2 = !LIN;    // generic two-slot kernel: rolled slot loop, row state in shared memory (measured faster for DD)
    static constexpr bool LIN2 = LIN;      // slot 1 holds the linear rows only (unrolled slots, closed-form contributions, 32-column dot products)
    using NodeData = DdNodeData;
    enum { RT_NONE = 0, RT_CBF, RT_FENP, RT_FENM, RT_BV, RT_BW };

#ifndef DCBF_DDL_NLINP
#define DCBF_DDL_NLINP 12
#endif
    static constexpr int NLINP = LIN ? DCBF_DDL_NLINP : 32;   // staged columns of the linear slot (twelve linear rows; 32: rows without a type stage zeros)
    static __device__ __forceinline__ const int *desc(const WarpTables *tab) { return tab->desc_dd; }
    static __device__ __forceinline__ const int *desc_lin(const WarpTables *tab) { return tab->desc_lin; }
    static __device__ __forceinline__ int rows_per_step(const dcbf_params &, int Ks) { return Ks + 4; }
    static __device__ __forceinline__ int class_start(int cls, int Ks, int) { return cls * Ks; }

    template <int NS>
    static __device__ __forceinline__ int setup(WarpShared<Self, NS> &sm, const dcbf_params &P, const BatchIn &in, int b, int lane, unsigned *mask_out = nullptr) {
        double rec[6];
        bool sel, is_c;
        const int Ks = stage_obstacles<Self, NS>(P, sm, in, b, lane, sm.x0[0], sm.x0[1], rec, sel, is_c, mask_out);
        if (lane == 0) { sm.goal[0] = sm.graw[0]; sm.goal[1] = sm.graw[1]; }   // no detour heuristic (MPC_DD_sig_step.py:144-168 is commented out)
        if (lane < 3) sm.nd.nodes[0][lane] = sm.x0[lane];
        if (lane < 6) { sm.nd.Jx[0][lane] = 0.0; sm.nd.Jy[0][lane] = 0.0; }
        {   // the columns of linear and of unused rows keep their zeros for the whole solve (see stage_row)
            double *st = &sm.ST[0][0];
            for (int t = lane; t < WarpShared<Self, NS>::NST * WarpShared<Self, NS>::RP; t += 32) st[t] = 0.0;
        }
        __syncwarp();
        return Ks;
    }

    // row order: all D-CBF rows step-major (3 Ks), then per step the four linear rows [fen+, fen-, bound v, bound w].  A row of
    // step s has zero derivatives with respect to the variables of later steps, so a dot product that involves a variable of step
    // c may start at row c * Ks (the linear rows of earlier steps it then sweeps are zero in that variable).  With K = 10 the
    // second slot of a lane holds linear rows only, whose staging is two numbers.
    static __device__ __forceinline__ RowDesc row_desc(int r, int Ks, int, int m) {
        RowDesc d;
        d.type = RT_NONE; d.step = 0; d.obs = 0; d.cls = 0;
        if (LIN) {   // D-CBF rows 0 .. 3 Ks - 1 (Ks <= 10) in slot 0, linear row w of step i at 32 + 4 i + w
            if (r < 32) { if (r < 3 * Ks) { d.type = RT_CBF; d.step = r / Ks; d.obs = r - d.step * Ks; } }
            else if (r < 44) { const int w = r - 32; d.step = w >> 2; d.type = RT_FENP + (w & 3); }
            return d;
        }
        if (r >= m) return d;
        if (r < 3 * Ks) { d.type = RT_CBF; d.step = r / Ks; d.obs = r - d.step * Ks; }
        else { const int w = r - 3 * Ks; d.step = w >> 2; d.type = RT_FENP + (w & 3); }
        return d;
    }
    // a linear row (value of the linear form cv * z_v + cw * z_w of its step; the direction pass calls it on dz)
    static __device__ __forceinline__ double eval_lin(const dcbf_params &P, const RowDesc &rd, const double *z) {
        const int i = rd.step;
        const double cv = (rd.type == RT_BW || rd.type == RT_NONE) ? 0.0 : 1.0;
        const double cw = rd.type == RT_FENP ? P.s_turn : (rd.type == RT_FENM ? -P.s_turn : (rd.type == RT_BW ? 1.0 : 0.0));
        return cv * z[2 * i] + cw * z[2 * i + 1];
    }
    // closed-form contribution of the linear rows to the lane's entry of the assembled system (descriptor: WarpTables::desc_lin)
    template <int NS>
    static __device__ __forceinline__ double lin_term(const WarpShared<Self, NS> &sm, const dcbf_params &P, int dl) {
        constexpr int RP = WarpShared<Self, NS>::RP;
        DCBF_ASSERT((((dl >> 8) & 0xff) < (WarpShared<Self, NS>::NST) && (dl >> 16) + 3 < RP && ((dl >> 16) & 1) == 0));
        const double *px = &sm.ST[0][0] + ((dl >> 8) & 0xff) * RP + (dl >> 16);
        const double2 xa = *reinterpret_cast<const double2 *>(px), xb = *reinterpret_cast<const double2 *>(px + 2);
        const double s1 = xa.x + xa.y, s2 = xa.x - xa.y, st = P.s_turn;
        return ((dl & 1) ? s1 : 0.0) + ((dl & 2) ? st * s2 : 0.0) + ((dl & 4) ? xb.x : 0.0) + ((dl & 8) ? xb.y : 0.0) + ((dl & 16) ? st * st * s1 : 0.0);
    }
    // MPC_DD_sig_step.py:127-141 (variable bounds as rows; split form of the coupling row)
    static __device__ __forceinline__ RowBnd row_bounds(const dcbf_params &P, const RowDesc &rd, int) {
        RowBnd b;
        b.lo = -1e300; b.hi = 1e300; b.has_lo = false; b.has_hi = false;
        if (rd.type == RT_CBF) { b.lo = 0.0; b.has_lo = true; }
        else if (rd.type == RT_FENP || rd.type == RT_FENM) { b.hi = P.bvx_max; b.has_hi = true; }
        else if (rd.type == RT_BV) { b.lo = P.bvx_min; b.hi = P.bvx_max; b.has_lo = b.has_hi = true; }
        else if (rd.type == RT_BW) { b.lo = -P.ang_max; b.hi = P.ang_max; b.has_lo = b.has_hi = true; }
        return b;
    }

    template <int NS, bool GRAD>
    static __device__ __forceinline__ void eval_row(const dcbf_params &P, const WarpShared<Self, NS> &sm, const RowDesc &rd, const double *z, RowEval &e) {
        e.c = 0.0;
        if (GRAD) { e.p0 = e.p1 = e.q0 = e.q1 = e.t_all = e.t_own = 0.0; e.hq0 = e.hq1 = e.hq2 = 0.0; }
        const int i = rd.step;
        if (rd.type == RT_CBF) {
            DCBF_ASSERT(rd.obs >= 0 && rd.obs < KsMax<NS>::v && i >= 0 && i < 3);
            eval_cbf<GRAD>(sm.obs[rd.obs], P.gamma - 1.0, sm.nd.nodes[i + 1][0], sm.nd.nodes[i + 1][1], sm.nd.nodes[i][0], sm.nd.nodes[i][1], e);
        } else if (rd.type != RT_NONE) {
            const double v = z[2 * i], w = z[2 * i + 1];
            const double cv = rd.type == RT_BW ? 0.0 : 1.0;
            const double cw = rd.type == RT_FENP ? P.s_turn : (rd.type == RT_FENM ? -P.s_turn : (rd.type == RT_BW ? 1.0 : 0.0));
            e.c = cv * v + cw * w;
            if (GRAD) { e.t_all = cv; e.t_own = cw; }
        }
    }

    // D-CBF rows: chain rule through the node Jacobians.  Linear rows: the gradient is (t_all, t_own) on (v_i, w_i); the other
    // entries of their columns were zeroed once per problem (setup) and nobody else writes them.
    template <int NS>
    static __device__ __forceinline__ void stage_row(WarpShared<Self, NS> &sm, const CtaShared &, const RowDesc &rd, const RowEval &e, double sig, double y, int r) {
        constexpr int RP = WarpShared<Self, NS>::RP;
        const int i = rd.step, kn = i + 1;
        double *col = &sm.ST[0][r];
        if (rd.type == RT_CBF) {
            const double *jx1 = sm.nd.Jx[kn], *jy1 = sm.nd.Jy[kn], *jx0 = sm.nd.Jx[i], *jy0 = sm.nd.Jy[i];
#pragma unroll
            for (int a = 0; a < 6; a++) {
                const double g = fma(e.p0, jx1[a], fma(e.p1, jy1[a], fma(e.q0, jx0[a], e.q1 * jy0[a])));
                col[a * RP] = g; col[(6 + a) * RP] = sig * g;
            }
            DCBF_ASSERT(r >= 0 && r < (int)(sizeof(sm.HQ) / sizeof(sm.HQ[0])));
            double *h = sm.HQ[r];
            h[0] = y * e.hq0; h[1] = y * e.hq1; h[2] = y * e.hq2; h[3] = y * e.p0; h[4] = y * e.p1; h[5] = y * e.q0; h[6] = y * e.q1;
        } else if (rd.type != RT_NONE) {
            col[(2 * i) * RP] = e.t_all; col[(2 * i + 1) * RP] = e.t_own;
            col[(6 + 2 * i) * RP] = sig * e.t_all; col[(6 + 2 * i + 1) * RP] = sig * e.t_own;
        }
    }

    // headings and their sines / cosines (lanes 0..2 = th_0..2), then nodes 1..3 with their Jacobians and objective terms
    template <int NS>
    static __device__ __noinline__ void nodes(const double *z, int lane, int wid, double sf, bool want_hess) {
        WarpShared<Self, NS> &sm = g_sm<Self, NS>[wid];
        const dcbf_params &P = g_cs.P;
        const double dt = g_cs.K.dt;
        if (lane < 3) {
            double th = sm.x0[2];
#pragma unroll
            for (int l = 0; l < 2; l++) th += l < lane ? z[2 * l + 1] : 0.0;
            double sn, cs;
            fsincos(th, &sn, &cs);
            sm.nd.trig[lane][0] = sn; sm.nd.trig[lane][1] = cs;
        }
        __syncwarp();
        if (lane < 3) {
            const int kn = lane + 1;
            double x = sm.x0[0], y = sm.x0[1], th = sm.x0[2];
            double jx[6], jy[6];
#pragma unroll
            for (int a = 0; a < 6; a++) { jx[a] = 0.0; jy[a] = 0.0; }
            // x_k = x_0 + dt sum_{l<k} v_l cos th_l;  d/dv_l = dt cos th_l;  d/dw_l = sum_{l<j<k} -dt v_j sin th_j
            double sx = 0.0, sy = 0.0;   // running sums of -dt v_j sin th_j / +dt v_j cos th_j over j > l (built backwards)
#pragma unroll
            for (int l = 2; l >= 0; l--) {
                if (l < kn) {
                    const double sn = sm.nd.trig[l][0], cs = sm.nd.trig[l][1], v = z[2 * l];
                    x = fma(dt * cs, v, x); y = fma(dt * sn, v, y); th += z[2 * l + 1];
                    jx[2 * l] = dt * cs; jy[2 * l] = dt * sn;
                    jx[2 * l + 1] = sx; jy[2 * l + 1] = sy;
                    sx -= dt * v * sn; sy += dt * v * cs;
                }
            }
            sm.nd.nodes[kn][0] = x; sm.nd.nodes[kn][1] = y; sm.nd.nodes[kn][2] = th;
#pragma unroll
            for (int a = 0; a < 6; a++) { sm.nd.Jx[kn][a] = jx[a]; sm.nd.Jy[kn][a] = jy[a]; }
            LipW::node_objective(P, sm.goal, P.w_q + (kn == 1 ? P.w_p : 0.0), x, y, th, sf, want_hess, sm.nd.nobj[kn]);
        } else if (lane == 3) {
            // control smoothness  t sum_i |u_i - u_{i-1}|^2,  u_{-1} = last_u  (MPC_DD_sig_step.py:351-369)
            double f = 0.0, g[6] = {0, 0, 0, 0, 0, 0};
            double pv = sm.nd.last_u[0], pw = sm.nd.last_u[1];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const double dv = z[2 * i] - pv, dw = z[2 * i + 1] - pw;
                f += P.w_t * (dv * dv + dw * dw);
                g[2 * i] += 2.0 * P.w_t * dv; g[2 * i + 1] += 2.0 * P.w_t * dw;
                if (i > 0) { g[2 * i - 2] -= 2.0 * P.w_t * dv; g[2 * i - 1] -= 2.0 * P.w_t * dw; }
                pv = z[2 * i]; pw = z[2 * i + 1];
            }
            sm.nd.nobj[0][0] = f;
#pragma unroll
            for (int a = 0; a < 6; a++) sm.nd.gsm[a] = g[a];
        }
        __syncwarp();
    }

    template <int NS>
    static __device__ __forceinline__ double objective(const WarpShared<Self, NS> &sm) {
        return sm.nd.nobj[0][0] + sm.nd.nobj[1][0] + sm.nd.nobj[2][0] + sm.nd.nobj[3][0];
    }
    template <int NS>
    static __device__ __forceinline__ double grad(const WarpShared<Self, NS> &sm, const CtaShared &, int lane, int ln) {
        const int a = ln < 6 ? ln : 5;
        double g = sm.nd.gsm[a];
#pragma unroll
        for (int kn = 1; kn <= 3; kn++) {
            const double et = ((a & 1) && (a >> 1) < kn) ? 1.0 : 0.0;   // d th_kn / d z_a
            g = fma(sm.nd.nobj[kn][1], sm.nd.Jx[kn][a], fma(sm.nd.nobj[kn][2], sm.nd.Jy[kn][a], fma(sm.nd.nobj[kn][3], et, g)));
        }
        return lane < 6 ? g : 0.0;
    }
    template <int NS>
    static __device__ __forceinline__ void rescale_objective_hessian(WarpShared<Self, NS> &sm, int lane, double sf) {
        if (lane < 3) {
#pragma unroll
            for (int c = 4; c < 10; c++) sm.nd.nobj[lane + 1][c] *= sf;
        }
    }

    // per-node second-order sources: Q_k = sum y Q_r (+ objective), C_k = sum y grad h (+ sf * objective gradient), and the
    // curvature coefficients of the positions (dd_add_second() in dcbf_core.cuh)
    template <int NS>
    static __device__ __forceinline__ void hess_sources(WarpShared<Self, NS> &sm, const dcbf_params &P, int lane, int Ks, int ms) {
        const double gm1 = P.gamma - 1.0;
        const int lp = lane - 16;
        if ((unsigned)lp < 15u) {   // lanes 16..30: node kn = lp / 5 + 1, component c = lp % 5  (qxx, qxy, qyy, cx, cy)
            const int kn = lp / 5 + 1, c = lp % 5;
            // rows of step kn - 1 see node kn as their far end (Q_r, h1); rows of step kn see it as their near end (gm1 Q_r, h0)
            double acc = 0.0;
            const int cf = c, cn = c < 3 ? c : c + 2;   // HQ slots: far end 0..2 / 3..4, near end 0..2 (x gm1) / 5..6
            if (LIN) {   // at most ten D-CBF rows per step: fixed trip count, rows beyond Ks masked (no loop bookkeeping)
                const double sc = kn < 3 ? (c < 3 ? gm1 : 1.0) : 0.0;
                const int r0 = (kn - 1) * Ks, r1 = kn < 3 ? kn * Ks : r0;
#pragma unroll
                for (int j = 0; j < 10; j++) {
                    const int jj = j < Ks ? j : 0;
                    DCBF_ASSERT(r0 + jj >= 0 && r1 + jj < 32 && cf < NHQ && cn < NHQ);
                    const double a0 = sm.HQ[r0 + jj][cf], a1 = sm.HQ[r1 + jj][cn];
                    acc += j < Ks ? fma(sc, a1, a0) : 0.0;
                }
            } else {
                for (int j = 0; j < Ks; j++) acc += sm.HQ[(kn - 1) * Ks + j][cf];
                if (kn < 3) {
                    const double sc = c < 3 ? gm1 : 1.0;   // q0 / q1 already carry gm1
                    for (int j = 0; j < Ks; j++) acc = fma(sc, sm.HQ[kn * Ks + j][cn], acc);
                }
            }
            if (c < 3) sm.nd.Q[kn][c] = acc + sm.nd.nobj[kn][4 + c];
            else sm.nd.Cc[kn][c - 3] = acc;   // the objective part is added by the lanes below (needs sf)
        }
    }
    template <int NS>
    static __device__ __forceinline__ void hess_curvature(WarpShared<Self, NS> &sm, int lane, double sf, const double *z) {
        if (lane >= 1 && lane < 3) {   // l = 1, 2
            const int l = lane;
            const double dt = g_cs.K.dt;
            double CX = 0.0, CY = 0.0;
            for (int kn = l + 1; kn < 4; kn++) { CX += sm.nd.Cc[kn][0] + sf * sm.nd.nobj[kn][1]; CY += sm.nd.Cc[kn][1] + sf * sm.nd.nobj[kn][2]; }
            const double sn = sm.nd.trig[l][0], cs = sm.nd.trig[l][1];
            sm.nd.cvw[l] = dt * (-CX * sn + CY * cs);
            sm.nd.cww[l] = -dt * z[2 * l] * (CX * cs + CY * sn);
        }
    }
    // Lagrangian Hessian of matrix entry e = tri(a, b), a >= b
    static __host__ __device__ constexpr int round_terms(int) { return 1; }
    template <int NS>
    static __device__ __forceinline__ double hess_entry(const WarpShared<Self, NS> &sm, const CtaShared &cs_, int e, double sf, int, int ab) {
        const int a = ab & 0xff, b = ab >> 8;   // the entry's variables (a >= b), decoded from the descriptor once per problem
        double acc = 2.0 * cs_.P.w_t * sf * sm.nd.sm_dd[e];
#pragma unroll
        for (int kn = 1; kn <= 3; kn++) {
            const double jxa = sm.nd.Jx[kn][a], jya = sm.nd.Jy[kn][a], jxb = sm.nd.Jx[kn][b], jyb = sm.nd.Jy[kn][b];
            const double qxx = sm.nd.Q[kn][0], qxy = sm.nd.Q[kn][1], qyy = sm.nd.Q[kn][2];
            acc += (qxx * jxa + qxy * jya) * jxb + (qxy * jxa + qyy * jya) * jyb;
            const double hxt = sm.nd.nobj[kn][7], hyt = sm.nd.nobj[kn][8], htt = sm.nd.nobj[kn][9];
            const double ua = hxt * jxa + hyt * jya, ub = hxt * jxb + hyt * jyb;
            const double ea = ((a & 1) && (a >> 1) < kn) ? 1.0 : 0.0, eb = ((b & 1) && (b >> 1) < kn) ? 1.0 : 0.0;
            acc += ea * ub + ua * eb + htt * ea * eb;
        }
        // curvature of the positions
        if (!(a & 1) && (b & 1) && (b >> 1) < (a >> 1)) acc += sm.nd.cvw[a >> 1];
        if ((a & 1) && (b & 1)) {
            const int la = a >> 1;
            if (la < 1) acc += sm.nd.cww[1];
            if (la < 2) acc += sm.nd.cww[2];
        }
        return acc;
    }
};

// slot loop of the driver: one row per lane and slot.  ROW_BEGIN binds rd / bb / the row state of slot s, ROW_END writes the state
// back (registers for NS == 1, sm.RS otherwise)
#define DCBF_ROW_BEGIN                                                                                         \
    const int r = s * 32 + lane;                                                                               \
    const RowDesc rd = ROLLED ? M::row_desc(r, Ks, ms, m) : rdA[ROLLED ? 0 : s];                                \
    const RowBnd bb = ROLLED ? M::row_bounds(P, rd, leg) : rbA[ROLLED ? 0 : s];                                 \
    double rs_, rzl_, rzu_, rds_, rel_, reu_;                                                                  \
    if (!ROLLED) { rs_ = rsA[ROLLED ? 0 : s]; rzl_ = rzlA[ROLLED ? 0 : s]; rzu_ = rzuA[ROLLED ? 0 : s];          \
                   rds_ = rdsA[ROLLED ? 0 : s]; rel_ = relA[ROLLED ? 0 : s]; reu_ = reuA[ROLLED ? 0 : s]; }      \
    else { rs_ = sm.RS[0][ROLLED ? r : 0]; rzl_ = sm.RS[ROLLED ? 1 : 0][ROLLED ? r : 0]; rzu_ = sm.RS[ROLLED ? 2 : 0][ROLLED ? r : 0]; \
           rds_ = sm.RS[ROLLED ? 3 : 0][ROLLED ? r : 0]; rel_ = sm.RS[ROLLED ? 4 : 0][ROLLED ? r : 0]; reu_ = sm.RS[ROLLED ? 5 : 0][ROLLED ? r : 0]; }
#define DCBF_ROW_END                                                                                           \
    if (!ROLLED) { rsA[ROLLED ? 0 : s] = rs_; rzlA[ROLLED ? 0 : s] = rzl_; rzuA[ROLLED ? 0 : s] = rzu_;          \
                   rdsA[ROLLED ? 0 : s] = rds_; relA[ROLLED ? 0 : s] = rel_; reuA[ROLLED ? 0 : s] = reu_; }      \
    else { sm.RS[0][ROLLED ? r : 0] = rs_; sm.RS[ROLLED ? 1 : 0][ROLLED ? r : 0] = rzl_; sm.RS[ROLLED ? 2 : 0][ROLLED ? r : 0] = rzu_; \
           sm.RS[ROLLED ? 3 : 0][ROLLED ? r : 0] = rds_; sm.RS[ROLLED ? 4 : 0][ROLLED ? r : 0] = rel_; sm.RS[ROLLED ? 5 : 0][ROLLED ? r : 0] = reu_; }
// descriptor / bounds only (value passes)
#define DCBF_ROW_DESC                                                                                          \
    const int r = s * 32 + lane;                                                                               \
    const RowDesc rd = ROLLED ? M::row_desc(r, Ks, ms, m) : rdA[ROLLED ? 0 : s];

// eight statistics reduced together: four sums (interleaved shuffle butterflies) and four extrema of non-negative numbers
// (hardware reductions); m1 carries the MINIMUM of the complementarity products
struct Stat8 { double s0, s1, s2, s3, m0, m1, m2, m3; };
// four sums with ten 64-bit shuffles instead of twenty: the first two butterfly rounds halve the number of values a lane carries
// (lanes keep the pair / the value their half is responsible for and send the other), three more rounds finish one value per
// 8-lane group, four broadcasts hand every lane all four totals
__device__ __forceinline__ void reduce4_sums(double &s0, double &s1, double &s2, double &s3, int lane) {
    const bool up = (lane & 16) != 0, b3 = (lane & 8) != 0;
    double a = up ? s2 : s0, b = up ? s3 : s1;
    a += __shfl_xor_sync(FULL, up ? s0 : s2, 16);
    b += __shfl_xor_sync(FULL, up ? s1 : s3, 16);
    double c = b3 ? b : a;
    c += __shfl_xor_sync(FULL, b3 ? a : b, 8);
    c += __shfl_xor_sync(FULL, c, 4);
    c += __shfl_xor_sync(FULL, c, 2);
    c += __shfl_xor_sync(FULL, c, 1);
    s0 = __shfl_sync(FULL, c, 0); s1 = __shfl_sync(FULL, c, 8); s2 = __shfl_sync(FULL, c, 16); s3 = __shfl_sync(FULL, c, 24);
}
__device__ __forceinline__ void reduce8_inline(Stat8 &t, int lane) {
    reduce4_sums(t.s0, t.s1, t.s2, t.s3, lane);
    t.m0 = wmax(t.m0); t.m1 = wmin(t.m1); t.m2 = wmax(t.m2); t.m3 = wmax(t.m3);
}

// ---------------------------------------------------------------------------------------------------------------
// the solver for one problem (all 32 lanes call it with identical arguments).  Inputs in sm: x0, graw, zc (+ model data).
// ---------------------------------------------------------------------------------------------------------------
template <class M, int NS>
__device__ void solve_warp(const dcbf_params &P, const BatchIn &in, int b, int lane, int wid, int leg, WState &S, double mu0) {
    using Sh = WarpShared<M, NS>;
    constexpr int RP = Sh::RP;
    constexpr int N = M::N, NK = N * (N + 1) / 2;
    constexpr int KQ_RHS_ = KQ_K + NK;   // = packed row N of the system: tri(N, j) = NK + j
    Sh &sm = g_sm<M, NS>[wid];
    const CtaShared &cs_ = g_cs;
    // P is the kernel's own parameter (constant bank): its fields are instruction operands of the inlined body, not values loaded
    // from the shared-memory copy and then carried in registers across the solve (the out-of-line functions read g_cs.P)
    // ---- problem setup -------------------------------------------------------------------------------------------
    const int Ks = M::template setup<NS>(sm, P, in, b, lane);
    const int ms = M::rows_per_step(P, Ks);   // rows per step
    const int m = 3 * ms;
    // Unrolled kernels (one slot; two slots for the LIP models): descriptor, bounds and row state live in registers.  Rolled
    // kernels: descriptor and bounds are recomputed per slot (a few integer instructions) and the state sits in sm.RS.
    constexpr bool ROLLED = Sh::ROLLED;
    constexpr int NREG = ROLLED ? 1 : NS, UNR = ROLLED ? 1 : NS;
    constexpr int DOTC = M::LIN2 ? 32 : 32 * NS;   // columns the dot products sweep (LIN2: the linear rows of slot 1 enter in closed form)
    static_assert(!M::LIN2 || (NS == 2 && !ROLLED), "the linear-row slot belongs to the unrolled two-slot kernel");
    RowDesc rdA[NREG];
    RowBnd rbA[NREG];
    double rsA[NREG], rzlA[NREG], rzuA[NREG], rdsA[NREG], relA[NREG], reuA[NREG];
    int nz_l = 0;
#pragma unroll UNR
    for (int s = 0; s < NS; s++) {
        const RowDesc d0 = M::row_desc(s * 32 + lane, Ks, ms, m);
        const RowBnd b0 = M::row_bounds(P, d0, leg);
        nz_l += (b0.has_lo ? 1 : 0) + (b0.has_hi ? 1 : 0);
        if (!ROLLED) {
            rdA[ROLLED ? 0 : s] = d0; rbA[ROLLED ? 0 : s] = b0;
            rsA[ROLLED ? 0 : s] = rzlA[ROLLED ? 0 : s] = rzuA[ROLLED ? 0 : s] = rdsA[ROLLED ? 0 : s] = relA[ROLLED ? 0 : s] = reuA[ROLLED ? 0 : s] = 0.0;
        }
    }
    const int nz = wsumi(nz_l), nrows = m;
    // the lane's dot products of the assembly rounds (packed: operand rows, first contributing step, output slot)
    int dsc[M::NROUND];
#pragma unroll
    for (int t = 0; t < M::NROUND; t++) dsc[t] = __ldg(M::desc(cs_.tab) + 32 * t + lane);
    int hab[M::NROUND];   // matrix entries: the pair of variables (a | b << 8) of the lane's entry (operand rows N + a and b)
#pragma unroll
    for (int t = 0; t < M::NROUND; t++) hab[t] = dsc[t] >= 0 ? (((dsc[t] & 0xff) - N) & 0xff) | (((dsc[t] >> 8) & 0xff) << 8) : 0;
    int dlin[M::NROUND];
#pragma unroll
    for (int t = 0; t < M::NROUND; t++) dlin[t] = M::LIN2 ? __ldg(M::desc_lin(cs_.tab) + 32 * t + lane) : 0;
    int ln = lane < N ? lane : N - 1;   // clamped lane: keeps the per-variable sections branch-free
    int rowbase = lane < N + 1 ? lane * (lane + 1) / 2 : 0;
    // ---- solver state ------------------------------------------------------------------------------------------------
    S.mu = mu0; S.sf = 1.0; sm.cold[C_ALPHA] = 0.0; sm.cold[C_ALPHA_Z] = 0.0; sm.cold[C_DELTA_LAST] = 0.0; sm.cold[C_LM_LAMBDA] = 1e-4;
    sm.cold[C_RESTO_TARGET] = 0.0; sm.cold[C_RESTO_ENTRY] = 0.0; sm.cold[C_THETA_MAX] = 1e300; sm.cold[C_THETA_MIN] = 0.0;
    S.nf = 0; S.iters = 0; S.acc_cnt = 0; S.status = -1; S.nstall = 0; S.tiny = 0;
    S.nresto = 0; sm.cold[C_V2_H1] = 0.0; sm.cold[C_V2_H2] = 0.0;
    sm.cold[C_NROWS_NZ] = (double)(nrows + nz); sm.cold[C_NZ] = (double)(nz > 0 ? nz : 1);
    S.phase = PH_MAIN; S.pending = false; S.reinit = true; S.first = true; sm.cold[C_OBJ] = 0.0; sm.cold[C_VIOL] = 0.0;
    const double tol = P.tol;
    const double *stf = &sm.ST[0][0];

    bool nodes_valid = false;      // the node data (with Hessian terms) already describe zc (staged by the accepted trial)
    bool carry_ok = false;         // sm.cold[C_CARRY_LOG] holds the sum of log(gaps) at the accepted trial point = barrier term of the next full pass
    for (;;) {
        if (LaneRefresh<M, NS>::v) {
            // 128-register kernels (16 warps per SM): the lane index is re-read at the top of every iteration, so nothing derived from
            // it is loop-invariant for the compiler.  Left alone, ~40 registers hold hoisted per-lane indices, predicates and
            // addresses across the whole solve -- and at 128 registers they spill to local memory, which misses the (then 28 KB) L1
            // and costs an L2 round trip each (long_scoreboard 12 % of the stall samples); recomputing them is ~170 integer
            // instructions per iteration.
            lane = lane_id();
            ln = lane < N ? lane : N - 1;
            rowbase = lane < N + 1 ? lane * (lane + 1) / 2 : 0;
            if (!ROLLED && DCBF_REFRESH_BOUNDS(NS)) {   // same for what is derived from the row bounds (relaxed bounds, pushes): opaque to the optimiser from here on
#pragma unroll
                for (int s = 0; s < NREG; s++) asm volatile("" : "+d"(rbA[s].lo), "+d"(rbA[s].hi));
            }
        }
        if (Wpc<M, NS>::v > 1) cta_tick(1);   // the warps of the CTA start every iteration together (shared instruction fetch)
        const bool resto = S.phase == PH_RESTO;
        if (!nodes_valid) M::template nodes<NS>(sm.zc, lane, wid, S.first ? 1.0 : (resto ? 0.0 : S.sf), true);
        nodes_valid = false;
        const double fobj = M::template objective<NS>(sm);
        const double grad_a = M::template grad<NS>(sm, cs_, lane, ln);   // lane a < N owns grad[a]
        if (S.first) {
            const double gmax = wmax(fabs(grad_a));
            S.sf = gmax > 100.0 ? fdiv(100.0, gmax) : 1.0;
            M::template rescale_objective_hessian<NS>(sm, lane, S.sf);   // the objective Hessian staged above used sf = 1
            __syncwarp();
        }
        const double sf_eff = resto ? 0.0 : S.sf;
        // ---- rows: evaluate, update row state, stage gradients and weights; statistics stay in registers ----------------------
        Stat8 st8;
        st8.s0 = st8.s1 = st8.s2 = st8.s3 = 0.0; st8.m0 = 0.0; st8.m1 = 1e300; st8.m2 = 0.0; st8.m3 = 0.0;
        double gap_prod = 1.0;   // unrolled slots: product of the lane's gaps (each in (0, ~1e2]), one logarithm after the loop
#pragma unroll UNR
        for (int s = 0; s < NS; s++) {
            DCBF_ROW_BEGIN
            RowEval e;
            bool lin = false;   // slot 1 of a LIN2 model: linear rows only (s is a constant after unrolling)
            if constexpr (M::LIN2) lin = s == 1;
            if (lin) { if constexpr (M::LIN2) e.c = M::eval_lin(P, rd, sm.zc); }
            else M::template eval_row<NS, true>(P, sm, rd, sm.zc, e);
            double sig = 0.0, w1 = 0.0, binv = 0.0, y = 0.0;
            double t_rc = 0.0, t_cmin = 1e300, t_cmax = 0.0, t_z = 0.0, t_log = 0.0, t_v2 = 0.0, t_v = 0.0;
            if (rd.type != 0) {
                double v = 0.0;
                if (bb.has_lo && e.c < bb.lo) v = e.c - bb.lo;
                if (bb.has_hi && e.c > bb.hi) v = e.c - bb.hi;
                t_v2 = v * v; t_v = fabs(v);
                if (resto) {
                    sig = v != 0.0 ? 1.0 : 0.0; w1 = v; y = v;
                } else {
                    const double lr = bb.has_lo ? relax_lo(bb.lo) : 0.0, hr = bb.has_hi ? relax_hi(bb.hi) : 0.0;
                    const bool upd = !S.reinit && S.pending;
                    if (S.reinit) {
                        double sv = e.c;
                        if (bb.has_lo && bb.has_hi) {
                            const double pl = fmin(DCBF_BOUND_PUSH * fmax(1.0, fabs(lr)), DCBF_BOUND_FRAC * (hr - lr));
                            const double pu = fmin(DCBF_BOUND_PUSH * fmax(1.0, fabs(hr)), DCBF_BOUND_FRAC * (hr - lr));
                            sv = fmin(fmax(sv, lr + pl), hr - pu);
                        } else if (bb.has_lo) sv = fmax(sv, lr + DCBF_BOUND_PUSH * fmax(1.0, fabs(lr)));
                        else if (bb.has_hi) sv = fmin(sv, hr - DCBF_BOUND_PUSH * fmax(1.0, fabs(hr)));
                        rs_ = sv; rzl_ = bb.has_lo ? 1.0 : 0.0; rzu_ = bb.has_hi ? 1.0 : 0.0;
                    } else if (S.pending) {
                        rs_ += sm.cold[C_ALPHA] * rds_;
                    }
                    const double rc = e.c - rs_;
                    // both sides in straight-line code (an absent side has gap 1 and multiplier 0): the lanes of a warp hold rows with a
                    // lower bound, an upper bound or both, so the two guarded blocks ran for the whole warp anyway -- plus their
                    // divergence bookkeeping
                    const double gl = bb.has_lo ? rs_ - lr : 1.0, gh = bb.has_hi ? hr - rs_ : 1.0;
                    const double il = frcp(gl), ih = frcp(gh);
                    if (upd) {   // multiplier step with the kappa_sigma safeguard (mu / gap = mu * inv)
                        const double ml = S.mu * il, mh = S.mu * ih, alpha_z = sm.cold[C_ALPHA_Z];
                        const double zl = fmax(fmin(rzl_ + alpha_z * rel_, DCBF_KAPPA_SIGMA * ml), ml * (1.0 / DCBF_KAPPA_SIGMA));
                        const double zu = fmax(fmin(rzu_ + alpha_z * reu_, DCBF_KAPPA_SIGMA * mh), mh * (1.0 / DCBF_KAPPA_SIGMA));
                        rzl_ = bb.has_lo ? zl : 0.0; rzu_ = bb.has_hi ? zu : 0.0;
                    }
                    sig = fma(rzl_, il, rzu_ * ih);
                    binv = (bb.has_lo ? il : 0.0) - (bb.has_hi ? ih : 0.0);
                    y = rzu_ - rzl_;
                    const double czl = gl * rzl_, czu = gh * rzu_;
                    t_cmin = fmin(bb.has_lo ? czl : 1e300, bb.has_hi ? czu : 1e300);
                    t_cmax = fmax(czl, czu);
                    t_z = rzl_ + rzu_;
                    const double lp = gl * gh;
                    rel_ = il; reu_ = ih;
                    if (!carry_ok) { if (ROLLED || NS == 1) t_log = flog(lp); else gap_prod *= lp; }
                    rds_ = rc;
                    t_rc = fabs(rc);
                    w1 = sig * rc;
                }
            }
            // stage the transposed row (rows beyond m stage zeros so that the dot products need no guards)
            if (!lin) M::template stage_row<NS>(sm, cs_, rd, e, sig, y, r);   // linear rows: only the four weights below are staged
            if (!lin || lane < M::NLINP) {   // (a linear slot stages its first NLINP columns only)
                DCBF_ASSERT(r >= 0 && r < RP - 2 && Ks >= 0 && Ks <= KsMax<NS>::v && m <= 32 * NS + (M::LIN2 ? 32 : 0));
                double *col = &sm.ST[2 * N][r];
                col[0] = sig; col[RP] = w1; col[2 * RP] = binv; col[3 * RP] = y;
            }
            st8.s0 += t_rc; st8.s1 += t_z; st8.s2 += t_log; st8.s3 += t_v2;
            st8.m0 = fmax(st8.m0, t_rc); st8.m1 = fmin(st8.m1, t_cmin); st8.m2 = fmax(st8.m2, t_cmax); st8.m3 = fmax(st8.m3, t_v);
            DCBF_ROW_END
        }
        if (!ROLLED && NS > 1 && !resto && !carry_ok) st8.s2 = flog(gap_prod);
        __syncwarp();   // the row branches reconverge here, before the shuffles
        reduce8_inline(st8, lane);
        const double st_theta = st8.s0, st_zsum = st8.s1, st_logsum = carry_ok ? sm.cold[C_CARRY_LOG] : st8.s2, st_v2 = st8.s3, st_pinf = st8.m0,
                     st_cmin = st8.m1, st_cmax = st8.m2, st_vmax = st8.m3;
        carry_ok = false;
        sm.cold[C_ST_THETA] = st_theta; sm.cold[C_ST_LOGSUM] = st_logsum; sm.cold[C_ST_V2] = st_v2; sm.cold[C_ST_VMAX] = st_vmax;
        // ---- second-order sources of the model --------------------------------------------------------------------------------
        M::template hess_sources<NS>(sm, P, lane, Ks, ms);
        __syncwarp();
        if (M::HAS_CURV) {
            M::template hess_curvature<NS>(sm, lane, sf_eff, sm.zc);
            __syncwarp();
        }
        // ---- condensed matrix and J^T vectors: rounds of dot products over row pairs ---------------------------------------------
#pragma unroll
        for (int t = 0; t < M::NROUND; t++) {
            const int d = dsc[t];
            if (d >= 0) {
                const double *pp = stf + (d & 0xff) * RP, *pq = stf + ((d >> 8) & 0xff) * RP;
                const int eo = d >> 20;
                DCBF_ASSERT((d & 0xff) < Sh::NST && ((d >> 8) & 0xff) < Sh::NST && eo >= 0 && eo < KQ_RHS_);
                // Fixed trip count over the whole padded row range (rows beyond m and rows of earlier steps hold zeros): the loop
                // unrolls completely.  A per-lane start row (rows of steps before the entry's class cannot contribute) executed
                // fewer FMAs but twice the instructions -- remainder ladders, divergence bookkeeping, address decoding per block.
                double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
                for (int r = 0; r < DOTC; r += 4) {
                    const double2 u = *reinterpret_cast<const double2 *>(pp + r), v = *reinterpret_cast<const double2 *>(pq + r);
                    const double2 u2 = *reinterpret_cast<const double2 *>(pp + r + 2), v2 = *reinterpret_cast<const double2 *>(pq + r + 2);
                    acc0 = fma(u.x, v.x, acc0); acc1 = fma(u.y, v.y, acc1);
                    acc2 = fma(u2.x, v2.x, acc2); acc3 = fma(u2.y, v2.y, acc3);
                }
                double acc = (acc0 + acc1) + (acc2 + acc3);
                if constexpr (M::LIN2) acc += M::template lin_term<NS>(sm, P, dlin[t]);
                if (M::round_terms(t) > 0 && eo >= KQ_K) acc += M::template hess_entry<NS>(sm, cs_, eo - KQ_K, sf_eff, M::round_terms(t), hab[t]);   // Lagrangian Hessian
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
            const double dinf = wmax(lane < N ? fabs(fma(S.sf, grad_a, q[2 * N + ln])) : 0.0);
            const double sd = fmax(100.0, fdiv(2.0 * st_zsum, sm.cold[C_NROWS_NZ])) * 0.01;
            const double sc = fmax(100.0, fdiv(st_zsum, sm.cold[C_NZ])) * 0.01;
            const double isd = frcp(sd), isc = frcp(sc);
            double E0;
            for (;;) {
                const double compm = fmax(fabs(st_cmax - S.mu), fabs(st_cmin - S.mu));
                E0 = fmax(fmax(dinf * isd, st_pinf), st_cmax * isc);
                const double Emu = fmax(fmax(dinf * isd, st_pinf), compm * isc);
                if (E0 <= tol) break;
                if (Emu <= P.kappa_eps * S.mu && S.mu > tol * 0.1 * (1.0 + 1e-12)) {
                    S.mu = fmax(tol * 0.1, fmin(DCBF_KAPPA_MU * S.mu, DCBF_MU_POW(S.mu)));
                    S.nf = 0;
                    continue;
                }
                break;
            }
            if (E0 <= tol) { S.status = 0; break; }
            if (E0 <= 1e-6 && st_vmax <= P.constr_viol_tol) { S.acc_cnt = S.acc_cnt + 1; if (S.acc_cnt >= 15) { S.status = 1; break; } } else S.acc_cnt = 0;
            if (S.iters >= P.max_iter) { S.status = -1; break; }
            if ((int)S.tiny >= P.tiny_count) {   // pinned by the fraction-to-boundary rule while still infeasible: restoration now (see ipm_iterate())
                S.tiny = 0;
                const int slot = S.nf < DCBF_FILT ? S.nf : (S.iters % DCBF_FILT);
                if (lane == 0) { sm.filt_th[slot] = (1.0 - 1e-5) * st_theta; sm.filt_ph[slot] = (S.sf * fobj - S.mu * st_logsum) - 1e-5 * st_theta; }
                if (S.nf < DCBF_FILT) S.nf = S.nf + 1;
                __syncwarp();
                S.phase = PH_RESTO; sm.cold[C_RESTO_ENTRY] = st_vmax; sm.cold[C_RESTO_TARGET] = fmax(0.1 * st_vmax, 1e-9); sm.cold[C_LM_LAMBDA] = 1e-4; S.acc_cnt = 0; S.nresto = 0;
                S.iters++;
                continue;
            }
            rhs_a = -S.sf * grad_a - q[ln] + S.mu * q[N + ln];
        } else {
            if (st_vmax <= sm.cold[C_RESTO_TARGET]) { S.phase = PH_MAIN; S.reinit = true; continue; }
            const double gn = wmax(lane < N ? fabs(q[ln]) : 0.0);
            const bool stationary = gn <= 1e-10 * fmax(1.0, st_vmax) || sm.cold[C_LM_LAMBDA] > 1e12;
            if (stationary) {
                if (st_vmax > P.constr_viol_tol) { S.status = 2; break; }
                if (sm.cold[C_RESTO_ENTRY] <= 1e-9 || S.nstall >= 1) { S.status = -2; break; }
                S.nstall = S.nstall + 1;   // see ipm_iterate()
                S.phase = PH_MAIN; S.reinit = true; continue;
            }
            if (S.iters >= P.max_iter) { S.status = -1; break; }
            rhs_a = -q[ln];
        }
        sm.KQ[KQ_RHS_ + (lane < N ? lane : N)] = rhs_a;   // packed row N of the system (lanes >= N write the dump slot)
        __syncwarp();
        // ---- restoration: Levenberg-Marquardt trials reuse the assembled K while lambda is escalated ---------------------------
        // ---- main phase: one factorisation with inertia correction by delta ----------------------------------------------------
        bool lm_accept = false;
        double v2t = 0.0, vmt = 0.0;
        double lam = resto ? sm.cold[C_LM_LAMBDA] : 0.0;   // Levenberg-Marquardt parameter of this restoration step (written back below)
        const int lbase = (lane < N + 1 ? lane : N + 1) * N, diag_ix = ln * (ln + 3) / 2;
        const double kd = sm.KQ[KQ_K + diag_ix];           // the lane's diagonal entry of the assembled matrix
        for (int rt = 0; rt < 20; rt++) {
            double shift = lam;
            bool ok = false;
            for (int tr = 0; tr < 48; tr++) {
                // lanes 0..N-1 own the rows of K, lane N the right-hand side (its "row" of the factor is L^-1 rhs).  Branch-free:
                // every lane runs the column arithmetic (idle lanes on harmless operands) and stores below the diagonal of ITS row
                // of the factor (lanes beyond N: the dump row), so the warp reaches each shuffle converged and a column costs no
                // per-lane index arithmetic.  A diagonal shift is written into the staged matrix first.
                if (shift != 0.0) {
                    sm.KQ[KQ_K + diag_ix] = kd + shift;
                    __syncwarp();
                }
                ok = true;
                if constexpr (PivotAhead<M, NS>::v) {
                    // Right-looking on the lane's own row (sv[k] = K[lane][k] minus the columns done so far, subtracted in the same order
                    // as a left-looking sweep would: identical bits).  The serial chain of a column is pivot -> shuffle -> rsqrt -> scale:
                    // the pivot of column j + 1 is formed on its own lane from that lane's own L[j+1][j] (piv), so the round trip of column j
                    // through shared memory (store, barrier, loads, trailing update of the other entries) runs beside the shuffle and the
                    // rsqrt of the next column instead of in front of them.
                    double sv[N];
#pragma unroll
                    for (int j = 0; j < N; j++) sv[j] = sm.KQ[KQ_K + rowbase + j];
                    double piv = sv[0], Lprev = 0.0;
#pragma unroll
                    for (int j = 0; j < N; j++) {
                        const double d = __shfl_sync(FULL, piv, j);
                        if (j > 0) {   // trailing update by column j - 1 (stored and synchronised at the end of the previous column)
#pragma unroll
                            for (int k = j; k < N; k++) sv[k] = fma(-Lprev, sm.Lf[k * N + (j - 1)], sv[k]);
                        }
                        if (!(d > 1e-14)) { ok = false; break; }
                        const double rinv = frsqrt(d);
                        const double Lj = sv[j] * rinv;
                        if (j + 1 < N) piv = fma(-Lj, Lj, sv[j + 1]);   // = the updated sv[j + 1] of lane j + 1 (L[j+1][j] is its own Lj)
                        if (lane > j) sm.Lf[lbase + j] = Lj;
                        sm.Lf[LF_DIAG + j] = rinv;   // diagonal as its reciprocal (every lane stores the same value)
                        __syncwarp();
                        Lprev = Lj;
                    }
                } else {   // left-looking sweep: fewer live registers; the multi-slot kernels are bound by instruction supply, not by this chain
                    double Lrow[N];
#pragma unroll
                    for (int j = 0; j < N; j++) {
                        double s_ = sm.KQ[KQ_K + rowbase + j];
#pragma unroll
                        for (int c = 0; c < j; c++) s_ = fma(-Lrow[c], sm.Lf[j * N + c], s_);
                        const double d = __shfl_sync(FULL, s_, j);
                        if (!(d > 1e-14)) { ok = false; break; }
                        const double rinv = frsqrt(d);
                        Lrow[j] = s_ * rinv;
                        if (lane > j) sm.Lf[lbase + j] = Lrow[j];
                        sm.Lf[LF_DIAG + j] = rinv;   // diagonal as its reciprocal (every lane stores the same value)
                        __syncwarp();
                    }
                }
                if (ok || resto) break;
                const double dl = sm.cold[C_DELTA_LAST];
                if (shift == 0.0) shift = dl == 0.0 ? 1e-4 : fmax(1e-20, dl * (1.0 / 3.0));
                else shift *= (dl == 0.0 ? 100.0 : 8.0);
            }
            if (!ok) {
                if (!resto) { S.status = -3; break; }
                lam *= DCBF_LM_UP;
                if (lam > 1e12) break;
                continue;
            }
            if (!resto && shift > 0.0) sm.cold[C_DELTA_LAST] = shift;
            // backward substitution: lane i reduces component i of y = L^-1 rhs (row N of the factor) by L[c][i] x_c -- zero for
            // i >= c, so there is no lane test -- and scales by its reciprocal pivot at the end
            {
                double bi = sm.Lf[N * N + ln];
#pragma unroll
                for (int c = N - 1; c >= 0; c--) {
                    const double xc = __shfl_sync(FULL, bi * sm.Lf[LF_DIAG + c], c);
                    bi = fma(-sm.Lf[c * N + ln], xc, bi);
                }
                sm.dz[lane < 15 ? lane : 15] = bi * sm.Lf[LF_DIAG + ln];
            }
            __syncwarp();
            if (!resto) break;
            // Levenberg-Marquardt trial at full step (violation only)
            sm.zt[ln] = sm.zc[ln] + sm.dz[ln];
            __syncwarp();
            M::template nodes<NS>(sm.zt, lane, wid, 0.0, false);
            v2t = 0.0; vmt = 0.0;
#pragma unroll UNR
            for (int s = 0; s < NS; s++) {
                DCBF_ROW_DESC
                if (rd.type == 0) continue;
                const RowBnd bb = ROLLED ? M::row_bounds(P, rd, leg) : rbA[ROLLED ? 0 : s];
                RowEval e;
                bool lin = false;
                if constexpr (M::LIN2) lin = s == 1;
                if (lin) { if constexpr (M::LIN2) e.c = M::eval_lin(P, rd, sm.zt); }
                else M::template eval_row<NS, false>(P, sm, rd, sm.zt, e);
                double v = 0.0;
                if (bb.has_lo && e.c < bb.lo) v = e.c - bb.lo;
                if (bb.has_hi && e.c > bb.hi) v = e.c - bb.hi;
                v2t += v * v; vmt = fmax(vmt, fabs(v));
            }
            __syncwarp();
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v2t += __shfl_xor_sync(FULL, v2t, o);
            vmt = wmax(vmt);
            if (v2t < sm.cold[C_ST_V2] * (1.0 - 1e-12)) { lm_accept = true; break; }
            lam *= DCBF_LM_UP;
            if (lam > 1e12) break;
        }
        if (S.status == -3) break;
        if (resto) {
            if (lm_accept) {
                const double dn = wmax(lane < N ? fabs(sm.dz[ln]) : 0.0);
                sm.zc[ln] = sm.zt[ln];
                S.iters++;
                lam = fmax(lam * DCBF_LM_DOWN, 1e-12);
                const double v2c = sm.cold[C_ST_V2];
                if (v2c - v2t <= 1e-4 * v2c) S.acc_cnt = S.acc_cnt < 31 ? S.acc_cnt + 1 : 31; else S.acc_cnt = 0;
                const double h1 = sm.cold[C_V2_H1], h2 = sm.cold[C_V2_H2];
                const bool crawl = S.nresto >= 2 && h2 - v2t <= P.resto_window * h2;
                __syncwarp();   // every lane has read the window before anyone shifts it
                sm.cold[C_V2_H2] = h1; sm.cold[C_V2_H1] = v2c;
                if (S.nresto < 3) S.nresto = S.nresto + 1;
                if ((dn < 1e-12 || S.acc_cnt >= 2 || crawl) && vmt > sm.cold[C_RESTO_TARGET]) lam = 1e13;
            }
            __syncwarp();   // (also: every lane has read the old lambda at the top of this step)
            sm.cold[C_LM_LAMBDA] = lam;
            continue;
        }
        // ---- direction pass (main phase): ds, dz_L, dz_U, step sizes -----------------------------------------------------------
        double amax = 1.0, az = 1.0, dphi = 0.0;
        {
            const double tau = fmax(0.99, 1.0 - S.mu);
            dphi = S.sf * grad_a * sm.dz[ln];
#pragma unroll UNR
            for (int s = 0; s < NS; s++) {
                DCBF_ROW_BEGIN
                if (rd.type == 0) continue;
                double jd = 0.0;
                bool lin = false;
                if constexpr (M::LIN2) lin = s == 1;
                if (lin) { if constexpr (M::LIN2) jd = M::eval_lin(P, rd, sm.dz); }
                else {
#pragma unroll
                    for (int a = 0; a < N; a++) jd = fma(sm.ST[a][r], sm.dz[a], jd);
                }
                const double d = jd + rds_;
                rds_ = d;
                // both sides in straight-line code (see the full pass); rel_ / reu_ hold the reciprocal gaps on entry
                const double il = rel_, ih = reu_;
                const double gl = bb.has_lo ? rs_ - relax_lo(bb.lo) : 1.0, gh = bb.has_hi ? relax_hi(bb.hi) - rs_ : 1.0;
                const double dzl = S.mu * il - rzl_ - rzl_ * il * d;
                const double dzu = S.mu * ih - rzu_ + rzu_ * ih * d;
                rel_ = dzl; reu_ = dzu;
                dphi += S.mu * d * ((bb.has_hi ? ih : 0.0) - (bb.has_lo ? il : 0.0));
                // fraction to the boundary of the slack: only the side the slack moves towards can bind, so one division serves both
                const bool down = d < 0.0;
                const double a12 = fdiv(down ? -tau * gl : tau * gh, d);
                amax = fmin(amax, (down ? bb.has_lo : (bb.has_hi && d > 0.0)) ? a12 : 1.0);
                const double z1 = fdiv(-tau * rzl_, dzl), z2 = fdiv(-tau * rzu_, dzu);
                az = fmin(az, fmin(bb.has_lo && dzl < 0.0 ? z1 : 1.0, bb.has_hi && dzu < 0.0 ? z2 : 1.0));
                DCBF_ROW_END
            }
            __syncwarp();
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dphi += __shfl_xor_sync(FULL, dphi, o);
            amax = wmin(amax); az = wmin(az);   // step sizes are positive
        }
        // ---- filter line search -------------------------------------------------------------------------------------------------
        const double theta = sm.cold[C_ST_THETA];
        const double phi = S.sf * sm.cold[C_OBJ] - S.mu * sm.cold[C_ST_LOGSUM];
        const double eps_phi = 10.0 * 2.2e-16 * fabs(phi);
        double alpha = amax;
        int accepted = 0;
        for (int ls = 0; ls < DCBF_LS_MAX; ls++, alpha *= 0.5) {
            sm.zt[ln] = fma(alpha, sm.dz[ln], sm.zc[ln]);
            __syncwarp();
            M::template nodes<NS>(sm.zt, lane, wid, S.sf, true);
            const double ft = M::template objective<NS>(sm);
            double th_t = 0.0, lg_t = 0.0, gp_t = 1.0;
            bool okv = true;
#pragma unroll UNR
            for (int s = 0; s < NS; s++) {
                DCBF_ROW_DESC
                if (rd.type == 0) continue;
                const RowBnd bb = ROLLED ? M::row_bounds(P, rd, leg) : rbA[ROLLED ? 0 : s];
                const double rs_ = ROLLED ? sm.RS[0][ROLLED ? r : 0] : rsA[ROLLED ? 0 : s], rds_ = ROLLED ? sm.RS[ROLLED ? 3 : 0][ROLLED ? r : 0] : rdsA[ROLLED ? 0 : s];
                RowEval e;
                bool lin = false;
                if constexpr (M::LIN2) lin = s == 1;
                if (lin) { if constexpr (M::LIN2) e.c = M::eval_lin(P, rd, sm.zt); }
                else M::template eval_row<NS, false>(P, sm, rd, sm.zt, e);
                const double stv = rs_ + alpha * rds_;
                th_t += fabs(e.c - stv);
                const double gl = bb.has_lo ? stv - relax_lo(bb.lo) : 1.0, gh = bb.has_hi ? relax_hi(bb.hi) - stv : 1.0;
                okv = okv && gl > 0.0 && gh > 0.0;
                if (ROLLED || NS == 1) lg_t += flog(gl * gh); else gp_t *= gl * gh;
            }
            if (!ROLLED && NS > 1) lg_t = flog(gp_t);   // (a trial with a non-positive gap is rejected by okv below, whatever this is)
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
            const bool sw = dphi < 0.0 && theta <= sm.cold[C_THETA_MIN] && switch_cond_fast(alpha, -dphi, theta);
            if (sw) { if (ph_t <= phi + 1e-8 * alpha * dphi + eps_phi) accepted = 1; }
            else if (th_t <= (1.0 - 1e-5) * theta || ph_t <= phi - 1e-5 * theta + eps_phi) accepted = 2;
            if (accepted) { sm.cold[C_CARRY_LOG] = lg_t; break; }
        }
        if (accepted != 1) {   // filter augmentation (also before entering restoration)
            const int slot = S.nf < DCBF_FILT ? S.nf : (S.iters % DCBF_FILT);
            if (lane == 0) { sm.filt_th[slot] = (1.0 - 1e-5) * theta; sm.filt_ph[slot] = phi - 1e-5 * theta; }
            if (S.nf < DCBF_FILT) S.nf = S.nf + 1;
            __syncwarp();
        }
        if (!accepted) {
            const double vm = sm.cold[C_ST_VMAX];
            S.phase = PH_RESTO; sm.cold[C_RESTO_ENTRY] = vm; sm.cold[C_RESTO_TARGET] = fmax(0.1 * vm, 1e-9); sm.cold[C_LM_LAMBDA] = 1e-4; S.acc_cnt = 0; S.nresto = 0;
            S.iters++;
            continue;
        }
        sm.zc[ln] = sm.zt[ln];
        __syncwarp();
        sm.cold[C_ALPHA] = alpha; sm.cold[C_ALPHA_Z] = az; S.pending = true;
        nodes_valid = true; carry_ok = true;   // the accepted trial staged the nodes (with Hessian terms) and the barrier sum of the new point
        if (alpha < P.tiny_alpha && sm.cold[C_ST_VMAX] > P.constr_viol_tol) S.tiny = S.tiny < 15 ? S.tiny + 1 : 15; else S.tiny = 0;
        S.iters++;
    }
    // every exit leaves the loop right after a full pass (or before any trial), so the node data are the rollout of the final iterate
}

template <class M, int NS>
__device__ __forceinline__ bool w_close(const dcbf_params &P, const WarpShared<M, NS> &sm) {
    bool close = false;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double dxg = sm.nd.nodes[i + 1][0] - sm.graw[0], dyg = sm.nd.nodes[i + 1][1] - sm.graw[1];
        if ((i == 0 || P.close_any) && sqrt(dxg * dxg + dyg * dyg) <= P.close_radius) close = true;
    }
    return close;
}

// per-CTA staging of the constants; zeroes the pad columns of the staged rows of every warp
template <class M, int NS>
__device__ __forceinline__ void stage_cta(const dcbf_params &P, const Consts &K, const WarpTables *tab, int lane, int wid) {
    constexpr int W = Wpc<M, NS>::v;
    WarpShared<M, NS> &sm = g_sm<M, NS>[wid];
    const int t0 = wid * 32 + lane;
    if (t0 == 0) { g_cs.P = P; g_cs.K = K; g_cs.tab = tab; }
    if constexpr (M::N == 9) {
        for (int t = t0; t < 60; t += 32 * W) (&g_lt.cab[0][0])[t] = __ldg(&tab->cab[0][0] + t);
        for (int t = t0; t < NHT * 48; t += 32 * W) (&g_lt.hc[0][0])[t] = __ldg(&tab->hc[0][0] + t);
        for (int t = t0; t < NHT * 48 / 4; t += 32 * W)
            reinterpret_cast<unsigned *>(&g_lt.hs[0][0])[t] = __ldg(reinterpret_cast<const unsigned *>(&tab->hs[0][0]) + t);
    }
    if constexpr (M::N == 6) { if (lane < 24) sm.nd.sm_dd[lane] = __ldg(&tab->sm_dd[lane]); }
    for (int t = lane; t < 2 * WarpShared<M, NS>::NST; t += 32) sm.ST[t >> 1][WarpShared<M, NS>::RP - 2 + (t & 1)] = 0.0;
    for (int t = lane; t < LF_LEN; t += 32) sm.Lf[t] = 0.0;   // diagonal and upper triangle of the factor stay zero
    __syncthreads();
}

// next problem of this warp: dynamic assignment from a counter (problems take 11..30+ iterations)
__device__ __forceinline__ int next_problem(int *counter, int lane) {
    int i = 0;
    if (lane == 0) i = atomicAdd(counter, 1);
    return __shfl_sync(FULL, i, 0);
}

#endif  // __CUDACC__

}  // namespace wp
}  // namespace dcbf
