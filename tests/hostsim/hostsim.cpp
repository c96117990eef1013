// hostsim.cpp -- TEST-ONLY host build of the per-lane solver code (csrc/dcbf_core.cuh, dcbf_lanes.cuh).
//
// There is no GPU in the build container; this file compiles the exact lane functions the CUDA kernels call with g++
// so that the algorithm can be debugged and checked against the oracle in the `-m "not gpu"` tests.  It is NOT part
// of the product: nothing in the package loads it, the C ABI (libdcbf_mpc.so) has no CPU path, and no benchmark
// number is ever taken from it.
#include <cstdlib>
#include <vector>

#include "../../mujoco_lip_mpc_simulation_b200/csrc/dcbf_lanes.cuh"
#include "../../mujoco_lip_mpc_simulation_b200/csrc/dcbf_warp.cuh"   // host part only: the constant tables of the warp kernels
#include "../../mujoco_lip_mpc_simulation_b200/csrc/dcbf_gen.cuh"    // scenario generator (same functions the kernels call)

using namespace dcbf;

static void prep(int F, int Kc, const double *cir, int Ke, const double *elp, std::vector<double> &cr, std::vector<double> &er) {
    cr.assign((size_t)DCBF_CIR_REC * F * (Kc > 0 ? Kc : 1), 0.0);
    er.assign((size_t)DCBF_ELP_REC * F * (Ke > 0 ? Ke : 1), 0.0);
    for (int t = 0; t < F * Kc; t++) prep_circle(cir + 3 * (size_t)t, cr.data() + DCBF_CIR_REC * (size_t)t);
    for (int t = 0; t < F * Ke; t++) prep_ellipse(elp + 5 * (size_t)t, er.data() + DCBF_ELP_REC * (size_t)t);
}

extern "C" {

int hostsim_solve(const dcbf_params *P, int B, const double *x0, const double *goal, const int32_t *leg, const int32_t *field,
                  int F, int Kc, const double *cir, int Ke, const double *elp, const double *warm, const double *last_u,
                  double *u, double *x_plan, double *p_plan, int32_t *status, int32_t *iters, double *obj, double *viol,
                  uint8_t *close2goal) {
    std::vector<double> cr, er;
    prep(F, Kc, cir, Ke, elp, cr, er);
    const Consts K = make_consts();
    BatchIn in = {x0, goal, warm, last_u, leg, field, cr.data(), er.data(), Kc, Ke};
    SolveOut out = {u, x_plan, p_plan, obj, viol, status, iters, close2goal};
    for (int b = 0; b < B; b++) {
        if (P->formulation == DCBF_DD) solve_dd_lane(*P, K, in, out, b);
        else solve_lip_lane(*P, K, in, out, b);
    }
    return 0;
}

int hostsim_eval(const dcbf_params *P, int B, const double *x0, const double *goal, const int32_t *leg, const int32_t *field,
                 int F, int Kc, const double *cir, int Ke, const double *elp, const double *last_u, const double *z,
                 const double *lambda, int m, double *f, double *grad, double *c, double *jac, double *cl, double *cu, double *hess) {
    std::vector<double> cr, er;
    prep(F, Kc, cir, Ke, elp, cr, er);
    const Consts K = make_consts();
    BatchIn in = {x0, goal, nullptr, last_u, leg, field, cr.data(), er.data(), Kc, Ke};
    EvalPtrs ev = {z, lambda, f, grad, c, jac, cl, cu, hess, m};
    for (int b = 0; b < B; b++) {
        if (P->formulation == DCBF_DD) eval_dd_lane(*P, K, in, ev, b);
        else eval_lip_lane(*P, K, in, ev, b);
    }
    return 0;
}

int hostsim_rollout(const dcbf_params *P, int B, int steps, const double *x0, const double *goal, const int32_t *leg,
                    const int32_t *field, int F, int Kc, const double *cir, int Ke, const double *elp, double *x_final,
                    int32_t *steps_done, int32_t *n_infeasible, int32_t *total_iters, double *traj) {
    std::vector<double> cr, er;
    prep(F, Kc, cir, Ke, elp, cr, er);
    const Consts K = make_consts();
    BatchIn in = {x0, goal, nullptr, nullptr, leg, field, cr.data(), er.data(), Kc, Ke};
    RolloutOut out = {x_final, traj, steps_done, n_infeasible, total_iters};
    for (int b = 0; b < B; b++) rollout_lip_lane(*P, K, in, out, steps, b);
    return 0;
}

// lean elementary functions of the kernels (dcbf_math.cuh), evaluated on the host for the accuracy test
int hostsim_math(int n, const double *a, double *sn, double *cs, const double *y, const double *x, double *at) {
    for (int i = 0; i < n; i++) { fsincos(a[i], sn + i, cs + i); at[i] = fatan2(y[i], x[i]); }   // frcp / fdiv / frsqrt are plain divisions on the host
    return 0;
}

int hostsim_log(int n, const double *x, double *out) {
    for (int i = 0; i < n; i++) out[i] = flog(x[i]);
    return 0;
}

// constant tables of the warp kernels + the dense feature map they are derived from
int hostsim_warp_tables(int *desc, double *hc, int *hs, double *cab, double *T) {
    const Consts K = make_consts();
    wp::WarpTables W;
    if (!wp::build_warp_tables(K, W)) return -1;
    for (int t = 0; t < 96; t++) desc[t] = W.desc[t];
    for (int t = 0; t < wp::NHT; t++) for (int e = 0; e < 48; e++) { hc[48 * t + e] = W.hc[t][e]; hs[48 * t + e] = W.hs[t][e]; }
    for (int c = 0; c < 10; c++) for (int j = 0; j < 6; j++) cab[6 * c + j] = W.cab[c][j];
    double Tm[24][9];
    wp::build_feature_map(K, Tm);
    for (int f = 0; f < 24; f++) for (int v = 0; v < 9; v++) T[9 * f + v] = Tm[f][v];
    return 0;
}

// closed-form descriptors of the typed linear slots (wp::LipL turn rows, wp::DdL linear rows)
int hostsim_lin_descriptors(int *lin_lip, int *lin_dd, int *desc_dd) {
    const Consts K = make_consts();
    wp::WarpTables W;
    if (!wp::build_warp_tables(K, W)) return -1;
    for (int t = 0; t < 96; t++) lin_lip[t] = W.desc_lin_lip[t];
    for (int t = 0; t < 64; t++) { lin_dd[t] = W.desc_lin[t]; desc_dd[t] = W.desc_dd[t]; }
    return 0;
}

// scenario generator of csrc/dcbf_gen.cuh on the host (the kernels call the same two functions)
int hostsim_philox(uint32_t *c, uint32_t k0, uint32_t k1) { gen::philox4x32_10(c, k0, k1); return 0; }

int hostsim_gen_fields(int F, uint64_t seed, int num, int mix, double margin, double radius, double half_gap, double safe_dis,
                       int stall, int max_restarts, double *cir, double *elp, int32_t *draws) {
    gen::FieldSpec S = {num, mix, margin, radius, half_gap, safe_dis, stall, max_restarts};
    const int Kc = mix ? (num + 1) / 2 : num, Ke = mix ? num / 2 : 0;
    for (int f = 0; f < F; f++) draws[f] = gen::make_field(S, seed, (uint32_t)f, cir + 3 * (size_t)Kc * f, elp + 5 * (size_t)Ke * f);
    return 0;
}

int hostsim_gen_states(int B, uint64_t seed, int dd, double gx, double gy, double bvy_max, const int32_t *field, int F, int Kc,
                       const double *cir, int Ke, const double *elp, double *x0, double *goal, int32_t *leg, double *warm,
                       double *last_u, int32_t *attempts) {
    std::vector<double> cr, er;
    prep(F, Kc, cir, Ke, elp, cr, er);
    gen::StateSpec S = {dd, gx, gy, 8.0, 0.05, 0.3, 0.4, 0.8, 0.15, bvy_max, 64};
    const int nx = dd ? 3 : 5, nw = dd ? 6 : 15;
    for (int b = 0; b < B; b++) {
        const int fld = field ? field[b] : 0;
        attempts[b] = gen::make_state(S, seed, (uint32_t)b, cr.data() + (size_t)fld * Kc * DCBF_CIR_REC, Kc, DCBF_CIR_REC,
                                      er.data() + (size_t)fld * Ke * DCBF_ELP_REC, Ke, DCBF_ELP_REC, x0 + (size_t)nx * b, goal + 2 * (size_t)b,
                                      leg + b, warm + (size_t)nw * b, dd ? last_u + 2 * (size_t)b : nullptr,
                                      [](double a, double *s, double *c) { *s = sin(a); *c = cos(a); },
                                      [](double y, double x) { return atan2(y, x); });
    }
    return 0;
}

#ifdef DCBF_COUNT
long hostsim_trials() { return dcbf::g_trials; }
#endif

}  // extern "C"
