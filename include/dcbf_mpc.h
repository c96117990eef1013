/*
 * dcbf_mpc.h -- C ABI of the B200-native batched D-CBF ALIP/LIP MPC solver.
 *
 * One shared object (libdcbf_mpc.so), plain pointers and sizes, no torch types.  Every entry point replaces a
 * piece of the reference planner that runs once per re-plan on the CPU (paths relative to the reference repo):
 *
 *   dcbf_default_params  <- constants hard-coded in MPCCBF.__init__ / LIP_Prob.__init__
 *                           (MPC_LIP_sig_step.py:16-44,340-353; MPC_LIP_modi.py:16-45,397-411;
 *                            MPC_DD_sig_step.py:14-40,323-338) and the Ipopt options at the call sites
 *                           (MPC_LIP_sig_step.py:266-275, MPC_LIP_modi.py:284-293, MPC_DD_sig_step.py:181-189)
 *   dcbf_set_fields      <- obstacle lists handed to the MPCCBF constructors (obs_cbf / cir_cbf, elp_cbf)
 *   dcbf_eval            <- LIP_Prob.objective / gradient / constraints / jacobian
 *                           (MPC_LIP_sig_step.py:372-496, MPC_LIP_modi.py:430-583, MPC_DD_sig_step.py:351-477)
 *                           plus the Lagrangian Hessian the reference never forms
 *   dcbf_solve           <- MPCCBF.solveMPCCBF + the plan re-roll of gen_control_test / gen_dd_control
 *                           (MPC_LIP_sig_step.py:89-111,184-278; MPC_LIP_modi.py:90-115,197-301,325-338;
 *                            MPC_DD_sig_step.py:70-99,123-193) i.e. the cyipopt.Problem(...).solve(u0) call
 *   dcbf_setup_info      <- MPCCBF.select_obs (MPC_LIP_modi.py:325-338) and the goal shift inside solveMPCCBF
 *                           (MPC_LIP_sig_step.py:229-253, MPC_LIP_modi.py:249-271), as the solve applies them
 *   dcbf_rollout         <- the plan -> apply -> re-plan loop of MPC_LIP_sig_step.py:565-575
 *   dcbf_solve_host      <- same as dcbf_solve for callers that hold host (numpy) buffers
 *   dcbf_alip_foot       <- the closed-form ALIP foot placement behind the DD re-plan (Logger.ALIP_gen_foot_input,
 *                           data_procs/logger_dd.py:356-363 -> ALIP.AMprediction / computeSw2CoM / computeStepping /
 *                           regulate_lateral_step / getTimedState, ALIP_plan/planner.py:188-261,346-370)
 *   dcbf_veldes_foot     <- MPCCBF.alip_des_vel + MPCCBF.cal_foot_with_veldes (MPC_LIP_sig_step.py:168-181), the velocity-tracking
 *                           foothold Logger.cal_foot_input places between re-plans (data_procs/logger.py:380-418)
 *   dcbf_heading_input   <- Logger.tube_func + Logger.avg_hd, the heading-rate input of that prediction
 *                           (data_procs/logger_mpc.py:208-215,278-300)
 *   dcbf_gen_fields /    <- rand_obs.gen_ran_obs_list (rand_obs.py:31-81) and the start state of the __main__ loop
 *   dcbf_gen_states         (MPC_LIP_sig_step.py:553-568), batched, for the 1 M-scenario configuration
 *   dcbf_tick            <- one control tick of Logger.gen_nex_foot_input (data_procs/logger_mpc.py:318-341):
 *                           LIP prediction to the end of the running step (MPCCBF.get_next_states,
 *                           MPC_LIP_modi.py:149-178), the warm-start rule, the re-plan, and the dense plan trajectory
 *                           pos_det of gen_control_test (MPC_LIP_modi.py:117-122, xk_track_det :304-322)
 *
 * Conventions
 *   - all floating point data is FP64, row-major, densely packed; index data is int32
 *   - "device" pointers must be valid on the context's device; "host" pointers are ordinary host memory
 *   - work is enqueued on the CUDA stream passed as `void* stream` (a cudaStream_t; NULL = default stream);
 *     device-pointer calls do not synchronise
 *   - return value: 0 on success, negative dcbf_status on error; nothing throws; there is no CPU fallback
 *   - plan variables use the reference's layouts: state x = (px, py, vx, vy, theta) [DD: (x, y, theta)],
 *     foot/turn plan p_k = (foot_x, foot_y, dtheta) (= W(u_k - A x_k), MPC_LIP_sig_step.py:302-306),
 *     decision vector u in R^15 with the representative u_k := x_{k+1} (DD: u = (v0, w0, v1, w1, v2, w2))
 *   - solver status keeps Ipopt's integers: 0 solved, 1 acceptable, 2 infeasible problem detected,
 *     -1 iteration cap, -2 restoration failed, -3 step computation failed, -13 invalid number
 */
#ifndef DCBF_MPC_H
#define DCBF_MPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCBF_ABI_VERSION 4
#define DCBF_MAX_OBS 16 /* circles and ellipses each, per field */

enum dcbf_formulation { DCBF_SIG_STEP = 0, DCBF_MODI = 1, DCBF_DD = 2 };

enum dcbf_status {
    DCBF_OK = 0,
    DCBF_ERR_ARG = -1,      /* NULL / out-of-range argument */
    DCBF_ERR_CUDA = -2,     /* a CUDA runtime call failed (see dcbf_last_error) */
    DCBF_ERR_NO_FIELDS = -3 /* dcbf_set_fields has not been called */
};

typedef struct dcbf_params {
    int32_t formulation; /* dcbf_formulation */
    int32_t max_iter;    /* interior-point iteration cap */
    int32_t select_obs;  /* 1: keep obstacles with dist^2 - r^2 <= detect_sq only (MPC_LIP_modi.py:325-338) */
    int32_t goal_shift;  /* 1: 15-degree detour heuristic (MPC_LIP_sig_step.py:229-253) */
    int32_t has_fen;     /* 1: speed/turn coupling row s*|dtheta| + v (MPC_LIP_modi.py:493) */
    int32_t close_any;   /* 1: close_2_goal if any step is inside close_radius (sig_step), 0: first step only */
    int32_t tiny_count;  /* this many accepted steps in a row shorter than tiny_alpha, with rows still violated, send the iterate to the
                            restoration phase (the fraction-to-boundary rule is pinning it: typical for infeasible problems) */
    int32_t reserved1;
    double w_p, w_q, w_r, w_t;                                        /* cost weights p, q, r, t */
    double gamma, s_turn;                                             /* D-CBF decay, turn coupling */
    double bvx_min, bvx_max, bvy_min, bvy_max, leg_sq, ang_max;       /* row bounds */
    double detect_sq, close_radius;
    double tol, constr_viol_tol, mu_init;                             /* Ipopt: tol, constr_viol_tol, mu_init */
    double tiny_alpha;                                                /* see tiny_count */
    double mu_warm;   /* first barrier parameter of a re-plan warm-started from the previous plan verbatim (dcbf_tick mode 0) */
    double mu_shift;  /* ... from the shifted plan [x_2, x_3, x_3] (dcbf_tick mode 1, dcbf_rollout after the first step), whose third
                         step still violates the velocity rows; a cold start uses mu_init */
    double resto_window; /* restoration phase: three consecutive accepted steps that together reduce the squared violation by less than
                            this fraction end the phase at the current point (a stationary point of the violation for the purpose of
                            the infeasibility verdict; the iterate crawls along a kink of the violation).  ABI version 4 */
    double kappa_eps;    /* barrier tolerance factor (Ipopt barrier_tol_factor, default 10 there): the barrier parameter is lowered once
                            the error of the barrier problem is below kappa_eps * mu.  ABI version 4 */
} dcbf_params;

typedef struct dcbf_ctx dcbf_ctx;

int dcbf_abi_version(void);
int dcbf_default_params(int formulation, dcbf_params *out);
int dcbf_create(const dcbf_params *params, int device, dcbf_ctx **out);
void dcbf_destroy(dcbf_ctx *ctx);
const char *dcbf_last_error(const dcbf_ctx *ctx);

/* Obstacle fields shared by the scenarios of a batch (device pointers, copied into the context):
 * cir[F][Kc][3] = (cx, cy, r) and elp[F][Ke][5] = (cx, cy, a, b, phi), already inflated by the safety margin. */
int dcbf_set_fields(dcbf_ctx *ctx, int32_t F, int32_t Kc, const double *cir_dev, int32_t Ke, const double *elp_dev,
                    void *stream);

/* Rows of one NLP in the reference order.  m = 3*(4+Kc) sig_step, 3*(5+Kc+Ke) modi, 3*(Kc+Ke+1) dd; n = 9 (dd: 6).
 * No obstacle selection is applied here (every obstacle of the field is a row). */
int dcbf_num_rows(const dcbf_ctx *ctx);
int dcbf_num_vars(const dcbf_ctx *ctx);

/* K1: callbacks at a point z[B][n] of the reduced space (z = (p0, p1, p2); dd: z = u).
 * goal[B][2] is used as given (no goal shift).  Outputs (any may be NULL): f[B], grad[B][n], c[B][m],
 * jac[B][m][n], cl[B][m], cu[B][m], hess[B][n][n] = Hessian of  f + sum_r lambda_r c_r  (lambda[B][m], NULL => 0). */
int dcbf_eval(dcbf_ctx *ctx, int32_t B, const double *x0, const double *goal, const int32_t *leg, const int32_t *field,
              const double *last_u, const double *z, const double *lambda, double *f, double *grad, double *c,
              double *jac, double *cl, double *cu, double *hess, void *stream);

/* K1+K2: one re-plan per scenario.  Inputs: x0[B][5|3], goal[B][2], leg[B] (+1/-1; dd: ignored, may be NULL),
 * field[B] (NULL => field 0 for every scenario), warm[B][15|6] = the reference's u0 (MPC_LIP_sig_step.py:185-189
 * builds it from the previous plan; the caller passes the final vector; LIP formulations: NULL => [x_k, x_k, x_k], the
 * reference's start vector for init_guess = None, formed on the device -- 120 bytes per scenario less to read, which is
 * most of what a host-buffer call moves over PCIe), last_u[B][2] (dd only).
 * Outputs (any may be NULL): u[B][15|6], x_plan[B][3][5|3], p_plan[B][3][3] (LIP only), status[B], iters[B],
 * obj[B], viol[B] (max row violation), close2goal[B] (uint8). */
int dcbf_solve(dcbf_ctx *ctx, int32_t B, const double *x0, const double *goal, const int32_t *leg, const int32_t *field,
               const double *warm, const double *last_u, double *u, double *x_plan, double *p_plan, int32_t *status,
               int32_t *iters, double *obj, double *viol, uint8_t *close2goal, void *stream);

/* The per-scenario problem setup that precedes every solve, made visible (device pointers): obstacle selection
 * (MPCCBF.select_obs, MPC_LIP_modi.py:325-338: keep an obstacle if dist^2 - r^2 <= detect_sq, r = max(a, b) for an ellipse)
 * and the detour goal (MPC_LIP_sig_step.py:229-253; MPC_LIP_modi.py:249-271 searches the selected circles only).
 * Outputs (any may be NULL): mask[B] -- bit j set = obstacle j of the scenario's field is a row of its NLP (circles 0..Kc-1,
 * then ellipses Kc..Kc+Ke-1; all ones without select_obs), count[B] = number of selected obstacles, goal_eff[B][2] = the goal the
 * NLP is solved with.  It runs the same setup code as dcbf_solve (of the kernel family DCBF_KERNEL selects). */
int dcbf_setup_info(dcbf_ctx *ctx, int32_t B, const double *x0, const double *goal, const int32_t *field, uint32_t *mask,
                    int32_t *count, double *goal_eff, void *stream);

/* K3: closed loop of `steps` re-plans per scenario without leaving the GPU (LIP formulations).  Each step:
 * solve, apply the first foot placement exactly (x <- x_plan[0]), flip the stance leg, warm start from the shifted
 * plan [x_2, x_3, x_3]; stop early on close_2_goal.  Outputs (any may be NULL): x_final[B][5], steps_done[B],
 * n_infeasible[B] (re-plans that ended with status 2), total_iters[B],
 * traj[B][steps][8] = (px, py, vx, vy, theta, foot_x, foot_y, status) after each step (NaN once stopped). */
int dcbf_rollout(dcbf_ctx *ctx, int32_t B, int32_t steps, const double *x0, const double *goal, const int32_t *leg,
                 const int32_t *field, double *x_final, int32_t *steps_done, int32_t *n_infeasible,
                 int32_t *total_iters, double *traj, void *stream);

/* dcbf_solve_host without the final wait, for callers with a stream of batches (ABI version 4): the work is enqueued on the
 * context's own stream and the call returns; dcbf_wait(ctx) returns when every call enqueued on the context has finished and its
 * results are in the caller's buffers.  Requires page-locked buffers on both sides (the copy-free path of dcbf_solve_host; DCBF_ERR_ARG
 * otherwise).  The buffers must stay untouched until dcbf_wait.  Two or three contexts used round-robin overlap the drain of one
 * batch with the head of the next (what the reference's per-tick loop over robots, main_sim_mpc.py:85-88, would batch). */
int dcbf_solve_host_async(dcbf_ctx *ctx, int32_t B, const double *x0, const double *goal, const int32_t *leg,
                          const int32_t *field, const double *warm, const double *last_u, double *u, double *x_plan,
                          double *p_plan, int32_t *status, int32_t *iters, double *obj, double *viol, uint8_t *close2goal);
int dcbf_wait(dcbf_ctx *ctx);

/* dcbf_solve for host buffers; returns when the results are in the caller's buffers.  Pageable buffers are staged through one
 * pinned block (one copy each way).  If every buffer is page-locked (cudaHostAlloc / cudaHostRegister) and the batch runs on a
 * warp kernel, nothing is copied: the kernels read the inputs from and write the results to the mapped host memory directly
 * (environment DCBF_ZEROCOPY=0 restores cudaMemcpyAsync).  Obstacle fields still come from dcbf_set_fields_host / dcbf_set_fields. */
int dcbf_set_fields_host(dcbf_ctx *ctx, int32_t F, int32_t Kc, const double *cir_host, int32_t Ke, const double *elp_host);
int dcbf_solve_host(dcbf_ctx *ctx, int32_t B, const double *x0, const double *goal, const int32_t *leg,
                    const int32_t *field, const double *warm, const double *last_u, double *u, double *x_plan,
                    double *p_plan, int32_t *status, int32_t *iters, double *obj, double *viol, uint8_t *close2goal);

/* One control tick per scenario (LIP formulations), everything on the device:
 *   x_next = A(t_rest) [pos, vel, hd] + B(t_rest) glo_p          LIP flow to the end of the running step; the heading row
 *                                                               of B is t_rest / dt (MPC_LIP_modi.py:149-178)
 *   warm   = prev_plan verbatim (mode 0), shifted [x_2, x_3, x_3] (mode 1: first tick of a step, logger_mpc.py:329-331)
 *            or [x_next, x_next, x_next] (mode 2: no previous plan, logger_mpc.py:326-327); mode NULL => 2 everywhere
 *   re-plan from x_next with the swing leg's sign leg[B] (the caller passes -leg_ind like logger_mpc.py:336)
 *   pos_det[B][126][2]: per planned step the start position and the LIP flow at t = 0, 0.01, ..., 0.40 s
 * Inputs: glo_pos[B][2], glo_vel[B][2], glo_hd[B], glo_p[B][3] = (stance foot x, y, heading input), t_rest[B],
 * goal[B][2], leg[B], field[B]|NULL, prev_plan[B][15]|NULL, mode[B]|NULL.  Outputs (any may be NULL): x_next[B][5],
 * warm[B][15] (the start vector that was used), then the outputs of dcbf_solve, then pos_det. */
int dcbf_tick(dcbf_ctx *ctx, int32_t B, const double *glo_pos, const double *glo_vel, const double *glo_hd,
              const double *glo_p, const double *t_rest, const double *goal, const int32_t *leg, const int32_t *field,
              const double *prev_plan, const uint8_t *mode, double *x_next, double *warm, double *u, double *x_plan,
              double *p_plan, int32_t *status, int32_t *iters, double *obj, double *viol, uint8_t *close2goal,
              double *pos_det, void *stream);

/* Angular-momentum LIP one-step foot placement for B scenarios (device pointers), the step behind the DD re-plan in
 * data_procs/logger_dd.py:356-363.  Inputs: x_alip[B][2] = (p_x, L_y) and y_alip[B][2] = (p_y, L_x) relative to the stance foot,
 * time[B] into the running step, support[B] (+1 right, -1 left), speed = desired forward speed read as speed[b * speed_stride]
 * (pass the `u` output of dcbf_solve of the DD formulation with speed_stride = 6 to chain the two on the stream).  Model constants:
 * H (CoM height), T (step time), m (mass), W (step width) -- ALIPParam / ALIP.__init__ (ALIP_plan/planner.py:15-61, 545).
 * Outputs (any may be NULL): foot[B][2] = (px_sp2sw, py_sp2sw) with the lateral regulation of planner.py:346-370,
 * am[B][2] = (Ly_est, Lx_est) (planner.py:210-230), next[B][4] = (p_x, L_y, p_y, L_x) at the end of the step
 * (getTimedState over T - time, planner.py:188-208).  Moving-platform (DRS) terms are zero (planner.py:47-50). */
int dcbf_alip_foot(dcbf_ctx *ctx, int32_t B, const double *x_alip, const double *y_alip, const double *time, const int32_t *support,
                   const double *speed, int32_t speed_stride, double H, double T, double m, double W, double *foot, double *am,
                   double *next, void *stream);

/* Velocity-tracking foot placement used between re-plans (LIP formulations; MPCCBF.alip_des_vel and MPCCBF.cal_foot_with_veldes,
 * MPC_LIP_sig_step.py:168-181, called by Logger.cal_foot_input, data_procs/logger.py:380-418).  Per scenario:
 *   vel_des = (sigma vx_max dt / 2, 0.5 (-0.5 leg step_gap) beta sinh(beta dt) / (cosh(beta dt) + 1)), sigma = beta coth(beta dt / 2)
 *             -- or vel_des_in[B][2] when that is not NULL (then leg may be NULL);
 *   foot    = B_vel^-1 (vel_des - (A x_state)[2:4]): the foothold after which the LIP step ends with velocity vel_des.
 * x_state[B][5] (may be NULL if foot is NULL), leg[B] = leg_ind (+1 / -1), step_gap = 0.3 in the reference (MPC_LIP_sig_step.py:41).
 * Outputs (any may be NULL): vel_des_out[B][2], foot[B][2]. */
int dcbf_veldes_foot(dcbf_ctx *ctx, int32_t B, const double *x_state, const int32_t *leg, const double *vel_des_in, double vx_max,
                     double step_gap, double *vel_des_out, double *foot, void *stream);

/* Heading input of the LIP prediction, the per-tick tail of Logger.update_n_record (data_procs/logger_mpc.py:278-281):
 *   nex_turn <- tube_func(nex_turn, cur_hd)      the last plan's turn, scaled by 0.4 inside the +-0.15 rad tube and 0.7 outside
 *                                                (logger_mpc.py:284-300), wrapped by angle_A_minus_B (:169-175)
 *   hd_input <- avg_hd(cur_hd)                   (nex_turn + the three heading increments of the last plan, the first taken
 *                                                from cur_hd) / 4   (logger_mpc.py:208-215)
 * cur_hd[B]; nex_turn[B] is read and overwritten; the plan headings of scenario b are mpc_hds[b*hds_stride + k*hds_step],
 * k = 0..2 (pass x_plan + 4 with stride 15 and step 5 to read them from the last dcbf_tick / dcbf_solve output); the result
 * goes to hd_input[b*out_stride] (pass glo_p + 2 with stride 3 to write the third input column of dcbf_tick in place).
 * After a re-plan the caller resets nex_turn from the new plan's first turn p_plan[b][0][2] (logger_mpc.py:341). */
int dcbf_heading_input(dcbf_ctx *ctx, int32_t B, const double *cur_hd, double *nex_turn, const double *mpc_hds,
                       int32_t hds_stride, int32_t hds_step, double *hd_input, int32_t out_stride, void *stream);

/* Scenario generation on the device (rand_obs.py:31-81: random_circle / random_obs / gen_ran_obs_list restated for F fields
 * at once).  Circles (x, y, r) with x, y in [0, margin), r in [0.35, radius], all rounded to two decimals, are rejection
 * sampled against the keep-out discs (10, 10, 0.3) and (0, 0, 1.0) and against each other until `num` of them satisfy
 * |c_i - c_j| >= r_i + r_j + 2 half_gap (rand_obs.py:31-54); with mix != 0 every odd-numbered one becomes an ellipse
 * (a = r, b in [a/2, a), phi a whole number of degrees; rand_obs.py:57-72).  Unlike the reference loop this one terminates:
 * a field that stalls for 2000 draws is restarted, at most 64 times; draws[F] (may be NULL) receives the candidates drawn
 * or -1 for a field that could not be built (its obstacles are NaN).  The draws are Philox4x32-10 outputs keyed by
 * (seed, field, draw number): the batch does not depend on the launch geometry and oracle/scenario_gen.py reproduces it
 * bit for bit.  Outputs: cir[F][Kc][3], elp[F][Ke][5] with Kc = num, Ke = 0 (mix = 0) or Kc = ceil(num/2), Ke = floor(num/2);
 * radii and semi-axes already inflated by safe_dis -- the layout dcbf_set_fields takes.  1 <= num <= 32. */
int dcbf_gen_fields(dcbf_ctx *ctx, int32_t F, uint64_t seed, int32_t num, int32_t mix, double margin, double radius,
                    double half_gap, double safe_dis, double *cir, double *elp, int32_t *draws, void *stream);

/* Start states for B scenarios on the context's current fields (SURVEY.md 8(d) distribution, the batched form of the
 * hand-written start of MPC_LIP_sig_step.py:553-568): position uniform in [0, 8)^2, redrawn (at most 64 times) until every
 * obstacle's level set is >= 0.05 there; heading = bearing to the goal + U(-0.3, 0.3); leg = +-1; body velocity
 * vx in [0.4, 0.8), |vy| in [0.15, bvy_max) with the sign of -leg (bvy_max <= 0: the formulation's bound).  field[B] may be
 * NULL (field 0).  Outputs (any may be NULL): x0[B][5|3], goal[B][2], leg[B], warm[B][15|6] (cold start [x0, x0, x0]; dd:
 * (0.8, 0) x 3), last_u[B][2] (dd), attempts[B] = positions tried, -1 if none was clear (x0 is NaN then). */
int dcbf_gen_states(dcbf_ctx *ctx, int32_t B, uint64_t seed, const int32_t *field, double goal_x, double goal_y,
                    double bvy_max, double *x0, double *goal, int32_t *leg, double *warm, double *last_u,
                    int32_t *attempts, void *stream);

/* Test hook: the lean FP64 elementary functions of the kernels (csrc/dcbf_math.cuh) evaluated on the device.
 * out[n][7] = (sin a, cos a, atan2(a, b), 1 / b, a / b, 1 / sqrt(|b|), log |b|); a, b, out are device pointers. */
int dcbf_math_probe(dcbf_ctx *ctx, int32_t n, const double *a, const double *b, double *out, void *stream);

/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t dcbf_launch_count(const dcbf_ctx *ctx);

/* FP64 FMA microbenchmark on the context's device: returns achieved TFLOP/s (2 flop per DFMA), <0 on error. */
double dcbf_fp64_peak_tflops(dcbf_ctx *ctx, int32_t repeats);

#ifdef __cplusplus
}
#endif
#endif /* DCBF_MPC_H */
