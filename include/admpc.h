/*
 * admpc.h -- C ABI of libadmpc_b200.so: batched SQP-RTI NMPC solver for the AD_MPC bicycle model on NVIDIA B200.
 *
 * Drop-in boundary.  Paths cited below are under
 *   /root/reference/data_driven_mpc/ros_gp_mpc/src/ad_mpc/            ($A)
 *   /root/reference/data_driven_mpc/ros_gp_mpc/src/ad_mpc/c_generated_code/   ($G)
 *
 * Two layers are exported:
 *   1. `sim_car_acados_*`  -- the acados-generated solver shim's lifecycle symbols, same names / argument meaning /
 *      return convention as $G/acados_solver_sim_car.h:116-150, plus flat string-keyed set/get that replace the
 *      libacados calls `ocp_nlp_cost_model_set / ocp_nlp_constraints_model_set / ocp_nlp_out_get / ocp_nlp_get`
 *      the reference reaches through acados_template ($A/ad_3d_optimizer.py:430,438,441-442,450,456,462-465;
 *      $G/main_sim_car.c:126-128,171-187,207-208).  One capsule == one MPC instance (a batch of 1 on the GPU).
 *   2. `admpc_batch_*`     -- the batched twin: B independent instances solved by one call.
 *
 * Plain pointers and sizes only.  All setters copy from caller memory, all getters copy into caller memory
 * ($G/main_sim_car.c:184-187 ownership rule).  Return value 0 == success everywhere; solve returns the acados
 * status {0 success, 1 failure(NaN), 2 maxiter, 3 minstep, 4 QP failure} ($G/main_sim_car.c:190-197).  Negative
 * values are library errors (ADMPC_E_*).  There is NO CPU fallback: every entry point that needs the device fails
 * with ADMPC_E_CUDA when no CUDA device is usable.
 */
#ifndef ADMPC_H_
#define ADMPC_H_

#ifdef __cplusplus
extern "C" {
#endif

#define ADMPC_NX 7
#define ADMPC_NU 2
#define ADMPC_NP 1
#define ADMPC_NY 9
#define ADMPC_NYN 7
#define ADMPC_NC 10        /* per-stage inequality rows [lbu0 lbu1 lbx | ubu0 ubu1 ubx | ls0 ls1 | us0 us1] */
#define ADMPC_NMAX 128
#define ADMPC_DZMAX 8
#define ADMPC_GPOUT_MAX 2

#define ADMPC_E_ARG (-1)       /* bad argument / unknown field / wrong length */
#define ADMPC_E_CUDA (-2)      /* CUDA runtime error (message via admpc_last_error) */
#define ADMPC_E_STATE (-3)     /* call order violated (e.g. solve before create) */
#define ADMPC_E_NCCL (-4)      /* NCCL unavailable or failed */
#define ADMPC_E_UNSUPPORTED (-5)

/* Problem + solver options.  Defaults (admpc_default_opts) reproduce the shipped configuration:
 * $G/acados_solver_sim_car.c:362 (dt), :393-399 (W), :481-485 (W_e), :455-473 (z,Z), :549-552 (lbu,ubu),
 * :595-596 (lbx,ubx), :657-665 (ERK 4 stages x 1 step), :688-693 (HPIPM BALANCE, iter_max 50);
 * $A/ad_3d.py:47-60 (vehicle). */
typedef struct admpc_opts {
    int N;                 /* shooting intervals (20 = $A/ad_3d_mpc.py:23-24 default; 40 = generated code) */
    int iter_max;          /* QP iteration limit */
    int gp_enabled;        /* set by admpc_batch_set_gp */
    int gp_nout, gp_M, gp_dz;
    int gp_stage0_trigger; /* 1: stage 0 evaluates the GP at gp_state (quad_mpc/quad_3d_optimizer.py:295,548-552) */
    int model_variant;     /* 0: Cartesian-pose model ($A/ad_3d_optimizer.py:268-310, shipped); 1: Frenet-frame variant
                              ($A/__pycache__/fren_ad_3d_optimizer.cpython-36.pyc; x = [s, e_y, e_psi, v_x, v_y, r, delta],
                              path curvature per shooting node via admpc_batch_set_kappa) */
    int gp_feat[ADMPC_DZMAX];      /* feature indices into [x(7);u(2)] (B_z, model_fitting/gp.py:609-630); >= 2 */
    int gp_row[ADMPC_GPOUT_MAX];   /* state rows receiving the GP outputs (B_x, utils/utils.py:773-786); in {3,4,5} */
    int gp_precision;      /* 0 (default): FP64 RBF kernel values; 1: opt-in FP32 exponential (ex2.approx on the fractional
                              part of the FP64 exponent, FP64 accumulation): relative error of every kernel value <= 2^-22,
                              see DESIGN.md 5 for the bound on A, B, b (model_fitting/gp.py:117-138) */
    int con_set;           /* inequality set: 0 = u0, u1 soft + steering angle hard ($A/ad_3d_optimizer.py:165-199, 10 rows per stage);
                              1 = the Frenet variant's own set (fren_ad_3d_optimizer pyc; structure pinned by $A/debug.json: 12
                              multipliers, 2 + 2 slacks per stage): u0 soft, u1 hard, e_y = x[1] in [lbx2, ubx2] hard, steering
                              angle soft with the penalties zl[1], zu[1].  Rows [lbu0 lbu1 lbx_ey lbx_delta | ub.. | ls0 ls1 |
                              us0 us1].  Needs model_variant = 1; solved by the dense thread-per-instance kernel */
    double dt;
    double W[9], We[7];
    double zl[2], zu[2], Zl[2], Zu[2];
    double lbu[2], ubu[2], lbx, ubx;
    double lbx2, ubx2;     /* con_set = 1: hard bound on e_y, stages 1..N-1 (default -2 / 2) */
    double mass, lf, lr, iz, cf2, cr2;
    double mu0, tol_stat, tol_eq, tol_ineq, tol_comp, alpha_min, lam_min, t_min, thr0, reg;
    double blend_min, blend_max;   /* kinematic/dynamic blend window on v_x: p = clamp((v_x - blend_min)/(blend_max - blend_min), 0, 1)
                                      ($A/ad_3d_optimizer.py:443, $A/ad_3d.py:62-64 = 100 / 110).  Used where the library
                                      itself produces the next solve's p (device closed loop, pose-only step without p);
                                      blend_max <= blend_min (default 0 / 0) keeps the p that was uploaded. */
} admpc_opts;

void admpc_default_opts(admpc_opts *o);
const char *admpc_last_error(void);
int admpc_device_count(void);

/* ----------------------------------------------------------------------------------------------------------
 * 1. single-instance shim: replaces $G/acados_solver_sim_car.h:116-150
 * -------------------------------------------------------------------------------------------------------- */
typedef struct sim_car_solver_capsule sim_car_solver_capsule;     /* opaque ($G/acados_solver_sim_car.h:78-114) */

sim_car_solver_capsule *sim_car_acados_create_capsule(void);                      /* .h:116 */
int sim_car_acados_free_capsule(sim_car_solver_capsule *capsule);                 /* .h:117 */
int sim_car_acados_create(sim_car_solver_capsule *capsule);                       /* .h:119  N=40, dt=0.05 */
int sim_car_acados_create_with_discretization(sim_car_solver_capsule *capsule, int n_time_steps,
                                              double *new_time_steps);            /* .h:128 (uniform steps only) */
int sim_car_acados_update_time_steps(sim_car_solver_capsule *capsule, int N, double *new_time_steps); /* .h:133 */
int sim_car_acados_update_qp_solver_cond_N(sim_car_solver_capsule *capsule, int qp_solver_cond_N);   /* .h:137: accepted
                                                              and ignored (no condensing here; the reference's own body exits) */
int sim_car_acados_update_params(sim_car_solver_capsule *capsule, int stage, double *value, int np);  /* .h:138 */
int sim_car_acados_solve(sim_car_solver_capsule *capsule);                        /* .h:139 */
int sim_car_acados_reset(sim_car_solver_capsule *capsule, int reset_qp_solver_mem); /* .h:121 */
int sim_car_acados_free(sim_car_solver_capsule *capsule);                         /* .h:140 */
void sim_car_acados_print_stats(sim_car_solver_capsule *capsule);                 /* .h:141 */
/* options used by the next create() on this capsule (the reference freezes them at code-generation time) */
int sim_car_acados_set_opts(sim_car_solver_capsule *capsule, const admpc_opts *opts);
/* nlp_solver_type: "SQP_RTI" (shipped, $A/ad_3d_optimizer.py:205 default) or "SQP" (point-reference mode,
 * $A/create_ros_ad_mpc.py:47-51); max_iter <= 0 and tol4 == NULL keep nlp_solver_max_iter 100 / tolerances 1e-6
 * (acados_models/sim_car_acados_ocp.json:868-873).  solve() then iterates to convergence; get_stat "sqp_iter". */
int sim_car_acados_set_nlp_solver(sim_car_solver_capsule *capsule, const char *type, int max_iter, const double *tol4);
/* flat field access replacing ocp_nlp_{cost_model,constraints_model,out}_set / ocp_nlp_out_get / ocp_nlp_get.
 * set fields: "yref" (9, or 7 at stage N), "lbx"/"ubx" (stage 0: 7 = x0; stages 1..N-1: 1), "p" (1), "x" (7), "u" (2),
 *             "pi" (7), "lam" / "t" (10, stage 0: 22), "sl" / "su" (2)  -- all seven fields of acados' load_iterate,
 *             "kappa" (1; Frenet variant only: path curvature at the node)
 * The typed getters of the reference (sim_car_acados_get_nlp_in/out/solver/config/opts/dims/plan, .h:143-150) return
 * libacados structs; they are replaced by this flat string-keyed access and are deliberately NOT exported.
 * get fields: "x" (7), "u" (2), "pi" (7), "lam" (10, stage 0: 22), "t" (same), "sl" (2), "su" (2)
 * stats: "sqp_iter"(int) "qp_iter"(int) "qp_stat"(int) "status"(int) "time_tot"(double, s) "kkt_norm_inf"(double) */
int sim_car_acados_set(sim_car_solver_capsule *capsule, int stage, const char *field, const double *value, int n);
int sim_car_acados_get(sim_car_solver_capsule *capsule, int stage, const char *field, double *out, int n);
int sim_car_acados_get_stat(sim_car_solver_capsule *capsule, const char *name, void *out);

/* ----------------------------------------------------------------------------------------------------------
 * 2. batched solver: B independent instances, SoA-resident in HBM
 *    Host layouts are instance-major (AoS), i.e. what numpy gives for x0[B,7], yref[B,N*9+7], ...
 * -------------------------------------------------------------------------------------------------------- */
typedef struct admpc_batch admpc_batch;

int admpc_batch_create(const admpc_opts *opts, int B, int device, admpc_batch **out);
int admpc_batch_free(admpc_batch *h);
int admpc_batch_size(const admpc_batch *h);
int admpc_batch_horizon(const admpc_batch *h);

/* GP model (model_fitting/gp.py:495-508 schema): X[nout][M][dz], alpha[nout][M] (=K^-1 y), ell[nout][dz],
 * sigma_f[nout], y_mean[nout]; feat[dz], rows[nout].  Pass nout = 0 to disable the GP again. */
int admpc_batch_set_gp(admpc_batch *h, int nout, int M, int dz, const int *feat, const int *rows,
                       const double *X, const double *alpha, const double *ell, const double *sigma_f,
                       const double *y_mean, int stage0_trigger);

/* GP ensemble (GPEnsemble, model_fitting/gp.py:536-770, homogeneous case): K cluster models with a leading model axis,
 * X[K][nout][M][dz], alpha[K][nout][M], ell[K][nout][dz], sigma_f[K][nout], y_mean[K][nout], centroids[K][dz] (cluster
 * means in feature space, in the order the reference sorts them: ascending first coordinate, gp.py:592-595).  All
 * models are staged in shared memory together (220 KB budget).  Every instance uses ONE model per solve:
 *   admpc_batch_select_gp   nearest centroid to z = B_z [xq; uq] (select_gp, gp.py:738-770); xq NULL = the current x0,
 *                           uq NULL = zeros (the reference queries with the reference state / target input,
 *                           quad_mpc/quad_3d_optimizer.py:452,491)
 *   admpc_batch_set_gp_index explicit choice (the reference's use_model argument); model 0 after set_gp_ensemble. */
int admpc_batch_set_gp_ensemble(admpc_batch *h, int K, int nout, int M, int dz, const int *feat, const int *rows,
                                const double *X, const double *alpha, const double *ell, const double *sigma_f,
                                const double *y_mean, const double *centroids, int stage0_trigger);
int admpc_batch_select_gp(admpc_batch *h, const double *xq /*[B][7] or NULL*/, const double *uq /*[B][2] or NULL*/);
int admpc_batch_set_gp_index(admpc_batch *h, const int *idx /*[B]*/);
int admpc_batch_get_gp_index(admpc_batch *h, int *idx /*[B]*/);

/* per-solve inputs ($A/ad_3d_optimizer.py:420-450).  Asynchronous on the handle's stream; pinned host memory
 * (admpc_host_alloc) makes the copies truly asynchronous. */
int admpc_batch_set_x0(admpc_batch *h, const double *x0 /*[B][7]*/);
int admpc_batch_set_yref(admpc_batch *h, const double *yref /*[B][N*9+7]*/);
int admpc_batch_set_p(admpc_batch *h, const double *p /*[B][N]*/);
int admpc_batch_set_p_scalar(admpc_batch *h, const double *p /*[B]*/);      /* same switch on all stages (:449-450) */
/* Frenet variant only (opts.model_variant == 1): path curvature kappa[B][N] at every shooting node (default 0; the
 * reference evaluates a B-spline kappa(s) inside the model, fren_ad_3d_optimizer bytecode).  With kappa = 0 the variant
 * coincides with the Cartesian model.  Kernels: gp_sweep_kernel<FR> + csrc/frenet.cu (two-pass preparation) and csrc/qp_mma_g.cu
 * (feedback on the FP64 tensor cores: both curvature forms, both constraint sets, N <= 127), full SQP mode
 * included; the device reference generator and the closed-loop plant are Cartesian-only (ADMPC_E_UNSUPPORTED). */
int admpc_batch_set_kappa(admpc_batch *h, const double *kappa);
/* Frenet variant, the reference's own semantics: kappa(s) as a spline of the arc length evaluated INSIDE the model at every
 * RK4 sub-stage, the Jacobian gaining its d kappa / d s column (bytecode: interpolant('kapparef_s', 'bspline', s_knots,
 * curv)).  Any spline is passed in piecewise-polynomial form: K cubic pieces per instance, breaks[B][K+1], coef[B][K][4]
 * with kappa(s) = c0 + c1 t + c2 t^2 + c3 t^3, t = s - breaks[j] (end pieces extrapolate); K <= 64.  K = 0 / NULL returns
 * to the per-node constants.  With a spline the column of s in A_k is dense; the default feedback kernel (qp_mma_g.cu) assumes
 * no trivial column, so both forms run at the same speed. */
int admpc_batch_set_kappa_spline(admpc_batch *h, int K, const double *breaks, const double *coef);
int admpc_batch_set_gp_state(admpc_batch *h, const double *gp_state /*[B][7] or NULL = x0*/);
/* iterate (initial guess / warm start).  reset zeroes it like $G/acados_solver_sim_car.c:819-852. */
int admpc_batch_set_iterate(admpc_batch *h, const double *x /*[B][(N+1)*7]*/, const double *u /*[B][N*2]*/);
/* multipliers and slacks of the iterate: together with set_iterate the seven fields acados' load_iterate restores
 * ($A/ad_3d_optimizer.py:454; layouts as in the getters below).  The RTI step replaces them by the QP's; the full-SQP
 * loop's first KKT check reads them.  Any pointer may be NULL (left unchanged). */
int admpc_batch_set_duals(admpc_batch *h, const double *pi /*[B][N*7]*/, const double *lam /*[B][N*10]*/,
                          const double *t /*[B][N*10]*/, const double *sl /*[B][N*2]*/, const double *su /*[B][N*2]*/);
int admpc_batch_reset(admpc_batch *h);

int admpc_batch_solve(admpc_batch *h);   /* one SQP-RTI iteration for all B instances (async launch) */
int admpc_batch_wait(admpc_batch *h);    /* block until everything queued on the handle has finished */

/* results (each waits for the solve, then copies D2H) */
int admpc_batch_get_u(admpc_batch *h, double *u /*[B][N*2]*/);
int admpc_batch_get_x(admpc_batch *h, double *x /*[B][(N+1)*7]*/);
int admpc_batch_get_pi(admpc_batch *h, double *pi /*[B][N*7]*/);
int admpc_batch_get_lam(admpc_batch *h, double *lam /*[B][N*10]*/);
int admpc_batch_get_t(admpc_batch *h, double *t /*[B][N*10]*/);
int admpc_batch_get_slacks(admpc_batch *h, double *sl /*[B][N*2]*/, double *su /*[B][N*2]*/);
int admpc_batch_get_status(admpc_batch *h, int *status /*[B]*/, int *qp_status /*[B] or NULL*/, int *qp_iter /*[B] or NULL*/);
/* linearisation of the last solve, for kernel-level parity tests: A[B][N][49] B[B][N][14] b[B][N][7] q[B][(N+1)*7] r[B][N*2] */
int admpc_batch_get_lin(admpc_batch *h, double *A, double *Bm, double *b, double *q, double *r);

/* whole RTI step through host buffers in ONE call: H2D(x0,yref,p) -> solve -> D2H(u,x,status).  This is the
 * replacement of the 3N+5 ctypes calls of $A/ad_3d_optimizer.py:420-465. Any output pointer may be NULL. */
int admpc_batch_solve_host(admpc_batch *h, const double *x0, const double *yref, const double *p_scalar,
                           double *u_out, double *x_out, int *status_out);
/* same, but returns as soon as everything is enqueued on the handle's stream (pinned host buffers required for true
 * asynchrony); complete with admpc_batch_wait.  Several handles (one stream each) pipeline H2D / solve / D2H. */
int admpc_batch_solve_host_async(admpc_batch *h, const double *x0, const double *yref, const double *p_scalar,
                                 double *u_out, double *x_out, int *status_out);

/* The same call with the chunk pipeline INSIDE the library: one handle owning `chunks` chunk handles (contiguous blocks of
 * the batch, one CUDA stream each), so that the H2D copy of chunk c+1, the solve of chunk c and the D2H copy of chunk c-1
 * overlap.  One call from one host thread; host buffers should be pinned (admpc_host_alloc) for real overlap.  This is
 * the end-to-end entry bench.py times ("e2e").  admpc_pipe_chunk exposes a chunk's handle ([lo, hi) = its instances)
 * for per-chunk settings that have no pipe-level call (GP ensembles, curvature, gp_state, getters). */
typedef struct admpc_pipe admpc_pipe;
int admpc_pipe_create(const admpc_opts *opts, int B, int device, int chunks, admpc_pipe **out);
int admpc_pipe_free(admpc_pipe *p);
int admpc_pipe_chunks(const admpc_pipe *p);
admpc_batch *admpc_pipe_chunk(admpc_pipe *p, int c, int *lo, int *hi);
int admpc_pipe_set_gp(admpc_pipe *p, int nout, int M, int dz, const int *feat, const int *rows, const double *X,
                      const double *alpha, const double *ell, const double *sigma_f, const double *y_mean, int stage0_trigger);
int admpc_pipe_set_iterate(admpc_pipe *p, const double *x /*[B][(N+1)*7]*/, const double *u /*[B][N*2]*/);
int admpc_pipe_set_track(admpc_pipe *p, int L, const double *traj, int H, double traj_dt, int anchor_at_closest);
int admpc_pipe_solve_host(admpc_pipe *p, const double *x0, const double *yref, const double *p_scalar,
                          double *u_out, double *x_out, int *status_out);          /* blocking */
int admpc_pipe_solve_host_async(admpc_pipe *p, const double *x0, const double *yref, const double *p_scalar,
                                double *u_out, double *x_out, int *status_out);    /* complete with admpc_pipe_wait */
int admpc_pipe_solve_pose(admpc_pipe *p, const double *x0, const double *p_scalar, double *u_out, double *x_out,
                          int *status_out);                                        /* refgen + solve per chunk, blocking */
int admpc_pipe_wait(admpc_pipe *p);
long long admpc_pipe_kernel_launches(const admpc_pipe *p);

/* Reference generation on the device (ad_mpc/ref_traj.py:89-171 + nodes/gp_ad_mpc_node.py:180-187 +
 * ad_mpc/ad_3d_optimizer.py:343-345,420-438): set_track takes the table RefTrajectory.set_traj builds
 * (traj[L][6] rows = [vel, x, y, psi, cdist, curv], ref_traj.py:84) with the generator's horizon H (>= N) and time step;
 * make_yref turns the CURRENT x0 of every instance (pose = x0[0..2]) into its yref, in place on the device. */
int admpc_batch_set_track(admpc_batch *h, int L, const double *traj /*[L][6]*/, int H, double traj_dt);
/* 0 (default): literal get_waypoints, arc lengths measured from the start of the track window (ref_traj.py:128-131);
 * 1: extension for a shared global track -- arc lengths start at each vehicle's closest waypoint (H <= 64). */
int admpc_batch_set_track_anchor(admpc_batch *h, int anchor_at_closest);
int admpc_batch_make_yref(admpc_batch *h);
/* pose-only step: H2D(x0, p) -> make_yref -> solve -> D2H(u, x, status), enqueued on the handle's stream (complete with
 * admpc_batch_wait).  Replaces get_waypoints + set_reference + run_optimization of one control step for B vehicles. */
int admpc_batch_solve_pose_async(admpc_batch *h, const double *x0, const double *p_scalar, double *u_out,
                                 double *x_out, int *status_out);
int admpc_batch_get_yref(admpc_batch *h, double *yref /*[B][N*9+7]*/);
int admpc_batch_get_waypoint_info(admpc_batch *h, double *s0 /*[B]*/, double *e_y0 /*[B]*/, double *e_psi0 /*[B]*/, int *stop);

/* After-solve logic of the reference, batched on the device: is_valid_command
 * (ad_mpc/ad_3d_optimizer.py:385-394), backup control (:469-476), safety counter (nodes/gp_ad_mpc_node.py:206-216);
 * advance != 0 also integrates the nominal model one step with the applied control and installs the result as the
 * next x0.  closed_loop runs `steps` control steps ([make_yref] -> solve -> postsolve) without a host round trip. */
int admpc_batch_postsolve(admpc_batch *h, int advance, int safe_threshold);
int admpc_batch_closed_loop(admpc_batch *h, int steps, int use_track, int safe_threshold, double *log_x /*[steps+1][B][7] or NULL*/);
int admpc_batch_get_loop_info(admpc_batch *h, int *valid, int *safe_count, int *cmd_ok, double *u_apply /*[B][2]*/, double *x0 /*[B][7]*/);

/* instrumentation: device time (CUDA events on the handle's stream) of the last solve and of its kernels.
 * name: "solve" | "prepare" | "qp" | "h2d" | "d2h" ; returns milliseconds in *ms. Needs admpc_batch_set_profiling(1)
 * for the per-kernel entries. */
int admpc_batch_set_profiling(admpc_batch *h, int on);
int admpc_batch_last_ms(admpc_batch *h, const char *name, float *ms);
long long admpc_batch_kernel_launches(const admpc_batch *h);   /* kernels launched by this handle so far */
/* stream-ordered event pair for external timing of a region on the handle's stream */
int admpc_batch_timer_start(admpc_batch *h);
int admpc_batch_timer_stop(admpc_batch *h, float *ms);   /* records, synchronises and returns elapsed ms */
int admpc_batch_flush_l2(admpc_batch *h);                /* writes a >L2-sized scratch buffer (bench hygiene) */

/* pinned host memory helpers */
void *admpc_host_alloc(unsigned long long bytes);
int admpc_host_free(void *p);

/* multi-GPU plumbing (one process per GPU).  The NCCL library is dlopen'ed; unique id bytes (128) are exchanged by
 * the caller's launcher (torch.distributed / MPI / file).  Broadcast ships the GP model from root to every rank's
 * handle; gather collects [u | x | status] blocks on root.  Every rank must hold the SAME number of instances (checked
 * in comm_init; gather / gather_enable return ADMPC_E_UNSUPPORTED otherwise -- pad the batch to a multiple of nranks).
 * bcast_gp is safe against local failures: the root's verdict travels in the header and all ranks agree (one 4-byte
 * all-reduce) before the blob moves, so either every rank has the model or every rank returns an error. */
int admpc_nccl_unique_id(void *id128);
int admpc_batch_comm_init(admpc_batch *h, const void *id128, int rank, int nranks);
int admpc_batch_bcast_gp(admpc_batch *h, int root, int nout, int M, int dz, const int *feat, const int *rows,
                         const double *X, const double *alpha, const double *ell, const double *sigma_f,
                         const double *y_mean, int stage0_trigger);
int admpc_batch_gather(admpc_batch *h, int root, double *u_all /*[nranks*B][N*2]*/, double *x_all, int *status_all);
/* collective, optional: FUSED gather towards root.  The root exports its gathered block through CUDA IPC, every rank maps
 * it, and from then on the QP kernels' epilogue writes each instance's [u | x | status] straight into the rank's slice
 * of the root's block (NVLink peer stores overlapped with the solve); admpc_batch_gather becomes a 4-byte all-reduce
 * (stream-ordered completion barrier).  The root's block is double-buffered: each gather call delivers one half and
 * flips the half later solves write to, so a rank that runs ahead never overwrites a block the root is still copying
 * out -- PROVIDED every rank calls admpc_batch_gather between two solves whose results are to be gathered (ranks that
 * solve twice or more without a gather must separate them from the root's read-out with admpc_batch_barrier).
 * Returns 1 = fused path active on all ranks, 0 = stays on NCCL send/recv (IPC unavailable or ADMPC_GATHER=nccl), < 0 error. */
int admpc_batch_gather_enable(admpc_batch *h, int root);
/* host copy of what the last gather left on the root's device (rank-major blocks); root only */
int admpc_batch_get_gathered(admpc_batch *h, double *u_all, double *x_all, int *status_all);
int admpc_batch_barrier(admpc_batch *h);

/* GP model update on the device (model_fitting/gp.py:283-289,305-311,361-363): for ONE output dimension builds
 * K = sigma_f exp(-1/2 |x_i/l - x_j/l|^2) + sigma_n^2 I, factorises it (blocked FP64 Cholesky) and returns
 * alpha = K^-1 y (what admpc_batch_set_gp consumes) and the negative log likelihood the reference minimises.
 * y must already have its mean removed (gp.py:343).  ADMPC_E_ARG when K is not positive definite. */
int admpc_gp_fit(int device, int M, int dz, const double *X /*[M][dz]*/, const double *y /*[M]*/, const double *ell /*[dz]*/,
                 double sigma_f, double sigma_n, double *alpha_out /*[M] or NULL*/, double *nll_out, float *ms_out);

/* GP posterior at n test points (next-row f4; reference: CustomGPRegression.predict(x, return_cov=True),
 * model_fitting/gp.py:402-441): mu = k_s K^-1 y + y_mean, cov = k(x*,x*) + 1e-8 I - k_s K^-1 k_s^T.  The reference forms
 * inv(K); here the blocked Cholesky of admpc_gp_fit runs on the augmented matrix and the Schur complement is the
 * covariance.  var_out [n] = diag(cov); cov_out [n][n] full covariance or NULL; y has its mean removed (gp.py:343). */
int admpc_gp_predict(int device, int M, int dz, const double *X /*[M][dz]*/, const double *y /*[M]*/, const double *ell,
                     double sigma_f, double sigma_n, double y_mean, int n, const double *Xtest /*[n][dz]*/,
                     double *mu_out, double *var_out, double *cov_out);

/* Full SQP mode (nlp_solver_type "SQP": $A/create_ros_ad_mpc.py:47-51 -> $A/ad_3d_optimizer.py:205; defaults
 * nlp_solver_max_iter 100 and tolerances 1e-6 from acados_models/sim_car_acados_ocp.json:868-873).  Repeats
 * { linearise; NLP KKT residual check; QP; full step } on the device until every instance has converged, failed or
 * used max_iter iterations; tol4 = {stat, eq, ineq, comp} (NULL: 1e-6 each).  Synchronous.  iterations_run (may be
 * NULL) receives the number of batch iterations executed. */
int admpc_batch_solve_sqp(admpc_batch *h, int max_iter, const double *tol4, int *iterations_run);
/* per-instance outcome of the last solve_sqp: acados status {0 converged, 1 NaN, 2 max_iter, 4 QP failure}, QPs
 * solved, NLP residual norms [B][4] of the last check.  Any pointer may be NULL. */
int admpc_batch_get_sqp_info(admpc_batch *h, int *status, int *sqp_iter, double *res);

/* FP64 peak probe: runs a dependent-free DFMA loop and returns achieved TFLOP/s (roofline denominator; there is
 * no FP64 entry in MEASURED_PEAKS.json). */
int admpc_measure_fp64_peak(int device, double *tflops);
/* Same loop with 0 / 2 / 4 / 8 independent integer multiply-adds issued per 8 DFMAs: the FP64 throughput that survives
 * the issue-slot pressure of index / address arithmetic (diagnostic for DESIGN.md's attainable-roofline estimate).
 * int_ops_per_8_dfma = 16 selects the operand-fetch probe instead: DFMAs with three distinct register operands. */
int admpc_measure_fp64_mix(int device, int int_ops_per_8_dfma, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
