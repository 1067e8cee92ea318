/*
 * rti_oracle.h -- CPU ORACLE (test infrastructure; NOT part of the shipped product path).
 *
 * Plain-C restatement of ONE SQP-RTI iteration of the AD_MPC bicycle-model NMPC
 * (reference: data_driven_mpc/ros_gp_mpc/src/ad_mpc/ad_3d_optimizer.py:135-209,268-310,396-480 and the
 * acados-generated shim c_generated_code/acados_solver_sim_car.c:343-699).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Parity pinning (see DESIGN.md "Oracle"):
 *   - model f / VDE          : pinned against the reference's own CasADi C (oracle/_ref/libsim_car_ref.so)
 *   - RK4+sensitivities, cost scaling, multiplier conventions: pinned against the acados golden iterate
 *                              src/ad_mpc/sim_car_iterate.json (tests/golden/)
 *   - GP posterior mean      : pinned against the reference's own numeric predict() (model_fitting/gp.py imported with
 *                              its symbolic dependencies stubbed; tests/golden/gp_reference.npz); the Jacobian by finite
 *                              differences of that mean
 *   - QP SOLUTION            : certified solver-independently (tests/test_qp_certificate_cpu.py: the oracle's answer is
 *                              the KKT point of its active set, by one dense numpy KKT solve -- sufficient for a convex QP)
 *   - IPM iterate path / iteration counts: UNPINNED by reference fixtures (acados/HPIPM are not vendored in the
 *                              reference) -- restated from the published algorithm, design choices in DESIGN.md.
 */
#ifndef RTI_ORACLE_H_
#define RTI_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NX 7
#define ORC_NU 2
#define ORC_NC 12      /* storage rows per stage; the rows in use and their stride = orc_con_rows(): 10 for the Cartesian set
                          [lbu0 lbu1 lbx | ubu0 ubu1 ubx | ls0 ls1 | us0 us1], 12 for the Frenet set (rti_oracle.c con_get) */
#define ORC_NMAX 128   /* max horizon */
#define ORC_DZMAX 8
#define ORC_GPOUT_MAX 4

typedef struct orc_opts {
    int N;                /* shooting intervals */
    int iter_max;         /* QP (IPM) iteration limit                 acados_solver_sim_car.c:692 */
    int gp_enabled;       /* 0: nominal model, 1: f + B_x mu(z)       quad_3d_optimizer.py:315 */
    int gp_nout;          /* number of GP outputs (<= ORC_GPOUT_MAX) */
    int gp_M;             /* training points per output */
    int gp_dz;            /* feature dimension (<= ORC_DZMAX) */
    int gp_stage0_trigger;/* 1: stage 0 evaluates the GP at gp_state (quad_3d_optimizer.py:295,548-552) */
    int model_backend;    /* 0: analytic restatement, 1: reference CasADi C via oracle/_ref (nominal only),
                             2: Frenet variant (SURVEY 8a A2'), curvature per shooting node via the *_frenet entry points */
    int gp_feat[ORC_DZMAX];      /* indices into [x(7); u(2)] selected by B_z   gp.py:609-630 */
    int gp_row[ORC_GPOUT_MAX];   /* state row each GP output is added to (B_x)  utils.py:773-786 */
    double dt;            /* interval length = cost scaling Ts        acados_solver_sim_car.c:362-366 */
    double W[9];          /* diag of LINEAR_LS stage weight [x;u]     acados_solver_sim_car.c:393-399 */
    double We[7];         /* diag of terminal weight                  acados_solver_sim_car.c:481-485 */
    double zl[2], zu[2], Zl[2], Zu[2];   /* slack penalties           acados_solver_sim_car.c:455-473 */
    double lbu[2], ubu[2];               /* soft input bounds         acados_solver_sim_car.c:549-552 */
    double lbx, ubx;                     /* hard bound on x[6], stages 1..N-1        .c:595-596 */
    double mass, lf, lr, iz, cf2, cr2;   /* vehicle: m, L_F, L_R, Iz, 2Cf, 2Cr       ad_3d.py:47-60 */
    double mu0, tol_stat, tol_eq, tol_ineq, tol_comp, alpha_min, lam_min, t_min, thr0, reg;
    int con_set;          /* 0: u0, u1 soft + delta hard (ad_3d_optimizer.py:165-199) ; 1: the Frenet variant's set: u0 soft, u1 hard,
                             e_y = x[1] in [lbx2, ubx2] hard, delta soft (fren_ad_3d_optimizer pyc, structure pinned by ad_mpc/debug.json) */
    double lbx2, ubx2;
} orc_opts;

/* GP model, one per output j (gp.py:495-508 pickle schema: x_train, k_inv_y, kernel_params{l,sigma_f}, y_mean):
 *   X      [nout][M][dz]   training inputs
 *   alpha  [nout][M]       K^-1 y
 *   ell    [nout][dz]      ARD length scales
 *   sigma_f[nout], y_mean[nout] */
typedef struct orc_gp {
    const double *X, *alpha, *ell, *sigma_f, *y_mean;
} orc_gp;

/* Per-instance iterate, acados store_iterate layout per stage (sim_car_iterate.json). */
typedef struct orc_iterate {
    double x[(ORC_NMAX + 1) * ORC_NX];
    double u[ORC_NMAX * ORC_NU];
    double pi[ORC_NMAX * ORC_NX];
    double lam[ORC_NMAX * ORC_NC];
    double t[ORC_NMAX * ORC_NC];
    double sl[ORC_NMAX * ORC_NU];
    double su[ORC_NMAX * ORC_NU];
} orc_iterate;

/* Linearisation of one instance (output of the preparation phase). Row-major. */
typedef struct orc_lin {
    double A[ORC_NMAX * 49];
    double B[ORC_NMAX * 14];
    double b[ORC_NMAX * 7];       /* phi(x_k,u_k) - x_{k+1} */
    double q[(ORC_NMAX + 1) * 7]; /* cost gradient wrt x */
    double r[ORC_NMAX * 2];       /* cost gradient wrt u */
} orc_lin;

typedef struct orc_stats {
    int status;      /* acados status of the RTI step: 0 ok, 1 NaN in linearisation, 4 QP failure */
    int qp_status;   /* acados-mapped QP status: 0 ok, 2 maxiter, 3 minstep, 1 NaN */
    int qp_iter;
    double res[4];   /* final inf-norm residuals: stat, eq, ineq, comp */
    double step_inf; /* inf-norm of the primal QP solution (dx,du) */
} orc_stats;

void orc_default_opts(orc_opts *o);
int orc_con_rows(const orc_opts *o);   /* inequality rows per stage of the configured constraint set = stride of lam / t */
/* constraint-side KKT relations of an iterate: max |t - constraint function|, max |slack stationarity|, max |lam t| */
void orc_con_check(const orc_opts *o, const struct orc_iterate *it, double out[3]);

/* model */
void orc_ode(const orc_opts *o, const orc_gp *gp, const double *x, const double *u, double p,
             const double *gp_state, double trigger, double *xdot);
void orc_model_jac(const orc_opts *o, const orc_gp *gp, const double *x, const double *u, double p,
                   const double *gp_state, double trigger, double *f, double *Jx /*7x7 rm*/, double *Ju /*7x2 rm*/);
void orc_gp_predict(const orc_opts *o, const orc_gp *gp, const double *z, double *mu, double *dmu /*[nout][dz]*/);
/* ERK4, 1 step of length dt, forward sensitivities: xn = phi(x,u), A = dphi/dx, B = dphi/du (row-major) */
int orc_rk4_sens(const orc_opts *o, const orc_gp *gp, const double *x, const double *u, double p,
                 const double *gp_state, double trigger, double *xn, double *A, double *B);

/* preparation: linearise around the iterate. yref: N rows of 9 then 7 terminal; p: N values. */
int orc_prepare(const orc_opts *o, const orc_gp *gp, const orc_iterate *it, const double *yref,
                const double *p, const double *gp_state, orc_lin *lin);

/* feedback: solve the QP by the OCP-structured primal-dual IPM; writes the QP solution. */
typedef struct orc_qpsol {
    double dx[(ORC_NMAX + 1) * 7], du[ORC_NMAX * 2], pi[ORC_NMAX * 7];
    double lam[ORC_NMAX * ORC_NC], t[ORC_NMAX * ORC_NC], sl[ORC_NMAX * 2], su[ORC_NMAX * 2];
} orc_qpsol;
int orc_qp_solve(const orc_opts *o, const orc_lin *lin, const orc_iterate *it, const double *x0,
                 orc_qpsol *sol, orc_stats *st);

/* one full RTI step (prepare + feedback + update) */
int orc_rti_step(const orc_opts *o, const orc_gp *gp, const double *x0, const double *yref,
                 const double *p, const double *gp_state, orc_iterate *it, orc_stats *st);

/* batch helpers (OpenMP over instances). AoS host layouts:
 *   x0[B][7], yref[B][N*9+7], p[B][N], gp_state[B][7] or NULL,
 *   xit[B][(N+1)*7], uit[B][N*2] in/out iterate (primal part; duals are not carried: QP is cold-started)
 *   status/qp_status/qp_iter[B] */
int orc_rti_batch(const orc_opts *o, const orc_gp *gp, int B, const double *x0, const double *yref,
                  const double *p, const double *gp_state, double *xit, double *uit, double *piout,
                  int *status, int *qp_status, int *qp_iter, int nthreads);

/* full SQP mode (nlp_solver_type "SQP": create_ros_ad_mpc.py:47-51; nlp_solver_max_iter 100, tolerances 1e-6:
 * acados_models/sim_car_acados_ocp.json:868-873).  tol = {stat, eq, ineq, comp}. */
void orc_nlp_residuals(const orc_opts *o, const orc_lin *lin, const orc_iterate *it, const double *x0, double res[4]);
int orc_sqp_solve(const orc_opts *o, const orc_gp *gp, const double *x0, const double *yref, const double *p,
                  const double *gp_state, orc_iterate *it, int max_iter, const double tol[4], int *sqp_iter,
                  double res_out[4]);
int orc_sqp_batch(const orc_opts *o, const orc_gp *gp, int B, const double *x0, const double *yref,
                  const double *p, const double *gp_state, double *xit, double *uit, int max_iter, const double *tol,
                  int *status, int *sqp_iter, double *res, int nthreads);

/* Frenet variant (model_backend = 2): kappa[N] per instance = path curvature at every shooting node */
void orc_set_kappa(double kappa);      /* for direct orc_model_jac / orc_rk4_sens calls (thread-local) */
/* kappa(s) as a piecewise cubic (K pieces, breaks[K+1], coef[K][4], lowest power first) evaluated inside the model at every
 * RK4 sub-stage, with its d kappa / d s Jacobian column; NULL / K = 0 switches back to the per-node constant */
void orc_set_kappa_spline(int K, const double *breaks, const double *coef);            /* thread-local, direct calls */
void orc_set_batch_kappa_spline(int K, const double *breaks, const double *coef);      /* [B][K+1], [B][K][4]: batch entry points */
int orc_prepare_frenet(const orc_opts *o, const orc_gp *gp, const orc_iterate *it, const double *yref,
                       const double *p, const double *kappa, const double *gp_state, orc_lin *lin);
int orc_rti_step_frenet(const orc_opts *o, const orc_gp *gp, const double *x0, const double *yref,
                        const double *p, const double *kappa, const double *gp_state, orc_iterate *it, orc_stats *st);
int orc_rti_batch_frenet(const orc_opts *o, const orc_gp *gp, int B, const double *x0, const double *yref,
                         const double *p, const double *kappa, const double *gp_state, double *xit, double *uit,
                         double *piout, int *status, int *qp_status, int *qp_iter, int nthreads);

int orc_sqp_batch_frenet(const orc_opts *o, const orc_gp *gp, int B, const double *x0, const double *yref,
                         const double *p, const double *kappa, const double *gp_state, double *xit, double *uit, int max_iter,
                         const double *tol, int *status, int *sqp_iter, double *res, int nthreads);

int orc_load_ref_model(const char *path); /* dlopen oracle/_ref/libsim_car_ref.so; 0 on success */

#ifdef __cplusplus
}
#endif
#endif
