/*
 * rti_oracle.c -- CPU ORACLE (test infrastructure; NOT part of the shipped product path).
 *
 * Straightforward dense C restatement of one SQP-RTI iteration of the reference controller.  Written for clarity,
 * not speed: dense 7x7 / 9x9 loops, no structure exploitation.  The CUDA product (ad_mpc_b200/csrc) is an
 * independent implementation that must reproduce these numbers.
 *
 * What each block follows (paths under /root/reference/data_driven_mpc/ros_gp_mpc/src/):
 *   orc_ode / orc_model_jac : ad_mpc/ad_3d_optimizer.py:268-310 (CasADi model), generated C
 *                             ad_mpc/c_generated_code/sim_car_model/sim_car_expl_ode_fun.c:51-291 and
 *                             sim_car_expl_vde_forw.c:120-2371 (checked against the compiled originals, oracle/_ref)
 *   orc_gp_predict          : model_fitting/gp.py:117-138 (kernel), :426-430 / :446-460 (mean), :140-165 (d/dz)
 *   GP injection            : quad_mpc/quad_3d_optimizer.py:289-327,548-552 ; utils/utils.py:773-808
 *   orc_rk4_sens            : acados sim_erk [EXT], options at acados_solver_sim_car.c:657-665 (4 stages, 1 step)
 *   orc_prepare             : acados LINEAR_LS cost / BGH bounds [EXT], data at acados_solver_sim_car.c:378-605
 *   orc_qp_solve            : HPIPM-style Mehrotra predictor-corrector IPM [EXT]; OCP-structured (Riccati) instead of
 *                             the reference's FULL_CONDENSING (acados_solver_sim_car.c:145) -- same QP, same solution
 *   orc_rti_step            : acados ocp_nlp_sqp_rti [EXT], options at acados_solver_sim_car.c:647-681
 * [EXT] = acados@91a01d4 / HPIPM / BLASFEO, pinned in requirements.txt:1 but not vendored; algorithm restated from
 * its published description and pinned by ad_mpc/sim_car_iterate.json (see tests/test_oracle_golden.py).
 */
#define _GNU_SOURCE
#include "rti_oracle.h"
#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* NaN-propagating max: a NaN residual must surface in the norms (fmax would drop it) */
static inline double nanmax(double a, double b) { return (a > b || a != a) ? a : b; }

#define NX ORC_NX
#define NU ORC_NU
#define NC ORC_NC

void orc_default_opts(orc_opts *o)
{
    memset(o, 0, sizeof(*o));
    o->N = 20;
    o->iter_max = 50;
    o->dt = 0.05;
    const double W[9] = {10, 10, 100, 0, 0, 0, 0, 1, 100};
    const double We[7] = {1e-5, 1e-5, 1e-4, 0, 0, 0, 0};
    memcpy(o->W, W, sizeof W);
    memcpy(o->We, We, sizeof We);
    for (int j = 0; j < 2; j++) { o->zl[j] = 10; o->zu[j] = 10; o->Zl[j] = 0; o->Zu[j] = 0; }
    o->lbu[0] = -10; o->ubu[0] = 5; o->lbu[1] = -3; o->ubu[1] = 3;
    o->lbx = -0.52; o->ubx = 0.52;
    o->con_set = 0; o->lbx2 = -2.0; o->ubx2 = 2.0;      /* e_y bound of the Frenet variant's set (con_set = 1) */
    /* ad_3d.py:47-60, evaluated literally (note the reference's 3.14195) */
    double mass = 1500, f_mass = 900, r_mass = mass - f_mass, L = 2.7;
    o->mass = mass;
    o->lf = L * (1 - f_mass / mass);
    o->lr = L * (1 - r_mass / mass);
    o->iz = o->lf * o->lr * (r_mass + f_mass);
    o->cf2 = 2 * (f_mass * 0.5 * 9.81 * 0.165 * 180 / 3.14195);
    o->cr2 = 2 * (r_mass * 0.5 * 9.81 * 0.165 * 180 / 3.14195);
    /* HPIPM BALANCE-like settings [EXT]; design choice, see DESIGN.md */
    o->mu0 = 10.0;
    o->tol_stat = o->tol_eq = o->tol_ineq = o->tol_comp = 1e-8;
    o->alpha_min = 1e-12;
    o->lam_min = 1e-16;
    o->t_min = 1e-16;
    o->thr0 = 0.1;
    o->reg = 1e-15;
    o->gp_row[0] = 4; o->gp_row[1] = 5;
    o->gp_feat[0] = 3; o->gp_feat[1] = 4; o->gp_feat[2] = 5; o->gp_feat[3] = 6;
}

/* ------------------------------------------------------------------ reference CasADi model (oracle/_ref) --- */
typedef int (*casadi_fn)(const double **, double **, int *, double *, int);
static casadi_fn ref_vde = 0, ref_ode = 0;

int orc_load_ref_model(const char *path)
{
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return -1;
    ref_vde = (casadi_fn)dlsym(h, "sim_car_expl_vde_forw");
    ref_ode = (casadi_fn)dlsym(h, "sim_car_expl_ode_fun");
    return (ref_vde && ref_ode) ? 0 : -2;
}

/* ------------------------------------------------------------------------------------------ GP (gp.py) ----- */
void orc_gp_predict(const orc_opts *o, const orc_gp *gp, const double *z, double *mu, double *dmu)
{
    const int M = o->gp_M, dz = o->gp_dz;
    for (int j = 0; j < o->gp_nout; j++) {
        const double *X = gp->X + (size_t)j * M * dz;
        const double *al = gp->alpha + (size_t)j * M;
        const double *ell = gp->ell + j * dz;
        double m = 0.0, g[ORC_DZMAX] = {0};
        for (int i = 0; i < M; i++) {
            double d2 = 0.0, diff[ORC_DZMAX];
            for (int d = 0; d < dz; d++) {
                diff[d] = z[d] - X[i * dz + d];
                d2 += diff[d] * diff[d] / (ell[d] * ell[d]);    /* gp.py:131-134 */
            }
            double k = gp->sigma_f[j] * exp(-0.5 * d2);         /* gp.py:138 (sigma_f, not squared) */
            double ka = k * al[i];
            m += ka;                                            /* gp.py:453 k_s^T K^-1 y */
            for (int d = 0; d < dz; d++) g[d] += -ka * diff[d] / (ell[d] * ell[d]);   /* gp.py:165 */
        }
        mu[j] = m + gp->y_mean[j];
        for (int d = 0; d < dz; d++) dmu[j * dz + d] = g[d];
    }
}

/* ------------------------------------------------------------------------------------------ model ---------- */
/* Frenet variant (model_backend == 2; SURVEY 8a A2': fren_ad_3d_optimizer, bytecode only): the curvature of the
 * reference path at the stage is a per-stage model parameter, handed down through this thread-local slot by the
 * preparation loop (the reference evaluates a B-spline kappa(s) inside the model). */
static _Thread_local double tls_kappa = 0.0;
void orc_set_kappa(double kappa) { tls_kappa = kappa; }
/* kappa(s) as the reference has it: a spline of the arc length evaluated INSIDE the model (CasADi
 * interpolant('kapparef_s', 'bspline', [s_knots], curv) in the bytecode), so that every RK4 sub-stage sees the
 * curvature at its own s and the Jacobian gains a d kappa / d s column.  Any spline reaches this code in piecewise-
 * polynomial form: K cubic pieces, breaks[K+1], coef[K][4] with kappa(s) = c0 + c1 t + c2 t^2 + c3 t^3, t = s - breaks[j];
 * outside [breaks[0], breaks[K]] the end pieces extrapolate.  Thread-local (one instance per thread); the batch entry
 * points install the instance's spline from the table set by orc_set_batch_kappa_spline. */
static _Thread_local int tls_sp_K = 0;
static _Thread_local const double *tls_sp_breaks = 0, *tls_sp_coef = 0;
void orc_set_kappa_spline(int K, const double *breaks, const double *coef)
{
    tls_sp_K = (breaks && coef) ? K : 0; tls_sp_breaks = breaks; tls_sp_coef = coef;
}
static int g_sp_K = 0;
static const double *g_sp_breaks = 0, *g_sp_coef = 0;          /* [B][K+1], [B][K][4] */
void orc_set_batch_kappa_spline(int K, const double *breaks, const double *coef)
{
    g_sp_K = (breaks && coef) ? K : 0; g_sp_breaks = breaks; g_sp_coef = coef;
}
static void batch_spline_install(int b)
{
    if (g_sp_K > 0) orc_set_kappa_spline(g_sp_K, g_sp_breaks + (size_t)b * (g_sp_K + 1), g_sp_coef + (size_t)b * g_sp_K * 4);
    else orc_set_kappa_spline(0, 0, 0);
}
static void kappa_eval(double s, double *kap, double *dkap)
{
    if (tls_sp_K <= 0) { *kap = tls_kappa; *dkap = 0.0; return; }
    int j = 0;
    while (j + 1 < tls_sp_K && s >= tls_sp_breaks[j + 1]) j++;
    const double t = s - tls_sp_breaks[j], *c = tls_sp_coef + 4 * j;
    *kap = ((c[3] * t + c[2]) * t + c[1]) * t + c[0];
    *dkap = (3.0 * c[3] * t + 2.0 * c[2]) * t + c[1];
}

void orc_model_jac(const orc_opts *o, const orc_gp *gp, const double *x, const double *u, double p,
                   const double *gp_state, double trigger, double *f, double *Jx, double *Ju)
{
    const double psi = x[2], vx = x[3], vy = x[4], r = x[5], dl = x[6];
    const double u0 = u[0], u1 = u[1];
    const double m = o->mass, lf = o->lf, lr = o->lr, iz = o->iz, cf2 = o->cf2, cr2 = o->cr2;
    const double L = lr + lf, c = lr / L, q = 1.0 - p;
    const double D = vx + 1e-99;                                   /* ad_3d_optimizer.py:290 */
    const double sp = sin(psi), cp = cos(psi), sd = sin(dl), cd = cos(dl);
    const double Ff = cf2 * (dl - (vy + lf * r) / D);              /* :290 */
    const double Fr = cr2 * (lr * r - vy) / D;                     /* :296 */

    f[0] = vx * cp - vy * sp;                                      /* :277 */
    f[1] = vx * sp + vy * cp;                                      /* :280 */
    f[2] = r;                                                      /* :283 */
    f[3] = p * (u0 - (1.0 / m) * Ff * sd + vy * r) + q * u0;       /* :291-293 */
    f[4] = p * ((1.0 / m) * (Fr + Ff * cd) - vx * r) + q * ((u1 * vx + dl * u0) * c);           /* :298-300 */
    f[5] = p * ((1.0 / iz) * (lf * Ff * cd - lr * Fr)) + q * ((u1 * vx + dl * u0) / L);         /* :305-307 */
    f[6] = u1;                                                     /* :310 */

    memset(Jx, 0, 49 * sizeof(double));
    memset(Ju, 0, 14 * sizeof(double));
    const double Ff_vx = cf2 * (vy + lf * r) / (D * D), Ff_vy = -cf2 / D, Ff_r = -cf2 * lf / D, Ff_d = cf2;
    const double Fr_vx = -Fr / D, Fr_vy = -cr2 / D, Fr_r = cr2 * lr / D;
#define JX(i, j) Jx[(i) * 7 + (j)]
#define JU(i, j) Ju[(i) * 2 + (j)]
    JX(0, 2) = -vx * sp - vy * cp; JX(0, 3) = cp;  JX(0, 4) = -sp;
    JX(1, 2) = vx * cp - vy * sp;  JX(1, 3) = sp;  JX(1, 4) = cp;
    JX(2, 5) = 1.0;
    JX(3, 3) = p * (-sd * Ff_vx / m);
    JX(3, 4) = p * (-sd * Ff_vy / m + r);
    JX(3, 5) = p * (-sd * Ff_r / m + vy);
    JX(3, 6) = -p * (Ff_d * sd + Ff * cd) / m;
    JU(3, 0) = 1.0;
    JX(4, 3) = p * ((Fr_vx + cd * Ff_vx) / m - r) + q * u1 * c;
    JX(4, 4) = p * (Fr_vy + cd * Ff_vy) / m;
    JX(4, 5) = p * ((Fr_r + cd * Ff_r) / m - vx);
    JX(4, 6) = p * (Ff_d * cd - Ff * sd) / m + q * u0 * c;
    JU(4, 0) = q * dl * c;  JU(4, 1) = q * vx * c;
    JX(5, 3) = p * (lf * cd * Ff_vx - lr * Fr_vx) / iz + q * u1 / L;
    JX(5, 4) = p * (lf * cd * Ff_vy - lr * Fr_vy) / iz;
    JX(5, 5) = p * (lf * cd * Ff_r - lr * Fr_r) / iz;
    JX(5, 6) = p * lf * (Ff_d * cd - Ff * sd) / iz + q * u0 / L;
    JU(5, 0) = q * dl / L;  JU(5, 1) = q * vx / L;
    JU(6, 1) = 1.0;

    if (o->model_backend == 2) {
        /* rows 0..2 in curvilinear coordinates x = [s, e_y, e_psi, ...] (A2'; literal form of the bytecode, including
         * the e_y*kappa factor in the heading-error rate):
         *   s'     = (vx cos e_psi - vy sin e_psi) / (1 - e_y kappa)
         *   e_y'   =  vx sin e_psi + vy cos e_psi
         *   e_psi' =  r - e_y kappa s'                                                             */
        double kap, dkap;
        kappa_eval(x[0], &kap, &dkap);
        const double ey = x[1];
        const double vt = vx * cp - vy * sp, vn = vx * sp + vy * cp, den = 1.0 - ey * kap;
        const double sd0 = vt / den;
        f[0] = sd0; f[1] = vn; f[2] = r - ey * kap * sd0;
        for (int j = 0; j < 7; j++) { JX(0, j) = 0.0; JX(1, j) = 0.0; JX(2, j) = 0.0; }
        JX(0, 1) = sd0 * kap / den; JX(0, 2) = -vn / den; JX(0, 3) = cp / den; JX(0, 4) = -sp / den;
        JX(1, 2) = vt; JX(1, 3) = sp; JX(1, 4) = cp;
        JX(2, 1) = -kap * sd0 - ey * kap * JX(0, 1);
        JX(2, 2) = -ey * kap * JX(0, 2);
        JX(2, 3) = -ey * kap * JX(0, 3);
        JX(2, 4) = -ey * kap * JX(0, 4);
        JX(2, 5) = 1.0;
        /* d / d s through kappa(s): zero for a per-node constant curvature */
        JX(0, 0) = sd0 * ey * dkap / den;
        JX(2, 0) = -ey * dkap * sd0 - ey * kap * JX(0, 0);
    }
    if (o->gp_enabled && gp) {
        /* gp_x = gp_state*trigger + x*(1-trigger)   quad_3d_optimizer.py:295 ; z = B_z [x;u]  gp.py:609-630 */
        double z[ORC_DZMAX], mu[ORC_GPOUT_MAX], dmu[ORC_GPOUT_MAX * ORC_DZMAX];
        for (int d = 0; d < o->gp_dz; d++) {
            int fi = o->gp_feat[d];
            if (fi < 7) z[d] = (gp_state ? gp_state[fi] : x[fi]) * trigger + x[fi] * (1.0 - trigger);
            else z[d] = u[fi - 7];
        }
        orc_gp_predict(o, gp, z, mu, dmu);
        for (int j = 0; j < o->gp_nout; j++) {
            int row = o->gp_row[j];
            f[row] += mu[j];                                       /* f + B_x mu   quad_3d_optimizer.py:315 */
            for (int d = 0; d < o->gp_dz; d++) {
                int fi = o->gp_feat[d];
                if (fi < 7) JX(row, fi) += (1.0 - trigger) * dmu[j * o->gp_dz + d];
                else JU(row, fi - 7) += dmu[j * o->gp_dz + d];
            }
        }
    }
#undef JX
#undef JU
}

void orc_ode(const orc_opts *o, const orc_gp *gp, const double *x, const double *u, double p,
             const double *gp_state, double trigger, double *xdot)
{
    double Jx[49], Ju[14];
    orc_model_jac(o, gp, x, u, p, gp_state, trigger, xdot, Jx, Ju);
}

/* variational differential equation: xdot = f, dSx = Jx Sx, dSu = Jx Su + Ju   (row-major dense)
 * mirrors the signature of sim_car_expl_vde_forw (sim_car_expl_vde_forw.c:120) */
static void vde_forw(const orc_opts *o, const orc_gp *gp, const double *x, const double *Sx, const double *Su,
                     const double *u, double p, const double *gp_state, double trigger, double *xdot,
                     double *dSx, double *dSu)
{
    if (o->model_backend == 1 && ref_vde && !o->gp_enabled) {
        /* the reference's own generated code: column-major dense inputs, CCS-sparse outputs
         * (sparsity tables sim_car_expl_vde_forw.c:107-118: dSx 42 nz = rows 0..5 of each column,
         *  dSu 13 nz = col 0 rows 0..5, col 1 rows 0..6) */
        double Sxc[49], Suc[14], o1[42], o2[13], w[244];
        int iw[3];
        for (int i = 0; i < 7; i++) for (int j = 0; j < 7; j++) Sxc[j * 7 + i] = Sx[i * 7 + j];
        for (int i = 0; i < 7; i++) for (int j = 0; j < 2; j++) Suc[j * 7 + i] = Su[i * 2 + j];
        const double *arg[12] = {x, Sxc, Suc, u, &p};
        double *res[10] = {xdot, o1, o2};
        ref_vde(arg, res, iw, w, 0);
        memset(dSx, 0, 49 * sizeof(double));
        memset(dSu, 0, 14 * sizeof(double));
        for (int j = 0; j < 7; j++) for (int i = 0; i < 6; i++) dSx[i * 7 + j] = o1[j * 6 + i];
        for (int i = 0; i < 6; i++) dSu[i * 2 + 0] = o2[i];
        for (int i = 0; i < 7; i++) dSu[i * 2 + 1] = o2[6 + i];
        return;
    }
    double Jx[49], Ju[14];
    orc_model_jac(o, gp, x, u, p, gp_state, trigger, xdot, Jx, Ju);
    for (int i = 0; i < 7; i++) {
        for (int j = 0; j < 7; j++) {
            double s = 0;
            for (int l = 0; l < 7; l++) s += Jx[i * 7 + l] * Sx[l * 7 + j];
            dSx[i * 7 + j] = s;
        }
        for (int j = 0; j < 2; j++) {
            double s = Ju[i * 2 + j];
            for (int l = 0; l < 7; l++) s += Jx[i * 7 + l] * Su[l * 2 + j];
            dSu[i * 2 + j] = s;
        }
    }
}

int orc_rk4_sens(const orc_opts *o, const orc_gp *gp, const double *x, const double *u, double p,
                 const double *gp_state, double trigger, double *xn, double *A, double *B)
{
    /* classic RK4 on [x; Sx; Su], S(0) = [I 0]   (verified against the golden iterate) */
    const double h = o->dt;
    const double a[4] = {0.0, 0.5, 0.5, 1.0}, bw[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
    double Sx0[49] = {0}, Su0[14] = {0};
    for (int i = 0; i < 7; i++) Sx0[i * 7 + i] = 1.0;
    double kx[7] = {0}, kSx[49] = {0}, kSu[14] = {0};
    double ax[7] = {0}, aSx[49] = {0}, aSu[14] = {0};
    for (int s = 0; s < 4; s++) {
        double xs[7], Sxs[49], Sus[14];
        for (int i = 0; i < 7; i++) xs[i] = x[i] + h * a[s] * kx[i];
        for (int i = 0; i < 49; i++) Sxs[i] = Sx0[i] + h * a[s] * kSx[i];
        for (int i = 0; i < 14; i++) Sus[i] = Su0[i] + h * a[s] * kSu[i];
        vde_forw(o, gp, xs, Sxs, Sus, u, p, gp_state, trigger, kx, kSx, kSu);
        for (int i = 0; i < 7; i++) ax[i] += bw[s] * kx[i];
        for (int i = 0; i < 49; i++) aSx[i] += bw[s] * kSx[i];
        for (int i = 0; i < 14; i++) aSu[i] += bw[s] * kSu[i];
    }
    int bad = 0;
    for (int i = 0; i < 7; i++) { xn[i] = x[i] + h * ax[i]; if (!isfinite(xn[i])) bad = 1; }
    for (int i = 0; i < 49; i++) { A[i] = Sx0[i] + h * aSx[i]; if (!isfinite(A[i])) bad = 1; }
    for (int i = 0; i < 14; i++) { B[i] = Su0[i] + h * aSu[i]; if (!isfinite(B[i])) bad = 1; }
    return bad;
}

/* ------------------------------------------------------------------------------------------ prepare -------- */
static int prepare_impl(const orc_opts *o, const orc_gp *gp, const orc_iterate *it, const double *yref,
                        const double *p, const double *kappa, const double *gp_state, orc_lin *lin);
int orc_prepare(const orc_opts *o, const orc_gp *gp, const orc_iterate *it, const double *yref,
                const double *p, const double *gp_state, orc_lin *lin)
{
    return prepare_impl(o, gp, it, yref, p, 0, gp_state, lin);
}
int orc_prepare_frenet(const orc_opts *o, const orc_gp *gp, const orc_iterate *it, const double *yref,
                       const double *p, const double *kappa, const double *gp_state, orc_lin *lin)
{
    return prepare_impl(o, gp, it, yref, p, kappa, gp_state, lin);
}
static int prepare_impl(const orc_opts *o, const orc_gp *gp, const orc_iterate *it, const double *yref,
                        const double *p, const double *kappa, const double *gp_state, orc_lin *lin)
{
    const int N = o->N;
    const double Ts = o->dt;
    int bad = 0;
    for (int k = 0; k < N; k++) {
        const double *xk = it->x + k * 7, *uk = it->u + k * 2;
        double xn[7];
        tls_kappa = kappa ? kappa[k] : 0.0;
        double trig = (o->gp_enabled && o->gp_stage0_trigger && k == 0) ? 1.0 : 0.0;
        bad |= orc_rk4_sens(o, gp, xk, uk, p[k], gp_state, trig, xn, lin->A + k * 49, lin->B + k * 14);
        for (int i = 0; i < 7; i++) lin->b[k * 7 + i] = xn[i] - it->x[(k + 1) * 7 + i];
        /* LINEAR_LS, Vx=[I;0], Vu=[0;I], scaling Ts: grad = Ts * W (y - yref) */
        for (int i = 0; i < 7; i++) lin->q[k * 7 + i] = Ts * o->W[i] * (xk[i] - yref[k * 9 + i]);
        for (int j = 0; j < 2; j++) lin->r[k * 2 + j] = Ts * o->W[7 + j] * (uk[j] - yref[k * 9 + 7 + j]);
    }
    /* terminal: scaling 1 */
    for (int i = 0; i < 7; i++) lin->q[N * 7 + i] = o->We[i] * (it->x[N * 7 + i] - yref[N * 9 + i]);
    return bad;
}

/* ------------------------------------------------------------------------------------------ QP (IPM) ------- */
/* ---- constraint sets ------------------------------------------------------------------------------------------
 * Bounded quantities q = 0..nb-1 of the stage vector z = [u0 u1 | x0..x6] (index idx[q]), each hard (two rows) or soft
 * (slack id soft[q]: two more rows).  Row order per stage = acados / HPIPM order, as in the reference's iterate dumps:
 *     [ lb(q = 0..nb-1) | ub(q = 0..nb-1) | ls(s = 0..ns-1) | us(s = 0..ns-1) ]
 *   con_set 0 (ad_3d_optimizer.py:165-199, sim_car_iterate.json): u0 soft, u1 soft, x6 (delta) hard           -> 10 rows
 *   con_set 1 (fren_ad_3d_optimizer pyc; ad_mpc/debug.json: 12 multipliers and 2 + 2 slacks per stage, lam_ls1 + lam_lbx_delta
 *              = Ts zl at an active steering bound, lam_ubu1 active with no slack row): u0 soft, u1 hard, x1 (e_y) hard,
 *              x6 (delta) soft                                                                                 -> 12 rows
 * Bounds on states do not exist at stage 0 (x0 is eliminated), nor do their slacks; no terminal constraints. */
typedef struct { int nb, ns, nc; int idx[4], soft[4]; double lo[4], hi[4]; int on0[ORC_NC]; } con_desc;
static inline int q_on(const con_desc *d, int k, int q) { return d->idx[q] < 2 || k >= 1; }
/* quantity a row belongs to */
static inline int row_q(const con_desc *d, int c)
{
    if (c < d->nb) return c;
    if (c < 2 * d->nb) return c - d->nb;
    int s = (c < 2 * d->nb + d->ns) ? c - 2 * d->nb : c - 2 * d->nb - d->ns;
    for (int q = 0; q < d->nb; q++) if (d->soft[q] == s) return q;
    return 0;
}
static inline int con_on(const con_desc *d, int k, int c) { return k >= 1 || d->on0[c]; }     /* rows in use at stage k */
static inline __attribute__((always_inline)) void con_get(const orc_opts *o, con_desc *d)
{
    memset(d, 0, sizeof(*d));
    if (o->con_set == 1) {
        d->nb = 4; d->ns = 2;
        d->idx[0] = 0; d->idx[1] = 1; d->idx[2] = 3; d->idx[3] = 8;
        d->soft[0] = 0; d->soft[1] = -1; d->soft[2] = -1; d->soft[3] = 1;
        d->lo[0] = o->lbu[0]; d->hi[0] = o->ubu[0]; d->lo[1] = o->lbu[1]; d->hi[1] = o->ubu[1];
        d->lo[2] = o->lbx2; d->hi[2] = o->ubx2; d->lo[3] = o->lbx; d->hi[3] = o->ubx;
    } else {
        d->nb = 3; d->ns = 2;
        d->idx[0] = 0; d->idx[1] = 1; d->idx[2] = 8;
        d->soft[0] = 0; d->soft[1] = 1; d->soft[2] = -1;
        d->lo[0] = o->lbu[0]; d->hi[0] = o->ubu[0]; d->lo[1] = o->lbu[1]; d->hi[1] = o->ubu[1];
        d->lo[2] = o->lbx; d->hi[2] = o->ubx;
    }
    d->nc = 2 * d->nb + 2 * d->ns;
    for (int c = 0; c < d->nc; c++) d->on0[c] = q_on(d, 0, row_q(d, c));
}
int orc_con_rows(const orc_opts *o) { con_desc d; con_get(o, &d); return d.nc; }
#define RL(q) (q)
#define RU(q) (d->nb + (q))
#define RLS(s) (2 * d->nb + (s))
#define RUS(s) (2 * d->nb + d->ns + (s))

typedef struct {
    /* Newton-step right-hand sides */
    double rgu[ORC_NMAX][2], rgx[ORC_NMAX + 1][7], rgsl[ORC_NMAX][2], rgsu[ORC_NMAX][2];
    double rb[ORC_NMAX][7], rd[ORC_NMAX][ORC_NC], rm[ORC_NMAX][ORC_NC];
    /* factorisation: barrier-modified diagonal of the stage Hessian over z = [u; x] */
    double Hd[ORC_NMAX][9];
    double K[ORC_NMAX][14], Luu[ORC_NMAX][3], P[ORC_NMAX + 1][49];
    /* step */
    double ddu[ORC_NMAX][2], ddx[ORC_NMAX + 1][7], dpi[ORC_NMAX][7], dlam[ORC_NMAX][ORC_NC], dt[ORC_NMAX][ORC_NC],
        dsl[ORC_NMAX][2], dsu[ORC_NMAX][2];
} ipm_ws;

/* value of the bounded quantity q of stage k in the current delta point */
static inline double q_val(const con_desc *d, int q, const double *du, const double *dx)
{
    return (d->idx[q] < 2) ? du[d->idx[q]] : dx[d->idx[q] - 2];
}
/* its linearisation point */
static inline double q_bar(const con_desc *d, int q, const orc_iterate *it, int k)
{
    return (d->idx[q] < 2) ? it->u[k * 2 + d->idx[q]] : it->x[k * 7 + d->idx[q] - 2];
}

/* dlo / dhi [k][q]: bounds in delta form around the iterate */
static inline __attribute__((always_inline)) void ipm_residuals(const orc_opts *o, const con_desc *d, const orc_lin *lin, const double (*dlo)[4],
                          const double (*dhi)[4], const orc_qpsol *s, ipm_ws *w, double res[4], double *mu)
{
    const int N = o->N, nc = d->nc;
    const double Ts = o->dt;
    double ng = 0, nb = 0, nd = 0, nm = 0, summ = 0;
    int ncon = 0;
    for (int k = 0; k <= N; k++) {
        const double *dx = s->dx + k * 7;
        if (k < N) {
            const double *A = lin->A + k * 49, *B = lin->B + k * 14;
            const double *du = s->du + k * 2, *pi = s->pi + k * 7, *lam = s->lam + k * nc, *t = s->t + k * nc;
            for (int j = 0; j < 2; j++) {
                double g = Ts * o->W[7 + j] * du[j] + lin->r[k * 2 + j];
                for (int l = 0; l < 7; l++) g += B[l * 2 + j] * pi[l];
                for (int q = 0; q < d->nb; q++) if (d->idx[q] == j) g += -lam[RL(q)] + lam[RU(q)];
                w->rgu[k][j] = g;
                ng = nanmax(ng, fabs(g));
            }
            for (int c = 0; c < nc; c++) w->rd[k][c] = 0.0;
            for (int q = 0; q < d->nb; q++) {
                if (!q_on(d, k, q)) continue;
                const double v = q_val(d, q, du, dx);
                const int sq = d->soft[q];
                if (sq >= 0) {
                    const double sl = s->sl[k * 2 + sq], su = s->su[k * 2 + sq];
                    w->rgsl[k][sq] = Ts * o->zl[sq] + Ts * o->Zl[sq] * sl - lam[RL(q)] - lam[RLS(sq)];
                    w->rgsu[k][sq] = Ts * o->zu[sq] + Ts * o->Zu[sq] * su - lam[RU(q)] - lam[RUS(sq)];
                    w->rd[k][RL(q)] = t[RL(q)] - (v - dlo[k][q] + sl);
                    w->rd[k][RU(q)] = t[RU(q)] - (dhi[k][q] - v + su);
                    w->rd[k][RLS(sq)] = t[RLS(sq)] - sl;
                    w->rd[k][RUS(sq)] = t[RUS(sq)] - su;
                    ng = nanmax(ng, nanmax(fabs(w->rgsl[k][sq]), fabs(w->rgsu[k][sq])));
                } else {
                    w->rd[k][RL(q)] = t[RL(q)] - (v - dlo[k][q]);
                    w->rd[k][RU(q)] = t[RU(q)] - (dhi[k][q] - v);
                }
            }
            for (int i = 0; i < 7; i++) {
                double v = lin->b[k * 7 + i] - s->dx[(k + 1) * 7 + i];
                for (int l = 0; l < 7; l++) v += A[i * 7 + l] * dx[l];
                for (int j = 0; j < 2; j++) v += B[i * 2 + j] * du[j];
                w->rb[k][i] = v;
                nb = nanmax(nb, fabs(v));
            }
            for (int c = 0; c < nc; c++) {
                if (!con_on(d, k, c)) { w->rm[k][c] = 0; continue; }
                w->rm[k][c] = lam[c] * t[c];
                nd = nanmax(nd, fabs(w->rd[k][c]));
                nm = nanmax(nm, fabs(w->rm[k][c]));
                summ += w->rm[k][c];
                ncon++;
            }
        }
        if (k >= 1) {
            for (int i = 0; i < 7; i++) {
                double Qd = (k < N) ? Ts * o->W[i] : o->We[i];
                double g = Qd * dx[i] + lin->q[k * 7 + i] - s->pi[(k - 1) * 7 + i];
                if (k < N) {
                    const double *A = lin->A + k * 49, *pi = s->pi + k * 7;
                    for (int l = 0; l < 7; l++) g += A[l * 7 + i] * pi[l];
                    for (int q = 0; q < d->nb; q++)
                        if (d->idx[q] == 2 + i) g += -s->lam[k * nc + RL(q)] + s->lam[k * nc + RU(q)];
                }
                w->rgx[k][i] = g;
                ng = nanmax(ng, fabs(g));
            }
        }
    }
    res[0] = ng; res[1] = nb; res[2] = nd; res[3] = nm;
    *mu = summ / ncon;
}

/* barrier scalings of one bounded quantity */
typedef struct { double Sl, Su, Dl, Du; } q_scal;
static inline void q_scaling(const orc_opts *o, const con_desc *d, int q, const double *lam, const double *t, q_scal *S)
{
    const double Ts = o->dt;
    const int sq = d->soft[q];
    S->Sl = lam[RL(q)] / t[RL(q)]; S->Su = lam[RU(q)] / t[RU(q)];
    if (sq >= 0) {
        const double Ssl = lam[RLS(sq)] / t[RLS(sq)], Ssu = lam[RUS(sq)] / t[RUS(sq)];
        S->Dl = Ts * o->Zl[sq] + S->Sl + Ssl; S->Du = Ts * o->Zu[sq] + S->Su + Ssu;
    } else { S->Dl = 0; S->Du = 0; }
}

/* Riccati factorisation for the barrier-modified Hessian (matrix part only). */
static inline __attribute__((always_inline)) void ipm_factor(const orc_opts *o, const con_desc *d, const orc_lin *lin, const orc_qpsol *s, ipm_ws *w)
{
    const int N = o->N, nc = d->nc;
    const double Ts = o->dt;
    /* barrier terms with soft-bound slacks eliminated */
    for (int k = 0; k < N; k++) {
        const double *lam = s->lam + k * nc, *t = s->t + k * nc;
        for (int z = 0; z < 9; z++) w->Hd[k][z] = (z < 2) ? Ts * o->W[7 + z] : Ts * o->W[z - 2];
        for (int q = 0; q < d->nb; q++) {
            if (!q_on(d, k, q)) continue;
            q_scal S; q_scaling(o, d, q, lam, t, &S);
            if (d->soft[q] >= 0) w->Hd[k][d->idx[q]] = w->Hd[k][d->idx[q]] + S.Sl * (1.0 - S.Sl / S.Dl) + S.Su * (1.0 - S.Su / S.Du);
            else w->Hd[k][d->idx[q]] = w->Hd[k][d->idx[q]] + (S.Sl + S.Su);
        }
    }
    double *P = w->P[N];
    memset(P, 0, 49 * sizeof(double));
    for (int i = 0; i < 7; i++) P[i * 7 + i] = o->We[i];
    for (int k = N - 1; k >= 0; k--) {
        const double *A = lin->A + k * 49, *B = lin->B + k * 14;
        const double *Pn = w->P[k + 1];
        double BA[7][9], PBA[7][9], G[9][9];
        for (int i = 0; i < 7; i++) {
            BA[i][0] = B[i * 2]; BA[i][1] = B[i * 2 + 1];
            for (int j = 0; j < 7; j++) BA[i][2 + j] = A[i * 7 + j];
        }
        for (int i = 0; i < 7; i++) for (int c = 0; c < 9; c++) {
            double v = 0;
            for (int l = 0; l < 7; l++) v += Pn[i * 7 + l] * BA[l][c];
            PBA[i][c] = v;
        }
        for (int a = 0; a < 9; a++) for (int c = 0; c < 9; c++) {
            double v = 0;
            for (int l = 0; l < 7; l++) v += BA[l][a] * PBA[l][c];
            G[a][c] = v;
        }
        for (int z = 0; z < 9; z++) G[z][z] += w->Hd[k][z];
        /* Cholesky of the 2x2 input block */
        double l00 = sqrt(G[0][0] + o->reg), l10 = G[1][0] / l00, l11 = sqrt(G[1][1] + o->reg - l10 * l10);
        w->Luu[k][0] = l00; w->Luu[k][1] = l10; w->Luu[k][2] = l11;
        /* K = -Guu^-1 Gux */
        for (int j = 0; j < 7; j++) {
            double y0 = G[0][2 + j] / l00, y1 = (G[1][2 + j] - l10 * y0) / l11;
            double k1 = y1 / l11, k0 = (y0 - l10 * k1) / l00;
            w->K[k][j] = -k0; w->K[k][7 + j] = -k1;
        }
        double *Pk = w->P[k];
        for (int i = 0; i < 7; i++) for (int j = 0; j < 7; j++)
            Pk[i * 7 + j] = G[2 + i][2 + j] + G[2 + i][0] * w->K[k][j] + G[2 + i][1] * w->K[k][7 + j];
        for (int i = 0; i < 7; i++) for (int j = 0; j < i; j++) {    /* symmetrise */
            double m = 0.5 * (Pk[i * 7 + j] + Pk[j * 7 + i]);
            Pk[i * 7 + j] = Pk[j * 7 + i] = m;
        }
    }
}

/* Solve for the Newton step given the complementarity right-hand side w->rm (vector part of the Riccati). */
static inline __attribute__((always_inline)) void ipm_solve(const orc_opts *o, const con_desc *d, const orc_lin *lin, const orc_qpsol *s, ipm_ws *w)
{
    const int N = o->N, nc = d->nc;
    double gl[ORC_NMAX][ORC_NC], rt[ORC_NMAX][2], qx[ORC_NMAX][7], cl[ORC_NMAX][2], cu[ORC_NMAX][2];
    q_scal S[ORC_NMAX][4];
    for (int k = 0; k < N; k++) {
        const double *lam = s->lam + k * nc, *t = s->t + k * nc;
        for (int c = 0; c < nc; c++) gl[k][c] = con_on(d, k, c) ? (w->rm[k][c] - lam[c] * w->rd[k][c]) / t[c] : 0.0;
        for (int j = 0; j < 2; j++) rt[k][j] = w->rgu[k][j];
        for (int i = 0; i < 7; i++) qx[k][i] = (k >= 1) ? w->rgx[k][i] : 0.0;
        for (int q = 0; q < d->nb; q++) {
            if (!q_on(d, k, q)) continue;
            q_scaling(o, d, q, lam, t, &S[k][q]);
            double *gq = (d->idx[q] < 2) ? &rt[k][d->idx[q]] : &qx[k][d->idx[q] - 2];
            const int sq = d->soft[q];
            if (sq >= 0) {
                cl[k][sq] = w->rgsl[k][sq] + gl[k][RL(q)] + gl[k][RLS(sq)];
                cu[k][sq] = w->rgsu[k][sq] + gl[k][RU(q)] + gl[k][RUS(sq)];
                *gq = *gq + (gl[k][RL(q)] - S[k][q].Sl * cl[k][sq] / S[k][q].Dl) - (gl[k][RU(q)] - S[k][q].Su * cu[k][sq] / S[k][q].Du);
            } else {
                *gq = *gq + gl[k][RL(q)] - gl[k][RU(q)];
            }
        }
    }
    /* backward vector recursion */
    double pv[ORC_NMAX + 1][7], kf[ORC_NMAX][2];
    for (int i = 0; i < 7; i++) pv[N][i] = w->rgx[N][i];
    for (int k = N - 1; k >= 0; k--) {
        const double *A = lin->A + k * 49, *B = lin->B + k * 14, *Pn = w->P[k + 1];
        double h[7], gu[2], gx[7];
        for (int i = 0; i < 7; i++) {
            double v = pv[k + 1][i];
            for (int l = 0; l < 7; l++) v += Pn[i * 7 + l] * w->rb[k][l];
            h[i] = v;
        }
        for (int j = 0; j < 2; j++) {
            double v = rt[k][j];
            for (int l = 0; l < 7; l++) v += B[l * 2 + j] * h[l];
            gu[j] = v;
        }
        for (int i = 0; i < 7; i++) {
            double v = qx[k][i];
            for (int l = 0; l < 7; l++) v += A[l * 7 + i] * h[l];
            gx[i] = v;
        }
        const double l00 = w->Luu[k][0], l10 = w->Luu[k][1], l11 = w->Luu[k][2];
        double y0 = gu[0] / l00, y1 = (gu[1] - l10 * y0) / l11;
        double k1 = y1 / l11, k0 = (y0 - l10 * k1) / l00;
        kf[k][0] = -k0; kf[k][1] = -k1;
        /* p_k = g_x + K^T g_u   (G_xu kf = -G_xu Guu^-1 g_u = K^T g_u) */
        for (int i = 0; i < 7; i++) pv[k][i] = gx[i] + w->K[k][i] * gu[0] + w->K[k][7 + i] * gu[1];
    }
    /* forward rollout */
    for (int i = 0; i < 7; i++) w->ddx[0][i] = 0.0;
    for (int k = 0; k < N; k++) {
        const double *A = lin->A + k * 49, *B = lin->B + k * 14, *Pn = w->P[k + 1];
        for (int j = 0; j < 2; j++) {
            double v = kf[k][j];
            for (int l = 0; l < 7; l++) v += w->K[k][j * 7 + l] * w->ddx[k][l];
            w->ddu[k][j] = v;
        }
        for (int i = 0; i < 7; i++) {
            double v = w->rb[k][i];
            for (int l = 0; l < 7; l++) v += A[i * 7 + l] * w->ddx[k][l];
            for (int j = 0; j < 2; j++) v += B[i * 2 + j] * w->ddu[k][j];
            w->ddx[k + 1][i] = v;
        }
        for (int i = 0; i < 7; i++) {
            double v = pv[k + 1][i];
            for (int l = 0; l < 7; l++) v += Pn[i * 7 + l] * w->ddx[k + 1][l];
            w->dpi[k][i] = v;
        }
    }
    /* recover slack, t and lambda steps */
    for (int k = 0; k < N; k++) {
        const double *lam = s->lam + k * nc, *t = s->t + k * nc;
        for (int c = 0; c < nc; c++) w->dt[k][c] = 0.0;
        for (int sq = 0; sq < 2; sq++) { w->dsl[k][sq] = 0.0; w->dsu[k][sq] = 0.0; }
        for (int q = 0; q < d->nb; q++) {
            if (!q_on(d, k, q)) continue;
            const double dv = q_val(d, q, w->ddu[k], w->ddx[k]);
            const int sq = d->soft[q];
            if (sq >= 0) {
                w->dsl[k][sq] = -(cl[k][sq] + S[k][q].Sl * dv) / S[k][q].Dl;
                w->dsu[k][sq] = -(cu[k][sq] - S[k][q].Su * dv) / S[k][q].Du;
                w->dt[k][RL(q)] = dv + w->dsl[k][sq] - w->rd[k][RL(q)];
                w->dt[k][RU(q)] = -dv + w->dsu[k][sq] - w->rd[k][RU(q)];
                w->dt[k][RLS(sq)] = w->dsl[k][sq] - w->rd[k][RLS(sq)];
                w->dt[k][RUS(sq)] = w->dsu[k][sq] - w->rd[k][RUS(sq)];
            } else {
                w->dt[k][RL(q)] = dv - w->rd[k][RL(q)];
                w->dt[k][RU(q)] = -dv - w->rd[k][RU(q)];
            }
        }
        for (int c = 0; c < nc; c++)
            w->dlam[k][c] = con_on(d, k, c) ? -(w->rm[k][c] + lam[c] * w->dt[k][c]) / t[c] : 0.0;
    }
}

static inline __attribute__((always_inline)) double ipm_alpha(const orc_opts *o, const con_desc *d, const orc_qpsol *s, const ipm_ws *w)
{
    const int nc = d->nc;
    double a = 1.0;
    for (int k = 0; k < o->N; k++) for (int c = 0; c < nc; c++) {
        if (!con_on(d, k, c)) continue;
        double l = s->lam[k * nc + c], t = s->t[k * nc + c], dl = w->dlam[k][c], dt = w->dt[k][c];
        if (dl < 0 && -l / dl < a) a = -l / dl;
        if (dt < 0 && -t / dt < a) a = -t / dt;
    }
    return a;
}

static inline __attribute__((always_inline)) void delta_bounds(const con_desc *d, const orc_iterate *it, int N, double (*dlo)[4], double (*dhi)[4])
{
    for (int k = 0; k < N; k++) for (int q = 0; q < d->nb; q++) {
        const double bar = q_bar(d, q, it, k);
        dlo[k][q] = d->lo[q] - bar;
        dhi[k][q] = d->hi[q] - bar;
    }
}

/* Constraint-side KKT relations of an iterate, model-independent (used to pin the row order / softness mapping of a constraint
 * set against the reference's own iterate dumps): out[0] = max |t - constraint function| over the rows in use,
 * out[1] = max |stationarity w.r.t. the slacks| (Ts z + Ts Z s - lam_bound - lam_slack), out[2] = max |lam t|. */
void orc_con_check(const orc_opts *o, const orc_iterate *it, double out[3])
{
    con_desc dd; con_get(o, &dd);
    const con_desc *d = &dd;
    const int N = o->N, nc = d->nc;
    const double Ts = o->dt;
    double e0 = 0, e1 = 0, e2 = 0;
    for (int k = 0; k < N; k++) {
        const double *lam = it->lam + k * nc, *t = it->t + k * nc;
        for (int q = 0; q < d->nb; q++) {
            if (!q_on(d, k, q)) continue;
            const double v = q_bar(d, q, it, k);
            const int sq = d->soft[q];
            const double sl = (sq >= 0) ? it->sl[k * 2 + sq] : 0.0, su = (sq >= 0) ? it->su[k * 2 + sq] : 0.0;
            e0 = nanmax(e0, fabs(t[RL(q)] - (v - d->lo[q] + sl)));
            e0 = nanmax(e0, fabs(t[RU(q)] - (d->hi[q] - v + su)));
            e2 = nanmax(e2, nanmax(fabs(lam[RL(q)] * t[RL(q)]), fabs(lam[RU(q)] * t[RU(q)])));
            if (sq >= 0) {
                e0 = nanmax(e0, nanmax(fabs(t[RLS(sq)] - sl), fabs(t[RUS(sq)] - su)));
                e1 = nanmax(e1, fabs(Ts * o->zl[sq] + Ts * o->Zl[sq] * sl - lam[RL(q)] - lam[RLS(sq)]));
                e1 = nanmax(e1, fabs(Ts * o->zu[sq] + Ts * o->Zu[sq] * su - lam[RU(q)] - lam[RUS(sq)]));
                e2 = nanmax(e2, nanmax(fabs(lam[RLS(sq)] * t[RLS(sq)]), fabs(lam[RUS(sq)] * t[RUS(sq)])));
            }
        }
    }
    out[0] = e0; out[1] = e1; out[2] = e2;
}

static inline __attribute__((always_inline)) int qp_solve_body(const orc_opts *o, const con_desc *d, const orc_lin *lin, const orc_iterate *it,
                                                              const double *x0, orc_qpsol *s, orc_stats *st)
{
    const int N = o->N;
    const int nc = d->nc;
    ipm_ws *w = (ipm_ws *)malloc(sizeof(ipm_ws));
    double dlo[ORC_NMAX][4], dhi[ORC_NMAX][4];
    memset(s, 0, sizeof(*s));
    delta_bounds(d, it, N, dlo, dhi);            /* bounds in delta form around the iterate */
    /* cold start (qp_solver_warm_start 0, sim_car_acados_ocp.json:885): primal 0 pushed thr0 inside its box */
    for (int i = 0; i < 7; i++) s->dx[i] = x0[i] - it->x[i];     /* lbx_0 = ubx_0 = x0, eliminated (nbxe_0) */
    for (int k = 0; k < N; k++) {
        double *t = s->t + k * nc, *lam = s->lam + k * nc;
        for (int c = 0; c < nc; c++) { t[c] = 1.0; lam[c] = 0.0; }
        for (int q = 0; q < d->nb; q++) {
            if (!q_on(d, k, q)) continue;
            double lo = dlo[k][q], hi = dhi[k][q];
            double v = 0.0;
            if (v - lo < o->thr0) {
                if (hi - v < o->thr0) v = 0.5 * (lo + hi);
                else v = lo + o->thr0;
            } else if (hi - v < o->thr0) v = hi - o->thr0;
            if (d->idx[q] < 2) s->du[k * 2 + d->idx[q]] = v; else s->dx[k * 7 + d->idx[q] - 2] = v;
            t[RL(q)] = fmax(o->thr0, v - lo);
            t[RU(q)] = fmax(o->thr0, hi - v);
            if (d->soft[q] >= 0) { t[RLS(d->soft[q])] = o->thr0; t[RUS(d->soft[q])] = o->thr0; }
        }
        for (int c = 0; c < nc; c++) if (con_on(d, k, c)) lam[c] = o->mu0 / t[c];
    }
    int status = 1, iter = 0;
    double res[4] = {0}, mu = 0;
    for (iter = 0;; iter++) {
        ipm_residuals(o, d, lin, dlo, dhi, s, w, res, &mu);
        if (!(isfinite(res[0]) && isfinite(res[1]) && isfinite(res[2]) && isfinite(res[3]))) { status = 3; break; }
        if (res[0] < o->tol_stat && res[1] < o->tol_eq && res[2] < o->tol_ineq && res[3] < o->tol_comp) {
            status = 0; break;
        }
        if (iter >= o->iter_max) { status = 1; break; }
        /* predictor (affine scaling) */
        ipm_factor(o, d, lin, s, w);
        ipm_solve(o, d, lin, s, w);
        double a_aff = ipm_alpha(o, d, s, w);
        double mu_aff = 0; int ncon = 0;
        for (int k = 0; k < N; k++) for (int c = 0; c < nc; c++) if (con_on(d, k, c)) {
            mu_aff += (s->lam[k * nc + c] + a_aff * w->dlam[k][c]) * (s->t[k * nc + c] + a_aff * w->dt[k][c]);
            ncon++;
        }
        mu_aff /= ncon;
        double sigma = mu_aff / mu; sigma = sigma * sigma * sigma;
        /* corrector */
        for (int k = 0; k < N; k++) for (int c = 0; c < nc; c++) if (con_on(d, k, c))
            w->rm[k][c] = s->lam[k * nc + c] * s->t[k * nc + c] + w->dlam[k][c] * w->dt[k][c] - sigma * mu;
        ipm_solve(o, d, lin, s, w);
        double alpha = ipm_alpha(o, d, s, w);
        if (alpha < o->alpha_min) { status = 2; break; }
        if (alpha < 1.0) alpha *= 0.995;
        for (int k = 0; k < N; k++) {
            for (int j = 0; j < 2; j++) {
                s->du[k * 2 + j] += alpha * w->ddu[k][j];
                s->sl[k * 2 + j] += alpha * w->dsl[k][j];
                s->su[k * 2 + j] += alpha * w->dsu[k][j];
            }
            for (int i = 0; i < 7; i++) {
                s->dx[(k + 1) * 7 + i] += alpha * w->ddx[k + 1][i];
                s->pi[k * 7 + i] += alpha * w->dpi[k][i];
            }
            for (int c = 0; c < nc; c++) if (con_on(d, k, c)) {
                s->lam[k * nc + c] = fmax(s->lam[k * nc + c] + alpha * w->dlam[k][c], o->lam_min);
                s->t[k * nc + c] = fmax(s->t[k * nc + c] + alpha * w->dt[k][c], o->t_min);
            }
        }
    }
    free(w);
    /* hpipm {0 ok,1 maxiter,2 minstep,3 nan} -> acados {0,2,3,1}   [EXT]; SURVEY 8 A7 */
    static const int map[4] = {0, 2, 3, 1};
    if (st) {
        st->qp_status = map[status];
        st->qp_iter = iter;
        memcpy(st->res, res, sizeof res);
        double m = 0;
        for (int i = 0; i < (N + 1) * 7; i++) m = fmax(m, fabs(s->dx[i]));
        for (int i = 0; i < N * 2; i++) m = fmax(m, fabs(s->du[i]));
        st->step_inf = m;
    }
    return status;
}
/* one specialisation per constraint set: the descriptor's structure is a compile-time constant inside each (the generic code then
 * costs what the hard-coded version did) */
static int qp_solve_set0(const orc_opts *o, const orc_lin *lin, const orc_iterate *it, const double *x0, orc_qpsol *s, orc_stats *st)
{
    con_desc d; orc_opts oo = *o; oo.con_set = 0; con_get(&oo, &d);
    return qp_solve_body(o, &d, lin, it, x0, s, st);
}
static int qp_solve_set1(const orc_opts *o, const orc_lin *lin, const orc_iterate *it, const double *x0, orc_qpsol *s, orc_stats *st)
{
    con_desc d; orc_opts oo = *o; oo.con_set = 1; con_get(&oo, &d);
    return qp_solve_body(o, &d, lin, it, x0, s, st);
}
int orc_qp_solve(const orc_opts *o, const orc_lin *lin, const orc_iterate *it, const double *x0,
                 orc_qpsol *s, orc_stats *st)
{
    return (o->con_set == 1) ? qp_solve_set1(o, lin, it, x0, s, st) : qp_solve_set0(o, lin, it, x0, s, st);
}

/* ------------------------------------------------------------------------------------------ RTI step ------- */
static int rti_step_impl(const orc_opts *o, const orc_gp *gp, const double *x0, const double *yref,
                         const double *p, const double *kappa, const double *gp_state, orc_iterate *it, orc_stats *st);
int orc_rti_step(const orc_opts *o, const orc_gp *gp, const double *x0, const double *yref,
                 const double *p, const double *gp_state, orc_iterate *it, orc_stats *st)
{
    return rti_step_impl(o, gp, x0, yref, p, 0, gp_state, it, st);
}
int orc_rti_step_frenet(const orc_opts *o, const orc_gp *gp, const double *x0, const double *yref,
                        const double *p, const double *kappa, const double *gp_state, orc_iterate *it, orc_stats *st)
{
    return rti_step_impl(o, gp, x0, yref, p, kappa, gp_state, it, st);
}
static int rti_step_impl(const orc_opts *o, const orc_gp *gp, const double *x0, const double *yref,
                         const double *p, const double *kappa, const double *gp_state, orc_iterate *it, orc_stats *st)
{
    const int N = o->N;
    orc_lin *lin = (orc_lin *)malloc(sizeof(orc_lin));
    orc_qpsol *sol = (orc_qpsol *)malloc(sizeof(orc_qpsol));
    orc_stats local;
    if (!st) st = &local;
    memset(st, 0, sizeof(*st));
    const double *gps = gp_state ? gp_state : x0;          /* quad_3d_optimizer.py:549 */
    int bad = prepare_impl(o, gp, it, yref, p, kappa, gps, lin);
    if (bad) { st->status = 1; free(lin); free(sol); return 1; }     /* ACADOS_FAILURE: NaN in linearisation */
    orc_qp_solve(o, lin, it, x0, sol, st);
    /* RTI tolerates QP maxiter; anything else is ACADOS_QP_FAILURE (4)  [EXT] */
    st->status = (st->qp_status == 0 || st->qp_status == 2) ? 0 : 4;
    if (st->status == 0) {
        /* full step (step_length 1, fixed_step): primal += delta; duals <- QP duals */
        for (int i = 0; i < (N + 1) * 7; i++) it->x[i] += sol->dx[i];
        for (int i = 0; i < N * 2; i++) it->u[i] += sol->du[i];
        memcpy(it->pi, sol->pi, sizeof(double) * N * 7);
        memcpy(it->lam, sol->lam, sizeof(double) * N * orc_con_rows(o));
        memcpy(it->t, sol->t, sizeof(double) * N * orc_con_rows(o));
        memcpy(it->sl, sol->sl, sizeof(double) * N * 2);
        memcpy(it->su, sol->su, sizeof(double) * N * 2);
    }
    free(lin); free(sol);
    return st->status;
}

/* ------------------------------------------------------------------------------------------ full SQP ------- */
/* NLP KKT residual norms (stat, eq, ineq, comp) of the iterate against a fresh linearisation: the IPM residual
 * function evaluated at a zero step with the iterate's own multipliers and slacks (acados ocp_nlp_res_compute [EXT]:
 * res_ineq = constraint function + t, res_comp = lam .* t). */
void orc_nlp_residuals(const orc_opts *o, const orc_lin *lin, const orc_iterate *it, const double *x0, double res[4])
{
    const int N = o->N;
    con_desc dd; con_get(o, &dd);
    const con_desc *d = &dd;
    ipm_ws *w = (ipm_ws *)malloc(sizeof(ipm_ws));
    orc_qpsol *s = (orc_qpsol *)calloc(1, sizeof(orc_qpsol));
    double dlo[ORC_NMAX][4], dhi[ORC_NMAX][4], mu;
    delta_bounds(d, it, N, dlo, dhi);
    memcpy(s->pi, it->pi, sizeof(double) * N * 7);
    memcpy(s->lam, it->lam, sizeof(double) * N * d->nc);
    memcpy(s->t, it->t, sizeof(double) * N * d->nc);
    memcpy(s->sl, it->sl, sizeof(double) * N * 2);
    memcpy(s->su, it->su, sizeof(double) * N * 2);
    ipm_residuals(o, d, lin, dlo, dhi, s, w, res, &mu);
    /* the initial-state mismatch is an equality residual of the NLP (x_0 = x0 is a constraint of the OCP) */
    for (int i = 0; i < 7; i++) res[1] = nanmax(res[1], fabs(x0[i] - it->x[i]));
    free(w); free(s);
}

/* Full SQP (nlp_solver_type "SQP", create_ros_ad_mpc.py:47-51 point-reference mode; acados ocp_nlp_sqp [EXT]):
 * repeat { linearise; stop with status 0 when the four NLP residuals are below tol; solve the QP; full step }.
 * Returns the acados status: 0 converged, 1 NaN in the linearisation, 2 max_iter reached, 4 QP failure. */
static int sqp_solve_impl(const orc_opts *o, const orc_gp *gp, const double *x0, const double *yref, const double *p,
                          const double *kappa, const double *gp_state, orc_iterate *it, int max_iter, const double tol[4],
                          int *sqp_iter, double res_out[4]);
int orc_sqp_solve(const orc_opts *o, const orc_gp *gp, const double *x0, const double *yref, const double *p,
                  const double *gp_state, orc_iterate *it, int max_iter, const double tol[4], int *sqp_iter,
                  double res_out[4])
{
    return sqp_solve_impl(o, gp, x0, yref, p, 0, gp_state, it, max_iter, tol, sqp_iter, res_out);
}
static int sqp_solve_impl(const orc_opts *o, const orc_gp *gp, const double *x0, const double *yref, const double *p,
                          const double *kappa, const double *gp_state, orc_iterate *it, int max_iter, const double tol[4],
                          int *sqp_iter, double res_out[4])
{
    const int N = o->N;
    orc_lin *lin = (orc_lin *)malloc(sizeof(orc_lin));
    orc_qpsol *sol = (orc_qpsol *)malloc(sizeof(orc_qpsol));
    const double *gps = gp_state ? gp_state : x0;
    int status = 2, iter = 0;
    double res[4] = {0, 0, 0, 0};
    for (iter = 0; iter < max_iter; iter++) {
        if (prepare_impl(o, gp, it, yref, p, kappa, gps, lin)) { status = 1; break; }
        orc_nlp_residuals(o, lin, it, x0, res);
        if (res[0] < tol[0] && res[1] < tol[1] && res[2] < tol[2] && res[3] < tol[3]) { status = 0; break; }
        orc_stats st;
        orc_qp_solve(o, lin, it, x0, sol, &st);
        if (!(st.qp_status == 0 || st.qp_status == 2)) { status = 4; break; }
        for (int i = 0; i < (N + 1) * 7; i++) it->x[i] += sol->dx[i];
        for (int i = 0; i < N * 2; i++) it->u[i] += sol->du[i];
        memcpy(it->pi, sol->pi, sizeof(double) * N * 7);
        memcpy(it->lam, sol->lam, sizeof(double) * N * orc_con_rows(o));
        memcpy(it->t, sol->t, sizeof(double) * N * orc_con_rows(o));
        memcpy(it->sl, sol->sl, sizeof(double) * N * 2);
        memcpy(it->su, sol->su, sizeof(double) * N * 2);
    }
    if (sqp_iter) *sqp_iter = iter;
    if (res_out) memcpy(res_out, res, sizeof res);
    free(lin); free(sol);
    return status;
}

int orc_sqp_batch_frenet(const orc_opts *o, const orc_gp *gp, int B, const double *x0, const double *yref,
                         const double *p, const double *kappa, const double *gp_state, double *xit, double *uit, int max_iter,
                         const double *tol, int *status, int *sqp_iter, double *res, int nthreads);
int orc_sqp_batch(const orc_opts *o, const orc_gp *gp, int B, const double *x0, const double *yref,
                  const double *p, const double *gp_state, double *xit, double *uit, int max_iter, const double *tol,
                  int *status, int *sqp_iter, double *res, int nthreads)
{
    return orc_sqp_batch_frenet(o, gp, B, x0, yref, p, 0, gp_state, xit, uit, max_iter, tol, status, sqp_iter, res, nthreads);
}
int orc_sqp_batch_frenet(const orc_opts *o, const orc_gp *gp, int B, const double *x0, const double *yref,
                         const double *p, const double *kappa, const double *gp_state, double *xit, double *uit, int max_iter,
                         const double *tol, int *status, int *sqp_iter, double *res, int nthreads)
{
    const int N = o->N;
    const size_t ny = (size_t)N * 9 + 7;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        orc_iterate *it = (orc_iterate *)malloc(sizeof(orc_iterate));
#pragma omp for schedule(dynamic, 4)
        for (int b = 0; b < B; b++) {
            memset(it, 0, sizeof(*it));
            memcpy(it->x, xit + (size_t)b * (N + 1) * 7, sizeof(double) * (N + 1) * 7);
            memcpy(it->u, uit + (size_t)b * N * 2, sizeof(double) * N * 2);
            int si = 0;
            double r4[4];
            batch_spline_install(b);
            int stt = sqp_solve_impl(o, gp, x0 + (size_t)b * 7, yref + b * ny, p + (size_t)b * N, kappa ? kappa + (size_t)b * N : 0,
                                     gp_state ? gp_state + (size_t)b * 7 : 0, it, max_iter, tol, &si, r4);
            memcpy(xit + (size_t)b * (N + 1) * 7, it->x, sizeof(double) * (N + 1) * 7);
            memcpy(uit + (size_t)b * N * 2, it->u, sizeof(double) * N * 2);
            if (status) status[b] = stt;
            if (sqp_iter) sqp_iter[b] = si;
            if (res) memcpy(res + (size_t)b * 4, r4, sizeof r4);
        }
        free(it);
    }
    return 0;
}

int orc_rti_batch_frenet(const orc_opts *o, const orc_gp *gp, int B, const double *x0, const double *yref,
                         const double *p, const double *kappa, const double *gp_state, double *xit, double *uit,
                         double *piout, int *status, int *qp_status, int *qp_iter, int nthreads);
int orc_rti_batch(const orc_opts *o, const orc_gp *gp, int B, const double *x0, const double *yref,
                  const double *p, const double *gp_state, double *xit, double *uit, double *piout,
                  int *status, int *qp_status, int *qp_iter, int nthreads)
{
    return orc_rti_batch_frenet(o, gp, B, x0, yref, p, 0, gp_state, xit, uit, piout, status, qp_status, qp_iter, nthreads);
}
/* kappa[B][N]: path curvature at every shooting node (Frenet variant), or NULL */
int orc_rti_batch_frenet(const orc_opts *o, const orc_gp *gp, int B, const double *x0, const double *yref,
                         const double *p, const double *kappa, const double *gp_state, double *xit, double *uit,
                         double *piout, int *status, int *qp_status, int *qp_iter, int nthreads)
{
    const int N = o->N;
    const size_t ny = (size_t)N * 9 + 7;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        orc_iterate *it = (orc_iterate *)malloc(sizeof(orc_iterate));
#pragma omp for schedule(dynamic, 16)
        for (int b = 0; b < B; b++) {
            memset(it, 0, sizeof(*it));
            memcpy(it->x, xit + (size_t)b * (N + 1) * 7, sizeof(double) * (N + 1) * 7);
            memcpy(it->u, uit + (size_t)b * N * 2, sizeof(double) * N * 2);
            orc_stats st;
            batch_spline_install(b);
            rti_step_impl(o, gp, x0 + (size_t)b * 7, yref + b * ny, p + (size_t)b * N, kappa ? kappa + (size_t)b * N : 0,
                          gp_state ? gp_state + (size_t)b * 7 : 0, it, &st);
            memcpy(xit + (size_t)b * (N + 1) * 7, it->x, sizeof(double) * (N + 1) * 7);
            memcpy(uit + (size_t)b * N * 2, it->u, sizeof(double) * N * 2);
            if (piout) memcpy(piout + (size_t)b * N * 7, it->pi, sizeof(double) * N * 7);
            if (status) status[b] = st.status;
            if (qp_status) qp_status[b] = st.qp_status;
            if (qp_iter) qp_iter[b] = st.qp_iter;
        }
        free(it);
    }
    return 0;
}
