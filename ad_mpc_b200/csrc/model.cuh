// model.cuh -- bicycle model f(x,u,p), its hand-derived sparse Jacobian and the GP residual sweep; shared by the
// preparation kernel (prepare.cu) and the closed-loop plant step (closedloop.cu).
// Reference: ad_mpc/ad_3d_optimizer.py:268-310 (model), model_fitting/gp.py:117-165,446-460 (GP), SURVEY Appendix B.
#pragma once
#include "common.cuh"

// exp(x) for x <= 0 (RBF kernel argument -0.5 * squared distance): Cody-Waite reduction x = n ln2 + r, |r| <= ln2/2,
// degree-13 Taylor/Horner (truncation 4e-18 relative), exponent-field scaling.  ~20 FP64-pipe instructions against
// ~30 for the generic libdevice exp (no special cases needed here); arguments below -700 are clamped
// (result 1e-304 instead of a denormal/zero: absolute difference < 1e-300).
__device__ __forceinline__ double exp_neg(double x)
{
    x = fmax(x, -700.0);
    const double SHIFT = 6755399441055744.0;                 // 1.5 * 2^52: round-to-nearest integer in the low bits
    const double t = fma(x, 1.4426950408889634, SHIFT);
    const int n = __double2loint(t);
    const double nf = t - SHIFT;
    double r = fma(nf, -6.93147180369123816490e-01, x);
    r = fma(nf, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;                       // 1/13!
    p = fma(p, r, 2.08767569878681e-09);                     // 1/12!
    p = fma(p, r, 2.505210838544172e-08);                    // 1/11!
    p = fma(p, r, 2.755731922398589e-07);                    // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);                   // 1/9!
    p = fma(p, r, 2.48015873015873e-05);                     // 1/8!
    p = fma(p, r, 1.984126984126984e-04);                    // 1/7!
    p = fma(p, r, 1.388888888888889e-03);                    // 1/6!
    p = fma(p, r, 8.333333333333333e-03);                    // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);                   // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);                   // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// 2^t for t <= 0 (RBF argument pre-scaled by -0.5*log2(e) on the host): n = rint(t), r = t - n exactly, |r| <= 1/2,
// 2^r by degree-13 Horner in r with coefficients ln2^k/k! (truncation 4e-18), exponent-field scaling.
// 17 FP64-pipe instructions; arguments below -1000 are clamped (2^-1000 = 1e-301 instead of a denormal/zero).
__device__ __forceinline__ double exp2_neg(double t)
{
    t = fmax(t, -1000.0);
    const double SHIFT = 6755399441055744.0;                 // 1.5 * 2^52
    const double tt = t + SHIFT;
    const int n = __double2loint(tt);
    const double r = t - (tt - SHIFT);
    double p = 1.3691488853904128e-12;
    p = fma(p, r, 2.5678435993488206e-11);
    p = fma(p, r, 4.4455382718708116e-10);
    p = fma(p, r, 7.054911620801123e-09);
    p = fma(p, r, 1.01780860092397e-07);
    p = fma(p, r, 1.321548679014431e-06);
    p = fma(p, r, 1.5252733804059841e-05);
    p = fma(p, r, 0.0001540353039338161);
    p = fma(p, r, 0.0013333558146428443);
    p = fma(p, r, 0.009618129107628477);
    p = fma(p, r, 0.05550410866482158);
    p = fma(p, r, 0.24022650695910072);
    p = fma(p, r, 0.6931471805599453);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// 2^t for t <= 0, table-driven: t = n + j/GP_TAB + f with |f| <= 1/(2 GP_TAB); 2^(j/GP_TAB) from a GP_TAB-entry
// shared-memory table (appended to the GP blob by the host), 2^f - 1 by a short Taylor/Horner polynomial (truncation
// < 4e-17 relative), 2^n through the exponent field.  GP_TAB = 256 (2 KB, the default): degree 4, 9 FP64-pipe instructions
// against 17 for exp2_neg (GP_TAB = 64: degree 5, 10 instructions; measured 1.172 -> 1.131 ms for the cfg3 preparation, the
// extra bank conflicts of the larger table do not show); the clamp (t < -1000 -> 2^-1000 = 1e-301 instead of a denormal /
// zero) is an integer compare.
struct Exp2Part { double f, T; int nsh; };
// front half: range reduction, table fetch, exponent increment (already shifted into the high word)
__device__ __forceinline__ Exp2Part exp2_front(double t, uint32_t tab)       // tab: shared-space address of the table
{
    Exp2Part e;
    const bool far = (unsigned)__double2hiint(t) > 0xC08F4000u;      // t < -1000: off the FP64 dependency chain
    const double SHIFT = 6755399441055744.0 / GP_TAB;                // 1.5 * 2^(52 - GP_TAB_BITS)
    const double tt = t + SHIFT;
    const int ti = __double2loint(tt);                               // rint(t * GP_TAB)
    e.f = t - (tt - SHIFT);                                          // |f| <= 1/(2 GP_TAB) for every finite t
    asm("ld.shared.f64 %0, [%1];" : "=d"(e.T) : "r"(tab + ((ti & (GP_TAB - 1)) << 3)));
    e.nsh = far ? -1000 * (1 << 20) : (int)((unsigned)(ti << (20 - GP_TAB_BITS)) & 0xfff00000u);
    return e;
}
// back half: 2^f - 1 polynomial, table factor, exponent field
__device__ __forceinline__ double exp2_back(const Exp2Part &e)
{
    const double f = e.f;
#if GP_TAB_BITS >= 8                                        // |f| <= 1/512: the f^5 term is below 4e-17 relative
    double p = 9.61812910762847716e-03;
#else
#if GP_TAB_BITS == 6
    double p = 1.33335581464284434e-03;
#elif GP_TAB_BITS == 5
    double p = 1.54035303933816100e-04;
    p = fma(p, f, 1.33335581464284434e-03);
#else
    double p = 1.52527338040598403e-05;
    p = fma(p, f, 1.54035303933816100e-04);
    p = fma(p, f, 1.33335581464284434e-03);
#endif
    p = fma(p, f, 9.61812910762847716e-03);
#endif
    p = fma(p, f, 5.55041086648215800e-02);
    p = fma(p, f, 2.40226506959100712e-01);
    p = fma(p, f, 6.93147180559945309e-01);
    const double r = fma(e.T, f * p, e.T);
    return __hiloint2double(__double2hiint(r) + e.nsh, __double2loint(r));
}
__device__ __forceinline__ double exp2_neg_tab(double t, uint32_t tab) { return exp2_back(exp2_front(t, tab)); }

// sin and cos for the heading / steering angles of the model (|x| far below 1e5): Cody-Waite reduction by pi/2 in three
// parts + degree-13 / degree-14 polynomials, everything inline.  The libdevice sincos carries an out-of-line slow path
// for huge arguments; a call inside the preparation kernel makes the register allocator spill around it.
__device__ __forceinline__ void sincos_mid(double x, double *s, double *c)
{
    const double SHIFT = 6755399441055744.0;
    const double kd = fma(x, 0.63661977236758134308, SHIFT);           // x * 2/pi, rounded to nearest integer
    const int k = __double2loint(kd);
    const double kf = kd - SHIFT;
    double r = fma(kf, -1.57079632679489655800e+00, x);                // pi/2 split in three parts (53 + 53 + 53 bits)
    r = fma(kf, -6.12323399573676603587e-17, r);
    r = fma(kf, 1.49745833459802707e-33, r);
    const double z = r * r;
    double ps = 1.58962301576546568060e-10;
    ps = fma(ps, z, -2.50507477628578072866e-08);
    ps = fma(ps, z, 2.75573136213857245213e-06);
    ps = fma(ps, z, -1.98412698295895385996e-04);
    ps = fma(ps, z, 8.33333333332211858878e-03);
    ps = fma(ps, z, -1.66666666666666307295e-01);
    const double sn = fma(ps * z, r, r);
    double pc = -1.13596475577881948265e-11;
    pc = fma(pc, z, 2.08757232129817482790e-09);
    pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05);
    pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
    const double cs = fma(fma(pc, z, -0.5), z, 1.0);
    const double s0 = (k & 1) ? cs : sn, c0 = (k & 1) ? sn : cs;
    *s = (k & 2) ? -s0 : s0;
    *c = ((k + 1) & 2) ? -c0 : c0;
}
// 1/x for normal x without the IEEE division's out-of-line slow path (hardware seed + two Newton steps, <= 1 ulp)
__device__ __forceinline__ double rcp_mid(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

struct Jac {
    // d f / d x : rows 0,1 x cols {psi,vx,vy}; row 2 = e_r; rows 3..5 x cols 2..6 ; row 6 = 0
    double j0[3], j1[3];
    double jr[3][5];
    // d f / d u : rows 3..5 x 2 ; row 6 = [0 1]
    double ju[3][2];
};


// GP posterior mean and its gradient w.r.t. the features, per output (model_fitting/gp.py:117-165,446-460)
struct GpOut { double m[ADMPC_GPOUT_MAX]; double g[ADMPC_GPOUT_MAX][ADMPC_DZMAX]; };

// 2^t for t <= 0 with the FP32 special-function unit (opt-in, admpc_opts.gp_precision = 1): t = n + f, n = rint(t) exact in
// FP64, 2^f by ex2.approx.ftz.f32 on the FP32 copy of f (|f| <= 1/2: input rounding 3e-8 * ln2 relative, MUFU.EX2 error
// <= 2 ulp of FP32), 2^n through the FP64 exponent field.  Relative error <= 2^-22 for every argument; 3 FP64-pipe
// instructions instead of 10.
__device__ __forceinline__ double exp2_neg_f32(double t)
{
    const bool far = (unsigned)__double2hiint(t) > 0xC08F4000u;      // t < -1000
    const double SHIFT = 6755399441055744.0;                         // 1.5 * 2^52
    const double tt = t + SHIFT;
    const int n = __double2loint(tt);
    const float ff = (float)(t - (tt - SHIFT));
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ff));
    const double r = (double)e;
    return __hiloint2double(__double2hiint(r) + (far ? -1000 * (1 << 20) : (n << 20)), __double2loint(r));
}

// PREC 0: FP64 table-driven exp2 (default) ; 1: FP32 exponential, FP64 everything else
template <int PREC>
__device__ __forceinline__ double gp_exp2(double t, uint32_t tab) { return PREC ? exp2_neg_f32(t) : exp2_neg_tab(t, tab); }

// gpsm: shared-memory copy of THIS instance's cluster model (nout output blocks); tab: shared-space address of the
// 2^(j/GP_TAB) table of the device exp2
template <int PREC>
__device__ __forceinline__ void gp_eval(const admpc_opts &o, const double *__restrict__ gpsm, int gp_stride, uint32_t tab,
                                        const double x[7], const double u[2], const double gpx[7], double trig, GpOut &O)
{
    const int dz = o.gp_dz, M = o.gp_M;
    const double u0 = u[0], u1 = u[1];
    double z[ADMPC_DZMAX];
#pragma unroll
    for (int d = 0; d < ADMPC_DZMAX; d++) {
        if (d < dz) {
            const int fi = o.gp_feat[d];
            double v = 0.0;
            // feature select without dynamic register indexing
#pragma unroll
            for (int s = 2; s < 7; s++) if (fi == s) v = gpx[s] * trig + x[s] * (1.0 - trig);
            if (fi == 7) v = u0;
            if (fi == 8) v = u1;
            z[d] = v;
        }
    }
#pragma unroll
    for (int j = 0; j < ADMPC_GPOUT_MAX; j++) {
        O.m[j] = 0.0;
#pragma unroll
        for (int d = 0; d < ADMPC_DZMAX; d++) O.g[j][d] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < ADMPC_GPOUT_MAX; j++) {
        if (j >= o.gp_nout) continue;
        // blob per output: M points of {a_0..a_{dz-1}, c, sigma_f*alpha} with a_d = log2(e) X_d / ell_d^2 and
        // c = -0.5 log2(e) sum_d X_d^2/ell_d^2, so that  log2 k(z, X_i) = q + c_i + a_i . z,
        // q = -0.5 log2(e) sum_d z_d^2/ell_d^2  (expanded square: 4 FMAs per point instead of 12 ops; the
        // cancellation costs ~1e-15 absolute in the exponent).  Tail: 1/ell_d^2 (dz values), y_mean.
        const double *blk = gpsm + (size_t)j * gp_stride;
        const double *w = blk + (size_t)M * (dz + 2);
        double wv[ADMPC_DZMAX];
#pragma unroll
        for (int d = 0; d < ADMPC_DZMAX; d++) wv[d] = (d < dz) ? w[d] : 0.0;
        double q = 0.0;
#pragma unroll
        for (int d = 0; d < ADMPC_DZMAX; d++) if (d < dz) q = fma(z[d] * wv[d], z[d], q);
        q *= -0.5 * 1.4426950408889634;
        double m = 0.0, G2[ADMPC_DZMAX];
#pragma unroll
        for (int d = 0; d < ADMPC_DZMAX; d++) G2[d] = 0.0;
        if (dz == 4) {
            // hot case: 4 features -> 48 B per point, three LDS.128.  (Two-point interleaving and software
            // pipelining in the source were both tried: ptxas re-serialises the chains, no gain.)
#pragma unroll 4
            for (int i = 0; i < M; i++) {
                const double2 *pt = reinterpret_cast<const double2 *>(blk + (size_t)i * 6);
                const double2 a01 = pt[0], a23 = pt[1], ca = pt[2];
                const double t = fma(a23.y, z[3], fma(a23.x, z[2], fma(a01.y, z[1], fma(a01.x, z[0], q + ca.x))));
                const double ka = gp_exp2<PREC>(t, tab) * ca.y;
                m += ka;
                G2[0] = fma(ka, a01.x, G2[0]); G2[1] = fma(ka, a01.y, G2[1]);
                G2[2] = fma(ka, a23.x, G2[2]); G2[3] = fma(ka, a23.y, G2[3]);
            }
        } else {
            for (int i = 0; i < M; i++) {
                const double *pt = blk + (size_t)i * (dz + 2);
                double t = q + pt[dz];
#pragma unroll
                for (int d = 0; d < ADMPC_DZMAX; d++) if (d < dz) t = fma(pt[d], z[d], t);
                const double ka = gp_exp2<PREC>(t, tab) * pt[dz + 1];
                m += ka;
#pragma unroll
                for (int d = 0; d < ADMPC_DZMAX; d++) if (d < dz) G2[d] = fma(ka, pt[d], G2[d]);
            }
        }
        // d mu / d z_d = -sum_i ka_i (z_d - X_id)/ell_d^2 = -(z_d/ell_d^2 * m - ln2 * G2_d)
#pragma unroll
        for (int d = 0; d < ADMPC_DZMAX; d++) if (d < dz) O.g[j][d] = fma(0.6931471805599453, G2[d], -(z[d] * wv[d]) * m);
        O.m[j] = m + w[dz];   // + y_mean
    }
}

// f + B_x mu(z) and its Jacobian rows (quad_mpc/quad_3d_optimizer.py:295,315; utils/utils.py:773-808)
__device__ __forceinline__ void gp_apply(const admpc_opts &o, double trig, const GpOut &O, double f[7], Jac &J)
{
    const int dz = o.gp_dz;
#pragma unroll
    for (int j = 0; j < ADMPC_GPOUT_MAX; j++) {
        if (j >= o.gp_nout) continue;
        const int row = o.gp_row[j] - 3;       // 0..2
#pragma unroll
        for (int rr = 0; rr < 3; rr++) {
            if (rr == row) {
                f[3 + rr] += O.m[j];
#pragma unroll
                for (int d = 0; d < ADMPC_DZMAX; d++) {
                    if (d < dz) {
                        const int fi = o.gp_feat[d];
#pragma unroll
                        for (int s = 2; s < 7; s++) if (fi == s) J.jr[rr][s - 2] += (1.0 - trig) * O.g[j][d];
                        if (fi == 7) J.ju[rr][0] += O.g[j][d];
                        if (fi == 8) J.ju[rr][1] += O.g[j][d];
                    }
                }
            }
        }
    }
}

template <bool GP>
// gpsm: shared-memory copy of THIS instance's cluster model (nout output blocks); tab: shared-space address of the
// 2^(j/GP_TAB) table of the device exp2
__device__ __forceinline__ void model_eval(const admpc_opts &o, const double *__restrict__ gpsm, int gp_stride, uint32_t tab,
                                           const double x[7], const double u[2], double p, const double gpx[7],
                                           double trig, double f[7], Jac &J)
{
    const double psi = x[2], vx = x[3], vy = x[4], r = x[5], dl = x[6];
    const double u0 = u[0], u1 = u[1];
    const double im = rcp_mid(o.mass), iiz = rcp_mid(o.iz), L = o.lr + o.lf, iL = rcp_mid(L), c = o.lr * iL, q = 1.0 - p;
    double sp, cp, sd, cd;
    sincos_mid(psi, &sp, &cp);
    sincos_mid(dl, &sd, &cd);
    const double D = vx + 1e-99;
    const double iD = rcp_mid(D);
    const double a = (vy + o.lf * r) * iD;
    const double Ff = o.cf2 * (dl - a);
    const double Fr = o.cr2 * (o.lr * r - vy) * iD;
    const double kin = u1 * vx + dl * u0;

    f[0] = vx * cp - vy * sp;
    f[1] = vx * sp + vy * cp;
    f[2] = r;
    f[3] = p * (u0 - im * Ff * sd + vy * r) + q * u0;
    f[4] = p * (im * (Fr + Ff * cd) - vx * r) + q * (kin * c);
    f[5] = p * (iiz * (o.lf * Ff * cd - o.lr * Fr)) + q * (kin * iL);
    f[6] = u1;

    J.j0[0] = -f[1]; J.j0[1] = cp; J.j0[2] = -sp;
    J.j1[0] = f[0];  J.j1[1] = sp; J.j1[2] = cp;

    const double Ff_vx = o.cf2 * a * iD, Ff_vy = -o.cf2 * iD, Ff_r = -o.cf2 * o.lf * iD, Ff_d = o.cf2;
    const double Fr_vx = -Fr * iD, Fr_vy = -o.cr2 * iD, Fr_r = o.cr2 * o.lr * iD;
    const double t1 = Ff_d * cd - Ff * sd;
    // row 3 (v_x)
    J.jr[0][0] = 0.0;
    J.jr[0][1] = p * (-sd * Ff_vx * im);
    J.jr[0][2] = p * (-sd * Ff_vy * im + r);
    J.jr[0][3] = p * (-sd * Ff_r * im + vy);
    J.jr[0][4] = -p * (Ff_d * sd + Ff * cd) * im;
    J.ju[0][0] = 1.0; J.ju[0][1] = 0.0;
    // row 4 (v_y)
    J.jr[1][0] = 0.0;
    J.jr[1][1] = p * ((Fr_vx + cd * Ff_vx) * im - r) + q * u1 * c;
    J.jr[1][2] = p * (Fr_vy + cd * Ff_vy) * im;
    J.jr[1][3] = p * ((Fr_r + cd * Ff_r) * im - vx);
    J.jr[1][4] = p * t1 * im + q * u0 * c;
    J.ju[1][0] = q * dl * c; J.ju[1][1] = q * vx * c;
    // row 5 (yaw rate)
    J.jr[2][0] = 0.0;
    J.jr[2][1] = p * (o.lf * cd * Ff_vx - o.lr * Fr_vx) * iiz + q * u1 * iL;
    J.jr[2][2] = p * (o.lf * cd * Ff_vy - o.lr * Fr_vy) * iiz;
    J.jr[2][3] = p * (o.lf * cd * Ff_r - o.lr * Fr_r) * iiz;
    J.jr[2][4] = p * o.lf * t1 * iiz + q * u0 * iL;
    J.ju[2][0] = q * dl * iL; J.ju[2][1] = q * vx * iL;

    if (GP) {
        GpOut G;
        gp_eval<0>(o, gpsm, gp_stride, tab, x, u, gpx, trig, G);
        gp_apply(o, trig, G, f, J);
    }
}

// ---- results of the GP sweep kernel (prepare.cu gp_sweep_kernel), read back by the sensitivity kernels of both model variants ------
// GP results of one RK4 stage, SoA rows of gpr: [(k*4 + s) * R + j*(1+dz) + {0: mean, 1+d: d mean / d z_d}][Bp], R = nout*(1+dz)
__device__ __forceinline__ int gpr_rows(const admpc_opts &o) { return o.gp_nout * (1 + o.gp_dz); }
// GP results of one RK4 stage (rows row0 .. row0 + R - 1 of gpr) of instance i
__device__ __forceinline__ void gpr_load(const Params &P, const admpc_opts &o, int row0, int dz, int i, GpOut &G)
{
    const double *in = P.gpr + (size_t)row0 * P.Bp + i;
#pragma unroll
    for (int j = 0; j < ADMPC_GPOUT_MAX; j++) {
        G.m[j] = 0.0;
#pragma unroll
        for (int d = 0; d < ADMPC_DZMAX; d++) G.g[j][d] = 0.0;
        if (j >= o.gp_nout) continue;
        G.m[j] = in[(size_t)(j * (1 + dz)) * P.Bp];
#pragma unroll
        for (int d = 0; d < ADMPC_DZMAX; d++) if (d < dz) G.g[j][d] = in[(size_t)(j * (1 + dz) + 1 + d) * P.Bp];
    }
}

// ---- Frenet variant (frenet.cu; shared with the GP sweep kernel of prepare.cu) --------------------------------------------------
// kappa(s) of instance i: piecewise cubic, SoA rows [breaks (K+1) | coef (K x 4, lowest power first)][Bp]; the end pieces
// extrapolate.  The reference evaluates CasADi's interpolant('kapparef_s', 'bspline', ...) at this place.
// j: piece of the previous evaluation of this thread (-1: none yet -> bisection) ; the arc length moves by less than one piece
// between the sub-stages of one interval, so every later call costs one or two break-point loads.
__device__ __forceinline__ void kappa_spline(const Params &P, int i, double s, double &kap, double &dkap, int &j)
{
    const int K = P.kap_K, Bp = P.Bp;
    const double *sp = P.kap_sp + i;
    if (j < 0) {                                   // largest j with breaks[j] <= s, clamped to [0, K-1]
        int lo = 0, hi = K - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s >= sp[(size_t)mid * Bp]) lo = mid; else hi = mid - 1;
        }
        j = lo;
    } else {
        while (j + 1 < K && s >= sp[(size_t)(j + 1) * Bp]) j++;
        while (j > 0 && s < sp[(size_t)j * Bp]) j--;
    }
    const double t = s - sp[(size_t)j * Bp];
    const double *c = sp + (size_t)(K + 1 + 4 * j) * Bp;
    const double c0 = c[0], c1 = c[(size_t)Bp], c2 = c[(size_t)2 * Bp], c3 = c[(size_t)3 * Bp];
    kap = ((c3 * t + c2) * t + c1) * t + c0;
    dkap = (3.0 * c3 * t + 2.0 * c2) * t + c1;
}

// pose rows of the Frenet model at (x, kappa): s', e_y', e_psi' (frenet.cu header); rows 3..6 are the Cartesian model's
__device__ __forceinline__ void frenet_pose_rows(const double x[7], double kap, double f[7])
{
    double sp, cp;
    sincos(x[2], &sp, &cp);
    const double ey = x[1], vx = x[3], vy = x[4], r = x[5];
    const double vt = vx * cp - vy * sp, vn = vx * sp + vy * cp, den = 1.0 - ey * kap;
    const double sd0 = vt / den;
    f[0] = sd0; f[1] = vn; f[2] = r - ey * kap * sd0;
}
