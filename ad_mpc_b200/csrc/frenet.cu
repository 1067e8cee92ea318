// frenet.cu -- Frenet-frame model variant (SURVEY.md 8a row A2'): the model, pass 2 of its preparation, and the dense
// thread-per-instance IPM that serves as cross-check and long-horizon fallback.
//
// Reference: ad_mpc/__pycache__/fren_ad_3d_optimizer.cpython-36.pyc (bytecode only; equations recovered in SURVEY 8a):
// state x = [s, e_y, e_psi, v_x, v_y, r, delta];  rows 3..6 are the Cartesian model's (ad_3d_optimizer.py:286-310),
//     s'     = (v_x cos e_psi - v_y sin e_psi) / (1 - e_y kappa)
//     e_y'   =  v_x sin e_psi + v_y cos e_psi
//     e_psi' =  r - e_y kappa s'                         (literal bytecode form, including the e_y*kappa factor)
// The reference evaluates a B-spline kappa(s) inside the model (compiled with placeholder knots).  Two forms here:
//   * admpc_batch_set_kappa_spline: kappa(s) as a per-instance piecewise cubic evaluated at every RK4 sub-stage, with the
//     d kappa / d s column in the Jacobian (the reference's semantics; A(:,0) is then dense);
//   * admpc_batch_set_kappa: a per-instance, per-shooting-node constant (d/ds = 0 inside one linearisation).
// With kappa = 0 the model IS the Cartesian one; tests pin this variant to the (reference-pinned) Cartesian path that way.
// Constraint sets: the shipped one (con_set = 0) and the variant's own (con_set = 1, structure pinned by ad_mpc/debug.json),
// see the descriptor below.
//
// Kernels:
//   prepare_dense_kernel<GP>  RK4 + forward sensitivities on the structurally non-zero block of the Frenet Jacobian; GP terms
//                             read back from gpr (pass 1 = gp_sweep_kernel<.., FR = true>, prepare.cu); writes the 80-double
//                             instance-major records qp_mma_g.cu pulls by TMA and, when something reads them, the SoA rows
//                             lin_d[k][79][Bp]
//   qp_dense_kernel           the same Mehrotra / Riccati IPM on dense stage matrices, one thread per instance, workspace in HBM,
//                             run-time constraint descriptor: independent cross-check (ADMPC_QP_VARIANT=1) and N = 128 fallback;
//                             ~12x the run time of the default kernel (qp_mma_g.cu)
//   nlp_res_dense_kernel      NLP KKT residuals of the full-SQP mode on the dense linearisation
#include "common.cuh"

#include "model.cuh"

#define AT(arr, row) (arr)[(size_t)(row) * Bp + i]
#define DL_A 0
#define DL_B 49
#define DL_b 63
#define DL_q 70
#define DL_r 77

__device__ __forceinline__ double nmaxd(double a, double b) { return (a > b || a != a) ? a : b; }

// dense f, Jx (7x7), Ju (7x2) of the Frenet variant at (x, u); rows 3..6 (+ GP) from the shared model code
// GPR: the GP mean / feature gradient of this stage point come from the sweep kernel's results (prepare.cu gp_sweep_kernel<FR>)
template <bool GPR>
__device__ __forceinline__ void frenet_eval(const admpc_opts &o, const GpOut &G, const double x[7],
                                            const double u[2], double p, double kap, double dkap, const double gpx[7], double trig,
                                            double f[7], double Jx[7][7], double Ju[7][2])
{
    Jac J;
    model_eval<false>(o, nullptr, 0, 0u, x, u, p, gpx, trig, f, J);
    if (GPR) gp_apply(o, trig, G, f, J);
#pragma unroll
    for (int r = 0; r < 7; r++) {
#pragma unroll
        for (int c = 0; c < 7; c++) Jx[r][c] = 0.0;
        Ju[r][0] = 0.0; Ju[r][1] = 0.0;
    }
#pragma unroll
    for (int rr = 0; rr < 3; rr++) {
#pragma unroll
        for (int l = 0; l < 5; l++) Jx[3 + rr][2 + l] = J.jr[rr][l];
        Ju[3 + rr][0] = J.ju[rr][0]; Ju[3 + rr][1] = J.ju[rr][1];
    }
    Ju[6][1] = 1.0;
    double sp, cp;
    sincos(x[2], &sp, &cp);
    const double ey = x[1], vx = x[3], vy = x[4], r = x[5];
    const double vt = vx * cp - vy * sp, vn = vx * sp + vy * cp, den = 1.0 - ey * kap;
    const double sd0 = vt / den;
    f[0] = sd0; f[1] = vn; f[2] = r - ey * kap * sd0;
    Jx[0][1] = sd0 * kap / den; Jx[0][2] = -vn / den; Jx[0][3] = cp / den; Jx[0][4] = -sp / den;
    Jx[1][2] = vt; Jx[1][3] = sp; Jx[1][4] = cp;
    Jx[2][1] = -kap * sd0 - ey * kap * Jx[0][1];
    Jx[2][2] = -ey * kap * Jx[0][2];
    Jx[2][3] = -ey * kap * Jx[0][3];
    Jx[2][4] = -ey * kap * Jx[0][4];
    Jx[2][5] = 1.0;
    // d / d s through kappa(s) (spline curvature evaluated inside the model); zero for a per-node constant
    Jx[0][0] = sd0 * ey * dkap / den;
    Jx[2][0] = -ey * dkap * sd0 - ey * kap * Jx[0][0];
}

// GP: pass 2 of the two-pass preparation -- the GP mean / feature gradient at the four RK4 stage points were written to gpr by
// gp_sweep_kernel<.., FR = true> (prepare.cu); the 126-double sensitivity state never shares the SM with a sweep.
template <bool GP>
__global__ void __launch_bounds__(128, 2) prepare_dense_kernel(const Params P)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (i >= P.B || P.lin_bad[i] == 2) return;
    const double h = o.dt, Ts = o.dt;
    double *lin = P.lin_d + (size_t)k * DL_ROWS * Bp;
    const bool soa = (P.skip_lin_d == 0);      // the SoA rows feed the dense / warp QP kernels and the SQP residual kernel only
    double x[7];
#pragma unroll
    for (int c = 0; c < 7; c++) x[c] = AT(P.xb, k * 7 + c);
    if (k == N) {
#pragma unroll
        for (int c = 0; c < 7; c++) if (soa) AT(lin, DL_q + c) = o.We[c] * (x[c] - AT(P.yref, N * 9 + c));
        if (P.lin_im) {
            // terminal record: q_N and x_N at the offsets of the stage records (qp_mma_g.cu W_LQ / W_XB)
            double *rec = P.lin_im + ((size_t)N * Bp + i) * 80;
#pragma unroll
            for (int c = 0; c < 7; c++) { rec[61 + c] = o.We[c] * (x[c] - AT(P.yref, N * 9 + c)); rec[70 + c] = x[c]; }
        }
        return;
    }
    double u[2], gpx[7];
    u[0] = AT(P.ub, k * 2 + 0); u[1] = AT(P.ub, k * 2 + 1);
    const double pk = AT(P.p, k);
    double kap = AT(P.kappa, k), dkap = 0.0;
    const double trig = (GP && o.gp_stage0_trigger && k == 0) ? 1.0 : 0.0;
#pragma unroll
    for (int c = 0; c < 7; c++) gpx[c] = 0.0;
    const int R = GP ? gpr_rows(o) : 0, dz = o.gp_dz;
    GpOut Gn;
    if (GP) gpr_load(P, o, (k * 4) * R, dz, i, Gn);

    // classic RK4 on [x | S], S = [S_x (7x7) | S_u (7x2)], S(0) = [I 0]
    double K[7][9], acc[7][9], kx[7], ax[7];
#pragma unroll
    for (int r = 0; r < 7; r++) {
        kx[r] = 0.0; ax[r] = 0.0;
#pragma unroll
        for (int c = 0; c < 9; c++) { K[r][c] = 0.0; acc[r][c] = 0.0; }
    }
    int kap_j = -1;                                  // spline piece of the previous sub-stage
#pragma unroll 1
    for (int s = 0; s < 4; s++) {
        const double as = (s == 0) ? 0.0 : ((s == 3) ? 1.0 : 0.5);
        const double bs = (s == 0 || s == 3) ? (1.0 / 6.0) : (1.0 / 3.0);
        const double ha = h * as;
        double xs[7], f[7], Jx[7][7], Ju[7][2];
#pragma unroll
        for (int c = 0; c < 7; c++) xs[c] = fma(ha, kx[c], x[c]);
        if (P.kap_K > 0) kappa_spline(P, i, xs[0], kap, dkap, kap_j);      // curvature at this sub-stage's own arc length
        const GpOut G = Gn;                          // loaded one stage ahead
        if (GP && s < 3) gpr_load(P, o, (k * 4 + s + 1) * R, dz, i, Gn);
        frenet_eval<GP>(o, G, xs, u, pk, kap, dkap, gpx, trig, f, Jx, Ju);
#pragma unroll
        for (int c = 0; c < 7; c++) { kx[c] = f[c]; ax[c] = fma(bs, f[c], ax[c]); }
#pragma unroll
        for (int c = 0; c < 9; c++) {
            // Stage input S_in[:,c] = S0[:,c] + ha K[:,c].  Structure kept out of the arithmetic: the velocity rows 3..5 see
            // x2..x6 and u only (their sensitivities w.r.t. s, e_y stay zero), the delta row is analytic, the pose rows do not
            // see u or delta directly.  (Terms dropped are exact zeros: results are those of the dense recursion bit for bit.)
            const bool vel = (c >= 2);
            double sv[7];
#pragma unroll
            for (int r = 0; r < 3; r++) sv[r] = ha * K[r][c] + ((r == c) ? 1.0 : 0.0);
#pragma unroll
            for (int r = 3; r < 6; r++) sv[r] = vel ? ha * K[r][c] + ((r == c) ? 1.0 : 0.0) : 0.0;
            sv[6] = (c == 6) ? 1.0 : ((c == 8) ? ha : 0.0);
            double kn[3];
            kn[0] = fma(Jx[0][2], sv[2], fma(Jx[0][1], sv[1], Jx[0][0] * sv[0]));
            kn[1] = Jx[1][2] * sv[2];
            kn[2] = fma(Jx[2][2], sv[2], fma(Jx[2][1], sv[1], Jx[2][0] * sv[0]));
            if (vel) {
                kn[0] = fma(Jx[0][4], sv[4], fma(Jx[0][3], sv[3], kn[0]));
                kn[1] = fma(Jx[1][4], sv[4], fma(Jx[1][3], sv[3], kn[1]));
                kn[2] = fma(Jx[2][5], sv[5], fma(Jx[2][4], sv[4], fma(Jx[2][3], sv[3], kn[2])));
            }
#pragma unroll
            for (int r = 0; r < 3; r++) { K[r][c] = kn[r]; acc[r][c] = fma(bs, kn[r], acc[r][c]); }
            if (vel) {
#pragma unroll
                for (int r = 3; r < 6; r++) {
                    double v = (c >= 7) ? Ju[r][c - 7] : 0.0;
#pragma unroll
                    for (int l = 2; l < 7; l++) v = fma(Jx[r][l], sv[l], v);
                    K[r][c] = v; acc[r][c] = fma(bs, v, acc[r][c]);
                }
            }
        }
    }
    // delta row: d delta+ / d delta = 1, d delta+ / d u1 = h (acc holds (A - I) / h, B / h)
#pragma unroll
    for (int c = 0; c < 9; c++) acc[6][c] = (c == 8) ? 1.0 : 0.0;
    bool bad = false;
#pragma unroll
    for (int c = 0; c < 7; c++) {
        const double xp = fma(h, ax[c], x[c]);
        bad |= !isfinite(xp);
        if (soa) AT(lin, DL_b + c) = xp - AT(P.xb, (k + 1) * 7 + c);
    }
#pragma unroll
    for (int r = 0; r < 7; r++) {
#pragma unroll
        for (int c = 0; c < 7; c++) {
            const double v = h * acc[r][c] + ((r == c) ? 1.0 : 0.0);
            bad |= !isfinite(v);
            if (soa) AT(lin, DL_A + r * 7 + c) = v;
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const double v = h * acc[r][7 + c];
            bad |= !isfinite(v);
            if (soa) AT(lin, DL_B + r * 2 + c) = v;
        }
    }
#pragma unroll
    for (int c = 0; c < 7; c++) if (soa) AT(lin, DL_q + c) = Ts * o.W[c] * (x[c] - AT(P.yref, k * 9 + c));
#pragma unroll
    for (int c = 0; c < 2; c++) if (soa) AT(lin, DL_r + c) = Ts * o.W[7 + c] * (u[c] - AT(P.yref, k * 9 + 7 + c));
    if (P.lin_im) {
        // instance-major record of the tensor-core QP kernel (qp_mma_g.cu): M = [B | A] rows 0..5 column-major (6 x 9, the column of s included: a
        // spline curvature makes it dense), b, q, r, the linearisation point and one pad; 640 bytes per (stage, instance)
        double *rec = P.lin_im + ((size_t)k * Bp + i) * 80;
#pragma unroll
        for (int c = 0; c < 9; c++)
#pragma unroll
            for (int r = 0; r < 6; r += 2) {
                const double v0 = (c < 2) ? h * acc[r][7 + c] : h * acc[r][c - 2] + ((r == c - 2) ? 1.0 : 0.0);
                const double v1 = (c < 2) ? h * acc[r + 1][7 + c] : h * acc[r + 1][c - 2] + ((r + 1 == c - 2) ? 1.0 : 0.0);
                *reinterpret_cast<double2 *>(rec + c * 6 + r) = make_double2(v0, v1);
            }
        double tail[26];                       // b (7) q (7) r (2) x (7) u (2) pad
#pragma unroll
        for (int c = 0; c < 7; c++) {
            tail[c] = fma(h, ax[c], x[c]) - AT(P.xb, (k + 1) * 7 + c);
            tail[7 + c] = Ts * o.W[c] * (x[c] - AT(P.yref, k * 9 + c));
            tail[16 + c] = x[c];
        }
        tail[14] = Ts * o.W[7] * (u[0] - AT(P.yref, k * 9 + 7)); tail[15] = Ts * o.W[8] * (u[1] - AT(P.yref, k * 9 + 8));
        tail[23] = u[0]; tail[24] = u[1]; tail[25] = 0.0;
#pragma unroll
        for (int c = 0; c < 26; c += 2) *reinterpret_cast<double2 *>(rec + 54 + c) = make_double2(tail[c], tail[c + 1]);
    }
    if (bad) P.lin_bad[i] = 1;
}

// ---------------------------------------------------------------------------------------------- dense IPM ---------
// Constraint sets (twin of oracle/rti_oracle.c con_get): bounded quantities q = 0..nb-1 of z = [u0 u1 | x0..x6], each hard
// (two rows) or soft (slack id soft[q], two more rows); row order [lb(q) | ub(q) | ls(s) | us(s)] as in the reference's iterate
// dumps.  con_set 0: u0, u1 soft + delta hard (10 rows) ; con_set 1, the Frenet variant's own set (structure pinned by
// ad_mpc/debug.json): u0 soft, u1 hard, e_y = x[1] hard, delta soft (12 rows).  State bounds and their slacks do not exist at
// stage 0 (x0 is eliminated).
#define NCD 12
struct ConDesc {
    int nb, ns, nc;
    int idx[4], soft[4];
    double lo[4], hi[4];
};
__device__ __forceinline__ void con_get(const admpc_opts &o, ConDesc &d)
{
    d.ns = 2;
    d.idx[0] = 0; d.idx[1] = 1;
    d.lo[0] = o.lbu[0]; d.hi[0] = o.ubu[0]; d.lo[1] = o.lbu[1]; d.hi[1] = o.ubu[1];
    if (o.con_set == 1) {
        d.nb = 4;
        d.idx[2] = 3; d.idx[3] = 8;
        d.soft[0] = 0; d.soft[1] = -1; d.soft[2] = -1; d.soft[3] = 1;
        d.lo[2] = o.lbx2; d.hi[2] = o.ubx2; d.lo[3] = o.lbx; d.hi[3] = o.ubx;
    } else {
        d.nb = 3;
        d.idx[2] = 8; d.idx[3] = 8;
        d.soft[0] = 0; d.soft[1] = 1; d.soft[2] = -1; d.soft[3] = -1;
        d.lo[2] = o.lbx; d.hi[2] = o.ubx; d.lo[3] = 0.0; d.hi[3] = 0.0;
    }
    d.nc = 2 * d.nb + 2 * d.ns;
}
__device__ __forceinline__ bool q_on(const ConDesc &d, int k, int q) { return d.idx[q] < 2 || k >= 1; }
__device__ __forceinline__ int row_q(const ConDesc &d, int c)
{
    if (c < d.nb) return c;
    if (c < 2 * d.nb) return c - d.nb;
    const int s = (c < 2 * d.nb + d.ns) ? c - 2 * d.nb : c - 2 * d.nb - d.ns;
    for (int q = 0; q < d.nb; q++) if (d.soft[q] == s) return q;
    return 0;
}
__device__ __forceinline__ bool con_on(const ConDesc &d, int k, int c) { return q_on(d, k, row_q(d, c)); }
#define RL(q) (q)
#define RU(q) (d.nb + (q))
#define RLS(s) (2 * d.nb + (s))
#define RUS(s) (2 * d.nb + d.ns + (s))
__device__ __forceinline__ constexpr int sy(int a, int b) { return (a >= b) ? (a * (a + 1) / 2 + b) : (b * (b + 1) / 2 + a); }
// current value of the iterate entry a bounded quantity refers to, and of a delta vector [du | dx] of stage k
__device__ __forceinline__ double q_bar(const Params &P, const ConDesc &d, int q, int k, int i)
{
    const int Bp = P.Bp;
    return (d.idx[q] < 2) ? AT(P.ub, k * 2 + d.idx[q]) : AT(P.xb, k * 7 + d.idx[q] - 2);
}
struct QScal { double Sl, Su, Dl, Du; };
__device__ __forceinline__ void q_scaling(const Params &P, const ConDesc &d, int q, int k, int i, QScal &S)
{
    const admpc_opts &o = P.o;
    const int Bp = P.Bp, nc = d.nc;
    const double Ts = o.dt;
    const int sq = d.soft[q];
    S.Sl = AT(P.lam, k * nc + RL(q)) / AT(P.t, k * nc + RL(q));
    S.Su = AT(P.lam, k * nc + RU(q)) / AT(P.t, k * nc + RU(q));
    if (sq >= 0) {
        const double Ssl = AT(P.lam, k * nc + RLS(sq)) / AT(P.t, k * nc + RLS(sq)), Ssu = AT(P.lam, k * nc + RUS(sq)) / AT(P.t, k * nc + RUS(sq));
        S.Dl = Ts * o.Zl[sq] + S.Sl + Ssl; S.Du = Ts * o.Zu[sq] + S.Su + Ssu;
    } else { S.Dl = 0.0; S.Du = 0.0; }
}

// residuals of the current point: fills rgu, rgx, rgsl, rgsu, rb, rd, rm; returns the four norms and the sum of lam t
__device__ void dn_residuals(const Params &P, const ConDesc &d, int i, double res[4], double &summ, int &ncon)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp, nc = d.nc;
    const double Ts = o.dt;
    double ng = 0, nb = 0, nd = 0, nm = 0;
    summ = 0; ncon = 0;
    for (int k = 0; k <= N; k++) {
        const double *lin = P.lin_d + (size_t)k * DL_ROWS * Bp;
        double dx[7];
#pragma unroll
        for (int a = 0; a < 7; a++) dx[a] = AT(P.dx, k * 7 + a);
        if (k < N) {
            double du[2], pi[7];
            du[0] = AT(P.du, k * 2); du[1] = AT(P.du, k * 2 + 1);
#pragma unroll
            for (int a = 0; a < 7; a++) pi[a] = AT(P.pi, k * 7 + a);
#pragma unroll
            for (int j = 0; j < 2; j++) {
                double g = Ts * o.W[7 + j] * du[j] + AT(lin, DL_r + j);
#pragma unroll
                for (int l = 0; l < 7; l++) g += AT(lin, DL_B + l * 2 + j) * pi[l];
                for (int q = 0; q < d.nb; q++) if (d.idx[q] == j) g += -AT(P.lam, k * nc + RL(q)) + AT(P.lam, k * nc + RU(q));
                AT(P.rgu, k * 2 + j) = g;
                ng = nmaxd(ng, fabs(g));
            }
            for (int c = 0; c < nc; c++) AT(P.rd, k * nc + c) = 0.0;
            for (int q = 0; q < d.nb; q++) {
                if (!q_on(d, k, q)) continue;
                const double v = (d.idx[q] < 2) ? ((d.idx[q] == 0) ? du[0] : du[1]) : AT(P.dx, k * 7 + d.idx[q] - 2);
                const double bar = q_bar(P, d, q, k, i);
                const double lo = d.lo[q] - bar, hi = d.hi[q] - bar;
                const int sq = d.soft[q];
                if (sq >= 0) {
                    const double sl = AT(P.sl, k * 2 + sq), su = AT(P.su, k * 2 + sq);
                    const double gsl = Ts * o.zl[sq] + Ts * o.Zl[sq] * sl - AT(P.lam, k * nc + RL(q)) - AT(P.lam, k * nc + RLS(sq));
                    const double gsu = Ts * o.zu[sq] + Ts * o.Zu[sq] * su - AT(P.lam, k * nc + RU(q)) - AT(P.lam, k * nc + RUS(sq));
                    AT(P.rgsl, k * 2 + sq) = gsl; AT(P.rgsu, k * 2 + sq) = gsu;
                    AT(P.rd, k * nc + RL(q)) = AT(P.t, k * nc + RL(q)) - (v - lo + sl);
                    AT(P.rd, k * nc + RU(q)) = AT(P.t, k * nc + RU(q)) - (hi - v + su);
                    AT(P.rd, k * nc + RLS(sq)) = AT(P.t, k * nc + RLS(sq)) - sl;
                    AT(P.rd, k * nc + RUS(sq)) = AT(P.t, k * nc + RUS(sq)) - su;
                    ng = nmaxd(ng, nmaxd(fabs(gsl), fabs(gsu)));
                } else {
                    AT(P.rd, k * nc + RL(q)) = AT(P.t, k * nc + RL(q)) - (v - lo);
                    AT(P.rd, k * nc + RU(q)) = AT(P.t, k * nc + RU(q)) - (hi - v);
                }
            }
#pragma unroll
            for (int r = 0; r < 7; r++) {
                double v = AT(lin, DL_b + r) - AT(P.dx, (k + 1) * 7 + r);
#pragma unroll
                for (int l = 0; l < 7; l++) v += AT(lin, DL_A + r * 7 + l) * dx[l];
                v += AT(lin, DL_B + r * 2) * du[0];
                v += AT(lin, DL_B + r * 2 + 1) * du[1];
                AT(P.rb, k * 7 + r) = v;
                nb = nmaxd(nb, fabs(v));
            }
            for (int c = 0; c < nc; c++) {
                if (!con_on(d, k, c)) { AT(P.rm, k * nc + c) = 0.0; continue; }
                const double rm = AT(P.lam, k * nc + c) * AT(P.t, k * nc + c);
                AT(P.rm, k * nc + c) = rm;
                nd = nmaxd(nd, fabs(AT(P.rd, k * nc + c)));
                nm = nmaxd(nm, fabs(rm));
                summ += rm;
                ncon++;
            }
        }
        if (k >= 1) {
#pragma unroll
            for (int a = 0; a < 7; a++) {
                const double Qd = (k < N) ? Ts * o.W[a] : o.We[a];
                double g = Qd * dx[a] + AT(lin, DL_q + a) - AT(P.pi, (k - 1) * 7 + a);
                if (k < N) {
#pragma unroll
                    for (int l = 0; l < 7; l++) g += AT(lin, DL_A + l * 7 + a) * AT(P.pi, k * 7 + l);
                    for (int q = 0; q < d.nb; q++)
                        if (d.idx[q] == 2 + a) g += -AT(P.lam, k * nc + RL(q)) + AT(P.lam, k * nc + RU(q));
                }
                AT(P.rgx, k * 7 + a) = g;
                ng = nmaxd(ng, fabs(g));
            }
        }
    }
    res[0] = ng; res[1] = nb; res[2] = nd; res[3] = nm;
}

// Riccati factorisation of the barrier-modified Hessian (matrix part): K, Luu (in Ginv), P (packed symmetric)
__device__ void dn_factor(const Params &P, const ConDesc &d, int i)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const double Ts = o.dt;
    double Pn[28];
#pragma unroll
    for (int a = 0; a < 28; a++) Pn[a] = 0.0;
#pragma unroll
    for (int a = 0; a < 7; a++) Pn[sy(a, a)] = o.We[a];
#pragma unroll
    for (int a = 0; a < 28; a++) AT(P.P, N * 28 + a) = Pn[a];
    for (int k = N - 1; k >= 0; k--) {
        const double *lin = P.lin_d + (size_t)k * DL_ROWS * Bp;
        // barrier-modified diagonal of the stage Hessian over z = [u; x] (soft-bound slacks eliminated)
        double Hd[9];
#pragma unroll
        for (int z = 0; z < 9; z++) Hd[z] = (z < 2) ? Ts * o.W[7 + z] : Ts * o.W[z - 2];
        for (int q = 0; q < d.nb; q++) {
            if (!q_on(d, k, q)) continue;
            QScal S; q_scaling(P, d, q, k, i, S);
            const double add2 = (d.soft[q] >= 0) ? S.Su * (1.0 - S.Su / S.Du) : 0.0;
            const double add1 = (d.soft[q] >= 0) ? S.Sl * (1.0 - S.Sl / S.Dl) : (S.Sl + S.Su);
#pragma unroll
            for (int z = 0; z < 9; z++) if (z == d.idx[q]) Hd[z] = Hd[z] + add1 + add2;
        }
        double BA[7][9], PBA[7][9], G[45];
#pragma unroll
        for (int r = 0; r < 7; r++) {
            BA[r][0] = AT(lin, DL_B + r * 2); BA[r][1] = AT(lin, DL_B + r * 2 + 1);
#pragma unroll
            for (int c = 0; c < 7; c++) BA[r][2 + c] = AT(lin, DL_A + r * 7 + c);
        }
#pragma unroll
        for (int r = 0; r < 7; r++)
#pragma unroll
            for (int c = 0; c < 9; c++) {
                double v = 0.0;
#pragma unroll
                for (int l = 0; l < 7; l++) v += Pn[sy(r, l)] * BA[l][c];
                PBA[r][c] = v;
            }
#pragma unroll
        for (int a = 0; a < 9; a++)
#pragma unroll
            for (int c = 0; c <= a; c++) {
                double v = 0.0;
#pragma unroll
                for (int l = 0; l < 7; l++) v += BA[l][a] * PBA[l][c];
                G[sy(a, c)] = v;
            }
#pragma unroll
        for (int z = 0; z < 9; z++) G[sy(z, z)] += Hd[z];
        const double l00 = sqrt(G[sy(0, 0)] + o.reg), l10 = G[sy(1, 0)] / l00, l11 = sqrt(G[sy(1, 1)] + o.reg - l10 * l10);
        AT(P.Ginv, k * 3 + 0) = l00; AT(P.Ginv, k * 3 + 1) = l10; AT(P.Ginv, k * 3 + 2) = l11;
        double K0[7], K1[7];
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const double y0 = G[sy(2 + j, 0)] / l00, y1 = (G[sy(2 + j, 1)] - l10 * y0) / l11;
            const double k1 = y1 / l11, k0 = (y0 - l10 * k1) / l00;
            K0[j] = -k0; K1[j] = -k1;
            AT(P.K, k * 14 + j) = -k0; AT(P.K, k * 14 + 7 + j) = -k1;
        }
#pragma unroll
        for (int a = 0; a < 7; a++)
#pragma unroll
            for (int c = 0; c <= a; c++) {
                // symmetrised Schur complement (the oracle averages the two triangles)
                const double v1 = G[sy(2 + a, 2 + c)] + G[sy(2 + a, 0)] * K0[c] + G[sy(2 + a, 1)] * K1[c];
                const double v2 = G[sy(2 + c, 2 + a)] + G[sy(2 + c, 0)] * K0[a] + G[sy(2 + c, 1)] * K1[a];
                Pn[sy(a, c)] = (a == c) ? v1 : 0.5 * (v1 + v2);
            }
#pragma unroll
        for (int a = 0; a < 28; a++) AT(P.P, k * 28 + a) = Pn[a];
    }
}

// barrier gradient of stage k: gl per row, the modified gradients rt (inputs) / qx (states), cl / cu per slack
__device__ __forceinline__ void dn_stage_grad(const Params &P, const ConDesc &d, int i, int k, double gl[NCD], double rt[2], double qx[7],
                                               double cl[2], double cu[2], QScal S[4])
{
    const int Bp = P.Bp, nc = d.nc;
    for (int c = 0; c < nc; c++)
        gl[c] = con_on(d, k, c) ? (AT(P.rm, k * nc + c) - AT(P.lam, k * nc + c) * AT(P.rd, k * nc + c)) / AT(P.t, k * nc + c) : 0.0;
    rt[0] = AT(P.rgu, k * 2); rt[1] = AT(P.rgu, k * 2 + 1);
#pragma unroll
    for (int a = 0; a < 7; a++) qx[a] = (k >= 1) ? AT(P.rgx, k * 7 + a) : 0.0;
    cl[0] = cl[1] = cu[0] = cu[1] = 0.0;
    for (int q = 0; q < d.nb; q++) {
        if (!q_on(d, k, q)) continue;
        q_scaling(P, d, q, k, i, S[q]);
        const int sq = d.soft[q];
        double add;
        double base = 0.0;
#pragma unroll
        for (int z = 0; z < 9; z++) if (z == d.idx[q]) base = (z < 2) ? rt[z & 1] : qx[(z >= 2) ? z - 2 : 0];
        if (sq >= 0) {
            const double c_l = AT(P.rgsl, k * 2 + sq) + gl[RL(q)] + gl[RLS(sq)];
            const double c_u = AT(P.rgsu, k * 2 + sq) + gl[RU(q)] + gl[RUS(sq)];
            if (sq == 0) { cl[0] = c_l; cu[0] = c_u; } else { cl[1] = c_l; cu[1] = c_u; }
            add = base + (gl[RL(q)] - S[q].Sl * c_l / S[q].Dl) - (gl[RU(q)] - S[q].Su * c_u / S[q].Du);
        } else {
            add = base + gl[RL(q)] - gl[RU(q)];
        }
#pragma unroll
        for (int z = 0; z < 9; z++) if (z == d.idx[q]) { if (z < 2) rt[z & 1] = add; else qx[(z >= 2) ? z - 2 : 0] = add; }
    }
}

// Newton step for the complementarity right-hand side in P.rm: ddu, ddx, dpi, dsl, dsu, dt, dlam; returns the
// fraction-to-boundary step and the two sums of the Mehrotra centering estimate
__device__ void dn_solve(const Params &P, const ConDesc &d, int i, double &alpha, double &s1, double &s2)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp, nc = d.nc;
    double pv[7];
#pragma unroll
    for (int a = 0; a < 7; a++) { pv[a] = AT(P.rgx, N * 7 + a); AT(P.pv, N * 7 + a) = pv[a]; }
    // backward vector recursion (stage barrier quantities recomputed, rt / qx not stored)
    for (int k = N - 1; k >= 0; k--) {
        const double *lin = P.lin_d + (size_t)k * DL_ROWS * Bp;
        double gl[NCD], rt[2], qx[7], cl[2], cu[2];
        QScal S[4];
        dn_stage_grad(P, d, i, k, gl, rt, qx, cl, cu, S);
        double Pn[28], hv[7], gu[2], gx[7];
#pragma unroll
        for (int a = 0; a < 28; a++) Pn[a] = AT(P.P, (k + 1) * 28 + a);
#pragma unroll
        for (int a = 0; a < 7; a++) {
            double v = pv[a];
#pragma unroll
            for (int l = 0; l < 7; l++) v += Pn[sy(a, l)] * AT(P.rb, k * 7 + l);
            hv[a] = v;
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            double v = rt[j];
#pragma unroll
            for (int l = 0; l < 7; l++) v += AT(lin, DL_B + l * 2 + j) * hv[l];
            gu[j] = v;
        }
#pragma unroll
        for (int a = 0; a < 7; a++) {
            double v = qx[a];
#pragma unroll
            for (int l = 0; l < 7; l++) v += AT(lin, DL_A + l * 7 + a) * hv[l];
            gx[a] = v;
        }
        const double l00 = AT(P.Ginv, k * 3 + 0), l10 = AT(P.Ginv, k * 3 + 1), l11 = AT(P.Ginv, k * 3 + 2);
        const double y0 = gu[0] / l00, y1 = (gu[1] - l10 * y0) / l11;
        const double k1 = y1 / l11, k0 = (y0 - l10 * k1) / l00;
        AT(P.kf, k * 2 + 0) = -k0; AT(P.kf, k * 2 + 1) = -k1;
#pragma unroll
        for (int a = 0; a < 7; a++) {
            pv[a] = gx[a] + AT(P.K, k * 14 + a) * gu[0] + AT(P.K, k * 14 + 7 + a) * gu[1];
            AT(P.pv, k * 7 + a) = pv[a];
        }
    }
    // forward roll-out + step recovery + ratio test
    double ddx[7];
#pragma unroll
    for (int a = 0; a < 7; a++) { ddx[a] = 0.0; AT(P.ddx, a) = 0.0; }
    alpha = 1.0; s1 = 0.0; s2 = 0.0;
    for (int k = 0; k < N; k++) {
        const double *lin = P.lin_d + (size_t)k * DL_ROWS * Bp;
        double ddu[2], nx[7];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            double v = AT(P.kf, k * 2 + j);
#pragma unroll
            for (int l = 0; l < 7; l++) v += AT(P.K, k * 14 + j * 7 + l) * ddx[l];
            ddu[j] = v;
            AT(P.ddu, k * 2 + j) = v;
        }
        // slack / t / lambda steps of this stage (need ddu_k and ddx_k)
        {
            double gl[NCD], rt[2], qx[7], cl[2], cu[2], dt[NCD];
            QScal S[4];
            dn_stage_grad(P, d, i, k, gl, rt, qx, cl, cu, S);
            for (int c = 0; c < nc; c++) dt[c] = 0.0;
            AT(P.dsl, k * 2) = 0.0; AT(P.dsl, k * 2 + 1) = 0.0; AT(P.dsu, k * 2) = 0.0; AT(P.dsu, k * 2 + 1) = 0.0;
            for (int q = 0; q < d.nb; q++) {
                if (!q_on(d, k, q)) continue;
                double dv = 0.0;
#pragma unroll
                for (int z = 0; z < 9; z++) if (z == d.idx[q]) dv = (z < 2) ? ddu[z & 1] : ddx[(z >= 2) ? z - 2 : 0];
                const int sq = d.soft[q];
                if (sq >= 0) {
                    const double c_l = (sq == 0) ? cl[0] : cl[1], c_u = (sq == 0) ? cu[0] : cu[1];
                    const double dsl = -(c_l + S[q].Sl * dv) / S[q].Dl, dsu = -(c_u - S[q].Su * dv) / S[q].Du;
                    AT(P.dsl, k * 2 + sq) = dsl; AT(P.dsu, k * 2 + sq) = dsu;
                    dt[RL(q)] = dv + dsl - AT(P.rd, k * nc + RL(q));
                    dt[RU(q)] = -dv + dsu - AT(P.rd, k * nc + RU(q));
                    dt[RLS(sq)] = dsl - AT(P.rd, k * nc + RLS(sq));
                    dt[RUS(sq)] = dsu - AT(P.rd, k * nc + RUS(sq));
                } else {
                    dt[RL(q)] = dv - AT(P.rd, k * nc + RL(q));
                    dt[RU(q)] = -dv - AT(P.rd, k * nc + RU(q));
                }
            }
            for (int c = 0; c < nc; c++) {
                const double lam = AT(P.lam, k * nc + c), t = AT(P.t, k * nc + c);
                const bool on = con_on(d, k, c);
                const double dl = on ? -(AT(P.rm, k * nc + c) + lam * dt[c]) / t : 0.0;
                AT(P.dt, k * nc + c) = dt[c]; AT(P.dlam, k * nc + c) = dl;
                if (!on) continue;
                if (dl < 0 && -lam / dl < alpha) alpha = -lam / dl;
                if (dt[c] < 0 && -t / dt[c] < alpha) alpha = -t / dt[c];
                s1 += lam * dt[c] + t * dl;
                s2 += dl * dt[c];
            }
        }
#pragma unroll
        for (int r = 0; r < 7; r++) {
            double v = AT(P.rb, k * 7 + r);
#pragma unroll
            for (int l = 0; l < 7; l++) v += AT(lin, DL_A + r * 7 + l) * ddx[l];
            v += AT(lin, DL_B + r * 2) * ddu[0];
            v += AT(lin, DL_B + r * 2 + 1) * ddu[1];
            nx[r] = v;
        }
#pragma unroll
        for (int a = 0; a < 7; a++) { ddx[a] = nx[a]; AT(P.ddx, (k + 1) * 7 + a) = nx[a]; }
#pragma unroll
        for (int a = 0; a < 7; a++) {
            double v = AT(P.pv, (k + 1) * 7 + a);
#pragma unroll
            for (int l = 0; l < 7; l++) v += AT(P.P, (k + 1) * 28 + sy(a, l)) * ddx[l];
            AT(P.dpi, k * 7 + a) = v;
        }
    }
}

__global__ void __launch_bounds__(64) qp_dense_kernel(const Params P)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.B) return;
    if (const int flag = P.lin_bad[i]) {
        if (flag == 1) { P.status[i] = 1; P.qp_status[i] = 0; P.qp_iter[i] = 0; }
        return;
    }
    ConDesc d;
    con_get(o, d);
    const int nc = d.nc;
    // cold start (identical to qp_ipm.cu): primal 0 pushed thr0 inside its box, t from the box, lam = mu0 / t
#pragma unroll
    for (int a = 0; a < 7; a++) AT(P.dx, a) = AT(P.x0, a) - AT(P.xb, a);
    for (int k = 0; k < N; k++) {
        for (int a = 0; a < 7; a++) { AT(P.pi, k * 7 + a) = 0.0; AT(P.dx, (k + 1) * 7 + a) = 0.0; }
        AT(P.du, k * 2) = 0.0; AT(P.du, k * 2 + 1) = 0.0;
        AT(P.sl, k * 2) = 0.0; AT(P.sl, k * 2 + 1) = 0.0; AT(P.su, k * 2) = 0.0; AT(P.su, k * 2 + 1) = 0.0;
    }
    for (int k = 0; k < N; k++) {
        for (int c = 0; c < nc; c++) { AT(P.t, k * nc + c) = 1.0; AT(P.lam, k * nc + c) = 0.0; }
        for (int q = 0; q < d.nb; q++) {
            if (!q_on(d, k, q)) continue;
            const double bar = q_bar(P, d, q, k, i);
            const double lo = d.lo[q] - bar, hi = d.hi[q] - bar;
            double v = 0.0;
            if (v - lo < o.thr0) {
                if (hi - v < o.thr0) v = 0.5 * (lo + hi);
                else v = lo + o.thr0;
            } else if (hi - v < o.thr0) v = hi - o.thr0;
            if (d.idx[q] < 2) AT(P.du, k * 2 + d.idx[q]) = v; else AT(P.dx, k * 7 + d.idx[q] - 2) = v;
            AT(P.t, k * nc + RL(q)) = fmax(o.thr0, v - lo);
            AT(P.t, k * nc + RU(q)) = fmax(o.thr0, hi - v);
            if (d.soft[q] >= 0) { AT(P.t, k * nc + RLS(d.soft[q])) = o.thr0; AT(P.t, k * nc + RUS(d.soft[q])) = o.thr0; }
        }
        for (int c = 0; c < nc; c++) if (con_on(d, k, c)) AT(P.lam, k * nc + c) = o.mu0 / AT(P.t, k * nc + c);
    }
    int status = 1, iter = 0;
    double res[4] = {0, 0, 0, 0};
    for (iter = 0;; iter++) {
        double summ;
        int ncon;
        dn_residuals(P, d, i, res, summ, ncon);
        if (!(isfinite(res[0]) && isfinite(res[1]) && isfinite(res[2]) && isfinite(res[3]))) { status = 3; break; }
        if (res[0] < o.tol_stat && res[1] < o.tol_eq && res[2] < o.tol_ineq && res[3] < o.tol_comp) { status = 0; break; }
        if (iter >= o.iter_max) { status = 1; break; }
        const double inv_nc = 1.0 / (double)ncon;
        const double mu = summ * inv_nc;
        double a_aff, s1, s2, alpha;
        dn_factor(P, d, i);
        dn_solve(P, d, i, a_aff, s1, s2);
        const double mu_aff = (summ + a_aff * s1 + a_aff * a_aff * s2) * inv_nc;
        double sigma = mu_aff / mu;
        sigma = sigma * sigma * sigma;
        for (int k = 0; k < N; k++)
            for (int c = 0; c < nc; c++)
                if (con_on(d, k, c))
                    AT(P.rm, k * nc + c) = AT(P.lam, k * nc + c) * AT(P.t, k * nc + c) + AT(P.dlam, k * nc + c) * AT(P.dt, k * nc + c) - sigma * mu;
        dn_solve(P, d, i, alpha, s1, s2);
        if (alpha < o.alpha_min) { status = 2; break; }
        if (alpha < 1.0) alpha *= 0.995;
        for (int k = 0; k < N; k++) {
            for (int j = 0; j < 2; j++) {
                AT(P.du, k * 2 + j) += alpha * AT(P.ddu, k * 2 + j);
                AT(P.sl, k * 2 + j) += alpha * AT(P.dsl, k * 2 + j);
                AT(P.su, k * 2 + j) += alpha * AT(P.dsu, k * 2 + j);
            }
            for (int a = 0; a < 7; a++) {
                AT(P.dx, (k + 1) * 7 + a) += alpha * AT(P.ddx, (k + 1) * 7 + a);
                AT(P.pi, k * 7 + a) += alpha * AT(P.dpi, k * 7 + a);
            }
            for (int c = 0; c < nc; c++) {
                if (!con_on(d, k, c)) continue;
                AT(P.lam, k * nc + c) = fmax(AT(P.lam, k * nc + c) + alpha * AT(P.dlam, k * nc + c), o.lam_min);
                AT(P.t, k * nc + c) = fmax(AT(P.t, k * nc + c) + alpha * AT(P.dt, k * nc + c), o.t_min);
            }
        }
    }
    const int qps = (status == 0) ? 0 : ((status == 1) ? 2 : ((status == 2) ? 3 : 1));
    P.qp_status[i] = qps;
    P.qp_iter[i] = iter;
    P.status[i] = (qps == 0 || qps == 2) ? 0 : 4;
    AT(P.res_out, 0) = res[0]; AT(P.res_out, 1) = res[1]; AT(P.res_out, 2) = res[2]; AT(P.res_out, 3) = res[3];
}

// NLP KKT residual check of the full-SQP loop on the dense linearisation (twin of nlp_res_kernel in sqp.cu)
__global__ void __launch_bounds__(128) nlp_res_dense_kernel(const Params P, int it, double tol0, double tol1, double tol2, double tol3,
                                                            int *active)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.B) return;
    const int flag = P.lin_bad[i];
    if (flag == 2) return;
    if (it > 0 && P.status[i] != 0) { P.sqp_status[i] = P.status[i]; P.sqp_iter[i] = it - 1; P.lin_bad[i] = 2; return; }
    if (flag == 1) { P.sqp_status[i] = 1; P.status[i] = 1; P.sqp_iter[i] = it; P.lin_bad[i] = 2; return; }
    ConDesc d;
    con_get(o, d);
    const int nc = d.nc;
    const double Ts = o.dt;
    double ng = 0, nb = 0, nd = 0, nm = 0;
    double pim[7];
#pragma unroll
    for (int a = 0; a < 7; a++) { pim[a] = 0.0; nb = nmaxd(nb, fabs(AT(P.x0, a) - AT(P.xb, a))); }
    for (int k = 0; k < N; k++) {
        const double *lin = P.lin_d + (size_t)k * DL_ROWS * Bp;
        double pi[7], gu[2], gx[7];
#pragma unroll
        for (int a = 0; a < 7; a++) pi[a] = AT(P.pib, k * 7 + a);
#pragma unroll
        for (int j = 0; j < 2; j++) {
            double g = AT(lin, DL_r + j);
#pragma unroll
            for (int l = 0; l < 7; l++) g += AT(lin, DL_B + l * 2 + j) * pi[l];
            gu[j] = g;
        }
#pragma unroll
        for (int a = 0; a < 7; a++) {
            double g = AT(lin, DL_q + a) - pim[a];
#pragma unroll
            for (int l = 0; l < 7; l++) g += AT(lin, DL_A + l * 7 + a) * pi[l];
            gx[a] = g;
        }
        for (int q = 0; q < d.nb; q++) {
            if (!q_on(d, k, q)) continue;
            const double ll = AT(P.lamb, k * nc + RL(q)), lu = AT(P.lamb, k * nc + RU(q));
            const double tl = AT(P.tb, k * nc + RL(q)), tu = AT(P.tb, k * nc + RU(q));
            const double bar = q_bar(P, d, q, k, i);
#pragma unroll
            for (int z = 0; z < 9; z++) if (z == d.idx[q]) { if (z < 2) gu[z & 1] += -ll + lu; else gx[(z >= 2) ? z - 2 : 0] += -ll + lu; }
            const int sq = d.soft[q];
            double sl = 0.0, su = 0.0;
            if (sq >= 0) {
                sl = AT(P.slb, k * 2 + sq); su = AT(P.sub, k * 2 + sq);
                const double lls = AT(P.lamb, k * nc + RLS(sq)), lus = AT(P.lamb, k * nc + RUS(sq));
                const double gsl = Ts * o.zl[sq] + Ts * o.Zl[sq] * sl - ll - lls;
                const double gsu = Ts * o.zu[sq] + Ts * o.Zu[sq] * su - lu - lus;
                ng = nmaxd(ng, nmaxd(fabs(gsl), fabs(gsu)));
                nd = nmaxd(nd, nmaxd(fabs(AT(P.tb, k * nc + RLS(sq)) - sl), fabs(AT(P.tb, k * nc + RUS(sq)) - su)));
                nm = nmaxd(nm, nmaxd(fabs(lls * AT(P.tb, k * nc + RLS(sq))), fabs(lus * AT(P.tb, k * nc + RUS(sq)))));
            }
            nd = nmaxd(nd, fabs(tl - (0.0 - (d.lo[q] - bar) + sl)));
            nd = nmaxd(nd, fabs(tu - ((d.hi[q] - bar) - 0.0 + su)));
            nm = nmaxd(nm, nmaxd(fabs(ll * tl), fabs(lu * tu)));
        }
        ng = nmaxd(ng, nmaxd(fabs(gu[0]), fabs(gu[1])));
        if (k >= 1) {
#pragma unroll
            for (int a = 0; a < 7; a++) ng = nmaxd(ng, fabs(gx[a]));
        }
#pragma unroll
        for (int a = 0; a < 7; a++) nb = nmaxd(nb, fabs(AT(lin, DL_b + a)));
#pragma unroll
        for (int a = 0; a < 7; a++) pim[a] = pi[a];
    }
    {
        const double *lin = P.lin_d + (size_t)N * DL_ROWS * Bp;
#pragma unroll
        for (int a = 0; a < 7; a++) ng = nmaxd(ng, fabs(AT(lin, DL_q + a) - pim[a]));
    }
    AT(P.nlp_res, 0) = ng; AT(P.nlp_res, 1) = nb; AT(P.nlp_res, 2) = nd; AT(P.nlp_res, 3) = nm;
    if (ng < tol0 && nb < tol1 && nd < tol2 && nm < tol3) {
        P.sqp_status[i] = 0; P.sqp_iter[i] = it; P.status[i] = 0; P.lin_bad[i] = 2;
        return;
    }
    P.sqp_iter[i] = it + 1;
    atomicAdd(active, 1);
}
void launch_nlp_res_dense(const Params &P, int it, const double tol[4], int *active, cudaStream_t s)
{
    nlp_res_dense_kernel<<<(P.B + 127) / 128, 128, 0, s>>>(P, it, tol[0], tol[1], tol[2], tol[3], active);
}

void launch_prepare_dense(const Params &P, cudaStream_t s)
{
    dim3 grid((P.B + 127) / 128, P.o.N + 1);
    if (P.o.gp_enabled) {
        launch_gp_sweep(P, s);                 // pass 1: GP mean / gradient at the RK4 stage points (shared with the Cartesian model)
        prepare_dense_kernel<true><<<grid, 128, 0, s>>>(P);
    } else {
        prepare_dense_kernel<false><<<grid, 128, 0, s>>>(P);
    }
}
void launch_qp_dense(const Params &P, cudaStream_t s) { qp_dense_kernel<<<(P.B + 63) / 64, 64, 0, s>>>(P); }
