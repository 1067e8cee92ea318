// qp_ipm.cu -- feedback phase of the SQP-RTI step: OCP-structured primal-dual interior-point QP solver,
// Mehrotra predictor-corrector, Riccati factorisation over the horizon.  v1: one thread per MPC instance, all
// per-stage data streamed from HBM/L2 in SoA layout (fully coalesced).
//
// Replaces HPIPM (full condensing + dense IPM, acados_solver_sim_car.c:145,688-693) [EXT]; the reference condenses,
// this kernel keeps the stage structure: same QP, same solution.
//
// Model-specific structure that is exploited: A_k = [e0 e1 | a(6x5); 0 0 | 0 0 0 0 1],
// B_k = [bm(6x2); 0 dt], Q, R diagonal, soft input bounds eliminated analytically, one hard bound on x[6].
#include "common.cuh"

#define AT(arr, row) (arr)[(size_t)(row) * Bp + i]

// NaN-propagating max for the residual norms (fmax would hide a diverged instance)
__device__ __forceinline__ double nmax1(double a, double b) { return (a > b || a != a) ? a : b; }

// symmetric 7x7 packed lower: idx(i,j), i>=j
__device__ __forceinline__ constexpr int sidx(int i, int j) { return (i >= j) ? (i * (i + 1) / 2 + j) : (j * (j + 1) / 2 + i); }

struct StageBar {          // barrier quantities of one stage (recomputed where needed)
    double Rt[2], Qt6;     // modified Hessian diagonals
    double rt[2], qt6;     // modified gradients
    double cl[2], cu[2], Dl[2], Du[2], Sl[2], Su[2];
};

// CORR: the complementarity rhs is first replaced by the Mehrotra-corrected one
template <bool CORR>
__device__ __forceinline__ void stage_barrier(const Params &P, int k, int i, int Bp, double sigmu, StageBar &sb)
{
    const admpc_opts &o = P.o;
    const double Ts = o.dt;
    double lam[NC], t[NC], g[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        lam[c] = AT(P.lam, k * NC + c);
        t[c] = AT(P.t, k * NC + c);
        double rm = AT(P.rm, k * NC + c);
        if (CORR) {
            const bool on = !((c == 2 || c == 5) && k == 0);
            rm = on ? rm + AT(P.dlam, k * NC + c) * AT(P.dt, k * NC + c) - sigmu : 0.0;
            AT(P.rm, k * NC + c) = rm;
        }
        g[c] = (rm - lam[c] * AT(P.rd, k * NC + c)) / t[c];
    }
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const double Sl = lam[j] / t[j], Su = lam[3 + j] / t[3 + j];
        const double Ssl = lam[6 + j] / t[6 + j], Ssu = lam[8 + j] / t[8 + j];
        const double Dl = Ts * o.Zl[j] + Sl + Ssl, Du = Ts * o.Zu[j] + Su + Ssu;
        sb.Sl[j] = Sl; sb.Su[j] = Su; sb.Dl[j] = Dl; sb.Du[j] = Du;
        sb.Rt[j] = Ts * o.W[7 + j] + Sl * (1.0 - Sl / Dl) + Su * (1.0 - Su / Du);
        sb.cl[j] = AT(P.rgsl, k * 2 + j) + g[j] + g[6 + j];
        sb.cu[j] = AT(P.rgsu, k * 2 + j) + g[3 + j] + g[8 + j];
        sb.rt[j] = AT(P.rgu, k * 2 + j) + (g[j] - Sl * sb.cl[j] / Dl) - (g[3 + j] - Su * sb.cu[j] / Du);
    }
    if (k >= 1) {
        sb.Qt6 = Ts * o.W[6] + lam[2] / t[2] + lam[5] / t[5];
        sb.qt6 = AT(P.rgx, k * 7 + 6) + g[2] - g[5];
    } else {
        sb.Qt6 = Ts * o.W[6];
        sb.qt6 = 0.0;
    }
}

// backward sweep. FACTOR: also (re)build K, Ginv, P (matrix part); always: vector part kf, pv.
template <bool FACTOR, bool CORR>
__device__ __forceinline__ void backward_pass(const Params &P, int i, double sigmu)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const double Ts = o.dt, hdt = o.dt;
    double Pm[28];      // P_{k+1}, packed symmetric (only when FACTOR)
    double pv[7];       // p_{k+1}
    if (FACTOR) {
#pragma unroll
        for (int a = 0; a < 28; a++) Pm[a] = 0.0;
#pragma unroll
        for (int a = 0; a < 7; a++) Pm[sidx(a, a)] = o.We[a];
#pragma unroll
        for (int a = 0; a < 28; a++) AT(P.P, N * 28 + a) = Pm[a];
    }
#pragma unroll
    for (int a = 0; a < 7; a++) { pv[a] = AT(P.rgx, N * 7 + a); AT(P.pv, N * 7 + a) = pv[a]; }

    for (int k = N - 1; k >= 0; k--) {
        StageBar sb;
        stage_barrier<CORR>(P, k, i, Bp, sigmu, sb);
        const double *lin = P.lin + (size_t)k * LIN_ROWS * Bp;
        // M = [B | A(:,2:7)] rows 0..5 ; row 6 = [0 dt 0 0 0 0 1]
        double Mx[6][7];
#pragma unroll
        for (int r = 0; r < 6; r++) {
            Mx[r][0] = AT(lin, LIN_B + r * 2 + 0);
            Mx[r][1] = AT(lin, LIN_B + r * 2 + 1);
#pragma unroll
            for (int c = 0; c < 5; c++) Mx[r][2 + c] = AT(lin, LIN_A + r * 5 + c);
        }
        double Kk[2][7], gi00, gi01, gi11;
        double hv[7];
        if (FACTOR) {
            double rb[7], Pb[7];
#pragma unroll
            for (int a = 0; a < 7; a++) rb[a] = AT(P.rb, k * 7 + a);
#pragma unroll
            for (int a = 0; a < 7; a++) {
                double v = 0.0;
#pragma unroll
                for (int l = 0; l < 7; l++) v = fma(Pm[sidx(a, l)], rb[l], v);
                Pb[a] = v;
                AT(P.Pb, k * 7 + a) = v;
                hv[a] = v + pv[a];
            }
            // G = diag(Rt,Qt) + [B A]^T P [B A], 9x9 symmetric; v index: 0,1 = u ; 2..8 = x0..x6
            // M-column m -> v index: m<2 ? m : m+2
            double G[45];   // packed lower 9x9: gidx(a,b) = a*(a+1)/2+b
#define GI(a, b) ((a) >= (b) ? ((a) * ((a) + 1) / 2 + (b)) : ((b) * ((b) + 1) / 2 + (a)))
            G[GI(2, 2)] = Pm[sidx(0, 0)]; G[GI(3, 2)] = Pm[sidx(1, 0)]; G[GI(3, 3)] = Pm[sidx(1, 1)];
#pragma unroll
            for (int m = 0; m < 7; m++) {
                const int vm = (m < 2) ? m : m + 2;
                const double m6 = (m == 1) ? hdt : ((m == 6) ? 1.0 : 0.0);
                double w[7];
#pragma unroll
                for (int a = 0; a < 7; a++) {
                    double v = Pm[sidx(a, 6)] * m6;
#pragma unroll
                    for (int l = 0; l < 6; l++) v = fma(Pm[sidx(a, l)], Mx[l][m], v);
                    w[a] = v;
                }
                G[GI(vm, 2)] = w[0];     // x0 column of [B A] is e0
                G[GI(vm, 3)] = w[1];
#pragma unroll
                for (int m2 = 0; m2 <= m; m2++) {
                    const int vm2 = (m2 < 2) ? m2 : m2 + 2;
                    const double m26 = (m2 == 1) ? hdt : ((m2 == 6) ? 1.0 : 0.0);
                    double v = m26 * w[6];
#pragma unroll
                    for (int l = 0; l < 6; l++) v = fma(Mx[l][m2], w[l], v);
                    G[GI(vm, vm2)] = v;
                }
            }
            G[GI(0, 0)] += sb.Rt[0]; G[GI(1, 1)] += sb.Rt[1];
#pragma unroll
            for (int a = 0; a < 6; a++) G[GI(2 + a, 2 + a)] += Ts * o.W[a];
            G[GI(8, 8)] += sb.Qt6;
            // inverse of the 2x2 input block
            const double g00 = G[GI(0, 0)] + o.reg, g11 = G[GI(1, 1)] + o.reg, g01 = G[GI(1, 0)];
            const double idet = 1.0 / (g00 * g11 - g01 * g01);
            gi00 = g11 * idet; gi01 = -g01 * idet; gi11 = g00 * idet;
            AT(P.Ginv, k * 3 + 0) = gi00; AT(P.Ginv, k * 3 + 1) = gi01; AT(P.Ginv, k * 3 + 2) = gi11;
#pragma unroll
            for (int j = 0; j < 7; j++) {
                const double a0 = G[GI(2 + j, 0)], a1 = G[GI(2 + j, 1)];
                Kk[0][j] = -(gi00 * a0 + gi01 * a1);
                Kk[1][j] = -(gi01 * a0 + gi11 * a1);
                AT(P.K, k * 14 + j) = Kk[0][j];
                AT(P.K, k * 14 + 7 + j) = Kk[1][j];
            }
            // P_k = Gxx + Gxu K   (lower triangle)
#pragma unroll
            for (int a = 0; a < 7; a++)
#pragma unroll
                for (int b = 0; b <= a; b++) {
                    const double v = G[GI(2 + a, 2 + b)] + G[GI(2 + a, 0)] * Kk[0][b] + G[GI(2 + a, 1)] * Kk[1][b];
                    Pm[sidx(a, b)] = v;
                    AT(P.P, k * 28 + sidx(a, b)) = v;
                }
#undef GI
        } else {
#pragma unroll
            for (int a = 0; a < 7; a++) hv[a] = AT(P.Pb, k * 7 + a) + pv[a];
            gi00 = AT(P.Ginv, k * 3 + 0); gi01 = AT(P.Ginv, k * 3 + 1); gi11 = AT(P.Ginv, k * 3 + 2);
#pragma unroll
            for (int j = 0; j < 7; j++) { Kk[0][j] = AT(P.K, k * 14 + j); Kk[1][j] = AT(P.K, k * 14 + 7 + j); }
        }
        // vector part
        double gu[2], gx[7];
        gu[0] = sb.rt[0];
        gu[1] = fma(hdt, hv[6], sb.rt[1]);
#pragma unroll
        for (int l = 0; l < 6; l++) { gu[0] = fma(Mx[l][0], hv[l], gu[0]); gu[1] = fma(Mx[l][1], hv[l], gu[1]); }
        if (k >= 1) {
            gx[0] = AT(P.rgx, k * 7 + 0) + hv[0];
            gx[1] = AT(P.rgx, k * 7 + 1) + hv[1];
#pragma unroll
            for (int c = 0; c < 5; c++) {
                double v = (c == 4) ? (sb.qt6 + hv[6]) : AT(P.rgx, k * 7 + 2 + c);
#pragma unroll
                for (int l = 0; l < 6; l++) v = fma(Mx[l][2 + c], hv[l], v);
                gx[2 + c] = v;
            }
        } else {
#pragma unroll
            for (int a = 0; a < 7; a++) gx[a] = 0.0;
        }
        const double kf0 = -(gi00 * gu[0] + gi01 * gu[1]), kf1 = -(gi01 * gu[0] + gi11 * gu[1]);
        AT(P.kf, k * 2 + 0) = kf0;
        AT(P.kf, k * 2 + 1) = kf1;
#pragma unroll
        for (int a = 0; a < 7; a++) {
            pv[a] = gx[a] + Kk[0][a] * gu[0] + Kk[1][a] * gu[1];
            AT(P.pv, k * 7 + a) = pv[a];
        }
    }
}

// forward sweep: Newton step for all variables, fraction-to-boundary step length and the three sums that give
// mu_aff(alpha) = (s0 + alpha*s1 + alpha^2*s2)/nc.
__device__ __forceinline__ void forward_pass(const Params &P, int i, double &alpha, double &s1, double &s2)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const double Ts = o.dt, hdt = o.dt;
    double dxk[7];
#pragma unroll
    for (int a = 0; a < 7; a++) { dxk[a] = 0.0; AT(P.ddx, a) = 0.0; }
    alpha = 1.0; s1 = 0.0; s2 = 0.0;
    for (int k = 0; k < N; k++) {
        const double *lin = P.lin + (size_t)k * LIN_ROWS * Bp;
        double du[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            double v = AT(P.kf, k * 2 + j);
#pragma unroll
            for (int l = 0; l < 7; l++) v = fma(AT(P.K, k * 14 + j * 7 + l), dxk[l], v);
            du[j] = v;
            AT(P.ddu, k * 2 + j) = v;
        }
        // constraint steps of this stage (uses dxk[6] and du)
        double lam[NC], t[NC], dtv[NC], gq[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) {
            lam[c] = AT(P.lam, k * NC + c);
            t[c] = AT(P.t, k * NC + c);
            gq[c] = (AT(P.rm, k * NC + c) - lam[c] * AT(P.rd, k * NC + c)) / t[c];
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const double Sl = lam[j] / t[j], Su = lam[3 + j] / t[3 + j];
            const double Ssl = lam[6 + j] / t[6 + j], Ssu = lam[8 + j] / t[8 + j];
            const double Dl = Ts * o.Zl[j] + Sl + Ssl, Du = Ts * o.Zu[j] + Su + Ssu;
            const double cl = AT(P.rgsl, k * 2 + j) + gq[j] + gq[6 + j];
            const double cu = AT(P.rgsu, k * 2 + j) + gq[3 + j] + gq[8 + j];
            const double dsl = -(cl + Sl * du[j]) / Dl;
            const double dsu = -(cu - Su * du[j]) / Du;
            AT(P.dsl, k * 2 + j) = dsl;
            AT(P.dsu, k * 2 + j) = dsu;
            dtv[j] = du[j] + dsl - AT(P.rd, k * NC + j);
            dtv[3 + j] = -du[j] + dsu - AT(P.rd, k * NC + 3 + j);
            dtv[6 + j] = dsl - AT(P.rd, k * NC + 6 + j);
            dtv[8 + j] = dsu - AT(P.rd, k * NC + 8 + j);
        }
        if (k >= 1) {
            dtv[2] = dxk[6] - AT(P.rd, k * NC + 2);
            dtv[5] = -dxk[6] - AT(P.rd, k * NC + 5);
        } else {
            dtv[2] = 0.0; dtv[5] = 0.0;
        }
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const bool on = !((c == 2 || c == 5) && k == 0);
            const double dl = on ? -(AT(P.rm, k * NC + c) + lam[c] * dtv[c]) / t[c] : 0.0;
            AT(P.dlam, k * NC + c) = dl;
            AT(P.dt, k * NC + c) = dtv[c];
            if (on) {
                if (dl < 0.0) alpha = fmin(alpha, -lam[c] / dl);
                if (dtv[c] < 0.0) alpha = fmin(alpha, -t[c] / dtv[c]);
                s1 += lam[c] * dtv[c] + t[c] * dl;
                s2 += dl * dtv[c];
            }
        }
        // dx_{k+1} = A dx_k + B du + rb
        double dxn[7];
#pragma unroll
        for (int r = 0; r < 6; r++) {
            double v = AT(P.rb, k * 7 + r) + ((r < 2) ? dxk[r] : 0.0);
            v = fma(AT(lin, LIN_B + r * 2 + 0), du[0], v);
            v = fma(AT(lin, LIN_B + r * 2 + 1), du[1], v);
#pragma unroll
            for (int c = 0; c < 5; c++) v = fma(AT(lin, LIN_A + r * 5 + c), dxk[2 + c], v);
            dxn[r] = v;
        }
        dxn[6] = AT(P.rb, k * 7 + 6) + dxk[6] + hdt * du[1];
#pragma unroll
        for (int a = 0; a < 7; a++) { dxk[a] = dxn[a]; AT(P.ddx, (k + 1) * 7 + a) = dxn[a]; }
        // dpi_k = P_{k+1} dx_{k+1} + p_{k+1}
#pragma unroll
        for (int a = 0; a < 7; a++) {
            double v = AT(P.pv, (k + 1) * 7 + a);
#pragma unroll
            for (int l = 0; l < 7; l++) v = fma(AT(P.P, (k + 1) * 28 + sidx(a, l)), dxn[l], v);
            AT(P.dpi, k * 7 + a) = v;
        }
    }
}

__global__ void __launch_bounds__(128) qp_ipm_kernel(const Params P)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.B) return;
    const double Ts = o.dt, hdt = o.dt;

    if (const int flag = P.lin_bad[i]) {     // 1: NaN/Inf in the linearisation: ACADOS_FAILURE, iterate untouched
        if (flag == 1) { P.status[i] = 1; P.qp_status[i] = 0; P.qp_iter[i] = 0; }
        return;                              // 2: finished instance of the full-SQP loop
    }
    // ---- cold start: primal at 0 pushed thr0 inside its box, t from the box, lam = mu0/t -------------------
#pragma unroll
    for (int a = 0; a < 7; a++) AT(P.dx, a) = AT(P.x0, a) - AT(P.xb, a);
    for (int k = 0; k < N; k++) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (j == 2 && k == 0) {
                AT(P.t, 2) = 1.0; AT(P.t, 5) = 1.0; AT(P.lam, 2) = 0.0; AT(P.lam, 5) = 0.0;
                continue;
            }
            const double cur = (j < 2) ? AT(P.ub, k * 2 + j) : AT(P.xb, k * 7 + 6);
            const double lo = ((j < 2) ? o.lbu[j] : o.lbx) - cur, hi = ((j < 2) ? o.ubu[j] : o.ubx) - cur;
            double v = 0.0;
            if (v - lo < o.thr0) {
                if (hi - v < o.thr0) v = 0.5 * (lo + hi);
                else v = lo + o.thr0;
            } else if (hi - v < o.thr0) v = hi - o.thr0;
            if (j < 2) AT(P.du, k * 2 + j) = v; else AT(P.dx, k * 7 + 6) = v;
            const double tl = fmax(o.thr0, v - lo), tu = fmax(o.thr0, hi - v);
            AT(P.t, k * NC + j) = tl; AT(P.t, k * NC + 3 + j) = tu;
            AT(P.lam, k * NC + j) = o.mu0 / tl; AT(P.lam, k * NC + 3 + j) = o.mu0 / tu;
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            AT(P.t, k * NC + 6 + j) = o.thr0; AT(P.t, k * NC + 8 + j) = o.thr0;
            AT(P.lam, k * NC + 6 + j) = o.mu0 / o.thr0; AT(P.lam, k * NC + 8 + j) = o.mu0 / o.thr0;
            AT(P.sl, k * 2 + j) = 0.0; AT(P.su, k * 2 + j) = 0.0;
        }
#pragma unroll
        for (int a = 0; a < 7; a++) AT(P.pi, k * 7 + a) = 0.0;
#pragma unroll
        for (int a = 0; a < 6; a++) AT(P.dx, (k + 1) * 7 + a) = 0.0;
        if (k + 1 == N) AT(P.dx, N * 7 + 6) = 0.0;
    }

    const double inv_nc = 1.0 / (double)(NC * N - 2);
    int status = 1, iter = 0;
    double res0 = 0, res1 = 0, res2 = 0, res3 = 0;
    for (iter = 0;; iter++) {
        // ---- residuals of the current point ------------------------------------------------------------
        double ng = 0, nb = 0, nd = 0, nm = 0, summ = 0;
        double dxk[7], pim[7];   // dx_k, pi_{k-1}
#pragma unroll
        for (int a = 0; a < 7; a++) { dxk[a] = AT(P.dx, a); pim[a] = 0.0; }
        for (int k = 0; k < N; k++) {
            const double *lin = P.lin + (size_t)k * LIN_ROWS * Bp;
            double du[2], pik[7], dxn[7], lam[NC], t[NC];
#pragma unroll
            for (int j = 0; j < 2; j++) du[j] = AT(P.du, k * 2 + j);
#pragma unroll
            for (int a = 0; a < 7; a++) { pik[a] = AT(P.pi, k * 7 + a); dxn[a] = AT(P.dx, (k + 1) * 7 + a); }
#pragma unroll
            for (int c = 0; c < NC; c++) { lam[c] = AT(P.lam, k * NC + c); t[c] = AT(P.t, k * NC + c); }
            double Mx[6][7];
#pragma unroll
            for (int r = 0; r < 6; r++) {
                Mx[r][0] = AT(lin, LIN_B + r * 2 + 0);
                Mx[r][1] = AT(lin, LIN_B + r * 2 + 1);
#pragma unroll
                for (int c = 0; c < 5; c++) Mx[r][2 + c] = AT(lin, LIN_A + r * 5 + c);
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                double g = Ts * o.W[7 + j] * du[j] + AT(lin, LIN_r + j) - lam[j] + lam[3 + j];
#pragma unroll
                for (int l = 0; l < 6; l++) g = fma(Mx[l][j], pik[l], g);
                if (j == 1) g = fma(hdt, pik[6], g);
                AT(P.rgu, k * 2 + j) = g;
                const double sl = AT(P.sl, k * 2 + j), su = AT(P.su, k * 2 + j);
                const double gsl = Ts * o.zl[j] + Ts * o.Zl[j] * sl - lam[j] - lam[6 + j];
                const double gsu = Ts * o.zu[j] + Ts * o.Zu[j] * su - lam[3 + j] - lam[8 + j];
                AT(P.rgsl, k * 2 + j) = gsl;
                AT(P.rgsu, k * 2 + j) = gsu;
                ng = nmax1(ng, nmax1(fabs(g), nmax1(fabs(gsl), fabs(gsu))));
                const double lo = o.lbu[j] - AT(P.ub, k * 2 + j), hi = o.ubu[j] - AT(P.ub, k * 2 + j);
                const double r0 = t[j] - (du[j] - lo + sl), r1 = t[3 + j] - (hi - du[j] + su);
                const double r2 = t[6 + j] - sl, r3 = t[8 + j] - su;
                AT(P.rd, k * NC + j) = r0; AT(P.rd, k * NC + 3 + j) = r1;
                AT(P.rd, k * NC + 6 + j) = r2; AT(P.rd, k * NC + 8 + j) = r3;
                nd = nmax1(nd, nmax1(nmax1(fabs(r0), fabs(r1)), nmax1(fabs(r2), fabs(r3))));
            }
            if (k >= 1) {
                const double cur = AT(P.xb, k * 7 + 6);
                const double r0 = t[2] - (dxk[6] - (o.lbx - cur)), r1 = t[5] - ((o.ubx - cur) - dxk[6]);
                AT(P.rd, k * NC + 2) = r0; AT(P.rd, k * NC + 5) = r1;
                nd = nmax1(nd, nmax1(fabs(r0), fabs(r1)));
            } else {
                AT(P.rd, 2) = 0.0; AT(P.rd, 5) = 0.0;
            }
            // dynamics residual
#pragma unroll
            for (int r = 0; r < 6; r++) {
                double v = AT(lin, LIN_b + r) - dxn[r] + ((r < 2) ? dxk[r] : 0.0);
                v = fma(Mx[r][0], du[0], v);
                v = fma(Mx[r][1], du[1], v);
#pragma unroll
                for (int c = 0; c < 5; c++) v = fma(Mx[r][2 + c], dxk[2 + c], v);
                AT(P.rb, k * 7 + r) = v;
                nb = nmax1(nb, fabs(v));
            }
            {
                const double v = AT(lin, LIN_b + 6) - dxn[6] + dxk[6] + hdt * du[1];
                AT(P.rb, k * 7 + 6) = v;
                nb = nmax1(nb, fabs(v));
            }
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const bool on = !((c == 2 || c == 5) && k == 0);
                const double m = on ? lam[c] * t[c] : 0.0;
                AT(P.rm, k * NC + c) = m;
                nm = nmax1(nm, fabs(m));
                summ += m;
            }
            // stationarity wrt x_k (k >= 1): Q dx + q + A^T pi_k - pi_{k-1} -/+ lam_x
            if (k >= 1) {
#pragma unroll
                for (int a = 0; a < 7; a++) {
                    double g = Ts * o.W[a] * dxk[a] + AT(lin, LIN_q + a) - pim[a];
                    if (a < 2) g += pik[a];
                    else {
#pragma unroll
                        for (int l = 0; l < 6; l++) g = fma(Mx[l][a], pik[l], g);
                        if (a == 6) g += pik[6] - lam[2] + lam[5];
                    }
                    AT(P.rgx, k * 7 + a) = g;
                    ng = nmax1(ng, fabs(g));
                }
            }
#pragma unroll
            for (int a = 0; a < 7; a++) { dxk[a] = dxn[a]; pim[a] = pik[a]; }
        }
        {   // terminal stage
            const double *lin = P.lin + (size_t)N * LIN_ROWS * Bp;
#pragma unroll
            for (int a = 0; a < 7; a++) {
                const double g = o.We[a] * dxk[a] + AT(lin, LIN_q + a) - pim[a];
                AT(P.rgx, N * 7 + a) = g;
                ng = nmax1(ng, fabs(g));
            }
        }
        const double mu = summ * inv_nc;
        res0 = ng; res1 = nb; res2 = nd; res3 = nm;
        if (!(isfinite(ng) && isfinite(nb) && isfinite(nd) && isfinite(nm))) { status = 3; break; }
        if (ng < o.tol_stat && nb < o.tol_eq && nd < o.tol_ineq && nm < o.tol_comp) { status = 0; break; }
        if (iter >= o.iter_max) { status = 1; break; }

        // ---- predictor ---------------------------------------------------------------------------------
        double a_aff, s1, s2;
        backward_pass<true, false>(P, i, 0.0);
        forward_pass(P, i, a_aff, s1, s2);
        const double mu_aff = (summ + a_aff * s1 + a_aff * a_aff * s2) * inv_nc;
        double sigma = mu_aff / mu;
        sigma = sigma * sigma * sigma;
        // ---- corrector ---------------------------------------------------------------------------------
        double alpha;
        backward_pass<false, true>(P, i, sigma * mu);
        forward_pass(P, i, alpha, s1, s2);
        if (alpha < o.alpha_min) { status = 2; break; }
        if (alpha < 1.0) alpha *= 0.995;
        // ---- update ------------------------------------------------------------------------------------
        for (int k = 0; k < N; k++) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                AT(P.du, k * 2 + j) += alpha * AT(P.ddu, k * 2 + j);
                AT(P.sl, k * 2 + j) += alpha * AT(P.dsl, k * 2 + j);
                AT(P.su, k * 2 + j) += alpha * AT(P.dsu, k * 2 + j);
            }
#pragma unroll
            for (int a = 0; a < 7; a++) {
                AT(P.dx, (k + 1) * 7 + a) += alpha * AT(P.ddx, (k + 1) * 7 + a);
                AT(P.pi, k * 7 + a) += alpha * AT(P.dpi, k * 7 + a);
            }
#pragma unroll
            for (int c = 0; c < NC; c++) {
                if ((c == 2 || c == 5) && k == 0) continue;
                AT(P.lam, k * NC + c) = fmax(AT(P.lam, k * NC + c) + alpha * AT(P.dlam, k * NC + c), o.lam_min);
                AT(P.t, k * NC + c) = fmax(AT(P.t, k * NC + c) + alpha * AT(P.dt, k * NC + c), o.t_min);
            }
        }
    }
    // hpipm {0 ok,1 maxiter,2 minstep,3 nan} -> acados {0,2,3,1}; RTI tolerates maxiter
    const int qps = (status == 0) ? 0 : ((status == 1) ? 2 : ((status == 2) ? 3 : 1));
    P.qp_status[i] = qps;
    P.qp_iter[i] = iter;
    P.status[i] = (qps == 0 || qps == 2) ? 0 : 4;
    AT(P.res_out, 0) = res0; AT(P.res_out, 1) = res1; AT(P.res_out, 2) = res2; AT(P.res_out, 3) = res3;
}

// update phase of the RTI step: full step on primals, duals <- QP duals (step_length 1, fixed_step;
// acados_solver_sim_car.c:647-681).  One thread per (instance, row).
__global__ void update_kernel(const Params P)
{
    const int Bp = P.Bp, N = P.o.N;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (i >= P.B) return;
    if (P.status[i] != 0 || P.lin_bad[i] == 2) return;     // failed QP, or finished instance of the full-SQP loop
    if (row < (N + 1) * 7) AT(P.xb, row) += AT(P.dx, row);
    if (row < N * 2) {
        AT(P.ub, row) += AT(P.du, row);
        AT(P.slb, row) = AT(P.sl, row);
        AT(P.sub, row) = AT(P.su, row);
    }
    if (row < N * 7) AT(P.pib, row) = AT(P.pi, row);
    if (row < N * con_rows(P.o)) { AT(P.lamb, row) = AT(P.lam, row); AT(P.tb, row) = AT(P.t, row); }
}

void launch_qp(const Params &P, cudaStream_t s)
{
    qp_ipm_kernel<<<(P.B + 127) / 128, 128, 0, s>>>(P);
}

void launch_update(const Params &P, cudaStream_t s)
{
    const int N = P.o.N;
    int rows = N * con_rows(P.o);
    if ((N + 1) * 7 > rows) rows = (N + 1) * 7;
    dim3 grid((P.B + 127) / 128, rows);
    update_kernel<<<grid, 128, 0, s>>>(P);
}
