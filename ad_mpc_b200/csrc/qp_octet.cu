// qp_octet.cu -- feedback phase, v2: EIGHT lanes per MPC instance ("octet"), four instances per warp.
//
// Why: the benchmark batches (4096..16384 instances per GPU) are far too small to fill 148 SMs with one thread per
// instance (16384 threads = 3.5 warps/SM).  Here the horizon-sequential Riccati sweeps are parallelised over the
// columns/rows of the 7x7 stage matrices (lane c <-> column c of [B | A(:,2:7)], lane 7 <-> the vector column), with
// warp shuffles (width 8) for the 2x2 pivot / gains and shared memory for the P and M broadcasts; the stage-local
// parts of the IPM (residuals, barrier terms, slack/multiplier steps, step length, update) run stage-parallel,
// lane s <-> stages s, s+8, s+16, ...
//
// Same algorithm, same operation order per entry as qp_ipm.cu (v1, kept as the large-batch / reference variant):
// HPIPM-style Mehrotra predictor-corrector on the OCP-structured QP [EXT], replacing FULL_CONDENSING_HPIPM
// (acados_solver_sim_car.c:145,688-693).
#include "common.cuh"

#define FULL 0xffffffffu
#define AT(arr, row) (arr)[(size_t)(row) * Bp + i]
#define OCT_PER_CTA 16

__device__ __forceinline__ constexpr int sidx(int i, int j) { return (i >= j) ? (i * (i + 1) / 2 + j) : (j * (j + 1) / 2 + i); }
__device__ __forceinline__ double shfl8(double v, int src) { return __shfl_sync(FULL, v, src, 8); }
__device__ __forceinline__ double sum8(double v)
{
    v += __shfl_xor_sync(FULL, v, 4, 8); v += __shfl_xor_sync(FULL, v, 2, 8); v += __shfl_xor_sync(FULL, v, 1, 8);
    return v;
}
// NaN-propagating max (fmax would drop NaNs and hide a diverged instance)
__device__ __forceinline__ double nmax(double a, double b) { return (a > b || a != a) ? a : b; }
__device__ __forceinline__ double max8(double v)
{
    v = nmax(v, __shfl_xor_sync(FULL, v, 4, 8)); v = nmax(v, __shfl_xor_sync(FULL, v, 2, 8)); v = nmax(v, __shfl_xor_sync(FULL, v, 1, 8));
    return v;
}
__device__ __forceinline__ double min8(double v)
{
    v = fmin(v, __shfl_xor_sync(FULL, v, 4, 8)); v = fmin(v, __shfl_xor_sync(FULL, v, 2, 8)); v = fmin(v, __shfl_xor_sync(FULL, v, 1, 8));
    return v;
}

// barrier-modified Hessian diagonal / gradient of one stage, soft-bound slacks eliminated
struct BarOut { double Rt[2], Qt6, rt[2], qt6; };
__device__ __forceinline__ void barrier_terms(const admpc_opts &o, int k, const double lam[NC], const double t[NC],
                                              const double rd[NC], const double rm[NC], const double rgsl[2],
                                              const double rgsu[2], const double rgu[2], double rgx6, BarOut &b)
{
    const double Ts = o.dt;
    double g[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) g[c] = (rm[c] - lam[c] * rd[c]) / t[c];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const double Sl = lam[j] / t[j], Su = lam[3 + j] / t[3 + j];
        const double Ssl = lam[6 + j] / t[6 + j], Ssu = lam[8 + j] / t[8 + j];
        const double Dl = Ts * o.Zl[j] + Sl + Ssl, Du = Ts * o.Zu[j] + Su + Ssu;
        b.Rt[j] = Ts * o.W[7 + j] + Sl * (1.0 - Sl / Dl) + Su * (1.0 - Su / Du);
        const double cl = rgsl[j] + g[j] + g[6 + j];
        const double cu = rgsu[j] + g[3 + j] + g[8 + j];
        b.rt[j] = rgu[j] + (g[j] - Sl * cl / Dl) - (g[3 + j] - Su * cu / Du);
    }
    if (k >= 1) {
        b.Qt6 = Ts * o.W[6] + lam[2] / t[2] + lam[5] / t[5];
        b.qt6 = rgx6 + g[2] - g[5];
    } else {
        b.Qt6 = Ts * o.W[6];
        b.qt6 = 0.0;
    }
}

// ---- stage-parallel: residuals of the current point + predictor barrier terms -------------------------------
__device__ __forceinline__ void pass_residual(const Params &P, int i, int s, bool act, double &ng, double &nb,
                                              double &nd, double &nm, double &summ)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const double Ts = o.dt, hdt = o.dt;
    ng = nb = nd = nm = summ = 0.0;
    for (int k = s; k <= N; k += 8) {
        if (k == N) {
            const double *lin = P.lin + (size_t)N * LIN_ROWS * Bp;
#pragma unroll
            for (int a = 0; a < 7; a++) {
                const double g = o.We[a] * AT(P.dx, N * 7 + a) + AT(lin, LIN_q + a) - AT(P.pi, (N - 1) * 7 + a);
                if (act) AT(P.rgx, N * 7 + a) = g;
                ng = nmax(ng, fabs(g));
            }
            continue;
        }
        const double *lin = P.lin + (size_t)k * LIN_ROWS * Bp;
        double du[2], pik[7], pim[7], dxk[7], dxn[7], lam[NC], t[NC], rd[NC], rm[NC], rgu[2], rgsl[2], rgsu[2];
#pragma unroll
        for (int j = 0; j < 2; j++) du[j] = AT(P.du, k * 2 + j);
#pragma unroll
        for (int a = 0; a < 7; a++) {
            pik[a] = AT(P.pi, k * 7 + a);
            pim[a] = (k >= 1) ? AT(P.pi, (k - 1) * 7 + a) : 0.0;
            dxk[a] = AT(P.dx, k * 7 + a);
            dxn[a] = AT(P.dx, (k + 1) * 7 + a);
        }
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
            rgu[j] = g;
            const double sl = AT(P.sl, k * 2 + j), su = AT(P.su, k * 2 + j);
            rgsl[j] = Ts * o.zl[j] + Ts * o.Zl[j] * sl - lam[j] - lam[6 + j];
            rgsu[j] = Ts * o.zu[j] + Ts * o.Zu[j] * su - lam[3 + j] - lam[8 + j];
            ng = nmax(ng, nmax(fabs(g), nmax(fabs(rgsl[j]), fabs(rgsu[j]))));
            const double cur = AT(P.ub, k * 2 + j);
            const double lo = o.lbu[j] - cur, hi = o.ubu[j] - cur;
            rd[j] = t[j] - (du[j] - lo + sl);
            rd[3 + j] = t[3 + j] - (hi - du[j] + su);
            rd[6 + j] = t[6 + j] - sl;
            rd[8 + j] = t[8 + j] - su;
            nd = nmax(nd, nmax(nmax(fabs(rd[j]), fabs(rd[3 + j])), nmax(fabs(rd[6 + j]), fabs(rd[8 + j]))));
        }
        if (k >= 1) {
            const double cur = AT(P.xb, k * 7 + 6);
            rd[2] = t[2] - (dxk[6] - (o.lbx - cur));
            rd[5] = t[5] - ((o.ubx - cur) - dxk[6]);
            nd = nmax(nd, nmax(fabs(rd[2]), fabs(rd[5])));
        } else {
            rd[2] = 0.0; rd[5] = 0.0;
        }
#pragma unroll
        for (int r = 0; r < 6; r++) {
            double v = AT(lin, LIN_b + r) - dxn[r] + ((r < 2) ? dxk[r] : 0.0);
            v = fma(Mx[r][0], du[0], v);
            v = fma(Mx[r][1], du[1], v);
#pragma unroll
            for (int c = 0; c < 5; c++) v = fma(Mx[r][2 + c], dxk[2 + c], v);
            if (act) AT(P.rb, k * 7 + r) = v;
            nb = nmax(nb, fabs(v));
        }
        {
            const double v = AT(lin, LIN_b + 6) - dxn[6] + dxk[6] + hdt * du[1];
            if (act) AT(P.rb, k * 7 + 6) = v;
            nb = nmax(nb, fabs(v));
        }
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const bool on = !((c == 2 || c == 5) && k == 0);
            rm[c] = on ? lam[c] * t[c] : 0.0;
            nm = nmax(nm, fabs(rm[c]));
            summ += rm[c];
        }
        double rgx6 = 0.0;
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
                if (act) AT(P.rgx, k * 7 + a) = g;
                if (a == 6) rgx6 = g;
                ng = nmax(ng, fabs(g));
            }
        }
        BarOut b;
        barrier_terms(o, k, lam, t, rd, rm, rgsl, rgsu, rgu, rgx6, b);
        if (act) {
#pragma unroll
            for (int c = 0; c < NC; c++) { AT(P.rd, k * NC + c) = rd[c]; AT(P.rm, k * NC + c) = rm[c]; }
#pragma unroll
            for (int j = 0; j < 2; j++) { AT(P.rgu, k * 2 + j) = rgu[j]; AT(P.rgsl, k * 2 + j) = rgsl[j]; AT(P.rgsu, k * 2 + j) = rgsu[j]; }
            AT(P.bar, k * 6 + 0) = b.Rt[0]; AT(P.bar, k * 6 + 1) = b.Rt[1]; AT(P.bar, k * 6 + 2) = b.Qt6;
            AT(P.bar, k * 6 + 3) = b.rt[0]; AT(P.bar, k * 6 + 4) = b.rt[1]; AT(P.bar, k * 6 + 5) = b.qt6;
        }
    }
    ng = max8(ng); nb = max8(nb); nd = max8(nd); nm = max8(nm); summ = sum8(summ);
}

// ---- sequential backward sweep, lanes <-> columns ----------------------------------------------------------
// lane c < 7 : column c of M = [B | A(:,2:7)]  (c = 0,1 -> u0,u1 ; c = 2..6 -> x2..x6)
// lane 7     : vector column (rb -> P rb -> h = P rb + p)
template <bool FACTOR>
__device__ __forceinline__ void pass_backward(const Params &P, int i, int s, bool act, double *Ps, double *Ms)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const double Ts = o.dt, hdt = o.dt;
    // terminal: P_N = diag(We), p_N = rgx_N
    double pv[7];           // lane 7: p_{k+1}; other lanes: unused
#pragma unroll
    for (int a = 0; a < 7; a++) pv[a] = AT(P.rgx, N * 7 + a);
    if (FACTOR) {
        for (int a = s; a < 28; a += 8) Ps[a] = 0.0;
        __syncwarp();
        if (s < 7) Ps[sidx(s, s)] = o.We[s];
        if (act) {
            for (int a = s; a < 28; a += 8) AT(P.P, N * 28 + a) = 0.0;
        }
        __syncwarp();
        if (act && s < 7) AT(P.P, N * 28 + sidx(s, s)) = o.We[s];
    }
    if (act && s < 7) AT(P.pv, N * 7 + s) = pv[s];

    for (int k = N - 1; k >= 0; k--) {
        const double *lin = P.lin + (size_t)k * LIN_ROWS * Bp;
        // own column (rows 0..6)
        double col[7];
        if (s < 7) {
#pragma unroll
            for (int r = 0; r < 6; r++) col[r] = (s < 2) ? AT(lin, LIN_B + r * 2 + s) : AT(lin, LIN_A + r * 5 + (s - 2));
            col[6] = (s == 1) ? hdt : ((s == 6) ? 1.0 : 0.0);
        } else {
            if (FACTOR) {
#pragma unroll
                for (int r = 0; r < 7; r++) col[r] = AT(P.rb, k * 7 + r);
            } else {
#pragma unroll
                for (int r = 0; r < 7; r++) col[r] = 0.0;
            }
        }
        // base gradient entry of this lane: lanes 0,1: rt ; lanes 2..5: rgx ; lane 6: qt6 ; lane 7: handled below
        double gbase = 0.0;
        if (s < 2) gbase = AT(P.bar, k * 6 + 3 + s);
        else if (s < 6) gbase = (k >= 1) ? AT(P.rgx, k * 7 + s) : 0.0;
        else if (s == 6) gbase = AT(P.bar, k * 6 + 5);
        double gx01[2] = {0.0, 0.0};
        if (s == 7 && k >= 1) { gx01[0] = AT(P.rgx, k * 7 + 0); gx01[1] = AT(P.rgx, k * 7 + 1); }

        double K0c = 0.0, K1c = 0.0;          // lane c>=2: column x_c of K ; lane 0: column x0 ; lane 1: column x1
        double gi00, gi01, gi11;
        double hv[7] = {0, 0, 0, 0, 0, 0, 0};
        if (FACTOR) {
            // w = P_{k+1} * col   (P broadcast from shared memory)
            double w[7];
            {
                double Pl[28];
#pragma unroll
                for (int a = 0; a < 28; a++) Pl[a] = Ps[a];
#pragma unroll
                for (int a = 0; a < 7; a++) {
                    double v = 0.0;
#pragma unroll
                    for (int l = 0; l < 7; l++) v = fma(Pl[sidx(a, l)], col[l], v);
                    w[a] = v;
                }
            }
            const double Pn00 = Ps[sidx(0, 0)], Pn10 = Ps[sidx(1, 0)], Pn11 = Ps[sidx(1, 1)];
            if (s == 7) {
#pragma unroll
                for (int a = 0; a < 7; a++) { if (act) AT(P.Pb, k * 7 + a) = w[a]; hv[a] = w[a] + pv[a]; }
            }
            // publish M columns, then every lane reads all of M
            if (s < 7) {
#pragma unroll
                for (int r = 0; r < 7; r++) Ms[s * 8 + r] = col[r];
            }
            __syncwarp();
            // Gm[c2] = M(:,c2)^T w  = G[c2][c] over the 7 M-indices ; rows x0,x1 of G are w[0], w[1]
            double Gm[7];
#pragma unroll
            for (int c2 = 0; c2 < 7; c2++) {
                double v = 0.0;
#pragma unroll
                for (int l = 0; l < 7; l++) v = fma(Ms[c2 * 8 + l], w[l], v);
                Gm[c2] = v;
            }
            // Hessian diagonal
            const double Rt0 = AT(P.bar, k * 6 + 0), Rt1 = AT(P.bar, k * 6 + 1), Qt6 = AT(P.bar, k * 6 + 2);
#pragma unroll
            for (int c2 = 0; c2 < 7; c2++) {
                if (c2 == s) {
                    const double d = (c2 == 0) ? Rt0 : (c2 == 1) ? Rt1 : (c2 == 6) ? Qt6 : Ts * o.W[c2];
                    Gm[c2] += d;
                }
            }
            // 2x2 pivot (redundantly on all lanes)
            const double g00 = shfl8(Gm[0], 0) + o.reg, g01 = shfl8(Gm[0], 1), g11 = shfl8(Gm[1], 1) + o.reg;
            const double idet = 1.0 / (g00 * g11 - g01 * g01);
            gi00 = g11 * idet; gi01 = -g01 * idet; gi11 = g00 * idet;
            // G[u0][x0], G[u1][x0], G[u0][x1], G[u1][x1]
            const double gu0x0 = shfl8(w[0], 0), gu1x0 = shfl8(w[0], 1), gu0x1 = shfl8(w[1], 0), gu1x1 = shfl8(w[1], 1);
            const double K0x0 = -(gi00 * gu0x0 + gi01 * gu1x0), K1x0 = -(gi01 * gu0x0 + gi11 * gu1x0);
            const double K0x1 = -(gi00 * gu0x1 + gi01 * gu1x1), K1x1 = -(gi01 * gu0x1 + gi11 * gu1x1);
            if (s >= 2 && s < 7) {
                K0c = -(gi00 * Gm[0] + gi01 * Gm[1]);
                K1c = -(gi01 * Gm[0] + gi11 * Gm[1]);
            } else if (s == 0) { K0c = K0x0; K1c = K1x0; }
            else if (s == 1) { K0c = K0x1; K1c = K1x1; }
            if (act) {
                if (s < 2) { AT(P.K, k * 14 + s) = K0c; AT(P.K, k * 14 + 7 + s) = K1c; }
                else if (s < 7) { AT(P.K, k * 14 + s) = K0c; AT(P.K, k * 14 + 7 + s) = K1c; }
                else { AT(P.Ginv, k * 3 + 0) = gi00; AT(P.Ginv, k * 3 + 1) = gi01; AT(P.Ginv, k * 3 + 2) = gi11; }
            }
            // Schur complement P_k = Gxx + Gxu K   (lower triangle), written to shared (next stage) and global
            __syncwarp();      // all lanes are done reading the old P
            // columns x_c, c = 2..6: rows a = c..6 need K(:,x_a) of lane a
#pragma unroll
            for (int a = 2; a < 7; a++) {
                const double Ka0 = shfl8(K0c, a), Ka1 = shfl8(K1c, a);
                if (s >= 2 && s <= a && s < 7) {
                    // P[x_a][x_s] = G[x_a][x_s] + K(:,x_a) . G[u][x_s]
                    const double v = Gm[a] + Ka0 * Gm[0] + Ka1 * Gm[1];
                    Ps[sidx(a, s)] = v;
                    if (act) AT(P.P, k * 28 + sidx(a, s)) = v;
                }
            }
            if (s >= 2 && s < 7) {
                // P[x_s][x0], P[x_s][x1] : G[x_s][x0] = w[0], G[x_s][u] = Gm[0..1]
                const double v0 = w[0] + Gm[0] * K0x0 + Gm[1] * K1x0;
                const double v1 = w[1] + Gm[0] * K0x1 + Gm[1] * K1x1;
                Ps[sidx(s, 0)] = v0; Ps[sidx(s, 1)] = v1;
                if (act) { AT(P.P, k * 28 + sidx(s, 0)) = v0; AT(P.P, k * 28 + sidx(s, 1)) = v1; }
            } else if (s == 0) {
                const double v00 = Pn00 + Ts * o.W[0] + gu0x0 * K0x0 + gu1x0 * K1x0;
                const double v10 = Pn10 + gu0x1 * K0x0 + gu1x1 * K1x0;
                Ps[sidx(0, 0)] = v00; Ps[sidx(1, 0)] = v10;
                if (act) { AT(P.P, k * 28 + sidx(0, 0)) = v00; AT(P.P, k * 28 + sidx(1, 0)) = v10; }
            } else if (s == 1) {
                const double v11 = Pn11 + Ts * o.W[1] + gu0x1 * K0x1 + gu1x1 * K1x1;
                Ps[sidx(1, 1)] = v11;
                if (act) AT(P.P, k * 28 + sidx(1, 1)) = v11;
            }
            __syncwarp();
        } else {
            if (s == 7) {
#pragma unroll
                for (int a = 0; a < 7; a++) hv[a] = AT(P.Pb, k * 7 + a) + pv[a];
            }
            gi00 = AT(P.Ginv, k * 3 + 0); gi01 = AT(P.Ginv, k * 3 + 1); gi11 = AT(P.Ginv, k * 3 + 2);
            if (s < 7) { K0c = AT(P.K, k * 14 + s); K1c = AT(P.K, k * 14 + 7 + s); }
        }
        // vector part: g_c = base + M(:,c)^T h
        double g = gbase;
#pragma unroll
        for (int l = 0; l < 7; l++) {
            const double hl = shfl8(hv[l], 7);
            if (s < 7) g = fma(col[l], hl, g);
            if (s == 7 && l < 2) gx01[l] += hv[l];
        }
        if (k == 0 && s >= 2) g = 0.0;
        const double gu0 = shfl8(g, 0), gu1 = shfl8(g, 1);
        const double kf0 = -(gi00 * gu0 + gi01 * gu1), kf1 = -(gi01 * gu0 + gi11 * gu1);
        if (act && s == 0) { AT(P.kf, k * 2 + 0) = kf0; AT(P.kf, k * 2 + 1) = kf1; }
        // p_k: lane c (2..6) -> entry x_c ; lanes 0,1 -> entries x0,x1 need gx01 of lane 7
        const double gx0 = shfl8(gx01[0], 7), gx1 = shfl8(gx01[1], 7);
        double pvc;
        if (s >= 2 && s < 7) pvc = g + K0c * gu0 + K1c * gu1;
        else if (s == 0) pvc = gx0 + K0c * gu0 + K1c * gu1;
        else if (s == 1) pvc = gx1 + K0c * gu0 + K1c * gu1;
        else pvc = 0.0;
        if (k == 0) pvc = 0.0;
        if (act && s < 7) AT(P.pv, k * 7 + s) = pvc;
        // hand p_k to lane 7 for the next stage
#pragma unroll
        for (int a = 0; a < 7; a++) {
            const double v = shfl8(pvc, a);
            if (s == 7) pv[a] = v;
        }
    }
}

// ---- sequential forward sweep, lanes <-> rows ----------------------------------------------------------------
__device__ __forceinline__ void pass_forward(const Params &P, int i, int s, bool act)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const double hdt = o.dt;
    double dxr = 0.0;                       // lane r<7: ddx_k[r]
    if (act && s < 7) AT(P.ddx, s) = 0.0;
    for (int k = 0; k < N; k++) {
        const double *lin = P.lin + (size_t)k * LIN_ROWS * Bp;
        // du = K dx + kf : lane c contributes K(:,c) dx[c], lane 7 contributes kf
        double p0, p1;
        if (s < 7) { p0 = AT(P.K, k * 14 + s) * dxr; p1 = AT(P.K, k * 14 + 7 + s) * dxr; }
        else { p0 = AT(P.kf, k * 2 + 0); p1 = AT(P.kf, k * 2 + 1); }
        const double du0 = sum8(p0), du1 = sum8(p1);
        if (act && s == 0) { AT(P.ddu, k * 2 + 0) = du0; AT(P.ddu, k * 2 + 1) = du1; }
        // dx+ row r
        double arow[5] = {0, 0, 0, 0, 0}, b0 = 0.0, b1 = 0.0, rbr = 0.0;
        if (s < 7) rbr = AT(P.rb, k * 7 + s);
        if (s < 6) {
#pragma unroll
            for (int c = 0; c < 5; c++) arow[c] = AT(lin, LIN_A + s * 5 + c);
            b0 = AT(lin, LIN_B + s * 2 + 0); b1 = AT(lin, LIN_B + s * 2 + 1);
        }
        double v = rbr + ((s < 2 || s == 6) ? dxr : 0.0);
        if (s == 6) v = fma(hdt, du1, v);
        v = fma(b0, du0, v);
        v = fma(b1, du1, v);
#pragma unroll
        for (int c = 0; c < 5; c++) {
            const double dxc = shfl8(dxr, 2 + c);
            v = fma(arow[c], dxc, v);
        }
        const double dxn = (s < 7) ? v : 0.0;
        if (act && s < 7) AT(P.ddx, (k + 1) * 7 + s) = dxn;
        // dpi_k[r] = p_{k+1}[r] + P_{k+1}[r][:] dx+
        double pr[7], acc = 0.0;
        if (s < 7) {
            acc = AT(P.pv, (k + 1) * 7 + s);
#pragma unroll
            for (int l = 0; l < 7; l++) {
                // packed symmetric row r: element (r,l)
                int idx = 0;
#pragma unroll
                for (int rr = 0; rr < 7; rr++) if (rr == s) idx = sidx(rr, l);
                pr[l] = AT(P.P, (k + 1) * 28 + idx);
            }
        }
#pragma unroll
        for (int l = 0; l < 7; l++) {
            const double dl = shfl8(dxn, l);
            if (s < 7) acc = fma(pr[l], dl, acc);
        }
        if (act && s < 7) AT(P.dpi, k * 7 + s) = acc;
        dxr = dxn;
    }
}

// ---- stage-parallel: slack / t / lambda steps of all stages, step length, mu_aff sums ----------------------------
__device__ __forceinline__ void pass_constraint_step(const Params &P, int i, int s, bool act, double &alpha,
                                                     double &s1, double &s2)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const double Ts = o.dt;
    alpha = 1.0; s1 = 0.0; s2 = 0.0;
    for (int k = s; k < N; k += 8) {
        double lam[NC], t[NC], rd[NC], rm[NC], dtv[NC], gq[NC], du[2];
#pragma unroll
        for (int c = 0; c < NC; c++) {
            lam[c] = AT(P.lam, k * NC + c); t[c] = AT(P.t, k * NC + c);
            rd[c] = AT(P.rd, k * NC + c); rm[c] = AT(P.rm, k * NC + c);
            gq[c] = (rm[c] - lam[c] * rd[c]) / t[c];
        }
        du[0] = AT(P.ddu, k * 2 + 0); du[1] = AT(P.ddu, k * 2 + 1);
        const double dx6 = AT(P.ddx, k * 7 + 6);
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const double Sl = lam[j] / t[j], Su = lam[3 + j] / t[3 + j];
            const double Ssl = lam[6 + j] / t[6 + j], Ssu = lam[8 + j] / t[8 + j];
            const double Dl = Ts * o.Zl[j] + Sl + Ssl, Du = Ts * o.Zu[j] + Su + Ssu;
            const double cl = AT(P.rgsl, k * 2 + j) + gq[j] + gq[6 + j];
            const double cu = AT(P.rgsu, k * 2 + j) + gq[3 + j] + gq[8 + j];
            const double dsl = -(cl + Sl * du[j]) / Dl;
            const double dsu = -(cu - Su * du[j]) / Du;
            if (act) { AT(P.dsl, k * 2 + j) = dsl; AT(P.dsu, k * 2 + j) = dsu; }
            dtv[j] = du[j] + dsl - rd[j];
            dtv[3 + j] = -du[j] + dsu - rd[3 + j];
            dtv[6 + j] = dsl - rd[6 + j];
            dtv[8 + j] = dsu - rd[8 + j];
        }
        if (k >= 1) { dtv[2] = dx6 - rd[2]; dtv[5] = -dx6 - rd[5]; }
        else { dtv[2] = 0.0; dtv[5] = 0.0; }
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const bool on = !((c == 2 || c == 5) && k == 0);
            const double dl = on ? -(rm[c] + lam[c] * dtv[c]) / t[c] : 0.0;
            if (act) { AT(P.dlam, k * NC + c) = dl; AT(P.dt, k * NC + c) = dtv[c]; }
            if (on) {
                if (dl < 0.0) alpha = fmin(alpha, -lam[c] / dl);
                if (dtv[c] < 0.0) alpha = fmin(alpha, -t[c] / dtv[c]);
                s1 += lam[c] * dtv[c] + t[c] * dl;
                s2 += dl * dtv[c];
            }
        }
    }
    alpha = min8(alpha); s1 = sum8(s1); s2 = sum8(s2);
}

// ---- stage-parallel: Mehrotra corrector rhs + barrier gradients ------------------------------------------------
__device__ __forceinline__ void pass_corrector_rhs(const Params &P, int i, int s, bool act, double sigmu)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    for (int k = s; k < N; k += 8) {
        double lam[NC], t[NC], rd[NC], rm[NC], rgu[2], rgsl[2], rgsu[2];
#pragma unroll
        for (int c = 0; c < NC; c++) {
            lam[c] = AT(P.lam, k * NC + c); t[c] = AT(P.t, k * NC + c); rd[c] = AT(P.rd, k * NC + c);
            const bool on = !((c == 2 || c == 5) && k == 0);
            rm[c] = on ? AT(P.rm, k * NC + c) + AT(P.dlam, k * NC + c) * AT(P.dt, k * NC + c) - sigmu : 0.0;
            if (act) AT(P.rm, k * NC + c) = rm[c];
        }
#pragma unroll
        for (int j = 0; j < 2; j++) { rgu[j] = AT(P.rgu, k * 2 + j); rgsl[j] = AT(P.rgsl, k * 2 + j); rgsu[j] = AT(P.rgsu, k * 2 + j); }
        const double rgx6 = (k >= 1) ? AT(P.rgx, k * 7 + 6) : 0.0;
        BarOut b;
        barrier_terms(o, k, lam, t, rd, rm, rgsl, rgsu, rgu, rgx6, b);
        if (act) { AT(P.bar, k * 6 + 3) = b.rt[0]; AT(P.bar, k * 6 + 4) = b.rt[1]; AT(P.bar, k * 6 + 5) = b.qt6; }
    }
}

__device__ __forceinline__ void pass_update(const Params &P, int i, int s, bool act, double alpha)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    if (!act) return;
    for (int k = s; k < N; k += 8) {
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

__global__ void __launch_bounds__(OCT_PER_CTA * 8) qp_octet_kernel(const Params P)
{
    __shared__ double Ps_all[OCT_PER_CTA][28];
    __shared__ __align__(16) double Ms_all[OCT_PER_CTA][56];
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const int oct = threadIdx.x >> 3, s = threadIdx.x & 7;
    const int i = blockIdx.x * OCT_PER_CTA + oct;        // < Bp by construction of the grid
    double *Ps = Ps_all[oct], *Ms = Ms_all[oct];
    const bool valid = i < P.B;
    const bool bad = valid ? (P.lin_bad[i] != 0) : false;
    // -1 = running ; 0 ok, 1 maxiter, 2 minstep, 3 nan (hpipm numbering)
    int status = (valid && !bad) ? -1 : 0;
    if (valid && bad && s == 0) { P.status[i] = 1; P.qp_status[i] = 0; P.qp_iter[i] = 0; }

    // ---- cold start (stage-parallel) -------------------------------------------------------------------------
    if (status < 0) {
        if (s < 7) AT(P.dx, s) = AT(P.x0, s) - AT(P.xb, s);
        for (int k = s; k < N; k += 8) {
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
    }
    __syncwarp();
    __threadfence_block();

    const double inv_nc = 1.0 / (double)(NC * N - 2);
    int iter = 0;
    double res0 = 0, res1 = 0, res2 = 0, res3 = 0;
    for (;;) {
        bool act = status < 0;
        double ng, nb, nd, nm, summ;
        pass_residual(P, i, s, act, ng, nb, nd, nm, summ);
        if (act) {
            res0 = ng; res1 = nb; res2 = nd; res3 = nm;
            if (!(isfinite(ng) && isfinite(nb) && isfinite(nd) && isfinite(nm))) status = 3;
            else if (ng < o.tol_stat && nb < o.tol_eq && nd < o.tol_ineq && nm < o.tol_comp) status = 0;
            else if (iter >= o.iter_max) status = 1;
        }
        act = status < 0;
        if (!__any_sync(FULL, act)) break;
        __syncwarp();
        const double mu = summ * inv_nc;
        // predictor
        pass_backward<true>(P, i, s, act, Ps, Ms);
        __syncwarp();
        pass_forward(P, i, s, act);
        __syncwarp();
        double a_aff, s1, s2;
        pass_constraint_step(P, i, s, act, a_aff, s1, s2);
        const double mu_aff = (summ + a_aff * s1 + a_aff * a_aff * s2) * inv_nc;
        double sigma = mu_aff / mu;
        sigma = sigma * sigma * sigma;
        __syncwarp();
        // corrector
        pass_corrector_rhs(P, i, s, act, sigma * mu);
        __syncwarp();
        pass_backward<false>(P, i, s, act, Ps, Ms);
        __syncwarp();
        pass_forward(P, i, s, act);
        __syncwarp();
        double alpha;
        pass_constraint_step(P, i, s, act, alpha, s1, s2);
        __syncwarp();
        if (act) {
            if (alpha < o.alpha_min) status = 2;
            else {
                if (alpha < 1.0) alpha *= 0.995;
                pass_update(P, i, s, true, alpha);
                iter++;
            }
        }
        __syncwarp();
    }
    if (valid && !bad && s == 0) {
        const int qps = (status == 0) ? 0 : ((status == 1) ? 2 : ((status == 2) ? 3 : 1));
        P.qp_status[i] = qps;
        P.qp_iter[i] = iter;
        P.status[i] = (qps == 0 || qps == 2) ? 0 : 4;
        AT(P.res_out, 0) = res0; AT(P.res_out, 1) = res1; AT(P.res_out, 2) = res2; AT(P.res_out, 3) = res3;
    }
}

void launch_qp_octet(const Params &P, cudaStream_t s)
{
    qp_octet_kernel<<<P.Bp / OCT_PER_CTA, OCT_PER_CTA * 8, 0, s>>>(P);
}
