// qp_smem.cu -- feedback phase, v3: octets (8 lanes per MPC instance) with the whole Riccati working set of the
// instance RESIDENT IN SHARED MEMORY for the duration of the interior-point solve.
//
// One CTA = one warp = 4 instances.  Per instance and stage, shared memory holds
//     M = [B | A(:,2:7)] (6x7, column-major), rb, K (2x7), Guu^-1, P rb, k_ff, barrier diagonals/gradients, r_x
// (87 doubles per stage; N=20 -> 14.2 KB per instance, 4 CTAs = 16 instances per SM).  The horizon-sequential
// sweeps (factorise + predictor, corrector, two roll-outs, one adjoint sweep for pi) never leave the SM: operands
// come from shared memory and warp shuffles.  Everything that is stage-local (residuals, barrier terms, slack /
// multiplier steps, step length, update) runs stage-parallel, lane s <-> stages s, s+8, ..., on a per-warp tile of
// global scratch ([row][4 instances], 32-byte sectors, L2-resident while the warp lives).
//
// Algorithm: HPIPM-style Mehrotra predictor-corrector IPM on the OCP-structured QP [EXT] (replaces
// FULL_CONDENSING_HPIPM, acados_solver_sim_car.c:145,688-693); identical maths to qp_ipm.cu (v1) except that the
// dynamics multipliers are recovered by the adjoint recursion  dpi_{k-1} = Qt_k dx_k + gt_k + A_k^T dpi_k  instead of
// P_{k+1} dx_{k+1} + p_{k+1} (so P never has to be stored per stage).  The RTI update (x += dx, ...) is fused into
// the epilogue.
#include "common.cuh"

#define FULL 0xffffffffu
#define ATS(arr, row) (arr)[(size_t)(row) * Bp + i]      // SoA interface arrays [row][Bp]
#define WS(off, row) ws[((off) + (row)) * 4 + sub]       // per-warp scratch tile [row][4]

// shared-memory stage record (doubles)
#define S_M 0       // 42: column c (0,1 = u0,u1 ; 2..6 = x2..x6) at c*6 + r, r < 6
#define S_RB 42     // 7
#define S_K 49      // 14: K0[0..6], K1[0..6]
#define S_GI 63     // 3
#define S_PB 66     // 7
#define S_KF 73     // 2
#define S_BAR 75    // 5: Rt0 Rt1 Qt6 rt0 rt1
#define S_GX 80     // 7: rgx0..rgx5, qt6   (overwritten by the adjoint base vector in the corrector roll-out)
#define S_STRIDE 87

__device__ __forceinline__ constexpr int sidx3(int i, int j) { return (i >= j) ? (i * (i + 1) / 2 + j) : (j * (j + 1) / 2 + i); }
__device__ __forceinline__ double shf8(double v, int src) { return __shfl_sync(FULL, v, src, 8); }
__device__ __forceinline__ double sum8s(double v)
{
    v += __shfl_xor_sync(FULL, v, 4, 8); v += __shfl_xor_sync(FULL, v, 2, 8); v += __shfl_xor_sync(FULL, v, 1, 8);
    return v;
}
__device__ __forceinline__ double nmx(double a, double b) { return (a > b || a != a) ? a : b; }   // NaN-propagating
__device__ __forceinline__ double max8s(double v)
{
    v = nmx(v, __shfl_xor_sync(FULL, v, 4, 8)); v = nmx(v, __shfl_xor_sync(FULL, v, 2, 8)); v = nmx(v, __shfl_xor_sync(FULL, v, 1, 8));
    return v;
}
__device__ __forceinline__ double min8s(double v)
{
    v = fmin(v, __shfl_xor_sync(FULL, v, 4, 8)); v = fmin(v, __shfl_xor_sync(FULL, v, 2, 8)); v = fmin(v, __shfl_xor_sync(FULL, v, 1, 8));
    return v;
}

struct WsOff {      // row offsets inside the scratch tile
    int du, dx, pi, lam, t, sl, su, rgu, rgsl, rgsu, rd, rm, rgx6, ddu, ddx, dlam, dt, dsl, dsu, rows;
    __host__ __device__ explicit WsOff(int N)
    {
        int o = 0;
        du = o; o += 2 * N; dx = o; o += 7 * (N + 1); pi = o; o += 7 * N; lam = o; o += NC * N; t = o; o += NC * N;
        sl = o; o += 2 * N; su = o; o += 2 * N; rgu = o; o += 2 * N; rgsl = o; o += 2 * N; rgsu = o; o += 2 * N;
        rd = o; o += NC * N; rm = o; o += NC * N; rgx6 = o; o += N; ddu = o; o += 2 * N; ddx = o; o += 7 * (N + 1);
        dlam = o; o += NC * N; dt = o; o += NC * N; dsl = o; o += 2 * N; dsu = o; o += 2 * N;
        rows = o;
    }
};
int qp_smem_ws_rows(int N) { return WsOff(N).rows; }

struct Bar3 { double Rt[2], Qt6, rt[2], qt6; };
__device__ __forceinline__ void barrier3(const admpc_opts &o, int k, const double lam[NC], const double t[NC],
                                         const double rd[NC], const double rm[NC], const double rgsl[2],
                                         const double rgsu[2], const double rgu[2], double rgx6, Bar3 &b)
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

struct Ctx {
    const Params *P;
    const WsOff *wo;
    double *ws;      // scratch tile of this warp
    double *sm;      // shared-memory region of this octet
    double *Ps;      // 28 doubles
    int i, sub, s, Bp, N;
};

// ---- stage-parallel: residuals of the current point + predictor barrier terms ---------------------------------
__device__ __forceinline__ void p_residual(const Ctx &c, double &ng, double &nb, double &nd, double &nm, double &summ)
{
    const Params &P = *c.P;
    const admpc_opts &o = P.o;
    const WsOff &wo = *c.wo;
    double *ws = c.ws, *sm = c.sm;
    const int N = c.N, Bp = c.Bp, i = c.i, sub = c.sub;
    const double Ts = o.dt, hdt = o.dt;
    ng = nb = nd = nm = summ = 0.0;
    for (int k = c.s; k <= N; k += 8) {
        if (k == N) {
            const double *lin = P.lin + (size_t)N * LIN_ROWS * Bp;
#pragma unroll
            for (int a = 0; a < 7; a++) {
                const double g = o.We[a] * WS(wo.dx, N * 7 + a) + ATS(lin, LIN_q + a) - WS(wo.pi, (N - 1) * 7 + a);
                sm[N * S_STRIDE + a] = g;          // terminal record holds r_x only
                ng = nmx(ng, fabs(g));
            }
            continue;
        }
        const double *lin = P.lin + (size_t)k * LIN_ROWS * Bp;
        double *st = sm + k * S_STRIDE;
        double du[2], pik[7], pim[7], dxk[7], dxn[7], lam[NC], t[NC], rd[NC], rm[NC], rgu[2], rgsl[2], rgsu[2];
#pragma unroll
        for (int j = 0; j < 2; j++) du[j] = WS(wo.du, k * 2 + j);
#pragma unroll
        for (int a = 0; a < 7; a++) {
            pik[a] = WS(wo.pi, k * 7 + a);
            pim[a] = (k >= 1) ? WS(wo.pi, (k - 1) * 7 + a) : 0.0;
            dxk[a] = WS(wo.dx, k * 7 + a);
            dxn[a] = WS(wo.dx, (k + 1) * 7 + a);
        }
#pragma unroll
        for (int cc = 0; cc < NC; cc++) { lam[cc] = WS(wo.lam, k * NC + cc); t[cc] = WS(wo.t, k * NC + cc); }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            double g = Ts * o.W[7 + j] * du[j] + ATS(lin, LIN_r + j) - lam[j] + lam[3 + j];
#pragma unroll
            for (int l = 0; l < 6; l++) g = fma(st[S_M + j * 6 + l], pik[l], g);
            if (j == 1) g = fma(hdt, pik[6], g);
            rgu[j] = g;
            const double sl = WS(wo.sl, k * 2 + j), su = WS(wo.su, k * 2 + j);
            rgsl[j] = Ts * o.zl[j] + Ts * o.Zl[j] * sl - lam[j] - lam[6 + j];
            rgsu[j] = Ts * o.zu[j] + Ts * o.Zu[j] * su - lam[3 + j] - lam[8 + j];
            ng = nmx(ng, nmx(fabs(g), nmx(fabs(rgsl[j]), fabs(rgsu[j]))));
            const double cur = ATS(P.ub, k * 2 + j);
            const double lo = o.lbu[j] - cur, hi = o.ubu[j] - cur;
            rd[j] = t[j] - (du[j] - lo + sl);
            rd[3 + j] = t[3 + j] - (hi - du[j] + su);
            rd[6 + j] = t[6 + j] - sl;
            rd[8 + j] = t[8 + j] - su;
            nd = nmx(nd, nmx(nmx(fabs(rd[j]), fabs(rd[3 + j])), nmx(fabs(rd[6 + j]), fabs(rd[8 + j]))));
        }
        if (k >= 1) {
            const double cur = ATS(P.xb, k * 7 + 6);
            rd[2] = t[2] - (dxk[6] - (o.lbx - cur));
            rd[5] = t[5] - ((o.ubx - cur) - dxk[6]);
            nd = nmx(nd, nmx(fabs(rd[2]), fabs(rd[5])));
        } else {
            rd[2] = 0.0; rd[5] = 0.0;
        }
        // dynamics residual rb = A dx + B du + b - dx+
#pragma unroll
        for (int r = 0; r < 6; r++) {
            double v = ATS(lin, LIN_b + r) - dxn[r] + ((r < 2) ? dxk[r] : 0.0);
            v = fma(st[S_M + 0 * 6 + r], du[0], v);
            v = fma(st[S_M + 1 * 6 + r], du[1], v);
#pragma unroll
            for (int cc = 0; cc < 5; cc++) v = fma(st[S_M + (2 + cc) * 6 + r], dxk[2 + cc], v);
            st[S_RB + r] = v;
            nb = nmx(nb, fabs(v));
        }
        {
            const double v = ATS(lin, LIN_b + 6) - dxn[6] + dxk[6] + hdt * du[1];
            st[S_RB + 6] = v;
            nb = nmx(nb, fabs(v));
        }
#pragma unroll
        for (int cc = 0; cc < NC; cc++) {
            const bool on = !((cc == 2 || cc == 5) && k == 0);
            rm[cc] = on ? lam[cc] * t[cc] : 0.0;
            nm = nmx(nm, fabs(rm[cc]));
            summ += rm[cc];
        }
        double rgx6 = 0.0;
        if (k >= 1) {
#pragma unroll
            for (int a = 0; a < 7; a++) {
                double g = Ts * o.W[a] * dxk[a] + ATS(lin, LIN_q + a) - pim[a];
                if (a < 2) g += pik[a];
                else {
#pragma unroll
                    for (int l = 0; l < 6; l++) g = fma(st[S_M + a * 6 + l], pik[l], g);
                    if (a == 6) g += pik[6] - lam[2] + lam[5];
                }
                if (a < 6) st[S_GX + a] = g; else rgx6 = g;
                ng = nmx(ng, fabs(g));
            }
        } else {
#pragma unroll
            for (int a = 0; a < 6; a++) st[S_GX + a] = 0.0;
        }
        Bar3 b;
        barrier3(o, k, lam, t, rd, rm, rgsl, rgsu, rgu, rgx6, b);
        st[S_BAR + 0] = b.Rt[0]; st[S_BAR + 1] = b.Rt[1]; st[S_BAR + 2] = b.Qt6;
        st[S_BAR + 3] = b.rt[0]; st[S_BAR + 4] = b.rt[1]; st[S_GX + 6] = b.qt6;
#pragma unroll
        for (int cc = 0; cc < NC; cc++) { WS(wo.rd, k * NC + cc) = rd[cc]; WS(wo.rm, k * NC + cc) = rm[cc]; }
#pragma unroll
        for (int j = 0; j < 2; j++) { WS(wo.rgu, k * 2 + j) = rgu[j]; WS(wo.rgsl, k * 2 + j) = rgsl[j]; WS(wo.rgsu, k * 2 + j) = rgsu[j]; }
        WS(wo.rgx6, k) = rgx6;
    }
    ng = max8s(ng); nb = max8s(nb); nd = max8s(nd); nm = max8s(nm); summ = sum8s(summ);
}

// ---- sequential backward sweep, lanes <-> columns; all operands in shared memory ----------------------------------
// lane c < 7 : column c of M ; lane 7 : vector column (rb -> P rb -> h = P rb + p)
template <bool FACTOR>
__device__ __forceinline__ void p_backward(const Ctx &c)
{
    const admpc_opts &o = c.P->o;
    double *sm = c.sm, *Ps = c.Ps;
    const int N = c.N, s = c.s;
    const double Ts = o.dt, hdt = o.dt;
    double pv[7];           // lane 7: p_{k+1}
#pragma unroll
    for (int a = 0; a < 7; a++) pv[a] = sm[N * S_STRIDE + a];
    if (FACTOR) {
        for (int a = s; a < 28; a += 8) Ps[a] = 0.0;
        __syncwarp();
        if (s < 7) Ps[sidx3(s, s)] = o.We[s];
        __syncwarp();
    }
    for (int k = N - 1; k >= 0; k--) {
        double *st = sm + k * S_STRIDE;
        double col[7];
        if (s < 7) {
#pragma unroll
            for (int r = 0; r < 6; r++) col[r] = st[S_M + s * 6 + r];
            col[6] = (s == 1) ? hdt : ((s == 6) ? 1.0 : 0.0);
        } else {
#pragma unroll
            for (int r = 0; r < 7; r++) col[r] = FACTOR ? st[S_RB + r] : 0.0;
        }
        // base gradient of this lane: 0,1 -> rt ; 2..5 -> r_x ; 6 -> qt6 ; lane 7 carries r_x[0..1]
        double gbase = 0.0;
        if (s < 2) gbase = st[S_BAR + 3 + s];
        else if (s < 7) gbase = st[S_GX + s];
        double gx01[2] = {0.0, 0.0};
        if (s == 7) { gx01[0] = st[S_GX + 0]; gx01[1] = st[S_GX + 1]; }

        double K0c = 0.0, K1c = 0.0;      // lane c>=2: column x_c of K ; lane 0: column x0 ; lane 1: column x1
        double gi00, gi01, gi11;
        double hv[7] = {0, 0, 0, 0, 0, 0, 0};
        if (FACTOR) {
            double w[7];
            {
                double Pl[28];
#pragma unroll
                for (int a = 0; a < 28; a++) Pl[a] = Ps[a];
#pragma unroll
                for (int a = 0; a < 7; a++) {
                    double v = 0.0;
#pragma unroll
                    for (int l = 0; l < 7; l++) v = fma(Pl[sidx3(a, l)], col[l], v);
                    w[a] = v;
                }
            }
            const double Pn00 = Ps[sidx3(0, 0)], Pn10 = Ps[sidx3(1, 0)], Pn11 = Ps[sidx3(1, 1)];
            if (s == 7) {
#pragma unroll
                for (int a = 0; a < 7; a++) { st[S_PB + a] = w[a]; hv[a] = w[a] + pv[a]; }
            }
            // Gm[c2] = M(:,c2)^T w = G[c2][c] over the 7 M-indices (row 6 of M is [0 dt 0 0 0 0 1])
            double Gm[7];
#pragma unroll
            for (int c2 = 0; c2 < 7; c2++) {
                double v = (c2 == 1) ? hdt * w[6] : ((c2 == 6) ? w[6] : 0.0);
#pragma unroll
                for (int l = 0; l < 6; l++) v = fma(st[S_M + c2 * 6 + l], w[l], v);
                Gm[c2] = v;
            }
            const double Rt0 = st[S_BAR + 0], Rt1 = st[S_BAR + 1], Qt6 = st[S_BAR + 2];
#pragma unroll
            for (int c2 = 0; c2 < 7; c2++) {
                if (c2 == s) Gm[c2] += (c2 == 0) ? Rt0 : (c2 == 1) ? Rt1 : (c2 == 6) ? Qt6 : Ts * o.W[c2];
            }
            const double g00 = shf8(Gm[0], 0) + o.reg, g01 = shf8(Gm[0], 1), g11 = shf8(Gm[1], 1) + o.reg;
            const double idet = 1.0 / (g00 * g11 - g01 * g01);
            gi00 = g11 * idet; gi01 = -g01 * idet; gi11 = g00 * idet;
            const double gu0x0 = shf8(w[0], 0), gu1x0 = shf8(w[0], 1), gu0x1 = shf8(w[1], 0), gu1x1 = shf8(w[1], 1);
            const double K0x0 = -(gi00 * gu0x0 + gi01 * gu1x0), K1x0 = -(gi01 * gu0x0 + gi11 * gu1x0);
            const double K0x1 = -(gi00 * gu0x1 + gi01 * gu1x1), K1x1 = -(gi01 * gu0x1 + gi11 * gu1x1);
            if (s >= 2 && s < 7) {
                K0c = -(gi00 * Gm[0] + gi01 * Gm[1]);
                K1c = -(gi01 * Gm[0] + gi11 * Gm[1]);
            } else if (s == 0) { K0c = K0x0; K1c = K1x0; }
            else if (s == 1) { K0c = K0x1; K1c = K1x1; }
            if (s < 7) { st[S_K + s] = K0c; st[S_K + 7 + s] = K1c; }
            else { st[S_GI + 0] = gi00; st[S_GI + 1] = gi01; st[S_GI + 2] = gi11; }
            __syncwarp();      // every lane has consumed the old P
#pragma unroll
            for (int a = 2; a < 7; a++) {
                const double Ka0 = shf8(K0c, a), Ka1 = shf8(K1c, a);
                if (s >= 2 && s <= a && s < 7) Ps[sidx3(a, s)] = Gm[a] + Ka0 * Gm[0] + Ka1 * Gm[1];
            }
            if (s >= 2 && s < 7) {
                Ps[sidx3(s, 0)] = w[0] + Gm[0] * K0x0 + Gm[1] * K1x0;
                Ps[sidx3(s, 1)] = w[1] + Gm[0] * K0x1 + Gm[1] * K1x1;
            } else if (s == 0) {
                Ps[sidx3(0, 0)] = Pn00 + Ts * o.W[0] + gu0x0 * K0x0 + gu1x0 * K1x0;
                Ps[sidx3(1, 0)] = Pn10 + gu0x1 * K0x0 + gu1x1 * K1x0;
            } else if (s == 1) {
                Ps[sidx3(1, 1)] = Pn11 + Ts * o.W[1] + gu0x1 * K0x1 + gu1x1 * K1x1;
            }
            __syncwarp();
        } else {
            if (s == 7) {
#pragma unroll
                for (int a = 0; a < 7; a++) hv[a] = st[S_PB + a] + pv[a];
            }
            gi00 = st[S_GI + 0]; gi01 = st[S_GI + 1]; gi11 = st[S_GI + 2];
            if (s < 7) { K0c = st[S_K + s]; K1c = st[S_K + 7 + s]; }
        }
        // vector part: g_c = base + M(:,c)^T h
        double g = gbase;
#pragma unroll
        for (int l = 0; l < 7; l++) {
            const double hl = shf8(hv[l], 7);
            if (s < 7) g = fma(col[l], hl, g);
            if (s == 7 && l < 2) gx01[l] += hv[l];
        }
        const double gu0 = shf8(g, 0), gu1 = shf8(g, 1);
        const double kf0 = -(gi00 * gu0 + gi01 * gu1), kf1 = -(gi01 * gu0 + gi11 * gu1);
        if (s == 7) { st[S_KF + 0] = kf0; st[S_KF + 1] = kf1; }
        const double gx0 = shf8(gx01[0], 7), gx1 = shf8(gx01[1], 7);
        double pvc;
        if (s >= 2 && s < 7) pvc = g + K0c * gu0 + K1c * gu1;
        else if (s == 0) pvc = gx0 + K0c * gu0 + K1c * gu1;
        else if (s == 1) pvc = gx1 + K0c * gu0 + K1c * gu1;
        else pvc = 0.0;
#pragma unroll
        for (int a = 0; a < 7; a++) {
            const double v = shf8(pvc, a);
            if (s == 7) pv[a] = v;
        }
    }
    __syncwarp();
}

// ---- sequential forward roll-out, lanes <-> rows -------------------------------------------------------------------
// ADJ: also leave the adjoint base vector  Qt_k dx_k + gt_k  in the r_x slot (consumed by p_adjoint)
template <bool ADJ>
__device__ __forceinline__ void p_forward(const Ctx &c, bool act)
{
    const admpc_opts &o = c.P->o;
    const WsOff &wo = *c.wo;
    double *ws = c.ws, *sm = c.sm;
    const int N = c.N, s = c.s, sub = c.sub;
    const double hdt = o.dt, Ts = o.dt;
    double dxr = 0.0;                        // lane r<7: ddx_k[r]
    if (act && s < 7) WS(wo.ddx, s) = 0.0;
    for (int k = 0; k < N; k++) {
        double *st = sm + k * S_STRIDE;
        double p0, p1;
        if (s < 7) { p0 = st[S_K + s] * dxr; p1 = st[S_K + 7 + s] * dxr; }
        else { p0 = st[S_KF + 0]; p1 = st[S_KF + 1]; }
        const double du0 = sum8s(p0), du1 = sum8s(p1);
        if (act && s == 7) { WS(wo.ddu, k * 2 + 0) = du0; WS(wo.ddu, k * 2 + 1) = du1; }
        if (ADJ && k >= 1 && s < 7) {
            const double Qd = (s == 6) ? st[S_BAR + 2] : Ts * o.W[s];
            st[S_GX + s] = fma(Qd, dxr, st[S_GX + s]);
        }
        double v = 0.0;
        if (s < 7) v = st[S_RB + s] + ((s < 2 || s == 6) ? dxr : 0.0);
        if (s == 6) v = fma(hdt, du1, v);
        if (s < 6) { v = fma(st[S_M + 0 * 6 + s], du0, v); v = fma(st[S_M + 1 * 6 + s], du1, v); }
#pragma unroll
        for (int cc = 0; cc < 5; cc++) {
            const double dxc = shf8(dxr, 2 + cc);
            if (s < 6) v = fma(st[S_M + (2 + cc) * 6 + s], dxc, v);
        }
        dxr = (s < 7) ? v : 0.0;
        if (act && s < 7) WS(wo.ddx, (k + 1) * 7 + s) = dxr;
    }
    if (ADJ && s < 7) sm[N * S_STRIDE + s] = fma(o.We[s], dxr, sm[N * S_STRIDE + s]);
    __syncwarp();
}

// ---- sequential adjoint sweep: dpi_{k-1} = base_k + A_k^T dpi_k ; pi += alpha dpi (lanes <-> columns/rows) ---------
__device__ __forceinline__ void p_adjoint(const Ctx &c, bool act, double alpha)
{
    const WsOff &wo = *c.wo;
    double *ws = c.ws, *sm = c.sm;
    const int N = c.N, s = c.s, sub = c.sub;
    double dpr = (s < 7) ? sm[N * S_STRIDE + s] : 0.0;       // lane r: dpi_{N-1}[r] = We dx_N + r_x,N
    for (int k = N - 1; k >= 0; k--) {
        if (act && s < 7) WS(wo.pi, k * 7 + s) += alpha * dpr;
        if (k == 0) break;
        const double *st = sm + k * S_STRIDE;
        // dpi_{k-1}[x_s] = base_k[s] + sum_l A_k[l][s] dpi_k[l]
        double v = (s < 7) ? st[S_GX + s] : 0.0;
        if (s < 2) v += dpr;                                 // columns 0,1 of A are e0,e1
        if (s == 6) v += dpr;                                // A[6][6] = 1
#pragma unroll
        for (int l = 0; l < 6; l++) {
            const double dl = shf8(dpr, l);
            if (s >= 2 && s < 7) v = fma(st[S_M + s * 6 + l], dl, v);
        }
        dpr = v;
    }
    __syncwarp();
}

// ---- stage-parallel: slack / t / lambda steps, step length, mu_aff sums ------------------------------------------
__device__ __forceinline__ void p_constraint_step(const Ctx &c, bool act, double &alpha, double &s1, double &s2)
{
    const admpc_opts &o = c.P->o;
    const WsOff &wo = *c.wo;
    double *ws = c.ws;
    const int N = c.N, sub = c.sub;
    const double Ts = o.dt;
    alpha = 1.0; s1 = 0.0; s2 = 0.0;
    for (int k = c.s; k < N; k += 8) {
        double lam[NC], t[NC], rd[NC], rm[NC], dtv[NC], gq[NC], du[2];
#pragma unroll
        for (int cc = 0; cc < NC; cc++) {
            lam[cc] = WS(wo.lam, k * NC + cc); t[cc] = WS(wo.t, k * NC + cc);
            rd[cc] = WS(wo.rd, k * NC + cc); rm[cc] = WS(wo.rm, k * NC + cc);
            gq[cc] = (rm[cc] - lam[cc] * rd[cc]) / t[cc];
        }
        du[0] = WS(wo.ddu, k * 2 + 0); du[1] = WS(wo.ddu, k * 2 + 1);
        const double dx6 = WS(wo.ddx, k * 7 + 6);
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const double Sl = lam[j] / t[j], Su = lam[3 + j] / t[3 + j];
            const double Ssl = lam[6 + j] / t[6 + j], Ssu = lam[8 + j] / t[8 + j];
            const double Dl = Ts * o.Zl[j] + Sl + Ssl, Du = Ts * o.Zu[j] + Su + Ssu;
            const double cl = WS(wo.rgsl, k * 2 + j) + gq[j] + gq[6 + j];
            const double cu = WS(wo.rgsu, k * 2 + j) + gq[3 + j] + gq[8 + j];
            const double dsl = -(cl + Sl * du[j]) / Dl;
            const double dsu = -(cu - Su * du[j]) / Du;
            if (act) { WS(wo.dsl, k * 2 + j) = dsl; WS(wo.dsu, k * 2 + j) = dsu; }
            dtv[j] = du[j] + dsl - rd[j];
            dtv[3 + j] = -du[j] + dsu - rd[3 + j];
            dtv[6 + j] = dsl - rd[6 + j];
            dtv[8 + j] = dsu - rd[8 + j];
        }
        if (k >= 1) { dtv[2] = dx6 - rd[2]; dtv[5] = -dx6 - rd[5]; }
        else { dtv[2] = 0.0; dtv[5] = 0.0; }
#pragma unroll
        for (int cc = 0; cc < NC; cc++) {
            const bool on = !((cc == 2 || cc == 5) && k == 0);
            const double dl = on ? -(rm[cc] + lam[cc] * dtv[cc]) / t[cc] : 0.0;
            if (act) { WS(wo.dlam, k * NC + cc) = dl; WS(wo.dt, k * NC + cc) = dtv[cc]; }
            if (on) {
                if (dl < 0.0) alpha = fmin(alpha, -lam[cc] / dl);
                if (dtv[cc] < 0.0) alpha = fmin(alpha, -t[cc] / dtv[cc]);
                s1 += lam[cc] * dtv[cc] + t[cc] * dl;
                s2 += dl * dtv[cc];
            }
        }
    }
    alpha = min8s(alpha); s1 = sum8s(s1); s2 = sum8s(s2);
}

// ---- stage-parallel: Mehrotra corrector rhs + barrier gradients -----------------------------------------------------
__device__ __forceinline__ void p_corrector_rhs(const Ctx &c, bool act, double sigmu)
{
    const admpc_opts &o = c.P->o;
    const WsOff &wo = *c.wo;
    double *ws = c.ws, *sm = c.sm;
    const int N = c.N, sub = c.sub;
    for (int k = c.s; k < N; k += 8) {
        double lam[NC], t[NC], rd[NC], rm[NC], rgu[2], rgsl[2], rgsu[2];
#pragma unroll
        for (int cc = 0; cc < NC; cc++) {
            lam[cc] = WS(wo.lam, k * NC + cc); t[cc] = WS(wo.t, k * NC + cc); rd[cc] = WS(wo.rd, k * NC + cc);
            const bool on = !((cc == 2 || cc == 5) && k == 0);
            rm[cc] = on ? WS(wo.rm, k * NC + cc) + WS(wo.dlam, k * NC + cc) * WS(wo.dt, k * NC + cc) - sigmu : 0.0;
            if (act) WS(wo.rm, k * NC + cc) = rm[cc];
        }
#pragma unroll
        for (int j = 0; j < 2; j++) { rgu[j] = WS(wo.rgu, k * 2 + j); rgsl[j] = WS(wo.rgsl, k * 2 + j); rgsu[j] = WS(wo.rgsu, k * 2 + j); }
        const double rgx6 = (k >= 1) ? WS(wo.rgx6, k) : 0.0;
        Bar3 b;
        barrier3(o, k, lam, t, rd, rm, rgsl, rgsu, rgu, rgx6, b);
        double *st = sm + k * S_STRIDE;
        st[S_BAR + 3] = b.rt[0]; st[S_BAR + 4] = b.rt[1]; st[S_GX + 6] = b.qt6;
    }
}

__device__ __forceinline__ void p_update(const Ctx &c, double alpha)
{
    const admpc_opts &o = c.P->o;
    const WsOff &wo = *c.wo;
    double *ws = c.ws;
    const int N = c.N, sub = c.sub;
    for (int k = c.s; k < N; k += 8) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
            WS(wo.du, k * 2 + j) += alpha * WS(wo.ddu, k * 2 + j);
            WS(wo.sl, k * 2 + j) += alpha * WS(wo.dsl, k * 2 + j);
            WS(wo.su, k * 2 + j) += alpha * WS(wo.dsu, k * 2 + j);
        }
#pragma unroll
        for (int a = 0; a < 7; a++) WS(wo.dx, (k + 1) * 7 + a) += alpha * WS(wo.ddx, (k + 1) * 7 + a);
#pragma unroll
        for (int cc = 0; cc < NC; cc++) {
            if ((cc == 2 || cc == 5) && k == 0) continue;
            WS(wo.lam, k * NC + cc) = fmax(WS(wo.lam, k * NC + cc) + alpha * WS(wo.dlam, k * NC + cc), o.lam_min);
            WS(wo.t, k * NC + cc) = fmax(WS(wo.t, k * NC + cc) + alpha * WS(wo.dt, k * NC + cc), o.t_min);
        }
    }
}

__global__ void __launch_bounds__(32) qp_smem_kernel(const Params P)
{
    extern __shared__ __align__(16) double smem_dyn[];
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const int lane = threadIdx.x, oct = lane >> 3, s = lane & 7;
    const int i = blockIdx.x * 4 + oct;                  // < Bp (grid = Bp / 4)
    const int per_oct = N * S_STRIDE + 8 + 28;
    const WsOff wo(N);
    Ctx c;
    c.P = &P; c.wo = &wo; c.i = i; c.sub = oct; c.s = s; c.Bp = Bp; c.N = N;
    c.sm = smem_dyn + (size_t)oct * per_oct;
    c.Ps = c.sm + N * S_STRIDE + 8;
    c.ws = P.ws + (size_t)blockIdx.x * wo.rows * 4;
    double *ws = c.ws, *sm = c.sm;
    const int sub = oct;
    const bool valid = i < P.B;
    const bool bad = valid ? (P.lin_bad[i] != 0) : false;
    int status = (valid && !bad) ? -1 : 0;     // -1 running ; 0 ok, 1 maxiter, 2 minstep, 3 nan (hpipm numbering)
    if (valid && bad && s == 0 && P.lin_bad[i] == 1) { P.status[i] = 1; P.qp_status[i] = 0; P.qp_iter[i] = 0; }   // 2: finished (full SQP)

    // ---- stage M into shared memory (lin is SoA [row][Bp]: 32-byte sector per row and warp) ----------------------
    for (int k = 0; k < N; k++) {
        const double *lin = P.lin + (size_t)k * LIN_ROWS * Bp;
        double *st = sm + k * S_STRIDE;
        for (int e = s; e < 42; e += 8) {
            // e = c*6 + r  (column-major);  c<2 -> B[r][c], else A[r][c-2]
            const int cc = e / 6, r = e - cc * 6;
            st[S_M + e] = (cc < 2) ? ATS(lin, LIN_B + r * 2 + cc) : ATS(lin, LIN_A + r * 5 + (cc - 2));
        }
    }
    // ---- cold start (stage-parallel) -----------------------------------------------------------------------------
    if (s < 7) WS(wo.dx, s) = ATS(P.x0, s) - ATS(P.xb, s);
    for (int k = s; k < N; k += 8) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (j == 2 && k == 0) {
                WS(wo.t, 2) = 1.0; WS(wo.t, 5) = 1.0; WS(wo.lam, 2) = 0.0; WS(wo.lam, 5) = 0.0;
                continue;
            }
            const double cur = (j < 2) ? ATS(P.ub, k * 2 + j) : ATS(P.xb, k * 7 + 6);
            const double lo = ((j < 2) ? o.lbu[j] : o.lbx) - cur, hi = ((j < 2) ? o.ubu[j] : o.ubx) - cur;
            double v = 0.0;
            if (v - lo < o.thr0) {
                if (hi - v < o.thr0) v = 0.5 * (lo + hi);
                else v = lo + o.thr0;
            } else if (hi - v < o.thr0) v = hi - o.thr0;
            if (j < 2) WS(wo.du, k * 2 + j) = v; else WS(wo.dx, k * 7 + 6) = v;
            const double tl = fmax(o.thr0, v - lo), tu = fmax(o.thr0, hi - v);
            WS(wo.t, k * NC + j) = tl; WS(wo.t, k * NC + 3 + j) = tu;
            WS(wo.lam, k * NC + j) = o.mu0 / tl; WS(wo.lam, k * NC + 3 + j) = o.mu0 / tu;
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            WS(wo.t, k * NC + 6 + j) = o.thr0; WS(wo.t, k * NC + 8 + j) = o.thr0;
            WS(wo.lam, k * NC + 6 + j) = o.mu0 / o.thr0; WS(wo.lam, k * NC + 8 + j) = o.mu0 / o.thr0;
            WS(wo.sl, k * 2 + j) = 0.0; WS(wo.su, k * 2 + j) = 0.0;
        }
#pragma unroll
        for (int a = 0; a < 7; a++) WS(wo.pi, k * 7 + a) = 0.0;
#pragma unroll
        for (int a = 0; a < 6; a++) WS(wo.dx, (k + 1) * 7 + a) = 0.0;
        if (k + 1 == N) WS(wo.dx, N * 7 + 6) = 0.0;
    }
    __syncwarp();

    const double inv_nc = 1.0 / (double)(NC * N - 2);
    int iter = 0;
    double res0 = 0, res1 = 0, res2 = 0, res3 = 0;
    for (;;) {
        bool act = status < 0;
        double ng, nb, nd, nm, summ;
        p_residual(c, ng, nb, nd, nm, summ);
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
        p_backward<true>(c);
        p_forward<false>(c, act);
        double a_aff, s1, s2;
        p_constraint_step(c, act, a_aff, s1, s2);
        const double mu_aff = (summ + a_aff * s1 + a_aff * a_aff * s2) * inv_nc;
        double sigma = mu_aff / mu;
        sigma = sigma * sigma * sigma;
        __syncwarp();
        // corrector
        p_corrector_rhs(c, act, sigma * mu);
        __syncwarp();
        p_backward<false>(c);
        p_forward<true>(c, act);
        double alpha;
        p_constraint_step(c, act, alpha, s1, s2);
        __syncwarp();
        bool step = false;
        if (act) {
            if (alpha < o.alpha_min) status = 2;
            else { if (alpha < 1.0) alpha *= 0.995; step = true; }
        }
        p_adjoint(c, step, alpha);
        if (step) { p_update(c, alpha); iter++; }
        __syncwarp();
    }
    // ---- epilogue: statuses + fused RTI update (full step; duals <- QP duals) -----------------------------------------
    const int qps = (status == 0) ? 0 : ((status == 1) ? 2 : ((status == 2) ? 3 : 1));
    const int nlp_status = (qps == 0 || qps == 2) ? 0 : 4;
    if (valid && !bad) {
        if (s == 0) {
            P.qp_status[i] = qps; P.qp_iter[i] = iter; P.status[i] = nlp_status;
            ATS(P.res_out, 0) = res0; ATS(P.res_out, 1) = res1; ATS(P.res_out, 2) = res2; ATS(P.res_out, 3) = res3;
        }
        if (nlp_status == 0) {
            for (int r = s; r < (N + 1) * 7; r += 8) ATS(P.xb, r) += WS(wo.dx, r);
            for (int r = s; r < N * 2; r += 8) {
                ATS(P.ub, r) += WS(wo.du, r);
                ATS(P.slb, r) = WS(wo.sl, r);
                ATS(P.sub, r) = WS(wo.su, r);
            }
            for (int r = s; r < N * 7; r += 8) ATS(P.pib, r) = WS(wo.pi, r);
            for (int r = s; r < N * NC; r += 8) { ATS(P.lamb, r) = WS(wo.lam, r); ATS(P.tb, r) = WS(wo.t, r); }
        }
    }
}

size_t qp_smem_bytes(int N) { return (size_t)4 * (N * S_STRIDE + 8 + 28) * sizeof(double); }

bool launch_qp_smem(const Params &P, cudaStream_t s)
{
    const size_t sm = qp_smem_bytes(P.o.N);
    if (sm > 227 * 1024) return false;
    static SmemGuard configured;
    if (configured.need(sm)) {
        cudaFuncSetAttribute(qp_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    }
    qp_smem_kernel<<<P.Bp / 4, 32, sm, s>>>(P);
    return true;
}
