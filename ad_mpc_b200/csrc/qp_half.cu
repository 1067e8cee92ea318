// qp_half.cu -- feedback phase, v5: HALF A WARP (16 lanes) PER MPC INSTANCE, two instances per warp, any N that fits.
//
// Everything an interior-point solve touches lives in shared memory for the whole solve: per shooting node one record
// of H_RS doubles holding the stage matrices M = [B | A(:,2:7)], the Riccati outputs (K, Guu^-1, P rb, k_ff), the
// right-hand-side vectors of the sweeps AND the node's slice of the IPM iterate (du, dx, pi, lam, t, slacks).  No
// per-lane IPM state in registers (the v4 kernel kept 72 doubles per lane and spilled), no global traffic inside the
// iteration except the cold reads of b_k / q_k / r_k.
//
// Lane roles inside a group of 16 (j = lane & 15):
//   * node role   : lane j owns nodes j, j+16, ... (residuals, barrier terms, step lengths, update): pure streaming
//                   shared memory -> registers -> shared memory, reductions by width-16 shuffles.
//   * factor role : j = (h, c), h = j >> 3 the row half, c = j & 7 the column: c < 7 column c of M, c = 7 the vector
//                   column (rb, p).  W(:,c) = P col is split by rows over the two halves (4 + 3 rows, one xor-8 exchange),
//                   G(d,c) = M(:,d)^T W(:,c) by rows d of the half.  Every operand that many lanes need (P, M, the u-rows
//                   of G, the gains) travels as a 16-byte BROADCAST load -- one shared-memory wavefront serves both
//                   instances of the warp -- instead of one shuffle per value; per-lane register blocking (4 x 7) gives
//                   each loaded value 4-7 uses.  The vector recursion p_k rides in the c = 7 lanes of the same stream.
//   * vector role : the corrector's backward sweep, the two roll-outs and the adjoint sweep: lane <-> row / column of
//                   the 7x7 stage matrix, the 7-vector of the recursion broadcast through a double-buffered slot.
//
// Algorithm: HPIPM-style Mehrotra predictor-corrector IPM on the OCP-structured QP [EXT], replacing
// FULL_CONDENSING_HPIPM (acados_solver_sim_car.c:145,688-693); identical maths to qp_warp.cu / the oracle
// (oracle/rti_oracle.c orc_qp_solve), results differ by rounding only.  The corrector's barrier gradient is formed as
// predictor value + the change caused by the complementarity right-hand side (no stage-resident residual copies needed).
#include "common.cuh"

#define FULL 0xffffffffu
#define ATS(arr, row) (arr)[(size_t)(row) * Bp + i]      // SoA interface arrays [row][Bp]

// ---- node record (doubles).  7-vectors start at even offsets (16-byte loads), the odd slots between them hold scalars.
#define H_M 0       // 42  column c (0,1 = u0,u1 ; 2..6 = x2..x6) at c*6 + r, r < 6
#define H_K0 42     // 7   first row of the gain K over the states x0..x6
#define H_GI0 49    //     Guu^-1 (0,0)
#define H_K1 50     // 7   second row
#define H_GI1 57    //     Guu^-1 (0,1)
#define H_RB 58     // 7   dynamics residual ; the corrector roll-out leaves ddx_{k+1} here
#define H_GI2 65    //     Guu^-1 (1,1)
#define H_PB 66     // 7   P_{k+1} rb_k ; the adjoint sweep leaves dpi_k here
#define H_KF0 73    //     k_ff
#define H_GX 74     // 7   rgx0..rgx5, qt6 ; the corrector roll-out leaves the adjoint base vector here
#define H_KF1 81
#define H_BAR 82    // 5   Rt0 Rt1 Qt6 rt0 rt1 ; the corrector roll-out leaves its ddu in rt0, rt1
#define H_DD 87     // 3   affine ddu0 ddu1 ddx_k[6]
#define H_DX 90     // 7   iterate: dx_k
#define H_XB6 97    //     steering angle of the linearisation point (bounds in delta form)
#define H_PI 98     // 7   iterate: pi_k
#define H_LAM 106   // 10  iterate: lam
#define H_T 116     // 10  iterate: t
#define H_DU 126    // 2
#define H_SL 128    // 2
#define H_SU 130    // 2
#define H_UB 132    // 2   input of the linearisation point
#define H_RS 134    // record stride: even, 2*H_RS mod 32 = 12 (node-parallel 8-byte accesses: 2-way conflicts at most)
// terminal record
#define T_DX 0      // 7
#define T_GX 8      // 7   r_x,N ; later We dx_N + r_x,N (adjoint start)
#define T_SIZE 16
// scratch of one instance
#define X_P 0       // 64  P_{k+1}, full symmetric, row a at a*8 (row 7 and column 7 are zero padding)
#define X_GU 64     // 16  (G[u0][c], G[u1][c]) for c = 0..6, (g_u0, g_u1) at c = 7
#define X_W01 80    // 4   G[u0][x0] G[u0][x1] | G[u1][x0] G[u1][x1]
#define X_HV 84     // 16  double-buffered broadcast of the 7-vector of the vector sweeps
#define X_SIZE 100

__device__ __forceinline__ double sel7h(const double *a, int idx)
{
    double v = 0.0;
#pragma unroll
    for (int c = 0; c < 7; c++) if (c == idx) v = a[c];
    return v;
}
// 1/x for positive normal x: hardware seed + two Newton steps (<= 1 ulp), no out-of-line slow path
__device__ __forceinline__ double rcp_h(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double nmaxh(double a, double b) { return (a > b || a != a) ? a : b; }   // NaN-propagating
__device__ __forceinline__ double gsum(double v)
{
#pragma unroll
    for (int o = 8; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o, 16);
    return v;
}
__device__ __forceinline__ double gmax(double v)
{
#pragma unroll
    for (int o = 8; o; o >>= 1) v = nmaxh(v, __shfl_xor_sync(FULL, v, o, 16));
    return v;
}
__device__ __forceinline__ double2 ldd2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void std2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }
__device__ __forceinline__ double dot6h(double2 a0, double2 a1, double2 a2, double2 b0, double2 b1, double2 b2)
{
    double v = a0.x * b0.x, w = a0.y * b0.y;
    v = fma(a1.x, b1.x, v); w = fma(a1.y, b1.y, w);
    v = fma(a2.x, b2.x, v); w = fma(a2.y, b2.y, w);
    return v + w;
}

// ---- constraint data of one node, streamed from its record ---------------------------------------------------------------
struct NodeCon {
    double lam[NC], t[NC], sl[2], su[2], du[2], dx6;
    double lo[2], hi[2], lox, hix;
};
__device__ __forceinline__ void load_con(const admpc_opts &o, const double *st, NodeCon &C)
{
#pragma unroll
    for (int c = 0; c < NC; c += 2) {
        const double2 l = ldd2(st + H_LAM + c), t = ldd2(st + H_T + c);
        C.lam[c] = l.x; C.lam[c + 1] = l.y; C.t[c] = t.x; C.t[c + 1] = t.y;
    }
    const double2 du = ldd2(st + H_DU), sl = ldd2(st + H_SL), su = ldd2(st + H_SU), ub = ldd2(st + H_UB);
    C.du[0] = du.x; C.du[1] = du.y; C.sl[0] = sl.x; C.sl[1] = sl.y; C.su[0] = su.x; C.su[1] = su.y;
    C.dx6 = st[H_DX + 6];
    const double xb6 = st[H_XB6];
    C.lo[0] = o.lbu[0] - ub.x; C.hi[0] = o.ubu[0] - ub.x;
    C.lo[1] = o.lbu[1] - ub.y; C.hi[1] = o.ubu[1] - ub.y;
    C.lox = o.lbx - xb6; C.hix = o.ubx - xb6;
}
struct NodeRes { double rd[NC], rgsl[2], rgsu[2]; };
__device__ __forceinline__ void node_res(const admpc_opts &o, bool k_ge1, const NodeCon &C, NodeRes &R)
{
    const double Ts = o.dt;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        R.rgsl[j] = Ts * o.zl[j] + Ts * o.Zl[j] * C.sl[j] - C.lam[j] - C.lam[6 + j];
        R.rgsu[j] = Ts * o.zu[j] + Ts * o.Zu[j] * C.su[j] - C.lam[3 + j] - C.lam[8 + j];
        R.rd[j] = C.t[j] - (C.du[j] - C.lo[j] + C.sl[j]);
        R.rd[3 + j] = C.t[3 + j] - (C.hi[j] - C.du[j] + C.su[j]);
        R.rd[6 + j] = C.t[6 + j] - C.sl[j];
        R.rd[8 + j] = C.t[8 + j] - C.su[j];
    }
    if (k_ge1) {
        R.rd[2] = C.t[2] - (C.dx6 - C.lox);
        R.rd[5] = C.hix - C.dx6;
        R.rd[5] = C.t[5] - R.rd[5];
    } else {
        R.rd[2] = 0.0; R.rd[5] = 0.0;
    }
}
// quantities shared by every pass over a node: 1/t, the barrier scalings and the slack-elimination pivots
struct NodeScal { double it[NC], Sl[2], Su[2], iDl[2], iDu[2]; };
__device__ __forceinline__ void node_scal(const admpc_opts &o, const NodeCon &C, NodeScal &S)
{
    const double Ts = o.dt;
#pragma unroll
    for (int c = 0; c < NC; c++) S.it[c] = rcp_h(C.t[c]);
#pragma unroll
    for (int j = 0; j < 2; j++) {
        S.Sl[j] = C.lam[j] * S.it[j]; S.Su[j] = C.lam[3 + j] * S.it[3 + j];
        const double Ssl = C.lam[6 + j] * S.it[6 + j], Ssu = C.lam[8 + j] * S.it[8 + j];
        S.iDl[j] = rcp_h(Ts * o.Zl[j] + S.Sl[j] + Ssl);
        S.iDu[j] = rcp_h(Ts * o.Zu[j] + S.Su[j] + Ssu);
    }
}
// slack / t / lambda steps of one node for a given primal step (duj, dx6) and complementarity right-hand side rm
struct NodeStep { double dsl[2], dsu[2], dtv[NC], dlv[NC]; };
__device__ __forceinline__ void node_step(bool k_ge1, const NodeCon &C, const NodeRes &R, const NodeScal &S,
                                          const double rm[NC], double du0, double du1, double dx6, NodeStep &D)
{
    double gq[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) gq[c] = (rm[c] - C.lam[c] * R.rd[c]) * S.it[c];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const double duj = (j == 0) ? du0 : du1;
        const double cl = R.rgsl[j] + gq[j] + gq[6 + j];
        const double cu = R.rgsu[j] + gq[3 + j] + gq[8 + j];
        D.dsl[j] = -(cl + S.Sl[j] * duj) * S.iDl[j];
        D.dsu[j] = -(cu - S.Su[j] * duj) * S.iDu[j];
        D.dtv[j] = duj + D.dsl[j] - R.rd[j];
        D.dtv[3 + j] = -duj + D.dsu[j] - R.rd[3 + j];
        D.dtv[6 + j] = D.dsl[j] - R.rd[6 + j];
        D.dtv[8 + j] = D.dsu[j] - R.rd[8 + j];
    }
    if (k_ge1) { D.dtv[2] = dx6 - R.rd[2]; D.dtv[5] = -dx6 - R.rd[5]; }
    else { D.dtv[2] = 0.0; D.dtv[5] = 0.0; }
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const bool on = !((c == 2 || c == 5) && !k_ge1);
        D.dlv[c] = on ? -(rm[c] + C.lam[c] * D.dtv[c]) * S.it[c] : 0.0;
    }
}
// ratio test without divisions: keep the smallest lam/(-dlam), t/(-dt) as a (num, den) pair
__device__ __forceinline__ void node_ratio(bool k_ge1, const NodeCon &C, const NodeStep &D, double &an, double &ad)
{
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const bool on = !((c == 2 || c == 5) && !k_ge1);
        if (on) {
            if (D.dlv[c] < 0.0 && C.lam[c] * ad < an * (-D.dlv[c])) { an = C.lam[c]; ad = -D.dlv[c]; }
            if (D.dtv[c] < 0.0 && C.t[c] * ad < an * (-D.dtv[c])) { an = C.t[c]; ad = -D.dtv[c]; }
        }
    }
}
__device__ __forceinline__ void group_ratio(double &an, double &ad)
{
#pragma unroll
    for (int off = 8; off; off >>= 1) {
        const double bn = __shfl_xor_sync(FULL, an, off, 16), bd = __shfl_xor_sync(FULL, ad, off, 16);
        if (bn * ad < an * bd) { an = bn; ad = bd; }
    }
    an = __shfl_sync(FULL, an, 0, 16); ad = __shfl_sync(FULL, ad, 0, 16);     // one representative pair for the group
}

// ---- factor sweep (predictor): Riccati factorisation + affine vector recursion, lanes (h, c) -------------------------------
__device__ __forceinline__ void hq_factor(const admpc_opts &o, double *rec, double *term, double *xs, int N, int j)
{
    const int h = j >> 3, c = j & 7;
    const double Ts = o.dt, hdt = o.dt;
    const int colOff = (c < 7) ? H_M + 6 * c : H_RB;
    const double m6c = (c == 1) ? hdt : (c == 6) ? 1.0 : 0.0;
    const bool vec = (c == 7);
    const double dconst = (c >= 2 && c < 6) ? Ts * sel7h(o.W, c) : 0.0;       // constant part of the diagonal of G at (c, c)
    const double pdiag = (c < 2) ? Ts * sel7h(o.W, c) : 0.0;                  // Q weight of the states x0, x1
    const int kidx0 = vec ? H_KF0 : H_K0 + c, kidx1 = vec ? H_KF1 : H_K1 + c;
    // terminal: P_N = diag(We), p_N = r_x,N
#pragma unroll
    for (int e = 0; e < 4; e++) xs[X_P + j * 4 + e] = 0.0;
    __syncwarp();
    if (j < 7) xs[X_P + j * 8 + j] = sel7h(o.We, j);
    double pv[4];
#pragma unroll
    for (int q = 0; q < 4; q++) pv[q] = (vec && 4 * h + q < 7) ? term[T_GX + 4 * h + q] : 0.0;
    __syncwarp();
    double *st = rec + (size_t)(N - 1) * H_RS;
    for (int k = N - 1; k >= 0; k--, st -= H_RS) {
        // ---- 1. W(a, c) = P(a, :) col_c for the rows a = 4h + q of this half ----------------------------------------------
        const double2 c01 = ldd2(st + colOff), c23 = ldd2(st + colOff + 2), c45 = ldd2(st + colOff + 4);
        const double col6 = vec ? st[H_RB + 6] : m6c;
        double wq[4], pold[2];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const double *pr = xs + X_P + (4 * h + q) * 8;
            const double2 p01 = ldd2(pr), p23 = ldd2(pr + 2), p45 = ldd2(pr + 4), p67 = ldd2(pr + 6);
            wq[q] = fma(p67.x, col6, dot6h(p01, p23, p45, c01, c23, c45));
            if (q < 2) pold[q] = (c == 0) ? p01.x : p01.y;       // P_{k+1}(q, c) for the (x0, x1) block (lanes h = 0, c < 2)
        }
        if (vec) {      // vector column: P rb is kept for the corrector, h = P rb + p continues as the column
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (4 * h + q < 7) st[H_PB + 4 * h + q] = wq[q];
                wq[q] += pv[q];
            }
        }
        // ---- 2. exchange the halves: w[0..6] = W(:, c) -----------------------------------------------------------------------
        double w[7];
        {
            double oq[4];
#pragma unroll
            for (int q = 0; q < 4; q++) oq[q] = __shfl_xor_sync(FULL, wq[q], 8);
#pragma unroll
            for (int a = 0; a < 4; a++) w[a] = h ? oq[a] : wq[a];
#pragma unroll
            for (int a = 0; a < 3; a++) w[4 + a] = h ? wq[a] : oq[a];
        }
        // ---- 3. G(d, c) = M(:, d)^T W(:, c) for d = 4h + q (+ diagonal) ; on the vector column: M(:, d)^T h ------------------
        const double2 bar01 = ldd2(st + H_BAR), bar23 = ldd2(st + H_BAR + 2);     // Rt0 Rt1 | Qt6 rt0
        const double rt1 = st[H_BAR + 4];
        const double dval = (c == 0) ? bar01.x : (c == 1) ? bar01.y : (c == 6) ? bar23.x : dconst;
        double Gq[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int d = (4 * h + q < 7) ? 4 * h + q : 6;
            const double *mc = st + H_M + 6 * d;
            const double2 m01 = ldd2(mc), m23 = ldd2(mc + 2), m45 = ldd2(mc + 4);
            const double m6d = (d == 1) ? hdt : (d == 6) ? 1.0 : 0.0;
            double v = m01.x * w[0], u = m01.y * w[1];
            v = fma(m23.x, w[2], v); u = fma(m23.y, w[3], u);
            v = fma(m45.x, w[4], v); u = fma(m45.y, w[5], u);
            v = fma(m6d, w[6], v) + u;
            Gq[q] = (4 * h + q == c) ? v + dval : v;
        }
        // ---- 4. the u-rows of G go to the broadcast slot ----------------------------------------------------------------------
        if (h == 0) {
            const double a0 = vec ? Gq[0] + bar23.y : Gq[0], a1 = vec ? Gq[1] + rt1 : Gq[1];     // vector column: g_u = rt + B^T h
            std2(xs + X_GU + 2 * c, a0, a1);
            if (c < 2) { xs[X_W01 + 2 * c] = w[0]; xs[X_W01 + 2 * c + 1] = w[1]; }
        }
        __syncwarp();
        // ---- 5. 2x2 pivot, gains ------------------------------------------------------------------------------------------------
        const double2 gA = ldd2(xs + X_GU), gB = ldd2(xs + X_GU + 2);               // G00 G10 | G01 G11
        double2 gc = ldd2(xs + X_GU + 2 * c);
        const double2 wa = ldd2(xs + X_W01), wb = ldd2(xs + X_W01 + 2);             // G[u0][x0] G[u0][x1] | G[u1][x0] G[u1][x1]
        if (c == 0) { gc.x = wa.x; gc.y = wb.x; }
        if (c == 1) { gc.x = wa.y; gc.y = wb.y; }
        const double g00 = gA.x + o.reg, g01 = gA.y, g11 = gB.y + o.reg;
        const double idet = rcp_h(g00 * g11 - g01 * g01);
        const double gi00 = g11 * idet, gi01 = -g01 * idet, gi11 = g00 * idet;
        const double K0c = -(gi00 * gc.x + gi01 * gc.y), K1c = -(gi01 * gc.x + gi11 * gc.y);
        if (h == 0) { st[kidx0] = K0c; st[kidx1] = K1c; }
        else if (vec) { st[H_GI0] = gi00; st[H_GI1] = gi01; st[H_GI2] = gi11; }
        __syncwarp();
        // ---- 6. Schur complement P_k(a, x_c) = base + K(:, a)^T G[u][x_c] ; vector column: p_k(a) ---------------------------------
        {
            const double2 k0a = ldd2(st + H_K0 + 4 * h), k0b = ldd2(st + H_K0 + 4 * h + 2);
            const double2 k1a = ldd2(st + H_K1 + 4 * h), k1b = ldd2(st + H_K1 + 4 * h + 2);
            const double2 gxa = ldd2(st + H_GX + 4 * h), gxb = ldd2(st + H_GX + 4 * h + 2);
            const double K0v[4] = {k0a.x, k0a.y, k0b.x, k0b.y}, K1v[4] = {k1a.x, k1a.y, k1b.x, k1b.y};
            const double gxv[4] = {gxa.x, gxa.y, gxb.x, gxb.y};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int a = 4 * h + q;
                double base;
                if (q < 2) base = h ? Gq[q] : ((c < 2) ? pold[q] + ((q == c) ? pdiag : 0.0) : w[q]);
                else base = Gq[q];
                if (vec) base += gxv[q];
                const double val = fma(K1v[q], gc.y, fma(K0v[q], gc.x, base));
                if (vec) pv[q] = (a < 7) ? val : 0.0;
                else if (a < 7) {
                    if (c >= 2) {
                        xs[X_P + a * 8 + c] = val;
                        if (a < 2) xs[X_P + c * 8 + a] = val;
                    } else if (a < 2) xs[X_P + a * 8 + c] = val;
                }
            }
        }
        __syncwarp();
    }
}

// ---- corrector backward sweep: vector part only ------------------------------------------------------------------------------
// lane v < 7: M-column v (g_v = base + M(:,v)^T h) ; lanes 7, 8: the states x0, x1.  p_k lives on lanes 7, 8, 2..6 (entries
// x0, x1, x2..x6); h = P rb + p is published through the double-buffered broadcast slot (one __syncwarp per stage).
__device__ __forceinline__ void hq_backward_vec(const admpc_opts &o, double *rec, double *term, double *xs, int N, int j)
{
    const double hdt = o.dt;
    const int v = (j < 9) ? j : 0;
    const int gb = (v < 2) ? H_BAR + 3 + v : (v < 7) ? H_GX + v : H_GX + (v - 7);
    const int mOff = H_M + ((v < 7) ? v : 0) * 6;
    const double m6 = (v == 1) ? hdt : (v == 6) ? 1.0 : 0.0;
    const bool isx01 = (v >= 7);
    const int sx = (v >= 2 && v < 7) ? v : (v == 8) ? 1 : 0;    // state index of the p entry this lane carries
    const bool owner = (j >= 2 && j < 9);
    const bool kfj = (j == 10);
    const int gio = kfj ? H_GI1 : H_GI0, gio2 = kfj ? H_GI2 : H_GI1;
    double pown = term[T_GX + sx];                              // p_N = r_x,N
    double *st = rec + (size_t)(N - 1) * H_RS;
    for (int k = N - 1; k >= 0; k--, st -= H_RS) {
        double *hb = xs + X_HV + (k & 1) * 8;
        const double hown = st[H_PB + sx] + pown;
        if (owner) hb[sx] = hown;
        __syncwarp();
        const double2 h01 = ldd2(hb), h23 = ldd2(hb + 2), h45 = ldd2(hb + 4);
        const double h6 = hb[6];
        const double2 m01 = ldd2(st + mOff), m23 = ldd2(st + mOff + 2), m45 = ldd2(st + mOff + 4);
        const double d = fma(m6, h6, dot6h(m01, m23, m45, h01, h23, h45));
        const double g = st[gb] + (isx01 ? hown : d);
        const double gu0 = __shfl_sync(FULL, g, 0, 16), gu1 = __shfl_sync(FULL, g, 1, 16);
        const double c0 = st[gio], c1 = st[gio2];               // lane 9: (gi00, gi01) ; lane 10: (gi01, gi11)
        const double kf = -(c0 * gu0 + c1 * gu1);
        if (j == 9) st[H_KF0] = kf;
        if (j == 10) st[H_KF1] = kf;
        pown = g + st[H_K0 + sx] * gu0 + st[H_K1 + sx] * gu1;
    }
    __syncwarp();
}

// ---- forward roll-out: lanes 0..5 rows of M, lane 6 the delta row, lanes 7 / 8 the gain rows ---------------------------------
// ADJ (corrector): ddu goes to the rt slots of the record, ddx_{k+1} to the rb slot, the adjoint base vector
// Qt_k ddx_k + gt_k to the r_x slot.  Predictor: (ddu, ddx_k[6]) to the DD slot only.
template <bool ADJ>
__device__ __forceinline__ void hq_forward(const admpc_opts &o, double *rec, double *term, double *xs, int N, int j)
{
    const double hdt = o.dt, Ts = o.dt;
    const int l7 = (j < 7) ? j : 6;
    const double wq_l = Ts * sel7h(o.W, l7), we_l = sel7h(o.We, l7);
    const double cself = (j < 2 || j == 6) ? 1.0 : 0.0, cdt = (j == 6) ? hdt : 0.0, mB = (j < 6) ? 1.0 : 0.0;
    const double is6 = (j == 6) ? 1.0 : 0.0;
    const int job = (j < 9) ? j : 8;                       // lanes 9..15 shadow lane 8
    const bool isK = (job >= 7);
    const int rbase = isK ? ((job == 7) ? H_K0 : H_K1) : H_M + ((job < 6) ? job : 5);
    const int rstr = isK ? 1 : 6;                          // element c of the lane's row at rbase + c * rstr
    const int offi = isK ? ((job == 7) ? H_KF0 : H_KF1) : H_RB + l7;
    const int o1 = rstr, o2 = 2 * rstr, o3 = 3 * rstr, o4 = 4 * rstr, o5 = 5 * rstr, o6 = 6 * rstr;
    double dxr = 0.0;
    double *st = rec;
    for (int k = 0; k < N; k++, st += H_RS) {
        const double *row = st + rbase;
        double *bx = xs + X_HV + (k & 1) * 8;
        if (j < 8) bx[j] = dxr;                        // lane 7 writes the zero pad
        __syncwarp();
        const double2 x01 = ldd2(bx), x23 = ldd2(bx + 2), x45 = ldd2(bx + 4);
        const double x6 = bx[6];
        double ta = row[o2] * x23.x, tb = row[o3] * x23.y;
        ta = fma(row[o4], x45.x, ta); tb = fma(row[o5], x45.y, tb);
        ta = fma(row[o6], x6, ta);
        const double e0 = row[0], e1 = row[o1], off = st[offi];
        const double duj = fma(e0, x01.x, off) + fma(e1, x01.y, tb) + ta;      // gain lanes: ddu_j = K_j . x + k_ff,j
        const double du0 = __shfl_sync(FULL, duj, 7, 16), du1 = __shfl_sync(FULL, duj, 8, 16);
        if (j == 7) {
            if (ADJ) { st[H_BAR + 3] = du0; st[H_BAR + 4] = du1; }
            else { st[H_DD + 0] = du0; st[H_DD + 1] = du1; st[H_DD + 2] = x6; }
        }
        if (ADJ && k >= 1) {
            const double Qd = fma(is6, st[H_BAR + 2] - wq_l, wq_l);
            const double nb = fma(Qd, dxr, st[H_GX + l7]);
            if (j < 7) st[H_GX + j] = nb;
        }
        const double d = fma(e0, du0, ta) + fma(e1, du1, tb);                  // state lanes: (A ddx)[j] + (B ddu)[j]
        double v = off + fma(cself, dxr, cdt * du1);
        v = fma(mB, d, v);
        dxr = (j < 7) ? v : 0.0;
        if (ADJ && j < 7) st[H_RB + j] = dxr;           // ddx_{k+1}
    }
    if (ADJ && j < 7) term[T_GX + j] = fma(we_l, dxr, term[T_GX + j]);
    __syncwarp();
}

// ---- adjoint sweep: dpi_{k-1} = base_k + A_k^T dpi_k ; leaves dpi_k in the P rb slot -----------------------------------------
__device__ __forceinline__ void hq_adjoint(double *rec, double *term, double *xs, int N, int j)
{
    const int l7 = (j < 7) ? j : 6;
    const int lc = (j >= 2 && j < 7) ? j : 2;
    const double cself = (j < 2 || j == 6) ? 1.0 : 0.0, mA = (j >= 2 && j < 7) ? 1.0 : 0.0;
    double dpr = (j < 7) ? term[T_GX + j] : 0.0;             // dpi_{N-1} = We dx_N + r_x,N
    double *st = rec + (size_t)(N - 1) * H_RS;
    for (int k = N - 1; k >= 0; k--, st -= H_RS) {
        double *hb = xs + X_HV + (k & 1) * 8;
        if (j < 7) { st[H_PB + j] = dpr; hb[j] = dpr; }
        if (k == 0) break;
        __syncwarp();
        const double2 q01 = ldd2(hb), q23 = ldd2(hb + 2), q45 = ldd2(hb + 4);
        const double *mc = st + H_M + lc * 6;
        const double d = dot6h(ldd2(mc), ldd2(mc + 2), ldd2(mc + 4), q01, q23, q45);
        const double v = st[H_GX + l7] + fma(cself, dpr, mA * d);  // A(:,0..1) = e0,e1 ; A[6][6] = 1
        dpr = (j < 7) ? v : 0.0;
    }
    __syncwarp();
}

// One CTA = one warp = two instances.
__global__ void __launch_bounds__(32) qp_half_kernel(const Params P)
{
    extern __shared__ __align__(16) double smh[];
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const int l = threadIdx.x, g = l >> 4, j = l & 15;
    const int iraw = blockIdx.x * 2 + g;
    const bool writer = iraw < P.B;                  // odd batch: the last group repeats the last instance, writes nothing
    const int i = writer ? iraw : P.B - 1;
    const int inst_doubles = N * H_RS + T_SIZE + X_SIZE;
    double *rec = smh + (size_t)g * inst_doubles;
    double *term = rec + (size_t)N * H_RS;
    double *xs = term + T_SIZE;
    const double Ts = o.dt, hdt = o.dt;
    const int NS = (N + 16) / 16;                    // node slots per lane: nodes j, j + 16, ... <= N
    const int flag = P.lin_bad[i];                   // 1: NaN/Inf in the linearisation ; 2: finished instance of the SQP loop
    bool done = (flag != 0);
    int status = 1, iter = 0;
    double res0 = 0, res1 = 0, res2 = 0, res3 = 0;

    // ---- stage the linearisation and the cold start ---------------------------------------------------------------------------
    // M: entry w = cc*6 + r of a stage (cc < 2: B(r,cc), else A(r,cc-2)); 16 lanes sweep the 42 entries in three rounds
    for (int k = 0; k < N; k++) {
        const double *lin = P.lin + (size_t)k * LIN_ROWS * Bp;
        double *st = rec + (size_t)k * H_RS;
#pragma unroll
        for (int rnd = 0; rnd < 3; rnd++) {
            const int w = j + 16 * rnd;
            if (w < 42) {
                const int cc = w / 6, r = w - cc * 6;
                const int s = (cc < 2) ? LIN_B + r * 2 + cc : LIN_A + r * 5 + (cc - 2);
                st[H_M + w] = ATS(lin, s);
            }
        }
    }
    for (int s = 0; s < NS; s++) {
        const int k = j + 16 * s;
        if (k > N) continue;
        if (k == N) {
#pragma unroll
            for (int a = 0; a < 7; a++) term[T_DX + a] = 0.0;
            continue;
        }
        double *st = rec + (size_t)k * H_RS;
        const double ub0 = ATS(P.ub, k * 2 + 0), ub1 = ATS(P.ub, k * 2 + 1), xb6 = ATS(P.xb, k * 7 + 6);
        st[H_UB] = ub0; st[H_UB + 1] = ub1; st[H_XB6] = xb6;
        double dx[7];
#pragma unroll
        for (int a = 0; a < 7; a++) dx[a] = (k == 0) ? ATS(P.x0, a) - ATS(P.xb, a) : 0.0;     // x0 eliminated (nbxe_0 = 7)
        double du[2] = {0.0, 0.0}, lam[NC], t[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) { lam[c] = 0.0; t[c] = 1.0; }
        // cold start: primal 0 pushed thr0 inside its box, t from the box, lam = mu0 / t
#pragma unroll
        for (int jj = 0; jj < 3; jj++) {
            if (jj == 2 && k == 0) continue;
            const double lo = (jj == 0) ? o.lbu[0] - ub0 : (jj == 1) ? o.lbu[1] - ub1 : o.lbx - xb6;
            const double hi = (jj == 0) ? o.ubu[0] - ub0 : (jj == 1) ? o.ubu[1] - ub1 : o.ubx - xb6;
            double v = 0.0;
            if (v - lo < o.thr0) {
                if (hi - v < o.thr0) v = 0.5 * (lo + hi);
                else v = lo + o.thr0;
            } else if (hi - v < o.thr0) v = hi - o.thr0;
            if (jj < 2) du[jj] = v; else dx[6] = v;
            const double tl = fmax(o.thr0, v - lo), tu = fmax(o.thr0, hi - v);
            t[jj] = tl; t[3 + jj] = tu;
            lam[jj] = o.mu0 / tl; lam[3 + jj] = o.mu0 / tu;
        }
#pragma unroll
        for (int jj = 0; jj < 2; jj++) {
            t[6 + jj] = o.thr0; t[8 + jj] = o.thr0;
            lam[6 + jj] = o.mu0 / o.thr0; lam[8 + jj] = o.mu0 / o.thr0;
        }
#pragma unroll
        for (int a = 0; a < 7; a++) { st[H_DX + a] = dx[a]; st[H_PI + a] = 0.0; }
#pragma unroll
        for (int c = 0; c < NC; c++) { st[H_LAM + c] = lam[c]; st[H_T + c] = t[c]; }
        st[H_DU] = du[0]; st[H_DU + 1] = du[1];
        st[H_SL] = 0.0; st[H_SL + 1] = 0.0; st[H_SU] = 0.0; st[H_SU + 1] = 0.0;
    }
    __syncwarp();

    const double inv_nc = 1.0 / (double)(NC * N - 2);
    for (int itn = 0;; itn++) {
        // ================= residuals of the current point + predictor barrier terms (node role) ==============================
        double ng = 0, nb = 0, nd = 0, nm = 0, summ = 0;
        for (int s = 0; s < NS; s++) {
            const int k = j + 16 * s;
            if (k > N) continue;
            const double *lin = P.lin + (size_t)k * LIN_ROWS * Bp;
            if (k == N) {
                const double *prev = rec + (size_t)(N - 1) * H_RS;
#pragma unroll
                for (int a = 0; a < 7; a++) {
                    const double gq = o.We[a] * term[T_DX + a] + ATS(lin, LIN_q + a) - prev[H_PI + a];
                    term[T_GX + a] = gq;
                    ng = nmaxh(ng, fabs(gq));
                }
                continue;
            }
            double *st = rec + (size_t)k * H_RS;
            double lq[7], lb[7], lr[2];
#pragma unroll
            for (int a = 0; a < 7; a++) { lq[a] = ATS(lin, LIN_q + a); lb[a] = ATS(lin, LIN_b + a); }
            lr[0] = ATS(lin, LIN_r + 0); lr[1] = ATS(lin, LIN_r + 1);
            NodeCon C;
            load_con(o, st, C);
            double pi[7], dx[7];
#pragma unroll
            for (int a = 0; a < 7; a++) { pi[a] = st[H_PI + a]; dx[a] = st[H_DX + a]; }
            NodeRes R;
            node_res(o, k >= 1, C, R);
            NodeScal S;
            node_scal(o, C, S);
            // stationarity w.r.t. u, dynamics residual, stationarity w.r.t. x: one pass over the columns of M
            double rgu[2], rgx[7], rbv[6];
            {
                const double *dxn = (k + 1 < N) ? st + H_RS + H_DX : term + T_DX;
#pragma unroll
                for (int r = 0; r < 6; r++) rbv[r] = lb[r] - dxn[r] + ((r < 2) ? dx[r] : 0.0);
                const double rb6 = lb[6] - dxn[6] + dx[6] + hdt * C.du[1];
                st[H_RB + 6] = rb6;
                nb = nmaxh(nb, fabs(rb6));
            }
#pragma unroll
            for (int cc = 0; cc < 7; cc++) {
                const double2 m01 = ldd2(st + H_M + cc * 6), m23 = ldd2(st + H_M + cc * 6 + 2), m45 = ldd2(st + H_M + cc * 6 + 4);
                const double mm[6] = {m01.x, m01.y, m23.x, m23.y, m45.x, m45.y};
                const double xv = (cc < 2) ? C.du[cc] : dx[cc];
                double gq = 0.0;
#pragma unroll
                for (int r = 0; r < 6; r++) { rbv[r] = fma(mm[r], xv, rbv[r]); gq = fma(mm[r], pi[r], gq); }
                if (cc < 2) rgu[cc] = gq; else rgx[cc] = gq;
            }
#pragma unroll
            for (int r = 0; r < 6; r++) { st[H_RB + r] = rbv[r]; nb = nmaxh(nb, fabs(rbv[r])); }
#pragma unroll
            for (int jj = 0; jj < 2; jj++) {
                double gq = Ts * o.W[7 + jj] * C.du[jj] + lr[jj] - C.lam[jj] + C.lam[3 + jj] + rgu[jj];
                if (jj == 1) gq = fma(hdt, pi[6], gq);
                rgu[jj] = gq;
                ng = nmaxh(ng, nmaxh(fabs(gq), nmaxh(fabs(R.rgsl[jj]), fabs(R.rgsu[jj]))));
                nd = nmaxh(nd, nmaxh(nmaxh(fabs(R.rd[jj]), fabs(R.rd[3 + jj])), nmaxh(fabs(R.rd[6 + jj]), fabs(R.rd[8 + jj]))));
            }
            if (k >= 1) nd = nmaxh(nd, nmaxh(fabs(R.rd[2]), fabs(R.rd[5])));
            double rm[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const bool on = !((c == 2 || c == 5) && k == 0);
                rm[c] = on ? C.lam[c] * C.t[c] : 0.0;
                nm = nmaxh(nm, fabs(rm[c]));
                summ += rm[c];
            }
            double rgx6 = 0.0;
            if (k >= 1) {
                const double *pim = st - H_RS + H_PI;
#pragma unroll
                for (int a = 0; a < 7; a++) {
                    double gq = Ts * o.W[a] * dx[a] + lq[a] - pim[a];
                    if (a < 2) gq += pi[a];
                    else {
                        gq += rgx[a];
                        if (a == 6) gq += pi[6] - C.lam[2] + C.lam[5];
                    }
                    if (a < 6) st[H_GX + a] = gq; else rgx6 = gq;
                    ng = nmaxh(ng, fabs(gq));
                }
            } else {
#pragma unroll
                for (int a = 0; a < 6; a++) st[H_GX + a] = 0.0;
            }
            // barrier-modified Hessian diagonal / gradient (soft-bound slacks eliminated)
            double gq[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) gq[c] = (rm[c] - C.lam[c] * R.rd[c]) * S.it[c];
#pragma unroll
            for (int jj = 0; jj < 2; jj++) {
                const double Rt = Ts * o.W[7 + jj] + S.Sl[jj] * (1.0 - S.Sl[jj] * S.iDl[jj]) + S.Su[jj] * (1.0 - S.Su[jj] * S.iDu[jj]);
                const double cl = R.rgsl[jj] + gq[jj] + gq[6 + jj];
                const double cu = R.rgsu[jj] + gq[3 + jj] + gq[8 + jj];
                const double rt = rgu[jj] + (gq[jj] - S.Sl[jj] * cl * S.iDl[jj]) - (gq[3 + jj] - S.Su[jj] * cu * S.iDu[jj]);
                st[H_BAR + jj] = Rt; st[H_BAR + 3 + jj] = rt;
            }
            if (k >= 1) {
                st[H_BAR + 2] = Ts * o.W[6] + C.lam[2] * S.it[2] + C.lam[5] * S.it[5];
                st[H_GX + 6] = rgx6 + gq[2] - gq[5];
            } else {
                st[H_BAR + 2] = Ts * o.W[6];
                st[H_GX + 6] = 0.0;
            }
        }
        ng = gmax(ng); nb = gmax(nb); nd = gmax(nd); nm = gmax(nm); summ = gsum(summ);
        if (!done) {
            res0 = ng; res1 = nb; res2 = nd; res3 = nm; iter = itn;
            if (!(isfinite(ng) && isfinite(nb) && isfinite(nd) && isfinite(nm))) { status = 3; done = true; }
            else if (ng < o.tol_stat && nb < o.tol_eq && nd < o.tol_ineq && nm < o.tol_comp) { status = 0; done = true; }
            else if (itn >= o.iter_max) { status = 1; done = true; }
        }
        if (__all_sync(FULL, done)) break;
        const double mu = summ * inv_nc;
        __syncwarp();

        // ================= predictor ========================================================================================
        hq_factor(o, rec, term, xs, N, j);
        hq_forward<false>(o, rec, term, xs, N, j);
        // affine step: step length, mu_aff, sigma (pass 1) ; corrected barrier gradient (pass 2)
        double an = 1.0, ad = 1.0, s1 = 0.0, s2 = 0.0;
        for (int s = 0; s < NS; s++) {
            const int k = j + 16 * s;
            if (k >= N) continue;
            const double *st = rec + (size_t)k * H_RS;
            NodeCon C; load_con(o, st, C);
            NodeRes R; node_res(o, k >= 1, C, R);
            NodeScal S; node_scal(o, C, S);
            double rm[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) rm[c] = ((c == 2 || c == 5) && k == 0) ? 0.0 : C.lam[c] * C.t[c];
            NodeStep D;
            node_step(k >= 1, C, R, S, rm, st[H_DD + 0], st[H_DD + 1], st[H_DD + 2], D);
            node_ratio(k >= 1, C, D, an, ad);
#pragma unroll
            for (int c = 0; c < NC; c++) {
                if ((c == 2 || c == 5) && k == 0) continue;
                s1 += C.lam[c] * D.dtv[c] + C.t[c] * D.dlv[c];
                s2 += D.dlv[c] * D.dtv[c];
            }
        }
        group_ratio(an, ad);
        s1 = gsum(s1); s2 = gsum(s2);
        const double a_aff = an * rcp_h(ad);
        const double mu_aff = (summ + a_aff * s1 + a_aff * a_aff * s2) * inv_nc;
        double sigma = mu_aff * rcp_h(mu);
        sigma = sigma * sigma * sigma;
        const double sigmu = sigma * mu;
        for (int s = 0; s < NS; s++) {
            const int k = j + 16 * s;
            if (k >= N) continue;
            double *st = rec + (size_t)k * H_RS;
            NodeCon C; load_con(o, st, C);
            NodeRes R; node_res(o, k >= 1, C, R);
            NodeScal S; node_scal(o, C, S);
            double rm[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) rm[c] = ((c == 2 || c == 5) && k == 0) ? 0.0 : C.lam[c] * C.t[c];
            NodeStep D;
            node_step(k >= 1, C, R, S, rm, st[H_DD + 0], st[H_DD + 1], st[H_DD + 2], D);
            // change of the barrier gradient caused by rm -> rm + dlam_aff dt_aff - sigma mu
            double e[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) e[c] = ((c == 2 || c == 5) && k == 0) ? 0.0 : (D.dlv[c] * D.dtv[c] - sigmu) * S.it[c];
#pragma unroll
            for (int jj = 0; jj < 2; jj++) {
                const double dl = e[jj] - S.Sl[jj] * (e[jj] + e[6 + jj]) * S.iDl[jj];
                const double du = e[3 + jj] - S.Su[jj] * (e[3 + jj] + e[8 + jj]) * S.iDu[jj];
                st[H_BAR + 3 + jj] += dl - du;
            }
            if (k >= 1) st[H_GX + 6] += e[2] - e[5];
        }
        __syncwarp();
        // ================= corrector ========================================================================================
        hq_backward_vec(o, rec, term, xs, N, j);
        hq_forward<true>(o, rec, term, xs, N, j);
        // final step: step length (pass 1), update of the constraint part of the iterate (pass 2)
        an = 1.0; ad = 1.0;
        double alpha = 0.0;
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            for (int s = 0; s < NS; s++) {
                const int k = j + 16 * s;
                if (k >= N) continue;
                double *st = rec + (size_t)k * H_RS;
                NodeCon C; load_con(o, st, C);
                NodeRes R; node_res(o, k >= 1, C, R);
                NodeScal S; node_scal(o, C, S);
                double rm[NC];
#pragma unroll
                for (int c = 0; c < NC; c++) rm[c] = ((c == 2 || c == 5) && k == 0) ? 0.0 : C.lam[c] * C.t[c];
                NodeStep D;
                node_step(k >= 1, C, R, S, rm, st[H_DD + 0], st[H_DD + 1], st[H_DD + 2], D);
#pragma unroll
                for (int c = 0; c < NC; c++) rm[c] = ((c == 2 || c == 5) && k == 0) ? 0.0 : rm[c] + D.dlv[c] * D.dtv[c] - sigmu;
                const double du0 = st[H_BAR + 3], du1 = st[H_BAR + 4];
                const double dx6 = (k >= 1) ? (st - H_RS)[H_RB + 6] : 0.0;
                node_step(k >= 1, C, R, S, rm, du0, du1, dx6, D);
                if (pass == 0) node_ratio(k >= 1, C, D, an, ad);
                else if (!done) {
                    st[H_DU] = C.du[0] + alpha * du0; st[H_DU + 1] = C.du[1] + alpha * du1;
#pragma unroll
                    for (int jj = 0; jj < 2; jj++) {
                        st[H_SL + jj] = C.sl[jj] + alpha * D.dsl[jj];
                        st[H_SU + jj] = C.su[jj] + alpha * D.dsu[jj];
                    }
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        if ((c == 2 || c == 5) && k == 0) continue;
                        st[H_LAM + c] = fmax(C.lam[c] + alpha * D.dlv[c], o.lam_min);
                        st[H_T + c] = fmax(C.t[c] + alpha * D.dtv[c], o.t_min);
                    }
                }
            }
            if (pass == 0) {
                group_ratio(an, ad);
                alpha = an * rcp_h(ad);
                if (!done && alpha < o.alpha_min) { status = 2; done = true; iter = itn; }
                if (alpha < 1.0) alpha *= 0.995;
            }
        }
        __syncwarp();
        // pi and dx wait for the adjoint sweep
        hq_adjoint(rec, term, xs, N, j);
        if (!done) {
            for (int s = 0; s < NS; s++) {
                const int k = j + 16 * s;
                if (k > N) continue;
                if (k < N) {
                    double *st = rec + (size_t)k * H_RS;
#pragma unroll
                    for (int a = 0; a < 7; a++) st[H_PI + a] += alpha * st[H_PB + a];
                }
                if (k >= 1) {
                    const double *prev = rec + (size_t)(k - 1) * H_RS;      // ddx_k was left in record k-1
                    double *dst = (k < N) ? rec + (size_t)k * H_RS + H_DX : term + T_DX;
#pragma unroll
                    for (int a = 0; a < 7; a++) dst[a] += alpha * prev[H_RB + a];
                }
            }
        }
        __syncwarp();
    }

    // ---- epilogue: statuses + fused RTI update (full step; duals <- QP duals) --------------------------------------------
    if (!writer) return;
    if (flag != 0) {
        if (j == 0 && flag == 1) { P.status[i] = 1; P.qp_status[i] = 0; P.qp_iter[i] = 0; }
        if (P.gat_x) {                               // fused gather: the (untouched) iterate still goes to the root's block
            for (int k = j; k <= N; k += 16) {
                for (int a = 0; a < 7; a++) P.gat_x[((size_t)i * (N + 1) + k) * 7 + a] = ATS(P.xb, k * 7 + a);
                if (k < N) for (int jj = 0; jj < 2; jj++) P.gat_u[((size_t)i * N + k) * 2 + jj] = ATS(P.ub, k * 2 + jj);
            }
            if (j == 0) P.gat_st[i] = (flag == 1) ? 1 : P.status[i];
        }
        return;
    }
    const int qps = (status == 0) ? 0 : ((status == 1) ? 2 : ((status == 2) ? 3 : 1));   // hpipm -> acados numbering
    const int nlp_status = (qps == 0 || qps == 2) ? 0 : 4;
    if (j == 0) {
        P.qp_status[i] = qps; P.qp_iter[i] = iter; P.status[i] = nlp_status;
        ATS(P.res_out, 0) = res0; ATS(P.res_out, 1) = res1; ATS(P.res_out, 2) = res2; ATS(P.res_out, 3) = res3;
    }
    const bool upd = (nlp_status == 0);
    for (int k = j; k <= N; k += 16) {
        const double *st = rec + (size_t)k * H_RS;
        const double *dxs = (k < N) ? st + H_DX : term + T_DX;
        if (upd || P.gat_x) {
#pragma unroll
            for (int a = 0; a < 7; a++) {
                double v = ATS(P.xb, k * 7 + a);
                if (upd) { v += dxs[a]; ATS(P.xb, k * 7 + a) = v; }
                if (P.gat_x) P.gat_x[((size_t)i * (N + 1) + k) * 7 + a] = v;
            }
            if (k < N) {
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    double v = ATS(P.ub, k * 2 + jj);
                    if (upd) { v += st[H_DU + jj]; ATS(P.ub, k * 2 + jj) = v; }
                    if (P.gat_x) P.gat_u[((size_t)i * N + k) * 2 + jj] = v;
                }
            }
            if (k == 0 && P.gat_x) P.gat_st[i] = nlp_status;
        }
        if (upd && k < N) {
#pragma unroll
            for (int jj = 0; jj < 2; jj++) {
                ATS(P.slb, k * 2 + jj) = st[H_SL + jj];
                ATS(P.sub, k * 2 + jj) = st[H_SU + jj];
            }
#pragma unroll
            for (int a = 0; a < 7; a++) ATS(P.pib, k * 7 + a) = st[H_PI + a];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const bool on = !((c == 2 || c == 5) && k == 0);
                ATS(P.lamb, k * NC + c) = on ? st[H_LAM + c] : 0.0;
                ATS(P.tb, k * NC + c) = on ? st[H_T + c] : 1.0;
            }
        }
    }
}

// false: the horizon does not fit the shared memory of one SM (two instances per CTA)
bool launch_qp_half(const Params &P, cudaStream_t s)
{
    const int N = P.o.N;
    const size_t sm = (size_t)2 * ((size_t)N * H_RS + T_SIZE + X_SIZE) * sizeof(double);
    if (sm > 227 * 1024) return false;
    static SmemGuard configured;
    if (configured.need(sm)) cudaFuncSetAttribute(qp_half_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    qp_half_kernel<<<(P.B + 1) / 2, 32, sm, s>>>(P);
    return true;
}
