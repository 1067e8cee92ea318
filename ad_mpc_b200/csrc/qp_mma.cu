// qp_mma.cu -- feedback phase, v7: ONE (N <= 31), TWO (N <= 63) OR FOUR (N <= 127) WARPS PER MPC INSTANCE, the whole solve resident in shared
// memory, the horizon-sequential Riccati sweeps on the FP64 TENSOR CORES (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4).
//
// Why: the v6 kernel (same residency, hand-distributed sweeps; profiles/r02_qp_rw_summary.md) spent 70 % of its time in five
// sweeps per IPM iteration whose 7x7 products were spread over the lanes by hand: every operand a lane does not own is a
// shared-memory broadcast or a shuffle (103 shared-memory wavefronts per factor stage, three round trips per stage on the
// critical path; the LSU pipe, not the FP64 pipe, was the busiest unit).  An m8n8k4 DMMA moves the operands inside the tensor
// core instead.  With nx = 7 (+1 homogeneous coordinate) everything is 8 x 8:
//
//   fragment convention cl(X): lane (g, t) = (lane >> 2, lane & 3) holds X[g][2t], X[g][2t+1]  (the accumulator layout).
//   cl(A B) = mm8(cl(A), cl(B^T)): the k index is split as {0,2,4,6} / {1,3,5,7} over two DMMAs, so that the SAME two registers
//   serve as A-fragment of X and as B-fragment of X^T.  A symmetric matrix is therefore its own B-fragment, and the whole chain
//       W^T = Mh^T Ph          (Ph = [[P, p],[p^T, .]] cost-to-go with its gradient, Mh = [[M7, rb],[0, 1]])
//       G   = Mh^T W           (Gram matrix, M7^T h in its last row / column)
//       Ph' = H - Gu^T Guu^-1 Gu   (rank-2 DMMA; H = G with the rows / columns of the inputs replaced by those of x0, x1)
//   runs accumulator -> operand with NO shared-memory traffic and NO __syncwarp: 5 DMMAs + 20 SHFL per stage.
//   The vector sweeps (two roll-outs, corrector backward sweep, adjoint sweep) are row-vector x matrix chains
//       x^T <- x^T Acl^T,   p^T <- h^T Acl + c^T,   dpi^T <- dpi^T A + base^T
//   whose result row (lanes 0..3) is already the A-fragment of the next stage; the closed-loop matrix Acl = A + B K is rebuilt
//   per stage by one rank-2 DMMA one stage ahead of the chain.  Critical path per stage: two chained DMMAs (52 cycles) instead
//   of 110..150 (store, __syncwarp, broadcast load, dot product, shuffle).
//   No lane-conditional blocks inside the sweeps: selects on values every lane computes, clamped prefetch pointers, 8-double
//   slots per vector (a divergent `if` costs a reconvergence stall per stage; the first correct build had them and was no faster
//   than v6).
//
// Measured on B200 (scripts/dmma_probe.cu): DMMA.8x8x4 26 cycles dependent, 0.25 / clk / SM (= the DFMA pipe: 37 TFLOP/s), so a
// padded 8x8x8 product costs what the hand-distributed one cost in pipe time, minus all of its operand traffic.
// cfg3 (B = 16384, N = 20): 1.55 ms against 2.13 (v6) and 3.10 (round 1); cfg4 (N = 40, two warps): 1.18 against 2.57.
//
// Node role: thread k owns node k (residuals, barrier terms, step lengths, update: shared memory -> registers -> shared memory
// with 16-byte accesses, one pass per phase); staging: one TMA bulk copy per instance-major stage record.  Algorithm: HPIPM-style
// Mehrotra predictor-corrector IPM on the OCP-structured QP [EXT], replacing FULL_CONDENSING_HPIPM
// (acados_solver_sim_car.c:145,688-693); identical maths to oracle/rti_oracle.c orc_qp_solve, results differ by rounding only.
#include "common.cuh"
#include "tma.cuh"

// ---- node record (doubles).  The first LIM_STRIDE doubles are the instance-major linearisation record written by the
// preparation kernel (common.cuh LIM_*), pulled in by ONE TMA bulk copy per stage.  Every 7-vector owns an 8-double slot at an
// even offset, so that a fragment row (lanes 0..3, 16 bytes each) is stored and loaded without special cases.
#define W_M 0       // 42  column c (0,1 = u0,u1 ; 2..6 = x2..x6) at c*6 + r, r < 6
#define W_LB 42     // 7   b_k
#define W_LQ 49     // 7   q_k
#define W_LR 56     // 2   r_k
#define W_XB 58     // 7   linearisation point x_k
#define W_UB 65     // 2   linearisation point u_k
#define W_K0 68     // 8   first row of (K | k_ff) over the states x0..x6
#define W_KF0 75
#define W_K1 76     // 8   second row
#define W_KF1 83
#define W_RB 84     // 8   dynamics residual, 1.0 (the homogeneous coordinate: never overwritten)
#define W_PB 92     // 8   P_{k+1} rb_k ; corrector backward sweep: h_k = P rb + p_{k+1} ; adjoint sweep: dpi_k
#define W_GX 100    // 8   rgx0..rgx5, qt6, 0.0 ; after the corrector roll-out the adjoint base vector
#define W_BAR 108   // 5   Rt0 Rt1 | rt0 rt1 | Qt6
#define W_GI0 113   // 3   Guu^-1 (0,0), (0,1), (1,1)
#define W_GI1 114
#define W_GI2 115
#define W_DX 116    // 7   iterate: dx_k
#define W_PI 123    // 7   iterate: pi_k
#define W_LAM 130   // 10  iterate: lam
#define W_T 140     // 10  iterate: t
#define W_DU 150    // 2
#define W_SL 152    // 2
#define W_SU 154    // 2
#define W_XA 158    // 8   roll-out: ddx_{k+1}, 1
#define W_RS 166    // record stride (2*W_RS mod 32 = 12: 16-byte node-parallel accesses are conflict-free)
// terminal record
#define T_DX 0      // 7
#define T_GX 8      // 8   r_x,N ; later We dx_N + r_x,N (adjoint start) ; 0.0
#define T_LQ 16     // 7   q_N
#define T_XB 24     // 7   x_N of the linearisation point
#define T_SIZE 32

#define NMX_INT
#include "qp_node.cuh"

// D(8x8) = A(8x4) B(4x8) + C on the FP64 tensor core: a = A[g][t], b = B[t][g], (c0, c1) = C[g][2t..2t+1]
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b, double c0, double c1)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
        : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
// cl(A B + C) from cl(A) = (ax, ay), cl(B^T) = (bx, by): the k index split as evens / odds over two chained DMMAs.  (Two
// independent DMMAs + a DADD are no faster: the pipe takes one DMMA per 16 cycles, so the second starts 16 cycles late anyway.)
__device__ __forceinline__ void mm8(double &d0, double &d1, double ax, double ay, double bx, double by, double c0, double c1)
{
    double e0, e1;
    dmma(e0, e1, ax, bx, c0, c1);
    dmma(d0, d1, ay, by, e0, e1);
}
#ifndef MMA_UF
#define MMA_UF 1        // unroll factor of the factor sweep (A/B on B200: 1 -> 1.555 ms, 2 -> 1.565 ms)
#endif
#ifndef MMA_UV
#define MMA_UV 2        // unroll factor of the vector sweeps (1 -> 1.599 ms, 2 -> 1.555 ms, 4 -> 1.575 ms)
#endif
#define MMA_PRAGMA_(x) _Pragma(#x)
#define MMA_UNROLL(n) MMA_PRAGMA_(unroll n)
__device__ __forceinline__ double shf(double v, int src) { return __shfl_sync(FULL, v, src); }

// ---- operand fragments of one stage, straight from the column-major 6x7 block M = [B | A(:,2:7)] of the record ---------------
// cl(Mh^T): lane (g, t) holds Mh[2t][g], Mh[2t+1][g] ; Mh = [[M7, rb], [0, 1]], M7 = M with the delta row (0, dt, 0, .., 0, 1)
struct MtFrag {
    int off; double c6; bool cst;
    __device__ __forceinline__ MtFrag(int g, int t, double hdt)
    {
        off = (g < 7) ? W_M + 6 * g + 2 * t : W_RB + 2 * t;        // g = 7: (rb[2t], rb[2t+1]), the slot ends with 1.0
        c6 = (g == 1) ? hdt : (g == 6) ? 1.0 : 0.0;
        cst = (t == 3 && g < 7);                                    // rows 6, 7 of a column of M7: (c6, 0)
    }
    __device__ __forceinline__ void load(const double *st, double &x, double &y) const
    {
        const double2 v = ldv(st + off);
        x = cst ? c6 : v.x;
        y = cst ? 0.0 : v.y;
    }
};
// cl(A0^T) restricted to the states: lane (g, t) holds A[2t][g], A[2t+1][g] with A = [e0 e1 M7(:,2:7)] ; column 7 zero
struct AtFrag {
    int off; double cx, cy; bool ld;
    __device__ __forceinline__ AtFrag(int g, int t)
    {
        ld = (g >= 2 && g < 7 && t < 3);
        off = ld ? W_M + 6 * g + 2 * t : W_M;
        cx = (g < 2) ? ((2 * t == g) ? 1.0 : 0.0) : ((g == 6 && t == 3) ? 1.0 : 0.0);
        cy = (g < 2 && 2 * t + 1 == g) ? 1.0 : 0.0;
    }
    __device__ __forceinline__ void load(const double *st, double &x, double &y) const
    {
        const double2 v = ldv(st + off);
        x = ld ? v.x : cx; y = ld ? v.y : cy;
    }
};
// rank-2 operands of the closed-loop matrix: bm = Bh[g][t] (B with the delta row, zero row 7), kh = Kh[t][g] = (K | k_ff), t < 2
struct ClFrag {
    int boff, koff; double bc; bool bl, kl;
    __device__ __forceinline__ ClFrag(int g, int t, double hdt)
    {
        bl = (t < 2 && g < 6);
        boff = bl ? W_M + 6 * t + g : W_M;
        bc = (g == 6 && t == 1) ? hdt : 0.0;
        kl = (t < 2);
        koff = W_K0 + 8 * (t & 1) + g;
    }
    __device__ __forceinline__ void load(const double *st, double &bm, double &kh) const
    {
        const double b = st[boff], k = st[koff];
        bm = bl ? b : bc; kh = kl ? k : 0.0;
    }
};

// ---- factor sweep (predictor): Riccati factorisation + affine vector recursion, all in fragments -----------------------------
// Dependent chain per stage: Ph -> W (2 DMMA) -> G (2 DMMA) -> shuffle of Guu -> determinant -> reciprocal -> Ph.  Everything
// else (operand loads of the next stage, the rows / columns of x0, x1 taken from W, the gains, the rank-2 DMMA) hangs off it.
__device__ __forceinline__ void mma_factor(const admpc_opts &o, double *rec, const double *term, int N, int l)
{
    asm volatile("" : "+r"(l));          // per-lane role constants are rebuilt here, not carried through the node role
    const int g = l >> 2, t = l & 3;
    const double Ts = o.dt, hdt = o.dt;
    const MtFrag mf(g, t, hdt);
    const bool g7 = (g == 7), g0 = (g == 0), t0 = (t == 0), t3 = (t == 3), glow = (g < 2);
    // H[g][2t..2t+1] = base + add.  base: own P (states x0, x1 among themselves), W handed over by the lanes (2t, 0), (2t+1, 0)
    // (rows x0, x1), own W (columns x0, x1), own G (the rest).  add: diagonal weights, gradient on the vector row / column.
    const bool catA = glow && t0, catB = glow && !t0, catD = !glow && !t0;
    const double dmG = catD ? 1.0 : 0.0;
    const double cdx = (g <= 5 && 2 * t == g) ? Ts * sel7w(o.W, g) : 0.0;     // Ts W[a] at (a, a), a <= 5
    const double cdy = (g <= 5 && 2 * t + 1 == g) ? Ts * sel7w(o.W, g) : 0.0;
    const bool q66 = (g == 6 && t3), vcol = t3 && g >= 2 && g < 7;
    const int gxo = g7 ? W_GX + 2 * t : W_GX + (g & 6);            // aligned pair holding what this lane adds
    const bool gsel = !g7 && (g & 1);
    const double mgx = g7 ? 1.0 : 0.0, mgy = (g7 || vcol) ? 1.0 : 0.0, mq = q66 ? 1.0 : 0.0;
    // G[u][.] on the lanes t = 0: the input rows of G (states x2.., vector column), W[a][u] for the states x0, x1
    const double dmU = glow ? 0.0 : 1.0, mrt = g7 ? 1.0 : 0.0;
    const int src0 = l & ~3;                                       // lane (g, 0)
    const bool tlow = (t < 2);
    const int kst = W_K0 + 8 * (t & 1) + g;
    const double reg = o.reg;
    // terminal: P_N = diag(We), p_N = r_x,N
    double px, py;
    {
        const double2 tg = ldv(term + T_GX + 2 * t);
        const double we = (g < 7) ? sel7w(o.We, g) : 0.0;
        px = g7 ? tg.x : ((2 * t == g) ? we : 0.0);
        py = g7 ? tg.y : ((2 * t + 1 == g) ? we : (t3 ? term[T_GX + ((g < 7) ? g : 0)] : 0.0));
    }
    const double *st = rec + (size_t)(N - 1) * W_RS;
    double mx, my;
    mf.load(st, mx, my);
    double2 bar01 = ldv(st + W_BAR), bar23 = ldv(st + W_BAR + 2), gxp = ldv(st + gxo);
    double qt6 = st[W_BAR + 4];
MMA_UNROLL(MMA_UF)
    for (int k = N - 1; k >= 0; k--, st -= W_RS) {
        ADMPC_ASSERT(st == rec + (size_t)k * W_RS);
        double *sw = const_cast<double *>(st);
        // ---- off the chain: what this stage adds to H, the pivot's diagonal ----------------------------------------------------
        const double gxa = gsel ? gxp.y : gxp.x;
        const double addx = fma(mq, qt6, fma(mgx, gxa, cdx));
        const double addy = fma(mgy, g7 ? gxp.y : gxa, cdy);
        const double r00 = bar01.x + reg, r11 = bar01.y + reg;
        const double ur0 = mrt * bar23.x, ur1 = mrt * bar23.y;
        // ---- 1. W^T = Mh^T Ph ; its last row is h = P rb + p -----------------------------------------------------------------
        double wx, wy;
        mm8(wx, wy, mx, my, px, py, 0.0, 0.0);
        // ---- 2. G = Mh^T W ----------------------------------------------------------------------------------------------------------
        double Gx, Gy;
        mm8(Gx, Gy, mx, my, wx, wy, 0.0, 0.0);
        // (meanwhile, from W alone)  P rb for the corrector ; rows / columns of the states x0, x1: G[u][x_a] = W[a][u], H[x_a][c] = W[a][c]
        const double pbx = wx - px, pby = wy - py;
        if (g7) stv(sw + W_PB + 2 * t, pbx, pby);
        const double rcv = __shfl_xor_sync(FULL, g0 ? wy : wx, 4);
        const double ua = g0 ? wx : rcv, ub = g0 ? rcv : wy;
        const double upx = glow ? ua : ur0, upy = glow ? ub : ur1;
        const double sx = wx + addx, sy = wy + addy, pax = px + addx, pay = py + addy;
        const double r0x = shf(sx, 8 * t), r1x = shf(sx, 8 * t + 4), r0y = shf(sy, 8 * t), r1y = shf(sy, 8 * t + 4);
        // (flat selects on values computed for every lane: no divergent code inside the sweep)
        const double rbx = g0 ? r0x : r0y, rby = g0 ? r1x : r1y;
        double hpx = catB ? rbx : sx, hpy = catB ? rby : sy;
        hpx = catA ? pax : hpx; hpy = catA ? pay : hpy;
        hpx = catD ? addx : hpx; hpy = catD ? addy : hpy;
        // next stage's operands
        const double *sn = (k > 0) ? st - W_RS : st;
        mf.load(sn, mx, my);
        bar01 = ldv(sn + W_BAR); bar23 = ldv(sn + W_BAR + 2); gxp = ldv(sn + gxo);
        qt6 = sn[W_BAR + 4];
        // ---- 3. 2x2 pivot -------------------------------------------------------------------------------------------------------------
        const double g00 = shf(Gx, 0) + r00, g01 = shf(Gy, 0), g11 = shf(Gy, 4) + r11;
        const double gu0 = shf(fma(dmU, Gx, upx), src0), gu1 = shf(fma(dmU, Gy, upy), src0);
        const double kta = fma(g11, gu0, -g01 * gu1), ktb = fma(g00, gu1, -g01 * gu0);      // rows of adj(Guu) Gu
        const double kt = t0 ? kta : ktb, gut = t0 ? gu0 : gu1;
        double Dx, Dy;
        dmma(Dx, Dy, tlow ? gut : 0.0, tlow ? kt : 0.0, 0.0, 0.0);                          // Gu^T adj(Guu) Gu
        const double idet = rcp_w(fma(g00, g11, -g01 * g01));
        // ---- 4. Schur complement --------------------------------------------------------------------------------------------------------
        px = fma(-idet, Dx, fma(dmG, Gx, hpx));
        py = fma(-idet, Dy, fma(dmG, Gy, hpy));
        // gains (K | k_ff) = -Guu^-1 Gu and Guu^-1 for the corrector
        const double kout = -idet * kt, gi0 = g11 * idet, gi1 = -g01 * idet, gi2 = g00 * idet;
        if (tlow) sw[kst] = kout;
        if (l == 2) { sw[W_GI0] = gi0; sw[W_GI1] = gi1; sw[W_GI2] = gi2; }
    }
    __syncwarp();
}

// ---- forward roll-out: [ddx_{k+1}; 1]^T = [ddx_k; 1]^T Acl^T, Acl = [[A + B K, B k_ff + rb], [0, 1]] ; ddx_{k+1}, 1 go to the
// XA slot of record k --------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_forward(const admpc_opts &o, double *rec, int N, int l)
{
    asm volatile("" : "+r"(l));
    const int g = l >> 2, t = l & 3;
    const double hdt = o.dt;
    const ClFrag cf(g, t, hdt);
    // cl(A0): lane (g, t) holds A0[g][2t], A0[g][2t+1] ; A0 = [[e0 e1 M7(:,2:7), rb], [0, 1]]
    const bool t0 = (t == 0), t3 = (t == 3), lo = (g < 6);
    const int a0o = (lo && !t0) ? (t3 ? W_M + 36 + g : W_M + 12 * t + g) : W_M;
    const int a1o = t3 ? W_RB + g : ((lo && !t0) ? W_M + 12 * t + 6 + g : W_M);      // (7, 3): the 1.0 of the rb slot
    const double c0x = t0 ? ((g == 0) ? 1.0 : 0.0) : ((g == 6 && t3) ? 1.0 : 0.0);
    const double c0y = (t0 && g == 1) ? 1.0 : 0.0;
    const bool l0x = lo && !t0, l0y = (lo && !t0) || t3;
    const bool row0 = (g == 0);
    double xx = 0.0, xy = (l == 3) ? 1.0 : 0.0;                  // ddx_0 = 0 (x0 is eliminated), homogeneous 1 ; rows g > 0 stay 0
    double *st = rec;
    double bm, kh, a0x, a0y, acx, acy;
    cf.load(st, bm, kh);
    { const double u = st[a0o], v = st[a1o]; a0x = l0x ? u : c0x; a0y = l0y ? v : c0y; }
    dmma(acx, acy, bm, kh, a0x, a0y);                              // cl(Acl) = Bh Kh + A0 of stage 0
    {
        const double *sn = st + W_RS;
        cf.load(sn, bm, kh);
        const double u = sn[a0o], v = sn[a1o];
        a0x = l0x ? u : c0x; a0y = l0y ? v : c0y;
    }
MMA_UNROLL(MMA_UV)
    for (int k = 0; k < N; k++, st += W_RS) {
        mm8(xx, xy, xx, xy, acx, acy, 0.0, 0.0);                   // the chain: two dependent DMMAs per stage
        dmma(acx, acy, bm, kh, a0x, a0y);                          // closed-loop matrix of stage k + 1, off the chain
        {
            const double *sn = (k + 2 < N) ? st + 2 * W_RS : st;
            ADMPC_ASSERT(sn >= rec && sn + W_RS <= rec + (size_t)N * W_RS);
            cf.load(sn, bm, kh);
            const double u = sn[a0o], v = sn[a1o];
            a0x = l0x ? u : c0x; a0y = l0y ? v : c0y;
        }
        if (l < 4) stv(st + W_XA + 2 * t, xx, xy);
    }
    __syncwarp();
}

// ---- corrector backward sweep: p_k^T = h_k^T (A + B K) + (gx + K^T rt)^T, h_k = P rb + p_{k+1} ; leaves h_k in the PB slot -----
__device__ __forceinline__ void mma_backward(const admpc_opts &o, double *rec, const double *term, int N, int l)
{
    asm volatile("" : "+r"(l));
    const int g = l >> 2, t = l & 3;
    const double hdt = o.dt;
    const ClFrag cf(g, t, hdt);
    const AtFrag af(g, t);
    const bool row0 = (g == 0);
    const double m0 = row0 ? 1.0 : 0.0;
    double px, py;
    { const double2 tg = ldv(term + T_GX + 2 * t); px = m0 * tg.x; py = m0 * tg.y; }                   // p_N = r_x,N
    double *st = rec + (size_t)(N - 1) * W_RS;
    double bm, kh, atx, aty, ctx, cty;
    cf.load(st, bm, kh); af.load(st, atx, aty);
    dmma(ctx, cty, kh, bm, atx, aty);                              // cl(Acl^T) = Kh^T Bh^T + A0^T of stage N-1 (row / column 7 never used)
    cf.load(st - W_RS, bm, kh); af.load(st - W_RS, atx, aty);
    double2 pb = ldv(st + W_PB + 2 * t), gx = ldv(st + W_GX + 2 * t), k0 = ldv(st + W_K0 + 2 * t), k1 = ldv(st + W_K1 + 2 * t);
    double2 rt = ldv(st + W_BAR + 2);
MMA_UNROLL(MMA_UV)
    for (int k = N - 1; k >= 0; k--, st -= W_RS) {
        const double cx = m0 * fma(k1.x, rt.y, fma(k0.x, rt.x, gx.x)), cy = m0 * fma(k1.y, rt.y, fma(k0.y, rt.x, gx.y));
        const double hx = fma(m0, pb.x, px), hy = fma(m0, pb.y, py);           // rows g > 0 stay 0
        mm8(px, py, hx, hy, ctx, cty, cx, cy);
        dmma(ctx, cty, kh, bm, atx, aty);                          // stage k - 1, off the chain
        if (l < 4) stv(st + W_PB + 2 * t, hx, hy);
        {
            const double *sn = (k > 0) ? st - W_RS : st, *s2 = (k > 1) ? st - 2 * W_RS : st;
            ADMPC_ASSERT(s2 >= rec && sn >= rec && st + W_RS <= rec + (size_t)N * W_RS);
            cf.load(s2, bm, kh); af.load(s2, atx, aty);
            pb = ldv(sn + W_PB + 2 * t); gx = ldv(sn + W_GX + 2 * t); k0 = ldv(sn + W_K0 + 2 * t); k1 = ldv(sn + W_K1 + 2 * t);
            rt = ldv(sn + W_BAR + 2);
        }
    }
    __syncwarp();
}

// ---- adjoint sweep: dpi_{k-1}^T = dpi_k^T A_k + base_k^T ; leaves dpi_k in the PB slot -----------------------------------------
__device__ __forceinline__ void mma_adjoint(double *rec, const double *term, int N, int l)
{
    asm volatile("" : "+r"(l));
    const int g = l >> 2, t = l & 3;
    const AtFrag af(g, t);
    const bool row0 = (g == 0);
    const double m0 = row0 ? 1.0 : 0.0;
    double px, py;
    { const double2 tg = ldv(term + T_GX + 2 * t); px = m0 * tg.x; py = m0 * tg.y; }                   // dpi_{N-1} = We dx_N + r_x,N
    double *st = rec + (size_t)(N - 1) * W_RS;
    double atx, aty;
    af.load(st, atx, aty);
    double2 gx = ldv(st + W_GX + 2 * t);
MMA_UNROLL(MMA_UV)
    for (int k = N - 1; k >= 0; k--, st -= W_RS) {
        if (l < 4) stv(st + W_PB + 2 * t, px, py);
        const double cx = m0 * gx.x, cy = m0 * gx.y;
        const double bx = atx, by = aty;
        {
            const double *sn = (k > 0) ? st - W_RS : st;
            af.load(sn, atx, aty); gx = ldv(sn + W_GX + 2 * t);
        }
        mm8(px, py, px, py, bx, by, cx, cy);                       // (the product of stage 0 is not used ; rows g > 0 stay 0)
    }
    __syncwarp();
}

// step of one node from the roll-out: ddx_k from the previous record's XA slot, ddu = K ddx + k_ff
__device__ __forceinline__ void node_dir(const double *st, int k, double ddx[7], double ddu[2])
{
    if (k >= 1) {
        const double *pv = st - W_RS + W_XA;
#pragma unroll
        for (int a = 0; a < 6; a += 2) { const double2 v = ldv(pv + a); ddx[a] = v.x; ddx[a + 1] = v.y; }
        ddx[6] = pv[6];
    } else {
#pragma unroll
        for (int a = 0; a < 7; a++) ddx[a] = 0.0;
    }
    double u0 = st[W_KF0], u1 = st[W_KF1];
#pragma unroll
    for (int a = 0; a < 6; a += 2) {
        const double2 k0 = ldv(st + W_K0 + a), k1 = ldv(st + W_K1 + a);
        u0 = fma(k0.y, ddx[a + 1], fma(k0.x, ddx[a], u0));
        u1 = fma(k1.y, ddx[a + 1], fma(k1.x, ddx[a], u1));
    }
    ddu[0] = fma(st[W_K0 + 6], ddx[6], u0);
    ddu[1] = fma(st[W_K1 + 6], ddx[6], u1);
}

#ifndef MMA_MINB
#define MMA_MINB 8
#endif
// NW warps per instance: thread k owns node k in the node role (N <= 32 NW - 1); the sweeps run on warp 0 while the others wait
// at the CTA barrier.  NW = 1: N <= 31, 8 instances per SM ; NW = 2: N <= 63, 4 instances per SM (BASELINE cfg4: N = 40) ;
// NW = 4: N <= 127 (long horizons: one or two instances per SM by shared memory).
template <int NW> __device__ __forceinline__ void bsync() { if (NW == 1) __syncwarp(); else __syncthreads(); }
template <int NW>
__global__ void __launch_bounds__(32 * NW, (NW == 1) ? MMA_MINB : (NW == 2) ? 4 : 1) qp_mma_kernel(const Params P)
{
    extern __shared__ __align__(16) double smr[];
    __shared__ double red[8 * NW];                   // cross-warp reductions (NW > 1): 8 slots per warp
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const int tid = threadIdx.x, l = tid & 31, wid = tid >> 5;
    const bool sweeper = (NW == 1) || wid == 0;
    const int i = blockIdx.x;                        // one instance per CTA
    double *rec = smr;
    double *term = rec + (size_t)N * W_RS;
    const double Ts = o.dt, hdt = o.dt;
#ifdef ADMPC_DEBUG
    {
        unsigned dyn;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        ADMPC_ASSERT((size_t)dyn >= ((size_t)N * W_RS + T_SIZE) * sizeof(double));
        ADMPC_ASSERT(i < P.B && N >= 2 && N <= 32 * NW - 1 && P.lin_im != nullptr);
        ADMPC_ASSERT((((size_t)(P.lin_im + ((size_t)0 * Bp + i) * LIM_STRIDE)) & 15) == 0);
    }
#endif
    const int flag = P.lin_bad[i];                   // 1: NaN/Inf in the linearisation ; 2: finished instance of the SQP loop
    if (flag) {
        if (tid == 0 && flag == 1) { P.status[i] = 1; P.qp_status[i] = 0; P.qp_iter[i] = 0; }
        if (P.gat_x) {                               // fused gather: the (untouched) iterate still goes to the root's block
            for (int k = tid; k <= N; k += 32 * NW) {
                for (int a = 0; a < 7; a++) P.gat_x[((size_t)i * (N + 1) + k) * 7 + a] = ATS(P.xb, k * 7 + a);
                if (k < N) for (int jj = 0; jj < 2; jj++) P.gat_u[((size_t)i * N + k) * 2 + jj] = ATS(P.ub, k * 2 + jj);
            }
            if (tid == 0) P.gat_st[i] = (flag == 1) ? 1 : P.status[i];
        }
        return;
    }

    // ---- stage the linearisation: one TMA bulk copy per stage record (M, b, q, r, x, u = 544 B), one mbarrier ---------------
    __shared__ uint64_t bar;
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&bar, (uint32_t)(N * LIM_STRIDE * sizeof(double)));
    }
    bsync<NW>();
    for (int k = tid; k < N; k += 32 * NW)
        tma_bulk_g2s(rec + (size_t)k * W_RS, P.lin_im + ((size_t)k * Bp + i) * LIM_STRIDE, LIM_STRIDE * sizeof(double), &bar);
    if (tid < 7) {                                   // terminal node: q_N and x_N only
        const double *rn = P.lin_im + ((size_t)N * Bp + i) * LIM_STRIDE;
        term[T_LQ + tid] = rn[LIM_Q + tid]; term[T_XB + tid] = rn[LIM_X + tid]; term[T_DX + tid] = 0.0;
        if (tid == 0) term[T_GX + 7] = 0.0;
    }
    double x0v[7];
#pragma unroll
    for (int a = 0; a < 7; a++) x0v[a] = (tid == 0) ? ATS(P.x0, a) : 0.0;
    mbar_wait(&bar, 0);
    // ---- cold start ------------------------------------------------------------------------------------------------------------
    {
        const int k = tid;
        if (k < N) {
            double *st = rec + (size_t)k * W_RS;
            const double ub0 = st[W_UB], ub1 = st[W_UB + 1], xb6 = st[W_XB + 6];
            double dx[7];
#pragma unroll
            for (int a = 0; a < 7; a++) dx[a] = (k == 0) ? x0v[a] - st[W_XB + a] : 0.0;     // x0 eliminated (nbxe_0 = 7)
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
            for (int a = 0; a < 7; a++) { st[W_DX + a] = dx[a]; st[W_PI + a] = 0.0; }
#pragma unroll
            for (int c = 0; c < NC; c += 2) { stv(st + W_LAM + c, lam[c], lam[c + 1]); stv(st + W_T + c, t[c], t[c + 1]); }
            stv(st + W_DU, du[0], du[1]);
            stv(st + W_SL, 0.0, 0.0); stv(st + W_SU, 0.0, 0.0);
            st[W_RB + 7] = 1.0; st[W_GX + 7] = 0.0;          // homogeneous coordinate / padding of the fragment rows
        }
    }
    bsync<NW>();

    const double inv_nc = 1.0 / (double)(NC * N - 2);
    const int k = tid;                               // node of this thread
    const bool act = (k < N);
    const int kc = act ? k : N - 1;                  // record the lanes without a node compute on (results masked)
    int status = 1, iter = 0;
    double res0 = 0, res1 = 0, res2 = 0, res3 = 0;
    for (iter = 0;; iter++) {
        // ================= residuals of the current point + predictor barrier terms (node role) ==============================
        double ng = 0, nb = 0, nd = 0, nm = 0, summ = 0;
        if (k == N) {
            const double *prev = rec + (size_t)(N - 1) * W_RS;
#pragma unroll
            for (int a = 0; a < 7; a++) {
                const double gq = o.We[a] * term[T_DX + a] + term[T_LQ + a] - prev[W_PI + a];
                term[T_GX + a] = gq;
                ng = nmx(ng, fabs(gq));
            }
        } else if (k < N) {
            double *st = rec + (size_t)k * W_RS;
            double lq[7], lb[7], lr[2];
            {
                double v[16];                       // b (7) q (7) r (2): 16 contiguous doubles from an even offset
#pragma unroll
                for (int a = 0; a < 16; a += 2) { const double2 t2 = ldv(st + W_LB + a); v[a] = t2.x; v[a + 1] = t2.y; }
#pragma unroll
                for (int a = 0; a < 7; a++) { lb[a] = v[a]; lq[a] = v[7 + a]; }
                lr[0] = v[14]; lr[1] = v[15];
            }
            double pi[7], dx[7];
            {
                double xp[14];
#pragma unroll
                for (int a = 0; a < 14; a += 2) { const double2 v = ldv(st + W_DX + a); xp[a] = v.x; xp[a + 1] = v.y; }
#pragma unroll
                for (int a = 0; a < 7; a++) { dx[a] = xp[a]; pi[a] = xp[7 + a]; }
            }
            const double2 duv = ldv(st + W_DU);
            // stationarity w.r.t. u, dynamics residual, stationarity w.r.t. x: one pass over the columns of M
            double rgu[2], rgx[7], rbv[6];
            {
                const double *dxn = (k + 1 < N) ? st + W_RS + W_DX : term + T_DX;
#pragma unroll
                for (int r = 0; r < 6; r++) rbv[r] = lb[r] - dxn[r] + ((r < 2) ? dx[r] : 0.0);
                const double rb6 = lb[6] - dxn[6] + dx[6] + hdt * duv.y;
                st[W_RB + 6] = rb6;
                nb = nmx(nb, fabs(rb6));
            }
#pragma unroll
            for (int cc = 0; cc < 7; cc++) {
                const double2 m01 = ldv(st + W_M + cc * 6), m23 = ldv(st + W_M + cc * 6 + 2), m45 = ldv(st + W_M + cc * 6 + 4);
                const double mm[6] = {m01.x, m01.y, m23.x, m23.y, m45.x, m45.y};
                const double xv = (cc == 0) ? duv.x : (cc == 1) ? duv.y : dx[cc];
                double gq = 0.0;
#pragma unroll
                for (int r = 0; r < 6; r++) { rbv[r] = fma(mm[r], xv, rbv[r]); gq = fma(mm[r], pi[r], gq); }
                if (cc < 2) rgu[cc] = gq; else rgx[cc] = gq;
            }
            stv(st + W_RB, rbv[0], rbv[1]); stv(st + W_RB + 2, rbv[2], rbv[3]); stv(st + W_RB + 4, rbv[4], rbv[5]);
#pragma unroll
            for (int r = 0; r < 6; r++) nb = nmx(nb, fabs(rbv[r]));
            // constraint data only now: nothing of it is live across the pass over M
            NCon C;
            load_ncon(o, st, C);
            NRes R;
            node_res_w(o, k >= 1, C, R);
            NScal S;
            node_scal_w(o, C, S);
#pragma unroll
            for (int jj = 0; jj < 2; jj++) {
                double gq = Ts * o.W[7 + jj] * C.du[jj] + lr[jj] - C.lam[jj] + C.lam[3 + jj] + rgu[jj];
                if (jj == 1) gq = fma(hdt, pi[6], gq);
                rgu[jj] = gq;
                ng = nmx(ng, nmx(fabs(gq), nmx(fabs(R.rgsl[jj]), fabs(R.rgsu[jj]))));
                nd = nmx(nd, nmx(nmx(fabs(R.rd[jj]), fabs(R.rd[3 + jj])), nmx(fabs(R.rd[6 + jj]), fabs(R.rd[8 + jj]))));
            }
            if (k >= 1) nd = nmx(nd, nmx(fabs(R.rd[2]), fabs(R.rd[5])));
            double rm[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const bool on = !((c == 2 || c == 5) && k == 0);
                rm[c] = on ? C.lam[c] * C.t[c] : 0.0;
                nm = nmx(nm, fabs(rm[c]));
                summ += rm[c];
            }
            double gx[7] = {0, 0, 0, 0, 0, 0, 0};
            if (k >= 1) {
                const double *pim = st - W_RS + W_PI;
#pragma unroll
                for (int a = 0; a < 7; a++) {
                    double gq = Ts * o.W[a] * dx[a] + lq[a] - pim[a];
                    if (a < 2) gq += pi[a];
                    else {
                        gq += rgx[a];
                        if (a == 6) gq += pi[6] - C.lam[2] + C.lam[5];
                    }
                    gx[a] = gq;
                    ng = nmx(ng, fabs(gq));
                }
            }
            // barrier-modified Hessian diagonal / gradient (soft-bound slacks eliminated)
            double gq[NC], Rt[2], rtv[2];
#pragma unroll
            for (int c = 0; c < NC; c++) gq[c] = (rm[c] - C.lam[c] * R.rd[c]) * S.it[c];
#pragma unroll
            for (int jj = 0; jj < 2; jj++) {
                Rt[jj] = Ts * o.W[7 + jj] + S.Sl[jj] * (1.0 - S.Sl[jj] * S.iDl[jj]) + S.Su[jj] * (1.0 - S.Su[jj] * S.iDu[jj]);
                const double cl = R.rgsl[jj] + gq[jj] + gq[6 + jj];
                const double cu = R.rgsu[jj] + gq[3 + jj] + gq[8 + jj];
                rtv[jj] = rgu[jj] + (gq[jj] - S.Sl[jj] * cl * S.iDl[jj]) - (gq[3 + jj] - S.Su[jj] * cu * S.iDu[jj]);
            }
            stv(st + W_BAR, Rt[0], Rt[1]); stv(st + W_BAR + 2, rtv[0], rtv[1]);
            if (k >= 1) {
                st[W_BAR + 4] = Ts * o.W[6] + C.lam[2] * S.it[2] + C.lam[5] * S.it[5];
                gx[6] = gx[6] + gq[2] - gq[5];
            } else {
                st[W_BAR + 4] = Ts * o.W[6];
            }
            stv(st + W_GX, gx[0], gx[1]); stv(st + W_GX + 2, gx[2], gx[3]); stv(st + W_GX + 4, gx[4], gx[5]);
            st[W_GX + 6] = gx[6];
        }
        // Termination tests as warp votes on the lanes' own partial norms ("every partial maximum is finite / below its tolerance"
        // is the same statement as the one about the maximum): the four norm reductions themselves are only needed for the
        // report, i.e. in the iteration that leaves the loop.
        summ = wsum32(summ);
        bool fin = __all_sync(FULL, isfinite(ng) && isfinite(nb) && isfinite(nd) && isfinite(nm));
        bool conv = __all_sync(FULL, ng < o.tol_stat && nb < o.tol_eq && nd < o.tol_ineq && nm < o.tol_comp);
        if (NW > 1) {
            if (l == 0) { red[wid * 8 + 4] = summ; red[wid * 8 + 5] = fin ? 1.0 : 0.0; red[wid * 8 + 6] = conv ? 1.0 : 0.0; }
            __syncthreads();
            summ = red[4]; fin = red[5] != 0.0; conv = red[6] != 0.0;
#pragma unroll
            for (int w = 1; w < NW; w++) { summ += red[w * 8 + 4]; fin = fin && red[w * 8 + 5] != 0.0; conv = conv && red[w * 8 + 6] != 0.0; }
            __syncthreads();
        }
        if (!fin || conv || iter >= o.iter_max) {
            ng = wmax32(ng); nb = wmax32(nb); nd = wmax32(nd); nm = wmax32(nm);
            if (NW > 1) {
                if (l == 0) { red[wid * 8] = ng; red[wid * 8 + 1] = nb; red[wid * 8 + 2] = nd; red[wid * 8 + 3] = nm; }
                __syncthreads();
                ng = red[0]; nb = red[1]; nd = red[2]; nm = red[3];
#pragma unroll
                for (int w = 1; w < NW; w++) { ng = nmx(ng, red[w * 8]); nb = nmx(nb, red[w * 8 + 1]); nd = nmx(nd, red[w * 8 + 2]); nm = nmx(nm, red[w * 8 + 3]); }
            }
            res0 = ng; res1 = nb; res2 = nd; res3 = nm;
            status = !fin ? 3 : (conv ? 0 : 1);
            break;
        }
        const double mu = summ * inv_nc;
        bsync<NW>();

        // ================= predictor (pass 0) and corrector (pass 1) ==========================================================
        // One rolled loop: the node-role arithmetic both passes share (constraint data, residuals, scalings, step of the node for
        // the direction the roll-out left) exists once in the instruction stream.  pass 0 runs with pr = 0, sigma mu = 0: rm = lam t.
        double an = 1.0, ad = 1.0, sigmu = 0.0;
        double pr[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) pr[c] = 0.0;
        NCon Cs;
        NScal Ss;
        NStep Ds;
        double duc[2] = {0.0, 0.0}, ddx[7];
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 0) {
                if (sweeper) mma_factor(o, rec, term, N, l);
            } else {
                if (sweeper) mma_backward(o, rec, term, N, l);
                bsync<NW>();
                // k_ff of the corrector: -Guu^-1 (rt + B^T h_k), node-parallel
                if (k < N) {
                    double *st = rec + (size_t)k * W_RS;
                    const double2 rt = ldv(st + W_BAR + 2);
                    const double2 h01 = ldv(st + W_PB), h23 = ldv(st + W_PB + 2), h45 = ldv(st + W_PB + 4);
                    const double h6 = st[W_PB + 6];
                    const double gu0 = rt.x + dot6v(ldv(st + W_M), ldv(st + W_M + 2), ldv(st + W_M + 4), h01, h23, h45);
                    const double gu1 = fma(hdt, h6, rt.y) + dot6v(ldv(st + W_M + 6), ldv(st + W_M + 8), ldv(st + W_M + 10), h01, h23, h45);
                    const double gi00 = st[W_GI0], gi01 = st[W_GI1], gi11 = st[W_GI2];
                    st[W_KF0] = -(gi00 * gu0 + gi01 * gu1);
                    st[W_KF1] = -(gi01 * gu0 + gi11 * gu1);
                }
            }
            bsync<NW>();
            if (sweeper) mma_forward(o, rec, N, l);
            bsync<NW>();
            NRes R;
            {
                const double *st = rec + (size_t)kc * W_RS;
                load_ncon(o, st, Cs);
                node_res_w(o, kc >= 1, Cs, R);
                node_scal_w(o, Cs, Ss);
                double rm[NC];
#pragma unroll
                for (int c = 0; c < NC; c++) rm[c] = ((c == 2 || c == 5) && kc == 0) ? 0.0 : Cs.lam[c] * Cs.t[c] + pr[c] - sigmu;
                node_dir(st, kc, ddx, duc);
                node_step_w(kc >= 1, Cs, R, Ss, rm, duc[0], duc[1], ddx[6], Ds);
            }
            if (pass == 0) {
                double s1 = 0.0, s2 = 0.0, m_aff = 1.0;
                double fa[3], fb[3];
                {
                    m_aff = node_ratio_aff(kc >= 1, Ss, Ds, m_aff);
                    double ea[NC], eb[NC];
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        const bool on = !((c == 2 || c == 5) && kc == 0);
                        pr[c] = on ? Ds.dlv[c] * Ds.dtv[c] : 0.0;
                        ea[c] = pr[c] * Ss.it[c];
                        eb[c] = on ? Ss.it[c] : 0.0;
                        if (on) {
                            s1 += Cs.lam[c] * Ds.dtv[c] + Cs.t[c] * Ds.dlv[c];
                            s2 += pr[c];
                        }
                    }
#pragma unroll
                    for (int jj = 0; jj < 2; jj++) {
                        fa[jj] = (ea[jj] - Ss.Sl[jj] * (ea[jj] + ea[6 + jj]) * Ss.iDl[jj]) - (ea[3 + jj] - Ss.Su[jj] * (ea[3 + jj] + ea[8 + jj]) * Ss.iDu[jj]);
                        fb[jj] = (eb[jj] - Ss.Sl[jj] * (eb[jj] + eb[6 + jj]) * Ss.iDl[jj]) - (eb[3 + jj] - Ss.Su[jj] * (eb[3 + jj] + eb[8 + jj]) * Ss.iDu[jj]);
                    }
                    fa[2] = ea[2] - ea[5];
                    fb[2] = eb[2] - eb[5];
                }
                // m_aff >= 1 on every lane: two REDUX instructions ; then ONE sum of the lanes' own a s1 + a^2 s2 instead of two sums
                m_aff = wmax_pos(act ? m_aff : 1.0);
                if (NW > 1) {
                    if (l == 0) red[wid * 8] = m_aff;
                    __syncthreads();
                    m_aff = red[0];
#pragma unroll
                    for (int w = 1; w < NW; w++) m_aff = fmax(m_aff, red[w * 8]);
                    __syncthreads();
                }
                const double a_aff = rcp_w(m_aff);           // min(1, min ratio) = 1 / max(1, max of the inverse ratios)
                double s12 = wsum32(act ? a_aff * (s1 + a_aff * s2) : 0.0);
                if (NW > 1) {
                    if (l == 0) red[wid * 8 + 1] = s12;
                    __syncthreads();
                    s12 = red[1];
#pragma unroll
                    for (int w = 1; w < NW; w++) s12 += red[w * 8 + 1];
                    __syncthreads();
                }
                const double mu_aff = (summ + s12) * inv_nc;
                double sigma = mu_aff * rcp_w(mu);
                sigma = sigma * sigma * sigma;
                sigmu = sigma * mu;
                if (k < N) {
                    double *st = rec + (size_t)k * W_RS;
                    const double2 rt = ldv(st + W_BAR + 2);
                    stv(st + W_BAR + 2, rt.x + fma(-sigmu, fb[0], fa[0]), rt.y + fma(-sigmu, fb[1], fa[1]));
                    if (k >= 1) st[W_GX + 6] += fma(-sigmu, fb[2], fa[2]);
                }
                bsync<NW>();
            }
        }
        if (k == N) {                                 // adjoint start: We ddx_N + r_x,N
            const double *pv = rec + (size_t)(N - 1) * W_RS + W_XA;
#pragma unroll
            for (int a = 0; a < 7; a++) term[T_GX + a] = fma(o.We[a], pv[a], term[T_GX + a]);
        } else if (k < N) {
            double *st = rec + (size_t)k * W_RS;
            if (k >= 1) {                             // adjoint base vector Qt_k ddx_k + gt_k
                const double qt6 = st[W_BAR + 4];
                double nbv[7];
#pragma unroll
                for (int a = 0; a < 7; a++) nbv[a] = fma((a == 6) ? qt6 : Ts * o.W[a], ddx[a], st[W_GX + a]);
                stv(st + W_GX, nbv[0], nbv[1]); stv(st + W_GX + 2, nbv[2], nbv[3]); stv(st + W_GX + 4, nbv[4], nbv[5]);
                st[W_GX + 6] = nbv[6];
            }
            node_ratio_lam(k >= 1, Cs, Ds, an, ad);
            const double mt = node_ratio_t(k >= 1, Ss, Ds, 1.0);
            if (ad < an * mt) { an = 1.0; ad = mt; }
        }
        // every lane's own bound an / ad (1 for the lanes without a row that limits the step), then the minimum by two REDUX
        // instructions instead of five shuffle rounds over (numerator, denominator) pairs
        double alpha = wmin_pos(an * rcp_w(ad));
        if (NW > 1) {
            if (l == 0) red[wid * 8] = alpha;
            __syncthreads();
            alpha = red[0];
#pragma unroll
            for (int w = 1; w < NW; w++) alpha = fmin(alpha, red[w * 8]);
            __syncthreads();
        }
        if (alpha < o.alpha_min) { status = 2; break; }
        if (alpha < 1.0) alpha *= 0.995;
        if (k < N) {
            double *st = rec + (size_t)k * W_RS;
            const NCon &C = Cs;
            const NStep &D = Ds;
            stv(st + W_DU, C.du[0] + alpha * duc[0], C.du[1] + alpha * duc[1]);
            stv(st + W_SL, C.sl[0] + alpha * D.dsl[0], C.sl[1] + alpha * D.dsl[1]);
            stv(st + W_SU, C.su[0] + alpha * D.dsu[0], C.su[1] + alpha * D.dsu[1]);
            double ln[NC], tn[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const bool on = !((c == 2 || c == 5) && k == 0);
                ln[c] = on ? fmax(C.lam[c] + alpha * D.dlv[c], o.lam_min) : C.lam[c];
                tn[c] = on ? fmax(C.t[c] + alpha * D.dtv[c], o.t_min) : C.t[c];
            }
#pragma unroll
            for (int c = 0; c < NC; c += 2) { stv(st + W_LAM + c, ln[c], ln[c + 1]); stv(st + W_T + c, tn[c], tn[c + 1]); }
        }
        bsync<NW>();
        // pi and dx wait for the adjoint sweep
        if (sweeper) mma_adjoint(rec, term, N, l);
        bsync<NW>();
        if (k <= N) {
            if (k < N) {
                double *st = rec + (size_t)k * W_RS;
#pragma unroll
                for (int a = 0; a < 7; a++) st[W_PI + a] += alpha * st[W_PB + a];
            }
            if (k >= 1) {
                const double *prev = rec + (size_t)(k - 1) * W_RS;      // ddx_k was left in record k-1
                double *dst = (k < N) ? rec + (size_t)k * W_RS + W_DX : term + T_DX;
#pragma unroll
                for (int a = 0; a < 7; a++) dst[a] += alpha * prev[W_XA + a];
            }
        }
        bsync<NW>();
    }

    // ---- epilogue: statuses + fused RTI update (full step; duals <- QP duals) --------------------------------------------
    const int qps = (status == 0) ? 0 : ((status == 1) ? 2 : ((status == 2) ? 3 : 1));   // hpipm -> acados numbering
    const int nlp_status = (qps == 0 || qps == 2) ? 0 : 4;
    if (tid == 0) {
        P.qp_status[i] = qps; P.qp_iter[i] = iter; P.status[i] = nlp_status;
        ATS(P.res_out, 0) = res0; ATS(P.res_out, 1) = res1; ATS(P.res_out, 2) = res2; ATS(P.res_out, 3) = res3;
    }
    const bool upd = (nlp_status == 0);
    if (k <= N) {
        ADMPC_ASSERT(soa_at(k * 7 + 6, (N + 1) * 7, i, Bp) < (size_t)(N + 1) * 7 * Bp);
        const double *st = rec + (size_t)k * W_RS;
        const double *dxs = (k < N) ? st + W_DX : term + T_DX;
        const double *xbs = (k < N) ? st + W_XB : term + T_XB;       // linearisation point: came in with the stage record
        if (upd || P.gat_x) {
#pragma unroll
            for (int a = 0; a < 7; a++) {
                double v = xbs[a];
                if (upd) { v += dxs[a]; ATS(P.xb, k * 7 + a) = v; }
                if (P.gat_x) P.gat_x[((size_t)i * (N + 1) + k) * 7 + a] = v;
            }
            if (k < N) {
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    double v = st[W_UB + jj];
                    if (upd) { v += st[W_DU + jj]; ATS(P.ub, k * 2 + jj) = v; }
                    if (P.gat_x) P.gat_u[((size_t)i * N + k) * 2 + jj] = v;
                }
            }
            if (k == 0 && P.gat_x) P.gat_st[i] = nlp_status;
        }
        if (upd && k < N) {
#pragma unroll
            for (int jj = 0; jj < 2; jj++) {
                ATS(P.slb, k * 2 + jj) = st[W_SL + jj];
                ATS(P.sub, k * 2 + jj) = st[W_SU + jj];
            }
#pragma unroll
            for (int a = 0; a < 7; a++) ATS(P.pib, k * 7 + a) = st[W_PI + a];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const bool on = !((c == 2 || c == 5) && k == 0);
                ATS(P.lamb, k * NC + c) = on ? st[W_LAM + c] : 0.0;
                ATS(P.tb, k * NC + c) = on ? st[W_T + c] : 1.0;
            }
        }
    }
}

// false: horizon outside the range of this kernel
bool launch_qp_mma(const Params &P, cudaStream_t s)
{
    const int N = P.o.N;
    if (N > 127 || !P.lin_im) return false;
    size_t sm = ((size_t)N * W_RS + T_SIZE) * sizeof(double);
#ifdef QPM_SMEM_PAD_ENV      // occupancy probe: ADMPC_SMEM_PAD = extra bytes of (unused) dynamic shared memory per CTA
    if (const char *e = getenv("ADMPC_SMEM_PAD")) sm += (size_t)atoi(e);
#endif
    if (N <= 31) {
        static SmemGuard configured;
        if (configured.need(sm)) cudaFuncSetAttribute(qp_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        qp_mma_kernel<1><<<P.B, 32, sm, s>>>(P);
    } else if (N <= 63) {
        static SmemGuard configured2;
        if (configured2.need(sm)) cudaFuncSetAttribute(qp_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        qp_mma_kernel<2><<<P.B, 64, sm, s>>>(P);
    } else {
        static SmemGuard configured4;
        if (configured4.need(sm)) cudaFuncSetAttribute(qp_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        qp_mma_kernel<4><<<P.B, 128, sm, s>>>(P);
    }
    return true;
}
