// qp_mma_f.cu -- feedback phase of the FRENET model variant (SURVEY 8a A2'; per-node curvature): the tensor-core kernel of
// qp_mma.cu re-derived for the Frenet stage structure.  ONE (N <= 31) OR TWO (N <= 63) WARPS PER INSTANCE, whole solve resident in
// shared memory, Riccati sweeps as FP64 DMMA fragment chains.
//
// Structure: in path coordinates x = [s, e_y, e_psi, v_x, v_y, r, delta] only the s column of A_k is trivial (nothing depends on
// s) and the delta row is (0, dt | 0 .. 0 1), so the stage matrix is M = [B | A(:,1:7)] (6 x 8) and the stage vector
// z = (u0, u1, x1..x6) fills all 8 columns of a tile.  State-space matrices (P, the closed-loop matrix) are laid out over the
// POSITIONS (x0, h, x1..x6): state a sits at PS(a) = a + 1 (a >= 1), s at 0, and position 1 carries the homogeneous coordinate
// in the vector sweeps -- so positions 2..7 of a state-space tile and of a z-space tile coincide and the Schur complement needs
// no permutation.  What does not fit is z plus a homogeneous column (9 > 8): the factor sweep therefore runs the gradient
// recursion as a second fragment chain of row-vector products next to the matrix chain,
//     matrix:  W^T = Ma^T P,   G = Ma^T W,   P' = H - Gu^T Guu^-1 Gu                      (5 DMMA per stage)
//     vector:  h^T = rb^T P + p^T,   gz^T = h^T Ma + g^T,   p' = gz_x + Gu^T k_ff          (4 DMMA per stage)
// both accumulator -> operand, no shared-memory operand traffic.  The roll-outs, the corrector's backward sweep and the adjoint
// sweep are the row-vector chains of qp_mma.cu over the positions.  7-vectors live in 8-double position-layout slots
// [v0, (h), v1..v6].  Node role and constraint rows are those of qp_mma.cu (con_set 0: u0, u1 soft, steering angle hard).
//
// Replaces qp_warp_f.cu (round-1 design: IPM state in registers, hand-distributed sweeps) as the default for the variant;
// identical maths to oracle/rti_oracle.c with model_backend = 2, results differ by rounding only.
#include "common.cuh"
#include "tma.cuh"

// instance-major linearisation record of the Frenet preparation kernel (frenet.cu): M 6x8 column-major, b, q, r, x, u
#define LIMF_M 0
#define LIMF_B 48
#define LIMF_Q 55
#define LIMF_R 62
#define LIMF_X 64
#define LIMF_U 71
#define LIMF_STRIDE 74
#define PS(a) (((a) == 0) ? 0 : (a) + 1)      // position of state a

// ---- node record (doubles); the first LIMF_STRIDE are pulled in by one TMA bulk copy per stage
#define W_M 0       // 48  column c (0,1 = u0,u1 ; 2..7 = x1..x6) at c*6 + r, r < 6 (next states x0..x5)
#define W_LB 48     // 7   b_k
#define W_LQ 55     // 7   q_k
#define W_LR 62     // 2   r_k
#define W_XB 64     // 7   linearisation point x_k
#define W_UB 71     // 2   linearisation point u_k
#define W_K0 74     // 8   first row of (K | k_ff), position layout: [K(x0), k_ff, K(x1..x6)]
#define W_KF0 75
#define W_K1 82     // 8   second row
#define W_KF1 83
#define W_RB 90     // 8   dynamics residual, position layout, slot 1 = 1.0 (homogeneous coordinate: never overwritten)
#define W_PB 98     // 8   P_{k+1} rb_k ; corrector backward sweep: h_k ; adjoint sweep: dpi_k   (position layout)
#define W_GX 106    // 8   gradient w.r.t. x (position layout, slot 1 = 0) ; after the corrector roll-out the adjoint base vector
#define W_BAR 114   // 5   Rt0 Rt1 | rt0 rt1 | Qt6
#define W_GI0 119   // 3   Guu^-1 (0,0), (0,1), (1,1)
#define W_GI1 120
#define W_GI2 121
#define W_DX 122    // 7   iterate: dx_k (state order)
#define W_PI 129    // 7   iterate: pi_k
#define W_LAM 136   // 10  iterate: lam
#define W_T 146     // 10  iterate: t
#define W_DU 156    // 2
#define W_SL 158    // 2
#define W_SU 160    // 2
#define W_XA 164    // 8   roll-out: [ddx_{k+1}(x0), 1, ddx_{k+1}(x1..x6)]
#define W_RS 174    // record stride (2*W_RS mod 32 = 28: 16-byte node-parallel accesses are conflict-free)
// terminal record
#define T_DX 0      // 7
#define T_GX 8      // 8   r_x,N ; later We dx_N + r_x,N (adjoint start) ; position layout
#define T_LQ 16     // 7   q_N
#define T_XB 24     // 7   x_N of the linearisation point
#define T_SIZE 32

#include "qp_node.cuh"

// D(8x8) = A(8x4) B(4x8) + C on the FP64 tensor core: a = A[g][t], b = B[t][g], (c0, c1) = C[g][2t..2t+1]
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b, double c0, double c1)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
        : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
// cl(A B + C) from cl(A) = (ax, ay), cl(B^T) = (bx, by): cl(X) = lane (g, t) holds X[g][2t], X[g][2t+1] (see qp_mma.cu)
__device__ __forceinline__ void mm8(double &d0, double &d1, double ax, double ay, double bx, double by, double c0, double c1)
{
    double e0, e1;
    dmma(e0, e1, ax, bx, c0, c1);
    dmma(d0, d1, ay, by, e0, e1);
}
#define MMAF_PRAGMA_(x) _Pragma(#x)
#define MMAF_UNROLL(n) MMAF_PRAGMA_(unroll n)
__device__ __forceinline__ double shf(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ int srow(int g) { return (g == 0) ? 0 : g - 1; }       // row of M that position g (0, 2..6) refers to

// ---- operand fragments of one stage ---------------------------------------------------------------------------------------------
// cl(Ma^T): lane (g, t) holds Ma[pos 2t][g], Ma[pos 2t+1][g] ; rows of Ma by position: (next x0, 0, next x1..x5, delta row),
// columns g = z.  delta row: (0, dt, 0, .., 0, 1).
struct MtFragF {
    int ox, oy; double cy; bool ly;
    __device__ __forceinline__ MtFragF(int g, int t, double hdt)
    {
        ox = W_M + 6 * g + ((t == 0) ? 0 : 2 * t - 1);
        ly = (t == 1 || t == 2);
        oy = ly ? W_M + 6 * g + 2 * t : W_M;
        cy = (t == 3) ? ((g == 1) ? hdt : (g == 7) ? 1.0 : 0.0) : 0.0;
    }
    __device__ __forceinline__ void load(const double *st, double &x, double &y) const
    {
        const double a = st[ox], b = st[oy];
        x = a; y = ly ? b : cy;
    }
};
// cl(A^T) over the positions, restricted to what the backward vector sweeps use: column g = 0 (s): e0 ; g = 1: zero ; g >= 2: the
// columns of Ma.  Row 1 (homogeneous) is zero everywhere: a junk entry at position 1 of the recursion vector never leaks.
struct AtFragF {
    MtFragF m; double cx; bool ld;
    __device__ __forceinline__ AtFragF(int g, int t, double hdt) : m(g, t, hdt)
    {
        ld = (g >= 2);
        cx = (g == 0 && t == 0) ? 1.0 : 0.0;
    }
    __device__ __forceinline__ void load(const double *st, double &x, double &y) const
    {
        double a, b;
        m.load(st, a, b);
        x = ld ? a : cx; y = ld ? b : 0.0;
    }
};
// rank-2 operands of the closed-loop matrix: bm = Bh[pos g][t] (B by position, delta row (0, dt), zero at position 1),
// kh = Kh[t][g] = (K | k_ff) in position layout, t < 2
struct ClFragF {
    int boff, koff; double bc; bool bl, kl;
    __device__ __forceinline__ ClFragF(int g, int t, double hdt)
    {
        bl = (t < 2 && g != 1 && g != 7);
        boff = bl ? W_M + 6 * t + srow(g) : W_M;
        bc = (g == 7 && t == 1) ? hdt : 0.0;
        kl = (t < 2);
        koff = W_K0 + 8 * (t & 1) + g;
    }
    __device__ __forceinline__ void load(const double *st, double &bm, double &kh) const
    {
        const double b = st[boff], k = st[koff];
        bm = bl ? b : bc; kh = kl ? k : 0.0;
    }
};

// ---- factor sweep (predictor): matrix chain + gradient chain ------------------------------------------------------------------
__device__ __forceinline__ void mmaf_factor(const admpc_opts &o, double *rec, const double *term, int N, int l)
{
    asm volatile("" : "+r"(l));
    const int g = l >> 2, t = l & 3;
    const double Ts = o.dt, hdt = o.dt;
    const MtFragF mf(g, t, hdt);
    const bool g0 = (g == 0), g1 = (g == 1), t0 = (t == 0), t3 = (t == 3), row0 = g0, tlow = (t < 2);
    // H[g][2t..2t+1]: own G for g >= 2, t >= 1 ; W[0][g] (and 0 at the homogeneous position) for g >= 2, t = 0 ; row 0 (state s):
    // own P(0,0) + weight at t = 0, W[0][c] handed over by the lanes (2t, 0), (2t+1, 0) otherwise ; row 1: zero.
    const bool catD = (g >= 2) && !t0, catC = (g >= 2) && t0, catA = g0 && t0, catB = g0 && !t0;
    const double dmG = catD ? 1.0 : 0.0;
    // constant part of the diagonal: Ts W[a] at position PS(a), a <= 5 (the steering angle's Qt6 and the inputs come from the record)
    const int sg = (g >= 2) ? g - 1 : 0;
    const double wdiag = Ts * sel7w(o.W, sg);
    const double cdx = ((g0 && t0) || (g >= 2 && g <= 6 && 2 * t == g)) ? wdiag : 0.0;
    const double cdy = (g >= 2 && g <= 6 && 2 * t + 1 == g) ? wdiag : 0.0;
    const double mq = (g == 7 && t3) ? 1.0 : 0.0;                 // H[7][7] += Qt6
    const double dmU = (g >= 2) ? 1.0 : 0.0;
    const int src0 = l & ~3;
    const int kst = W_K0 + 8 * (t & 1) + g;
    const bool kwr = tlow && !g1;                                  // position 1 of the gain rows is k_ff (written by the gradient chain)
    const double reg = o.reg;
    const double m0 = row0 ? 1.0 : 0.0;
    // terminal: P_N = diag(We) over the positions, p_N = r_x,N
    double px, py, pvx, pvy;
    {
        const double we = (g1) ? 0.0 : sel7w(o.We, sg);
        px = (2 * t == g) ? we : 0.0;
        py = (2 * t + 1 == g) ? we : 0.0;
        const double2 tg = ldv(term + T_GX + 2 * t);
        pvx = m0 * tg.x; pvy = m0 * tg.y;
    }
    const double *st = rec + (size_t)(N - 1) * W_RS;
    double mx, my;
    mf.load(st, mx, my);
    double2 bar01 = ldv(st + W_BAR), bar23 = ldv(st + W_BAR + 2), gxp = ldv(st + W_GX + 2 * t), rbp = ldv(st + W_RB + 2 * t);
    double qt6 = st[W_BAR + 4];
MMAF_UNROLL(1)
    for (int k = N - 1; k >= 0; k--, st -= W_RS) {
        ADMPC_ASSERT(st == rec + (size_t)k * W_RS);
        double *sw = const_cast<double *>(st);
        const double addx = cdx, addy = fma(mq, qt6, cdy);         // the steering angle sits at position 7: second entry of lane (7, 3)
        const double r00 = bar01.x + reg, r11 = bar01.y + reg;
        // gradient row over z: (rt0, rt1, gx1..gx6) ; gx0 (state s) separately
        const double grx = m0 * (t0 ? bar23.x : gxp.x), gry = m0 * (t0 ? bar23.y : gxp.y);
        const double gx0 = gxp.x;                                  // lanes t = 0 hold (gx0, 0)
        // ---- matrix chain 1: W^T = Ma^T P ; gradient chain 1: h^T = rb^T P + p^T ---------------------------------------------
        double wx, wy, hx, hy;
        mm8(wx, wy, mx, my, px, py, 0.0, 0.0);
        mm8(hx, hy, m0 * rbp.x, m0 * rbp.y, px, py, pvx, pvy);     // (slot 1 of rb is 1.0: multiplied by the zero row 1 of P)
        if (l < 4) stv(sw + W_PB + 2 * t, hx - pvx, hy - pvy);     // P rb for the corrector
        // ---- matrix chain 2: G = Ma^T W ; gradient chain 2: gz^T = h^T Ma + g^T ------------------------------------------------
        double Gx, Gy, zx, zy;
        mm8(Gx, Gy, mx, my, wx, wy, 0.0, 0.0);
        mm8(zx, zy, hx, hy, mx, my, grx, gry);
        // (from W alone) row / column of the state s: G[u][s] = W[0][u], H[s][c] = W[0][c]
        const double w4 = shf(wx, 4);                              // W[0][u1] (lane (1, 0))
        const double w0 = shf(wx, 0);                              // W[0][u0]
        const double r0x = shf(wx, 8 * t), r1x = shf(wx, 8 * t + 4);
        double hpx = catB ? r0x : (catC ? wx : 0.0), hpy = catB ? r1x : 0.0;
        hpx = catA ? px : hpx;
        hpx += addx; hpy += addy;
        hpx = g1 ? 0.0 : hpx; hpy = g1 ? 0.0 : hpy;
        const double upx = g0 ? w0 : 0.0, upy = g0 ? w4 : 0.0;
        // next stage's operands
        const double *sn = (k > 0) ? st - W_RS : st;
        mf.load(sn, mx, my);
        bar01 = ldv(sn + W_BAR); bar23 = ldv(sn + W_BAR + 2); gxp = ldv(sn + W_GX + 2 * t); rbp = ldv(sn + W_RB + 2 * t);
        qt6 = sn[W_BAR + 4];
        // ---- 2x2 pivot ----------------------------------------------------------------------------------------------------------------
        const double g00 = shf(Gx, 0) + r00, g01 = shf(Gy, 0), g11 = shf(Gy, 4) + r11;
        const double gu0 = shf(fma(dmU, Gx, upx), src0), gu1 = shf(fma(dmU, Gy, upy), src0);
        const double kta = fma(g11, gu0, -g01 * gu1), ktb = fma(g00, gu1, -g01 * gu0);
        const double kt = t0 ? kta : ktb, gut = t0 ? gu0 : gu1;
        double Dx, Dy;
        dmma(Dx, Dy, tlow ? gut : 0.0, tlow ? kt : 0.0, 0.0, 0.0);
        const double idet = rcp_w(fma(g00, g11, -g01 * g01));
        // ---- gradient chain 3: k_ff, p' = gz_x + Gu^T k_ff (lanes 0..3) -----------------------------------------------------------
        const double zu0 = shf(zx, 0), zu1 = shf(zy, 0);           // gz over the inputs
        const double kf0 = -idet * fma(g11, zu0, -g01 * zu1), kf1 = -idet * fma(g00, zu1, -g01 * zu0);
        const double G1x = shf(Gx, 4 + t), G1y = shf(Gy, 4 + t);   // G[u1][2t..2t+1] for the lanes (0, t)
        // positions 2t, 2t+1 >= 2: gz + G[u0][c] k_ff0 + G[u1][c] k_ff1 ; position 0: gx0 + h0 + W[0][u] k_ff ; position 1: 0
        const double pnx = t0 ? fma(w4, kf1, fma(w0, kf0, gx0 + hx)) : fma(G1x, kf1, fma(Gx, kf0, zx));
        const double pny = t0 ? 0.0 : fma(G1y, kf1, fma(Gy, kf0, zy));
        pvx = m0 * pnx; pvy = m0 * pny;
        // ---- Schur complement ------------------------------------------------------------------------------------------------------------
        px = fma(-idet, Dx, fma(dmG, Gx, hpx));
        py = fma(-idet, Dy, fma(dmG, Gy, hpy));
        // gains and Guu^-1 for the corrector
        const double kout = -idet * kt, gi0 = g11 * idet, gi1 = -g01 * idet, gi2 = g00 * idet;
        if (kwr) sw[kst] = kout;
        if (l == 1) { sw[W_KF0] = kf0; sw[W_KF1] = kf1; }
        if (l == 2) { sw[W_GI0] = gi0; sw[W_GI1] = gi1; sw[W_GI2] = gi2; }
    }
    __syncwarp();
}

// ---- forward roll-out over the positions: xh = [ddx(s), 1, ddx(x1..x6)], xh_{k+1}^T = xh_k^T Acl^T ----------------------------
__device__ __forceinline__ void mmaf_forward(const admpc_opts &o, double *rec, int N, int l)
{
    asm volatile("" : "+r"(l));
    const int g = l >> 2, t = l & 3;
    const double hdt = o.dt;
    const ClFragF cf(g, t, hdt);
    // cl(A0): lane (g, t) holds A0[g][2t], A0[g][2t+1] ; A0 over the positions: column 0 = e0, column 1 = (rb | 1), columns 2..7 = Ma
    const bool t0 = (t == 0), t3 = (t == 3), mrow = (g != 1 && g != 7);
    const int a0o = (mrow && !t0) ? W_M + 12 * t + srow(g) : W_M;
    const int a1o = t0 ? W_RB + g : ((mrow && !t0) ? W_M + 12 * t + 6 + srow(g) : W_M);
    const double c0x = (t0 && g == 0) ? 1.0 : 0.0;
    const double c0y = (g == 7 && t3) ? 1.0 : 0.0;
    const bool l0x = mrow && !t0, l0y = t0 || (mrow && !t0);
    double xx = 0.0, xy = (l == 0) ? 1.0 : 0.0;                  // ddx_0 = 0, homogeneous 1 at position 1 ; rows g > 0 stay 0
    double *st = rec;
    double bm, kh, a0x, a0y, acx, acy;
    cf.load(st, bm, kh);
    { const double u = st[a0o], v = st[a1o]; a0x = l0x ? u : c0x; a0y = l0y ? v : c0y; }
    dmma(acx, acy, bm, kh, a0x, a0y);
    {
        const double *sn = st + W_RS;
        cf.load(sn, bm, kh);
        const double u = sn[a0o], v = sn[a1o];
        a0x = l0x ? u : c0x; a0y = l0y ? v : c0y;
    }
MMAF_UNROLL(2)
    for (int k = 0; k < N; k++, st += W_RS) {
        mm8(xx, xy, xx, xy, acx, acy, 0.0, 0.0);
        dmma(acx, acy, bm, kh, a0x, a0y);
        {
            const double *sn = (k + 2 < N) ? st + 2 * W_RS : st;
            cf.load(sn, bm, kh);
            const double u = sn[a0o], v = sn[a1o];
            a0x = l0x ? u : c0x; a0y = l0y ? v : c0y;
        }
        if (l < 4) stv(st + W_XA + 2 * t, xx, xy);
    }
    __syncwarp();
}

// ---- corrector backward sweep: p_k^T = h_k^T (A + B K) + (gx + K^T rt)^T, h_k = P rb + p_{k+1} ; leaves h_k in the PB slot -----
__device__ __forceinline__ void mmaf_backward(const admpc_opts &o, double *rec, const double *term, int N, int l)
{
    asm volatile("" : "+r"(l));
    const int g = l >> 2, t = l & 3;
    const double hdt = o.dt;
    const ClFragF cf(g, t, hdt);
    const AtFragF af(g, t, hdt);
    const double m0 = (g == 0) ? 1.0 : 0.0;
    double px, py;
    { const double2 tg = ldv(term + T_GX + 2 * t); px = m0 * tg.x; py = m0 * tg.y; }
    double *st = rec + (size_t)(N - 1) * W_RS;
    double bm, kh, atx, aty, ctx, cty;
    cf.load(st, bm, kh); af.load(st, atx, aty);
    dmma(ctx, cty, kh, bm, atx, aty);
    cf.load(st - W_RS, bm, kh); af.load(st - W_RS, atx, aty);
    double2 pb = ldv(st + W_PB + 2 * t), gx = ldv(st + W_GX + 2 * t), k0 = ldv(st + W_K0 + 2 * t), k1 = ldv(st + W_K1 + 2 * t);
    double2 rt = ldv(st + W_BAR + 2);
MMAF_UNROLL(2)
    for (int k = N - 1; k >= 0; k--, st -= W_RS) {
        const double cx = m0 * fma(k1.x, rt.y, fma(k0.x, rt.x, gx.x)), cy = m0 * fma(k1.y, rt.y, fma(k0.y, rt.x, gx.y));
        const double hx = fma(m0, pb.x, px), hy = fma(m0, pb.y, py);
        mm8(px, py, hx, hy, ctx, cty, cx, cy);
        dmma(ctx, cty, kh, bm, atx, aty);
        if (l < 4) stv(st + W_PB + 2 * t, hx, hy);
        {
            const double *sn = (k > 0) ? st - W_RS : st, *s2 = (k > 1) ? st - 2 * W_RS : st;
            cf.load(s2, bm, kh); af.load(s2, atx, aty);
            pb = ldv(sn + W_PB + 2 * t); gx = ldv(sn + W_GX + 2 * t); k0 = ldv(sn + W_K0 + 2 * t); k1 = ldv(sn + W_K1 + 2 * t);
            rt = ldv(sn + W_BAR + 2);
        }
    }
    __syncwarp();
}

// ---- adjoint sweep: dpi_{k-1}^T = dpi_k^T A_k + base_k^T ; leaves dpi_k in the PB slot -----------------------------------------
__device__ __forceinline__ void mmaf_adjoint(const admpc_opts &o, double *rec, const double *term, int N, int l)
{
    asm volatile("" : "+r"(l));
    const int g = l >> 2, t = l & 3;
    const AtFragF af(g, t, o.dt);
    const double m0 = (g == 0) ? 1.0 : 0.0;
    double px, py;
    { const double2 tg = ldv(term + T_GX + 2 * t); px = m0 * tg.x; py = m0 * tg.y; }
    double *st = rec + (size_t)(N - 1) * W_RS;
    double atx, aty;
    af.load(st, atx, aty);
    double2 gx = ldv(st + W_GX + 2 * t);
MMAF_UNROLL(2)
    for (int k = N - 1; k >= 0; k--, st -= W_RS) {
        if (l < 4) stv(st + W_PB + 2 * t, px, py);
        const double cx = m0 * gx.x, cy = m0 * gx.y;
        const double bx = atx, by = aty;
        {
            const double *sn = (k > 0) ? st - W_RS : st;
            af.load(sn, atx, aty); gx = ldv(sn + W_GX + 2 * t);
        }
        mm8(px, py, px, py, bx, by, cx, cy);
    }
    __syncwarp();
}

// step of one node from the roll-out: ddx_k from the previous record's XA slot (position layout), ddu = K ddx + k_ff
__device__ __forceinline__ void node_dir(const double *st, int k, double ddx[7], double ddu[2])
{
    if (k >= 1) {
        const double *pv = st - W_RS + W_XA;
        const double2 a = ldv(pv + 2), b = ldv(pv + 4), c = ldv(pv + 6);
        ddx[0] = pv[0]; ddx[1] = a.x; ddx[2] = a.y; ddx[3] = b.x; ddx[4] = b.y; ddx[5] = c.x; ddx[6] = c.y;
    } else {
#pragma unroll
        for (int a = 0; a < 7; a++) ddx[a] = 0.0;
    }
    double u0 = st[W_KF0], u1 = st[W_KF1];
    u0 = fma(st[W_K0], ddx[0], u0); u1 = fma(st[W_K1], ddx[0], u1);
#pragma unroll
    for (int a = 1; a < 7; a += 2) {
        const double2 k0 = ldv(st + W_K0 + a + 1), k1 = ldv(st + W_K1 + a + 1);
        u0 = fma(k0.y, ddx[a + 1], fma(k0.x, ddx[a], u0));
        u1 = fma(k1.y, ddx[a + 1], fma(k1.x, ddx[a], u1));
    }
    ddu[0] = u0; ddu[1] = u1;
}

#ifndef MMAF_MINB
#define MMAF_MINB 8
#endif
// NW warps per instance: thread k owns node k in the node role (N <= 32 NW - 1); the sweeps run on warp 0 while the others wait
// at the CTA barrier.  NW = 1: N <= 31, 8 instances per SM ; NW = 2: N <= 63, 4 instances per SM (BASELINE cfg4: N = 40).
template <int NW> __device__ __forceinline__ void bsync() { if (NW == 1) __syncwarp(); else __syncthreads(); }
template <int NW>
__global__ void __launch_bounds__(32 * NW, (NW == 1) ? MMAF_MINB : 4) qp_mma_f_kernel(const Params P)
{
    extern __shared__ __align__(16) double smr[];
    __shared__ double red[16];                       // cross-warp reductions (NW = 2)
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
        ADMPC_ASSERT((((size_t)(P.lin_im + ((size_t)0 * Bp + i) * LIMF_STRIDE)) & 15) == 0);
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
        mbar_expect_tx(&bar, (uint32_t)(N * LIMF_STRIDE * sizeof(double)));
    }
    bsync<NW>();
    for (int k = tid; k < N; k += 32 * NW)
        tma_bulk_g2s(rec + (size_t)k * W_RS, P.lin_im + ((size_t)k * Bp + i) * LIMF_STRIDE, LIMF_STRIDE * sizeof(double), &bar);
    if (tid < 7) {                                   // terminal node: q_N and x_N only
        const double *rn = P.lin_im + ((size_t)N * Bp + i) * LIMF_STRIDE;
        term[T_LQ + tid] = rn[LIMF_Q + tid]; term[T_XB + tid] = rn[LIMF_X + tid]; term[T_DX + tid] = 0.0;
        if (tid == 0) term[T_GX + 1] = 0.0;
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
            st[W_RB + 1] = 1.0; st[W_GX + 1] = 0.0; st[W_PB + 1] = 0.0;     // homogeneous coordinate / unused position of the fragment rows
        }
    }
    bsync<NW>();

    const double inv_nc = 1.0 / (double)(NC * N - 2);
    const int k = tid;                               // node of this thread
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
                term[T_GX + PS(a)] = gq;
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
            NCon C;
            load_ncon(o, st, C);
            double pi[7], dx[7];
            {
                double xp[14];
#pragma unroll
                for (int a = 0; a < 14; a += 2) { const double2 v = ldv(st + W_DX + a); xp[a] = v.x; xp[a + 1] = v.y; }
#pragma unroll
                for (int a = 0; a < 7; a++) { dx[a] = xp[a]; pi[a] = xp[7 + a]; }
            }
            NRes R;
            node_res_w(o, k >= 1, C, R);
            NScal S;
            node_scal_w(o, C, S);
            // stationarity w.r.t. u, dynamics residual, stationarity w.r.t. x: one pass over the columns of M
            double rgu[2], rgx[7], rbv[6], rb6s;
            {
                const double *dxn = (k + 1 < N) ? st + W_RS + W_DX : term + T_DX;
#pragma unroll
                for (int r = 0; r < 6; r++) rbv[r] = lb[r] - dxn[r] + ((r < 1) ? dx[r] : 0.0);      // only the s column of A is trivial
                const double rb6 = lb[6] - dxn[6] + dx[6] + hdt * C.du[1];
                rb6s = rb6;
                nb = nmx(nb, fabs(rb6));
            }
#pragma unroll
            for (int cc = 0; cc < 8; cc++) {                  // columns u0, u1, x1..x6
                const double2 m01 = ldv(st + W_M + cc * 6), m23 = ldv(st + W_M + cc * 6 + 2), m45 = ldv(st + W_M + cc * 6 + 4);
                const double mm[6] = {m01.x, m01.y, m23.x, m23.y, m45.x, m45.y};
                const double xv = (cc < 2) ? C.du[cc] : dx[(cc >= 2) ? cc - 1 : 0];
                double gq = 0.0;
#pragma unroll
                for (int r = 0; r < 6; r++) { rbv[r] = fma(mm[r], xv, rbv[r]); gq = fma(mm[r], pi[r], gq); }
                if (cc < 2) rgu[cc] = gq; else rgx[(cc >= 2) ? cc - 1 : 0] = gq;
            }
            st[W_RB] = rbv[0]; stv(st + W_RB + 2, rbv[1], rbv[2]); stv(st + W_RB + 4, rbv[3], rbv[4]); stv(st + W_RB + 6, rbv[5], rb6s);
#pragma unroll
            for (int r = 0; r < 6; r++) nb = nmx(nb, fabs(rbv[r]));
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
                    if (a < 1) gq += pi[a];
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
            st[W_GX] = gx[0]; stv(st + W_GX + 2, gx[1], gx[2]); stv(st + W_GX + 4, gx[3], gx[4]); stv(st + W_GX + 6, gx[5], gx[6]);
        }
        ng = wmax32(ng); nb = wmax32(nb); nd = wmax32(nd); nm = wmax32(nm); summ = wsum32(summ);
        if (NW == 2) {
            if (l == 0) { red[wid * 8] = ng; red[wid * 8 + 1] = nb; red[wid * 8 + 2] = nd; red[wid * 8 + 3] = nm; red[wid * 8 + 4] = summ; }
            __syncthreads();
            ng = nmx(red[0], red[8]); nb = nmx(red[1], red[9]); nd = nmx(red[2], red[10]); nm = nmx(red[3], red[11]); summ = red[4] + red[12];
            __syncthreads();
        }
        res0 = ng; res1 = nb; res2 = nd; res3 = nm;
        if (!(isfinite(ng) && isfinite(nb) && isfinite(nd) && isfinite(nm))) { status = 3; break; }
        if (ng < o.tol_stat && nb < o.tol_eq && nd < o.tol_ineq && nm < o.tol_comp) { status = 0; break; }
        if (iter >= o.iter_max) { status = 1; break; }
        const double mu = summ * inv_nc;
        bsync<NW>();

        // ================= predictor ========================================================================================
        if (sweeper) mmaf_factor(o, rec, term, N, l);
        bsync<NW>();
        if (sweeper) mmaf_forward(o, rec, N, l);
        bsync<NW>();
        // affine step: step length, mu_aff ; the complementarity products and the two linear functionals the corrected
        // barrier gradient needs stay in registers of the node's lane
        double an = 1.0, ad = 1.0, s1 = 0.0, s2 = 0.0, m_aff = 1.0;
        double pr[NC], fa[3], fb[3];
#pragma unroll
        for (int c = 0; c < NC; c++) pr[c] = 0.0;
#pragma unroll
        for (int c = 0; c < 3; c++) { fa[c] = 0.0; fb[c] = 0.0; }
        if (k < N) {
            const double *st = rec + (size_t)k * W_RS;
            NCon C; load_ncon(o, st, C);
            NRes R; node_res_w(o, k >= 1, C, R);
            NScal S; node_scal_w(o, C, S);
            double rm[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) rm[c] = ((c == 2 || c == 5) && k == 0) ? 0.0 : C.lam[c] * C.t[c];
            double ddx[7], ddu[2];
            node_dir(st, k, ddx, ddu);
            NStep D;
            node_step_w(k >= 1, C, R, S, rm, ddu[0], ddu[1], ddx[6], D);
            m_aff = node_ratio_aff(k >= 1, S, D, m_aff);
            double ea[NC], eb[NC];       // change of g = (rm - lam rd)/t caused by rm -> rm + dlam dt - sigma mu: ea - sigma mu eb
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const bool on = !((c == 2 || c == 5) && k == 0);
                pr[c] = on ? D.dlv[c] * D.dtv[c] : 0.0;
                ea[c] = pr[c] * S.it[c];
                eb[c] = on ? S.it[c] : 0.0;
                if (on) {
                    s1 += C.lam[c] * D.dtv[c] + C.t[c] * D.dlv[c];
                    s2 += pr[c];
                }
            }
#pragma unroll
            for (int jj = 0; jj < 2; jj++) {
                fa[jj] = (ea[jj] - S.Sl[jj] * (ea[jj] + ea[6 + jj]) * S.iDl[jj]) - (ea[3 + jj] - S.Su[jj] * (ea[3 + jj] + ea[8 + jj]) * S.iDu[jj]);
                fb[jj] = (eb[jj] - S.Sl[jj] * (eb[jj] + eb[6 + jj]) * S.iDl[jj]) - (eb[3 + jj] - S.Su[jj] * (eb[3 + jj] + eb[8 + jj]) * S.iDu[jj]);
            }
            fa[2] = ea[2] - ea[5];
            fb[2] = eb[2] - eb[5];
        }
        m_aff = wmaxf32(m_aff);
        s1 = wsum32(s1); s2 = wsum32(s2);
        if (NW == 2) {
            if (l == 0) { red[wid * 8] = m_aff; red[wid * 8 + 1] = s1; red[wid * 8 + 2] = s2; }
            __syncthreads();
            m_aff = fmax(red[0], red[8]); s1 = red[1] + red[9]; s2 = red[2] + red[10];
            __syncthreads();
        }
        const double a_aff = rcp_w(m_aff);               // min(1, min ratio) = 1 / max(1, max of the inverse ratios)
        const double mu_aff = (summ + a_aff * s1 + a_aff * a_aff * s2) * inv_nc;
        double sigma = mu_aff * rcp_w(mu);
        sigma = sigma * sigma * sigma;
        const double sigmu = sigma * mu;
        if (k < N) {
            double *st = rec + (size_t)k * W_RS;
            const double2 rt = ldv(st + W_BAR + 2);
            stv(st + W_BAR + 2, rt.x + fma(-sigmu, fb[0], fa[0]), rt.y + fma(-sigmu, fb[1], fa[1]));
            if (k >= 1) st[W_GX + 7] += fma(-sigmu, fb[2], fa[2]);
        }
        bsync<NW>();
        // ================= corrector ========================================================================================
        if (sweeper) mmaf_backward(o, rec, term, N, l);
        bsync<NW>();
        // k_ff of the corrector: -Guu^-1 (rt + B^T h_k), node-parallel
        if (k < N) {
            double *st = rec + (size_t)k * W_RS;
            const double2 rt = ldv(st + W_BAR + 2);
            const double2 ha = ldv(st + W_PB + 2), hb = ldv(st + W_PB + 4), hc = ldv(st + W_PB + 6);     // h in position layout
            const double2 h01 = make_double2(st[W_PB], ha.x), h23 = make_double2(ha.y, hb.x), h45 = make_double2(hb.y, hc.x);
            const double h6 = hc.y;
            const double gu0 = rt.x + dot6v(ldv(st + W_M), ldv(st + W_M + 2), ldv(st + W_M + 4), h01, h23, h45);
            const double gu1 = fma(hdt, h6, rt.y) + dot6v(ldv(st + W_M + 6), ldv(st + W_M + 8), ldv(st + W_M + 10), h01, h23, h45);
            const double gi00 = st[W_GI0], gi01 = st[W_GI1], gi11 = st[W_GI2];
            st[W_KF0] = -(gi00 * gu0 + gi01 * gu1);
            st[W_KF1] = -(gi01 * gu0 + gi11 * gu1);
        }
        bsync<NW>();
        if (sweeper) mmaf_forward(o, rec, N, l);
        bsync<NW>();
        // final step: step length, then the update of the constraint part of the iterate from the same registers
        an = 1.0; ad = 1.0;
        NCon Cs;
        NStep Ds;
        double duc[2] = {0.0, 0.0};
        if (k == N) {                                 // adjoint start: We ddx_N + r_x,N
            const double *pv = rec + (size_t)(N - 1) * W_RS + W_XA;
#pragma unroll
            for (int a = 0; a < 7; a++) term[T_GX + PS(a)] = fma(o.We[a], pv[PS(a)], term[T_GX + PS(a)]);
        } else if (k < N) {
            double *st = rec + (size_t)k * W_RS;
            load_ncon(o, st, Cs);
            NRes R; node_res_w(o, k >= 1, Cs, R);
            NScal S; node_scal_w(o, Cs, S);
            double rm[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) rm[c] = ((c == 2 || c == 5) && k == 0) ? 0.0 : Cs.lam[c] * Cs.t[c] + pr[c] - sigmu;
            double ddx[7];
            node_dir(st, k, ddx, duc);
            if (k >= 1) {                             // adjoint base vector Qt_k ddx_k + gt_k
                const double qt6 = st[W_BAR + 4];
                double nbv[7];
#pragma unroll
                for (int a = 0; a < 7; a++) nbv[a] = fma((a == 6) ? qt6 : Ts * o.W[a], ddx[a], st[W_GX + PS(a)]);
                st[W_GX] = nbv[0]; stv(st + W_GX + 2, nbv[1], nbv[2]); stv(st + W_GX + 4, nbv[3], nbv[4]); stv(st + W_GX + 6, nbv[5], nbv[6]);
            }
            node_step_w(k >= 1, Cs, R, S, rm, duc[0], duc[1], ddx[6], Ds);
            node_ratio_lam(k >= 1, Cs, Ds, an, ad);
            const double mt = node_ratio_t(k >= 1, S, Ds, 1.0);
            if (ad < an * mt) { an = 1.0; ad = mt; }
        }
        warp_ratio(an, ad);
        if (NW == 2) {
            if (l == 0) { red[wid * 8] = an; red[wid * 8 + 1] = ad; }
            __syncthreads();
            an = red[0]; ad = red[1];
            if (red[8] * ad < an * red[9]) { an = red[8]; ad = red[9]; }
            __syncthreads();
        }
        double alpha = an * rcp_w(ad);
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
        if (sweeper) mmaf_adjoint(o, rec, term, N, l);
        bsync<NW>();
        if (k <= N) {
            if (k < N) {
                double *st = rec + (size_t)k * W_RS;
#pragma unroll
                for (int a = 0; a < 7; a++) st[W_PI + a] += alpha * st[W_PB + PS(a)];
            }
            if (k >= 1) {
                const double *prev = rec + (size_t)(k - 1) * W_RS;      // ddx_k was left in record k-1
                double *dst = (k < N) ? rec + (size_t)k * W_RS + W_DX : term + T_DX;
#pragma unroll
                for (int a = 0; a < 7; a++) dst[a] += alpha * prev[W_XA + PS(a)];
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

// false: horizon outside the range of this kernel, or no instance-major records on this handle
bool launch_qp_mma_f(const Params &P, cudaStream_t s)
{
    const int N = P.o.N;
    if (N > 63 || !P.lin_im) return false;
    const size_t sm = ((size_t)N * W_RS + T_SIZE) * sizeof(double);
    if (N <= 31) {
        static SmemGuard configured;
        if (configured.need(sm)) cudaFuncSetAttribute(qp_mma_f_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        qp_mma_f_kernel<1><<<P.B, 32, sm, s>>>(P);
    } else {
        static SmemGuard configured2;
        if (configured2.need(sm)) cudaFuncSetAttribute(qp_mma_f_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        qp_mma_f_kernel<2><<<P.B, 64, sm, s>>>(P);
    }
    return true;
}
