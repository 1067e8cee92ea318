// qp_mma_g.cu -- feedback phase of the FRENET model variant (SURVEY 8a A2'): ONE (N <= 31), TWO (N <= 63) OR FOUR (N <= 127) WARPS PER INSTANCE,
// whole IPM solve resident in shared memory, Riccati sweeps as FP64 DMMA fragment chains -- the design of qp_mma.cu for a stage
// with no exploitable column.  Covers the variant as the reference defines it: curvature as a spline kappa(s) inside the model
// (the column of s in A_k is then dense; nothing of A but the delta row is trivial) and the variant's own constraint set
// (admpc.h con_set = 1: acceleration soft, steering rate hard, e_y hard, steering angle soft; structure pinned by
// ad_mpc/debug.json), as well as per-node curvature / the shipped constraint set (con_set = 0).
//
// Structure: x = [s, e_y, e_psi, v_x, v_y, r, delta], A 7x7 dense except the delta row (0 .. 0 1 | 0, dt), so the stage vector
// z = (u0, u1, x0..x6) has NINE entries -- one more than a DMMA tile.  The state space (7 + one homogeneous coordinate for the
// vector sweeps = 8) still fits, so the factor sweep never forms the 9x9 Gram matrix: the state block goes through the matrix
// chain, and everything that touches the two inputs rides in the spare rows of the gradient chain's tile,
//     matrix:  W^T = A^T P,   Gxx = A^T W + Q,   P' = Gxx - Gux^T Guu^-1 Gux                          (5 DMMA per stage)
//     vector:  Y  = [rb^T; b0^T; b1^T] P + [p^T; 0; 0]          = [h^T; (P b0)^T; (P b1)^T]
//              Y2 = Y A + [gx^T; 0; 0]                           = [gx'^T; Gux(0,:); Gux(1,:)]
//              Y3 = Y B + [rt^T; Rt0 0; 0 Rt1]                   = [gu^T; Guu]
//              p'^T = gx'^T - det^-1 gu^T adj(Guu) Gux                                                 (7 DMMA per stage)
// all accumulator -> operand, no shared-memory operand traffic.  The roll-outs, the corrector's backward sweep and the adjoint
// sweep are the row-vector chains of qp_mma.cu over (x0..x6, 1).  The node role is written over a compile-time descriptor of the
// bounded quantities (twin of con_get in oracle/rti_oracle.c and frenet.cu), rows [lb(q) | ub(q) | ls(s) | us(s)].
//
// History: a first tensor-core kernel for this variant (qp_mma_f, z = (u, x1..x6) with the trivial column of s split off and a
// permuted position layout) covered per-node curvature and con_set = 0 only and measured 2.00 ms at B = 16384, N = 20; this
// kernel runs the same case in 1.89 ms and replaced it.  Cross-checks kept: qp_warp_f.cu (round-1 warp kernel, per-node
// curvature, con_set = 0) and the dense thread-per-instance kernel of frenet.cu (everything, also the N = 128 fallback);
// identical maths to oracle/rti_oracle.c with model_backend = 2, results differ by rounding only.
#include "common.cuh"
#include "tma.cuh"

// instance-major linearisation record of the Frenet preparation kernel (frenet.cu): M = [B | A] rows 0..5 column-major (6 x 9),
// b, q, r, x, u, one pad.  Only M is staged (one TMA bulk copy per stage); b, q, r are read once per IPM iteration and the
// linearisation point twice per solve by the node lanes straight from this record (L2-resident: the CTA has just pulled it) --
// 16 + 10 doubles less per stage in shared memory are the eighth resident instance per SM.
#define LIMG_STRIDE 80
#define G_LB 54     // 7   b_k
#define G_LQ 61     // 7   q_k
#define G_LR 68     // 2   r_k
#define G_XB 70     // 7   linearisation point x_k
#define G_UB 77     // 2   linearisation point u_k (+ 1 pad)

// ---- node record in shared memory (doubles)
#define W_M 0       // 54  column c (0,1 = u0,u1 ; 2..8 = x0..x6) at c*6 + r, r < 6 (next states x0..x5) ; the TMA'd part
#define W_BND 54    // 4   linearisation-point value of every bounded quantity (u0, u1, then the bounded states)
#define W_K0 58     // 8   first row of (K | k_ff)
#define W_KF0 65
#define W_K1 66     // 8   second row
#define W_KF1 73
#define W_RB 74     // 8   dynamics residual, slot 7 = 1.0 (homogeneous coordinate: never overwritten)
#define W_PB 82     // 8   P_{k+1} rb_k ; corrector backward sweep: h_k ; adjoint sweep: dpi_k
#define W_GX 90     // 8   gradient w.r.t. x (slot 7 = 0) ; after the corrector roll-out the adjoint base vector
#define W_BAR 98    // 6   Rt0 Rt1 | rt0 rt1 | Qt1 Qt6
#define W_GI0 104   // 3   Guu^-1 (0,0), (0,1), (1,1)
#define W_GI1 105
#define W_GI2 106
#define W_DX 108    // 7   iterate: dx_k
#define W_PI 115    // 7   iterate: pi_k
#define W_LAM 122   // 12  iterate: lam (10 rows used by con_set 0)
#define W_T 134     // 12  iterate: t
#define W_DU 146    // 2
#define W_SL 148    // 2
#define W_SU 150    // 2
#define W_XA 152    // 8   roll-out: [ddx_{k+1}, 1]
#define W_RS 162    // record stride (W_RS / 2 odd: 16-byte node-parallel accesses are conflict-free)
// terminal record
#define T_DX 0      // 7
#define T_GX 8      // 8   r_x,N ; later We dx_N + r_x,N (adjoint start)
#define T_LQ 16     // 7   q_N
#define T_XB 24     // 7   x_N of the linearisation point
#define T_SIZE 32

#define NMX_INT      // residual-norm maxima as integer maxima (qp_node.cuh)
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
#ifndef MMAG_UF
#define MMAG_UF 1        // unroll factor of the factor sweep
#endif
#ifndef MMAG_UV
#define MMAG_UV 2        // unroll factor of the vector sweeps
#endif
#define MMAG_PRAGMA_(x) _Pragma(#x)
#define MMAG_UNROLL(n) MMAG_PRAGMA_(unroll n)
__device__ __forceinline__ double shf(double v, int src) { return __shfl_sync(FULL, v, src); }

// ---- constraint sets at compile time --------------------------------------------------------------------------------------------
// bounded quantities q of z = [u0 u1 | x0..x6]: index into z, slack id (-1: hard)
template <int CS> struct CSet;
template <> struct CSet<0> { static constexpr int NQ = 3, NR = 10; };
template <> struct CSet<1> { static constexpr int NQ = 4, NR = 12; };
template <int CS> __device__ __forceinline__ constexpr int q_idx(int q) { return (CS == 0) ? ((q < 2) ? q : 8) : ((q < 2) ? q : ((q == 2) ? 3 : 8)); }
template <int CS> __device__ __forceinline__ constexpr int q_soft(int q) { return (CS == 0) ? ((q < 2) ? q : -1) : ((q == 0) ? 0 : ((q == 3) ? 1 : -1)); }
template <int CS> __device__ __forceinline__ constexpr int RL(int q) { return q; }
template <int CS> __device__ __forceinline__ constexpr int RU(int q) { return CSet<CS>::NQ + q; }
template <int CS> __device__ __forceinline__ constexpr int RLS(int s) { return 2 * CSet<CS>::NQ + s; }
template <int CS> __device__ __forceinline__ constexpr int RUS(int s) { return 2 * CSet<CS>::NQ + 2 + s; }
// quantity a row belongs to
template <int CS> __device__ __forceinline__ constexpr int row_q(int c)
{
    constexpr int NQ = CSet<CS>::NQ;
    if (c < NQ) return c;
    if (c < 2 * NQ) return c - NQ;
    const int s = (c < 2 * NQ + 2) ? c - 2 * NQ : c - 2 * NQ - 2;
    for (int q = 0; q < NQ; q++) if (q_soft<CS>(q) == s) return q;
    return 0;
}
template <int CS> __device__ __forceinline__ bool row_on(int c, bool kge1) { return q_idx<CS>(row_q<CS>(c)) < 2 || kge1; }
template <int CS> __device__ __forceinline__ double q_lo(const admpc_opts &o, int q)
{
    const int z = q_idx<CS>(q);
    return (z == 0) ? o.lbu[0] : (z == 1) ? o.lbu[1] : (z == 8) ? o.lbx : o.lbx2;
}
template <int CS> __device__ __forceinline__ double q_hi(const admpc_opts &o, int q)
{
    const int z = q_idx<CS>(q);
    return (z == 0) ? o.ubu[0] : (z == 1) ? o.ubu[1] : (z == 8) ? o.ubx : o.ubx2;
}

// ---- node-role arithmetic over the descriptor (the hard-coded set of qp_node.cuh generalised) ----------------------------------
template <int CS> struct GCon {
    double lam[CSet<CS>::NR], t[CSet<CS>::NR], sl[2], su[2], v[CSet<CS>::NQ], lo[CSet<CS>::NQ], hi[CSet<CS>::NQ];
};
template <int CS> __device__ __forceinline__ void load_gcon(const admpc_opts &o, const double *st, GCon<CS> &C)
{
    constexpr int NQ = CSet<CS>::NQ, NR = CSet<CS>::NR;
#pragma unroll
    for (int c = 0; c < NR; c += 2) {
        const double2 l = ldv(st + W_LAM + c), t = ldv(st + W_T + c);
        C.lam[c] = l.x; C.lam[c + 1] = l.y; C.t[c] = t.x; C.t[c + 1] = t.y;
    }
    const double2 du = ldv(st + W_DU), sl = ldv(st + W_SL), su = ldv(st + W_SU);
    C.sl[0] = sl.x; C.sl[1] = sl.y; C.su[0] = su.x; C.su[1] = su.y;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const int z = q_idx<CS>(q);
        const double bar = st[W_BND + q];
        C.v[q] = (z == 0) ? du.x : (z == 1) ? du.y : st[W_DX + z - 2];
        C.lo[q] = q_lo<CS>(o, q) - bar; C.hi[q] = q_hi<CS>(o, q) - bar;
    }
}
template <int CS> struct GRes { double rd[CSet<CS>::NR], rgsl[2], rgsu[2]; };
template <int CS> __device__ __forceinline__ void node_res_g(const admpc_opts &o, bool kge1, const GCon<CS> &C, GRes<CS> &R)
{
    constexpr int NQ = CSet<CS>::NQ;
    const double Ts = o.dt;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const int s = q_soft<CS>(q);
        const bool on = q_idx<CS>(q) < 2 || kge1;
        if (s >= 0) {
            R.rgsl[s] = on ? Ts * o.zl[s] + Ts * o.Zl[s] * C.sl[s] - C.lam[RL<CS>(q)] - C.lam[RLS<CS>(s)] : 0.0;
            R.rgsu[s] = on ? Ts * o.zu[s] + Ts * o.Zu[s] * C.su[s] - C.lam[RU<CS>(q)] - C.lam[RUS<CS>(s)] : 0.0;
            R.rd[RL<CS>(q)] = on ? C.t[RL<CS>(q)] - (C.v[q] - C.lo[q] + C.sl[s]) : 0.0;
            R.rd[RU<CS>(q)] = on ? C.t[RU<CS>(q)] - (C.hi[q] - C.v[q] + C.su[s]) : 0.0;
            R.rd[RLS<CS>(s)] = on ? C.t[RLS<CS>(s)] - C.sl[s] : 0.0;
            R.rd[RUS<CS>(s)] = on ? C.t[RUS<CS>(s)] - C.su[s] : 0.0;
        } else {
            R.rd[RL<CS>(q)] = on ? C.t[RL<CS>(q)] - (C.v[q] - C.lo[q]) : 0.0;
            R.rd[RU<CS>(q)] = on ? C.t[RU<CS>(q)] - (C.hi[q] - C.v[q]) : 0.0;
        }
    }
}
// 1/t, the barrier scalings and the slack-elimination pivots
template <int CS> struct GScal { double it[CSet<CS>::NR], Sl[CSet<CS>::NQ], Su[CSet<CS>::NQ], iDl[2], iDu[2]; };
template <int CS> __device__ __forceinline__ void node_scal_g(const admpc_opts &o, bool kge1, const GCon<CS> &C, GScal<CS> &S)
{
    constexpr int NQ = CSet<CS>::NQ, NR = CSet<CS>::NR;
    const double Ts = o.dt;
#pragma unroll
    for (int c = 0; c < NR; c++) S.it[c] = rcp_w(C.t[c]);
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const int s = q_soft<CS>(q);
        const bool on = q_idx<CS>(q) < 2 || kge1;
        S.Sl[q] = C.lam[RL<CS>(q)] * S.it[RL<CS>(q)]; S.Su[q] = C.lam[RU<CS>(q)] * S.it[RU<CS>(q)];
        if (s >= 0) {
            const double Ssl = C.lam[RLS<CS>(s)] * S.it[RLS<CS>(s)], Ssu = C.lam[RUS<CS>(s)] * S.it[RUS<CS>(s)];
            // a quantity that does not exist at this node (state bound at node 0) has lam = 0, t = 1: keep the pivot finite
            S.iDl[s] = rcp_w(on ? Ts * o.Zl[s] + S.Sl[q] + Ssl : 1.0);
            S.iDu[s] = rcp_w(on ? Ts * o.Zu[s] + S.Su[q] + Ssu : 1.0);
        }
    }
}
// barrier-modified Hessian diagonal entry of quantity q (to be added to the cost weight)
template <int CS> __device__ __forceinline__ double q_hess(const GScal<CS> &S, int q)
{
    const int s = q_soft<CS>(q);
    return (s >= 0) ? S.Sl[q] * (1.0 - S.Sl[q] * S.iDl[s]) + S.Su[q] * (1.0 - S.Su[q] * S.iDu[s]) : S.Sl[q] + S.Su[q];
}
// linear functional of a per-row vector e that the elimination of t, lam and the slacks sends into the gradient of quantity q;
// rs = 1 adds the slack stationarity residuals (the gradient itself), rs = 0 for corrections of it
template <int CS> __device__ __forceinline__ double q_grad(const GRes<CS> &R, const GScal<CS> &S, const double *e, int q, double rs)
{
    const int s = q_soft<CS>(q);
    if (s >= 0) {
        const double cl = rs * R.rgsl[s] + e[RL<CS>(q)] + e[RLS<CS>(s)];
        const double cu = rs * R.rgsu[s] + e[RU<CS>(q)] + e[RUS<CS>(s)];
        return (e[RL<CS>(q)] - S.Sl[q] * cl * S.iDl[s]) - (e[RU<CS>(q)] - S.Su[q] * cu * S.iDu[s]);
    }
    return e[RL<CS>(q)] - e[RU<CS>(q)];
}
// slack / t / lambda steps of one node for a given primal step dv[q] and complementarity right-hand side rm
template <int CS> struct GStep { double dsl[2], dsu[2], dtv[CSet<CS>::NR], dlv[CSet<CS>::NR]; };
template <int CS>
__device__ __forceinline__ void node_step_g(bool kge1, const GCon<CS> &C, const GRes<CS> &R, const GScal<CS> &S, const double *rm,
                                            const double *dv, GStep<CS> &D)
{
    constexpr int NQ = CSet<CS>::NQ, NR = CSet<CS>::NR;
    double gq[NR];
#pragma unroll
    for (int c = 0; c < NR; c++) gq[c] = (rm[c] - C.lam[c] * R.rd[c]) * S.it[c];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const int s = q_soft<CS>(q);
        const bool on = q_idx<CS>(q) < 2 || kge1;
        if (s >= 0) {
            const double cl = R.rgsl[s] + gq[RL<CS>(q)] + gq[RLS<CS>(s)];
            const double cu = R.rgsu[s] + gq[RU<CS>(q)] + gq[RUS<CS>(s)];
            const double dsl = on ? -(cl + S.Sl[q] * dv[q]) * S.iDl[s] : 0.0;
            const double dsu = on ? -(cu - S.Su[q] * dv[q]) * S.iDu[s] : 0.0;
            D.dsl[s] = dsl; D.dsu[s] = dsu;
            D.dtv[RL<CS>(q)] = on ? dv[q] + dsl - R.rd[RL<CS>(q)] : 0.0;
            D.dtv[RU<CS>(q)] = on ? -dv[q] + dsu - R.rd[RU<CS>(q)] : 0.0;
            D.dtv[RLS<CS>(s)] = on ? dsl - R.rd[RLS<CS>(s)] : 0.0;
            D.dtv[RUS<CS>(s)] = on ? dsu - R.rd[RUS<CS>(s)] : 0.0;
        } else {
            D.dtv[RL<CS>(q)] = on ? dv[q] - R.rd[RL<CS>(q)] : 0.0;
            D.dtv[RU<CS>(q)] = on ? -dv[q] - R.rd[RU<CS>(q)] : 0.0;
        }
    }
#pragma unroll
    for (int c = 0; c < NR; c++) D.dlv[c] = row_on<CS>(c, kge1) ? -(rm[c] + C.lam[c] * D.dtv[c]) * S.it[c] : 0.0;
}
template <int N_> __device__ __forceinline__ double max_n(const double *w)
{
    double m = w[0];
#pragma unroll
    for (int c = 1; c < N_; c++) m = fmax(m, w[c]);
    return m;
}
// ratio tests through the reciprocals the barrier terms already hold (qp_node.cuh node_ratio_t / node_ratio_aff)
template <int CS> __device__ __forceinline__ double node_ratio_t_g(bool kge1, const GScal<CS> &S, const GStep<CS> &D, double m)
{
    constexpr int NR = CSet<CS>::NR;
    double w[NR];
#pragma unroll
    for (int c = 0; c < NR; c++) w[c] = row_on<CS>(c, kge1) ? -D.dtv[c] * S.it[c] : 0.0;
    return fmax(m, max_n<NR>(w));
}
template <int CS> __device__ __forceinline__ double node_ratio_aff_g(bool kge1, const GScal<CS> &S, const GStep<CS> &D, double m)
{
    constexpr int NR = CSet<CS>::NR;
    double w[NR];
#pragma unroll
    for (int c = 0; c < NR; c++) {
        const double q = D.dtv[c] * S.it[c];
        w[c] = row_on<CS>(c, kge1) ? fmax(-q, 1.0 + q) : 0.0;
    }
    return fmax(m, max_n<NR>(w));
}
template <int CS> __device__ __forceinline__ void node_ratio_lam_g(bool kge1, const GCon<CS> &C, const GStep<CS> &D, double &an, double &ad)
{
#pragma unroll
    for (int c = 0; c < CSet<CS>::NR; c++)
        if (row_on<CS>(c, kge1) && D.dlv[c] < 0.0 && C.lam[c] * ad < an * (-D.dlv[c])) { an = C.lam[c]; ad = -D.dlv[c]; }
}

// ---- operand fragments of one stage ---------------------------------------------------------------------------------------------
// pair of adjacent record entries or a constant pair
struct PairFrag {
    int off; bool ld; double cx;
    __device__ __forceinline__ void load(const double *st, double &x, double &y) const
    {
        const double2 v = ldv(st + off);
        x = ld ? v.x : cx; y = ld ? v.y : 0.0;
    }
};
// cl(A^T) over (x0..x6, 1): lane (g, t) holds A[2t][g], A[2t+1][g] ; row 6 of A is e6^T, row / column 7 are zero (a junk entry at
// slot 7 of a recursion vector never leaks)
__device__ __forceinline__ PairFrag at_frag(int g, int t)
{
    PairFrag f;
    f.ld = (g <= 6 && t < 3);
    f.off = f.ld ? W_M + 6 * (2 + g) + 2 * t : W_M;
    f.cx = (g == 6 && t == 3) ? 1.0 : 0.0;
    return f;
}
// cl(V), V = [rb^T; b0^T; b1^T; 0 ..] (rows of the gradient chain's tile) ; b1 has dt in the delta row
__device__ __forceinline__ PairFrag v_frag(int g, int t, double hdt)
{
    PairFrag f;
    f.ld = (g == 0) || ((g == 1 || g == 2) && t < 3);
    f.off = (g == 0) ? W_RB + 2 * t : (f.ld ? W_M + 6 * (g - 1) + 2 * t : W_M);
    f.cx = (g == 2 && t == 3) ? hdt : 0.0;
    return f;
}
// cl(Bh^T), Bh = [B | 0 ..] (7 x 2 in an 8 x 8 tile)
__device__ __forceinline__ PairFrag bt_frag(int g, int t, double hdt)
{
    PairFrag f;
    f.ld = (g < 2 && t < 3);
    f.off = f.ld ? W_M + 6 * g + 2 * t : W_M;
    f.cx = (g == 1 && t == 3) ? hdt : 0.0;
    return f;
}
// rank-2 operands of the closed-loop matrix: bm = Bh[g][t], kh = (K | k_ff)[t][g], t < 2 ; KFF = false drops the k_ff column
struct ClFragG {
    int boff, koff; double bc; bool bl, kl;
    __device__ __forceinline__ ClFragG(int g, int t, double hdt, bool kff)
    {
        bl = (t < 2 && g <= 5);
        boff = bl ? W_M + 6 * t + g : W_M;
        bc = (g == 6 && t == 1) ? hdt : 0.0;
        kl = (t < 2) && (kff || g <= 6);
        koff = W_K0 + 8 * (t & 1) + g;
    }
    __device__ __forceinline__ void load(const double *st, double &bm, double &kh) const
    {
        const double b = st[boff], k = st[koff];
        bm = bl ? b : bc; kh = kl ? k : 0.0;
    }
};

// ---- factor sweep (predictor): matrix chain + gradient / input chain ----------------------------------------------------------
__device__ __forceinline__ void mmag_factor(const admpc_opts &o, double *rec, const double *term, int N, int l)
{
    asm volatile("" : "+r"(l));
    const int g = l >> 2, t = l & 3;
    const double Ts = o.dt, hdt = o.dt;
    const PairFrag af = at_frag(g, t), vf = v_frag(g, t, hdt), bf = bt_frag(g, t, hdt);
    const bool t0 = (t == 0), tlow = (t < 2), odd = (g & 1);
    const double m0 = (g == 0) ? 1.0 : 0.0;
    // diagonal of the stage Hessian: Ts W[g] except e_y (g = 1) and the steering angle (g = 6), which come with their barrier
    // terms from the record
    const double wdiag = Ts * sel7w(o.W, (g <= 6) ? g : 0);
    const double cdx = (2 * t == g && g != 6) ? wdiag : 0.0;
    const double cdy = (2 * t + 1 == g && g != 1 && g != 7) ? wdiag : 0.0;
    const double mq6 = (g == 6 && t == 3) ? 1.0 : 0.0, mq1 = (g == 1 && t0) ? 1.0 : 0.0;
    // accumulator of Y3 = Y B: row 0 = rt, rows 1, 2 = diag(Rt) + reg
    const double c3a = (l == 0) ? 1.0 : 0.0, c3b = (l == 4) ? 1.0 : 0.0, c3c = (l == 8) ? 1.0 : 0.0;
    const int sA = 4 + (g >> 1), sB = 8 + (g >> 1);
    const int kst = W_K0 + 8 * (t & 1) + g;
    const double reg = o.reg;
    // terminal: P_N = diag(We), p_N = r_x,N
    double px, py, pvx, pvy;
    {
        const double we = (g <= 6) ? sel7w(o.We, g) : 0.0;
        px = (2 * t == g) ? we : 0.0;
        py = (2 * t + 1 == g) ? we : 0.0;
        const double2 tg = ldv(term + T_GX + 2 * t);
        pvx = m0 * tg.x; pvy = m0 * tg.y;
    }
    const double *st = rec + (size_t)(N - 1) * W_RS;
    double atx, aty, vx, vy, btx, bty;
    af.load(st, atx, aty); vf.load(st, vx, vy); bf.load(st, btx, bty);
    double2 bar01 = ldv(st + W_BAR), bar23 = ldv(st + W_BAR + 2), bar45 = ldv(st + W_BAR + 4), gxp = ldv(st + W_GX + 2 * t);
MMAG_UNROLL(MMAG_UF)
    for (int k = N - 1; k >= 0; k--, st -= W_RS) {
        ADMPC_ASSERT(st == rec + (size_t)k * W_RS);
        double *sw = const_cast<double *>(st);
        const double addx = fma(mq6, bar45.y, cdx), addy = fma(mq1, bar45.x, cdy);
        const double c3x = fma(c3a, bar23.x, c3b * (bar01.x + reg)), c3y = fma(c3a, bar23.y, c3c * (bar01.y + reg));
        const double c2x = m0 * gxp.x, c2y = m0 * gxp.y;
        // ---- W^T = A^T P ; Y = V P + p --------------------------------------------------------------------------------------
        double wx, wy, yx, yy;
        mm8(wx, wy, atx, aty, px, py, 0.0, 0.0);
        mm8(yx, yy, vx, vy, px, py, pvx, pvy);
        if (l < 4) stv(sw + W_PB + 2 * t, yx - pvx, yy - pvy);      // P rb for the corrector
        // ---- Gxx = A^T W ; Y2 = Y A + gx ; Y3 = Y B + (rt ; Rt) -------------------------------------------------------------
        double Gx, Gy, y2x, y2y, y3x, y3y;
        mm8(Gx, Gy, atx, aty, wx, wy, addx, addy);
        mm8(y2x, y2y, yx, yy, atx, aty, c2x, c2y);
        mm8(y3x, y3y, yx, yy, btx, bty, c3x, c3y);
        // next stage's operands
        const double *sn = (k > 0) ? st - W_RS : st;
        af.load(sn, atx, aty); vf.load(sn, vx, vy); bf.load(sn, btx, bty);
        bar01 = ldv(sn + W_BAR); bar23 = ldv(sn + W_BAR + 2); bar45 = ldv(sn + W_BAR + 4); gxp = ldv(sn + W_GX + 2 * t);
        // ---- 2x2 pivot ----------------------------------------------------------------------------------------------------------
        const double g00 = shf(y3x, 4), g01 = shf(y3y, 4), g11 = shf(y3y, 8), gu0 = shf(y3x, 0), gu1 = shf(y3y, 0);
        const double a0 = shf(y2x, sA), a1 = shf(y2y, sA), b0 = shf(y2x, sB), b1 = shf(y2y, sB);
        const double G0 = odd ? a1 : a0, G1 = odd ? b1 : b0;          // Gux[0][g], Gux[1][g]
        const double kta = fma(g11, G0, -g01 * G1), ktb = fma(g00, G1, -g01 * G0);
        const double kt = t0 ? kta : ktb, gut = t0 ? G0 : G1, gul = t0 ? gu0 : gu1;
        double Dx, Dy, ex, ey;
        dmma(Dx, Dy, tlow ? gut : 0.0, tlow ? kt : 0.0, 0.0, 0.0);               // Gux^T adj(Guu) Gux
        dmma(ex, ey, tlow ? m0 * gul : 0.0, tlow ? kt : 0.0, 0.0, 0.0);          // row 0: gu^T adj(Guu) Gux
        const double idet = rcp_w(fma(g00, g11, -g01 * g01));
        // ---- Schur complement, gradient -------------------------------------------------------------------------------------
        px = fma(-idet, Dx, Gx);
        py = fma(-idet, Dy, Gy);
        pvx = m0 * fma(-idet, ex, y2x); pvy = m0 * fma(-idet, ey, y2y);
        // gains (k_ff in slot 7 of the rows) and Guu^-1 for the corrector
        const double kf = t0 ? fma(g11, gu0, -g01 * gu1) : fma(g00, gu1, -g01 * gu0);
        const double kout = -idet * ((g == 7) ? kf : kt);
        if (tlow) sw[kst] = kout;
        if (l == 2) { sw[W_GI0] = g11 * idet; sw[W_GI1] = -g01 * idet; sw[W_GI2] = g00 * idet; }
    }
    __syncwarp();
}

// ---- forward roll-out: xh = [ddx, 1], xh_{k+1}^T = xh_k^T Acl^T, Acl = [[A + B K, B k_ff + rb], [0, 1]] -----------------------
__device__ __forceinline__ void mmag_forward(const admpc_opts &o, double *rec, int N, int l)
{
    asm volatile("" : "+r"(l));
    const int g = l >> 2, t = l & 3;
    const double hdt = o.dt;
    const ClFragG cf(g, t, hdt, true);
    // cl(A0), A0 = [[A, rb], [0, 1]]: lane (g, t) holds A0[g][2t], A0[g][2t+1]
    const bool t3 = (t == 3), arow = (g <= 5);
    const int a0o = arow ? W_M + 6 * (2 + 2 * t) + g : W_M;
    const int a1o = (arow && !t3) ? W_M + 6 * (3 + 2 * t) + g : ((g <= 6 && t3) ? W_RB + g : W_M);
    const double c0x = (g == 6 && t3) ? 1.0 : 0.0;
    const double c0y = (g == 7 && t3) ? 1.0 : 0.0;
    const bool l0x = arow, l0y = arow || (g == 6 && t3);
    double xx = 0.0, xy = (l == 3) ? 1.0 : 0.0;                  // ddx_0 = 0, homogeneous 1 in slot 7 ; rows g > 0 stay 0
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
MMAG_UNROLL(MMAG_UV)
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
__device__ __forceinline__ void mmag_backward(const admpc_opts &o, double *rec, const double *term, int N, int l)
{
    asm volatile("" : "+r"(l));
    const int g = l >> 2, t = l & 3;
    const double hdt = o.dt;
    const ClFragG cf(g, t, hdt, false);
    const PairFrag af = at_frag(g, t);
    const double m0 = (g == 0) ? 1.0 : 0.0, m0y = (g == 0 && t != 3) ? 1.0 : 0.0;
    double px, py;
    { const double2 tg = ldv(term + T_GX + 2 * t); px = m0 * tg.x; py = m0 * tg.y; }
    double *st = rec + (size_t)(N - 1) * W_RS;
    double bm, kh, atx, aty, ctx, cty;
    cf.load(st, bm, kh); af.load(st, atx, aty);
    dmma(ctx, cty, kh, bm, atx, aty);
    cf.load(st - W_RS, bm, kh); af.load(st - W_RS, atx, aty);
    double2 pb = ldv(st + W_PB + 2 * t), gx = ldv(st + W_GX + 2 * t), k0 = ldv(st + W_K0 + 2 * t), k1 = ldv(st + W_K1 + 2 * t);
    double2 rt = ldv(st + W_BAR + 2);
MMAG_UNROLL(MMAG_UV)
    for (int k = N - 1; k >= 0; k--, st -= W_RS) {
        const double cx = m0 * fma(k1.x, rt.y, fma(k0.x, rt.x, gx.x)), cy = m0y * fma(k1.y, rt.y, fma(k0.y, rt.x, gx.y));
        const double hx = fma(m0, pb.x, px), hy = fma(m0y, pb.y, py);
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
__device__ __forceinline__ void mmag_adjoint(double *rec, const double *term, int N, int l)
{
    asm volatile("" : "+r"(l));
    const int g = l >> 2, t = l & 3;
    const PairFrag af = at_frag(g, t);
    const double m0 = (g == 0) ? 1.0 : 0.0;
    double px, py;
    { const double2 tg = ldv(term + T_GX + 2 * t); px = m0 * tg.x; py = m0 * tg.y; }
    double *st = rec + (size_t)(N - 1) * W_RS;
    double atx, aty;
    af.load(st, atx, aty);
    double2 gx = ldv(st + W_GX + 2 * t);
MMAG_UNROLL(MMAG_UV)
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

// step of one node from the roll-out: ddx_k from the previous record's XA slot, ddu = K ddx + k_ff
__device__ __forceinline__ void node_dir_g(const double *st, int k, double ddx[7], double ddu[2])
{
    if (k >= 1) {
        const double *pv = st - W_RS + W_XA;
        const double2 a = ldv(pv), b = ldv(pv + 2), c = ldv(pv + 4);
        ddx[0] = a.x; ddx[1] = a.y; ddx[2] = b.x; ddx[3] = b.y; ddx[4] = c.x; ddx[5] = c.y; ddx[6] = pv[6];
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
    u0 = fma(st[W_K0 + 6], ddx[6], u0); u1 = fma(st[W_K1 + 6], ddx[6], u1);
    ddu[0] = u0; ddu[1] = u1;
}

#ifndef MMAG_MINB
#define MMAG_MINB 8
#endif
// NW warps per instance: thread k owns node k in the node role (N <= 32 NW - 1); the sweeps run on warp 0 while the others wait
// at the CTA barrier.  NW = 1: N <= 31, 8 instances per SM at N = 20 ; NW = 2: N <= 63 ; NW = 4: N <= 127.
template <int NW> __device__ __forceinline__ void bsync() { if (NW == 1) __syncwarp(); else __syncthreads(); }
template <int NW, int CS>
__global__ void __launch_bounds__(32 * NW, (NW == 1) ? MMAG_MINB : (NW == 2) ? 4 : 1) qp_mma_g_kernel(const Params P)
{
    constexpr int NQ = CSet<CS>::NQ, NR = CSet<CS>::NR;
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
        ADMPC_ASSERT((((size_t)(P.lin_im + ((size_t)0 * Bp + i) * LIMG_STRIDE)) & 15) == 0);
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

    // ---- stage the stage matrices: one TMA bulk copy per stage (M, 432 B), one mbarrier ------------------------------------------
    __shared__ uint64_t bar;
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&bar, (uint32_t)(N * G_LB * sizeof(double)));
    }
    bsync<NW>();
    for (int k = tid; k < N; k += 32 * NW)
        tma_bulk_g2s(rec + (size_t)k * W_RS, P.lin_im + ((size_t)k * Bp + i) * LIMG_STRIDE, G_LB * sizeof(double), &bar);
    if (tid < 7) {                                   // terminal node: q_N and x_N only
        const double *rn = P.lin_im + ((size_t)N * Bp + i) * LIMG_STRIDE;
        term[T_LQ + tid] = rn[G_LQ + tid]; term[T_XB + tid] = rn[G_XB + tid]; term[T_DX + tid] = 0.0;
        if (tid == 0) term[T_GX + 7] = 0.0;
    }
    double x0v[7];
#pragma unroll
    for (int a = 0; a < 7; a++) x0v[a] = (tid == 0) ? ATS(P.x0, a) : 0.0;
    mbar_wait(&bar, 0);
    // ---- cold start: primal 0 pushed thr0 inside its box, t from the box, lam = mu0 / t ------------------------------------------
    {
        const int k = tid;
        if (k < N) {
            double *st = rec + (size_t)k * W_RS;
            const double *gr = P.lin_im + ((size_t)k * Bp + i) * LIMG_STRIDE;               // this node's record in HBM / L2
            double xbv[7], ubv[2];
            {
                double v[10];                       // x (7) u (2) pad from an even offset
#pragma unroll
                for (int a = 0; a < 10; a += 2) { const double2 t2 = ldv(gr + G_XB + a); v[a] = t2.x; v[a + 1] = t2.y; }
#pragma unroll
                for (int a = 0; a < 7; a++) xbv[a] = v[a];
                ubv[0] = v[7]; ubv[1] = v[8];
            }
            double dx[7];
#pragma unroll
            for (int a = 0; a < 7; a++) dx[a] = (k == 0) ? x0v[a] - xbv[a] : 0.0;           // x0 eliminated (nbxe_0 = 7)
            double du[2] = {0.0, 0.0}, lam[NR], t[NR];
#pragma unroll
            for (int c = 0; c < NR; c++) { lam[c] = 0.0; t[c] = 1.0; }
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const int z = q_idx<CS>(q), s = q_soft<CS>(q);
                if (z >= 2 && k == 0) continue;
                const double bar_ = (z < 2) ? ubv[z] : xbv[z - 2];
                const double lo = q_lo<CS>(o, q) - bar_, hi = q_hi<CS>(o, q) - bar_;
                double v = 0.0;
                if (v - lo < o.thr0) {
                    if (hi - v < o.thr0) v = 0.5 * (lo + hi);
                    else v = lo + o.thr0;
                } else if (hi - v < o.thr0) v = hi - o.thr0;
                if (z < 2) du[z] = v; else dx[z - 2] = v;
                const double tl = fmax(o.thr0, v - lo), tu = fmax(o.thr0, hi - v);
                t[RL<CS>(q)] = tl; t[RU<CS>(q)] = tu;
                lam[RL<CS>(q)] = o.mu0 / tl; lam[RU<CS>(q)] = o.mu0 / tu;
                if (s >= 0) {
                    t[RLS<CS>(s)] = o.thr0; t[RUS<CS>(s)] = o.thr0;
                    lam[RLS<CS>(s)] = o.mu0 / o.thr0; lam[RUS<CS>(s)] = o.mu0 / o.thr0;
                }
            }
#pragma unroll
            for (int a = 0; a < 7; a++) { st[W_DX + a] = dx[a]; st[W_PI + a] = 0.0; }
#pragma unroll
            for (int c = 0; c < NR; c += 2) { stv(st + W_LAM + c, lam[c], lam[c + 1]); stv(st + W_T + c, t[c], t[c + 1]); }
            stv(st + W_DU, du[0], du[1]);
            stv(st + W_SL, 0.0, 0.0); stv(st + W_SU, 0.0, 0.0);
            {
                double bnd[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int q = 0; q < NQ; q++) { const int z = q_idx<CS>(q); bnd[q] = (z < 2) ? ubv[z] : xbv[z - 2]; }
                stv(st + W_BND, bnd[0], bnd[1]); stv(st + W_BND + 2, bnd[2], bnd[3]);
            }
            st[W_RB + 7] = 1.0; st[W_GX + 7] = 0.0; st[W_PB + 7] = 0.0;     // homogeneous coordinate / unused slot of the fragment rows
        }
    }
    bsync<NW>();

    const double inv_nc = 1.0 / (double)((CS == 0) ? 10 * N - 2 : 12 * N - 6);
    const int k = tid;                               // node of this thread
    const bool kge1 = (k >= 1), act = (k < N);
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
            // b, q, r of this node from its global record (issued first: an L2 round trip that the pass over M covers)
            const double *gr = P.lin_im + ((size_t)k * Bp + i) * LIMG_STRIDE;
            double lb[7], lq[7], lr[2];
            {
                double v[16];                       // b (7) q (7) r (2): 16 contiguous doubles from an even offset
#pragma unroll
                for (int a = 0; a < 16; a += 2) { const double2 t2 = ldv(gr + G_LB + a); v[a] = t2.x; v[a + 1] = t2.y; }
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
            const double du[2] = {duv.x, duv.y};
            // stationarity w.r.t. u, dynamics residual, stationarity w.r.t. x: one pass over the columns of M
            double rgu[2], rgx[7], rbv[6], rb6s;
            {
                const double *dxn = (k + 1 < N) ? st + W_RS + W_DX : term + T_DX;
#pragma unroll
                for (int r = 0; r < 6; r++) rbv[r] = -dxn[r];
                rb6s = dx[6] - dxn[6] + hdt * du[1];
            }
#pragma unroll
            for (int cc = 0; cc < 9; cc++) {                  // columns u0, u1, x0..x6
                const double2 m01 = ldv(st + W_M + cc * 6), m23 = ldv(st + W_M + cc * 6 + 2), m45 = ldv(st + W_M + cc * 6 + 4);
                const double mm[6] = {m01.x, m01.y, m23.x, m23.y, m45.x, m45.y};
                const double xv = (cc < 2) ? du[cc] : dx[(cc >= 2) ? cc - 2 : 0];
                double gq = 0.0;
#pragma unroll
                for (int r = 0; r < 6; r++) { rbv[r] = fma(mm[r], xv, rbv[r]); gq = fma(mm[r], pi[r], gq); }
                if (cc < 2) rgu[cc] = gq; else rgx[(cc >= 2) ? cc - 2 : 0] = gq;
            }
#pragma unroll
            for (int r = 0; r < 6; r++) rbv[r] += lb[r];          // (b last: its load is the one that comes from L2)
            rb6s += lb[6];
            stv(st + W_RB, rbv[0], rbv[1]); stv(st + W_RB + 2, rbv[2], rbv[3]); stv(st + W_RB + 4, rbv[4], rbv[5]); st[W_RB + 6] = rb6s;
#pragma unroll
            for (int r = 0; r < 6; r++) nb = nmx(nb, fabs(rbv[r]));
            nb = nmx(nb, fabs(rb6s));
            rgx[6] += pi[6];
            rgu[1] = fma(hdt, pi[6], rgu[1]);
            // constraint data only now: nothing of it is live across the pass over M
            GCon<CS> C;
            load_gcon<CS>(o, st, C);
            GRes<CS> R;
            node_res_g<CS>(o, kge1, C, R);
            GScal<CS> S;
            node_scal_g<CS>(o, kge1, C, S);
            // multipliers of the bounded quantities, complementarity, norms
            double rm[NR];
#pragma unroll
            for (int c = 0; c < NR; c++) {
                const bool on = row_on<CS>(c, kge1);
                rm[c] = on ? C.lam[c] * C.t[c] : 0.0;
                nm = nmx(nm, fabs(rm[c]));
                nd = nmx(nd, fabs(R.rd[c]));
                summ += rm[c];
            }
#pragma unroll
            for (int s = 0; s < 2; s++) ng = nmx(ng, nmx(fabs(R.rgsl[s]), fabs(R.rgsu[s])));
#pragma unroll
            for (int jj = 0; jj < 2; jj++) rgu[jj] += Ts * o.W[7 + jj] * du[jj] + lr[jj];
            double gx[7] = {0, 0, 0, 0, 0, 0, 0};
            if (kge1) {
                const double *pim = st - W_RS + W_PI;
#pragma unroll
                for (int a = 0; a < 7; a++) gx[a] = Ts * o.W[a] * dx[a] + lq[a] - pim[a] + rgx[a];
            }
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const int z = q_idx<CS>(q);
                const double dl = -C.lam[RL<CS>(q)] + C.lam[RU<CS>(q)];
                if (z < 2) rgu[z] += dl; else if (kge1) gx[z - 2] += dl;
            }
            ng = nmx(ng, nmx(fabs(rgu[0]), fabs(rgu[1])));
#pragma unroll
            for (int a = 0; a < 7; a++) ng = nmx(ng, fabs(gx[a]));
            // barrier-modified Hessian diagonal / gradient (soft-bound slacks eliminated)
            double gq[NR], Rt[2], rtv[2], Qt1 = Ts * o.W[1], Qt6 = Ts * o.W[6];
#pragma unroll
            for (int c = 0; c < NR; c++) gq[c] = (rm[c] - C.lam[c] * R.rd[c]) * S.it[c];
            Rt[0] = Ts * o.W[7]; Rt[1] = Ts * o.W[8];
            rtv[0] = rgu[0]; rtv[1] = rgu[1];
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const int z = q_idx<CS>(q);
                const double hq = q_hess<CS>(S, q), gb = q_grad<CS>(R, S, gq, q, 1.0);
                if (z < 2) { Rt[z] += hq; rtv[z] += gb; }
                else if (kge1) {
                    if (z == 3) Qt1 += hq; else Qt6 += hq;
                    gx[z - 2] += gb;
                }
            }
            stv(st + W_BAR, Rt[0], Rt[1]); stv(st + W_BAR + 2, rtv[0], rtv[1]); stv(st + W_BAR + 4, Qt1, Qt6);
            stv(st + W_GX, gx[0], gx[1]); stv(st + W_GX + 2, gx[2], gx[3]); stv(st + W_GX + 4, gx[4], gx[5]); st[W_GX + 6] = gx[6];
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
        // the direction the roll-out left) exists once in the instruction stream -- the iteration body is streamed from the
        // instruction cache by every warp, its size is time.  pass 0 runs with pr = 0, sigma mu = 0: rm = lam t.
        double an = 1.0, ad = 1.0, sigmu = 0.0;
        double pr[NR];
#pragma unroll
        for (int c = 0; c < NR; c++) pr[c] = 0.0;
        GCon<CS> Cs;
        GScal<CS> Ss;
        GStep<CS> Ds;
        double duc[2] = {0.0, 0.0}, ddx[7];
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 0) {
                if (sweeper) mmag_factor(o, rec, term, N, l);
            } else {
                if (sweeper) mmag_backward(o, rec, term, N, l);
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
            if (sweeper) mmag_forward(o, rec, N, l);
            bsync<NW>();
            // step of this node along the direction of the roll-out.  Unconditional: the lanes without a node work on the last
            // record and are masked where it matters (a conditional assignment would keep the previous pass's values of every
            // one of these registers alive through the sweeps)
            GRes<CS> R;
            {
                const double *st = rec + (size_t)kc * W_RS;
                load_gcon<CS>(o, st, Cs);
                node_res_g<CS>(o, kge1, Cs, R);
                node_scal_g<CS>(o, kge1, Cs, Ss);
                double rm[NR];
#pragma unroll
                for (int c = 0; c < NR; c++) rm[c] = row_on<CS>(c, kge1) ? Cs.lam[c] * Cs.t[c] + pr[c] - sigmu : 0.0;
                double dv[NQ];
                node_dir_g(st, kc, ddx, duc);
#pragma unroll
                for (int q = 0; q < NQ; q++) { const int z = q_idx<CS>(q); dv[q] = (z < 2) ? duc[z] : ddx[z - 2]; }
                node_step_g<CS>(kge1, Cs, R, Ss, rm, dv, Ds);
            }
            if (pass == 0) {
                // affine step: step length, mu_aff ; the complementarity products stay in registers of the node's lane, the two
                // linear functionals the corrected barrier gradient needs go into the record
                double s1 = 0.0, s2 = 0.0, m_aff = 1.0;
                double fa[NQ], fb[NQ];
                {
                    m_aff = node_ratio_aff_g<CS>(kge1, Ss, Ds, m_aff);
                    double ea[NR], eb[NR];   // change of g = (rm - lam rd)/t caused by rm -> rm + dlam dt - sigma mu: ea - sigma mu eb
#pragma unroll
                    for (int c = 0; c < NR; c++) {
                        const bool on = row_on<CS>(c, kge1);
                        pr[c] = on ? Ds.dlv[c] * Ds.dtv[c] : 0.0;
                        ea[c] = pr[c] * Ss.it[c];
                        eb[c] = on ? Ss.it[c] : 0.0;
                        if (on) {
                            s1 += Cs.lam[c] * Ds.dtv[c] + Cs.t[c] * Ds.dlv[c];
                            s2 += pr[c];
                        }
                    }
#pragma unroll
                    for (int q = 0; q < NQ; q++) { fa[q] = q_grad<CS>(R, Ss, ea, q, 0.0); fb[q] = q_grad<CS>(R, Ss, eb, q, 0.0); }
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
                    if (kge1) {
#pragma unroll
                        for (int q = 2; q < NQ; q++) st[W_GX + q_idx<CS>(q) - 2] += fma(-sigmu, fb[q], fa[q]);
                    }
                }
                bsync<NW>();
            }
        }
        // final step: step length, then the update of the constraint part of the iterate from the same registers
        if (k == N) {                                 // adjoint start: We ddx_N + r_x,N
            const double *pv = rec + (size_t)(N - 1) * W_RS + W_XA;
#pragma unroll
            for (int a = 0; a < 7; a++) term[T_GX + a] = fma(o.We[a], pv[a], term[T_GX + a]);
        } else if (k < N) {
            double *st = rec + (size_t)k * W_RS;
            if (kge1) {                               // adjoint base vector Qt_k ddx_k + gt_k
                const double2 qt = ldv(st + W_BAR + 4);
                double nbv[7];
#pragma unroll
                for (int a = 0; a < 7; a++) nbv[a] = fma((a == 1) ? qt.x : ((a == 6) ? qt.y : Ts * o.W[a]), ddx[a], st[W_GX + a]);
                stv(st + W_GX, nbv[0], nbv[1]); stv(st + W_GX + 2, nbv[2], nbv[3]); stv(st + W_GX + 4, nbv[4], nbv[5]); st[W_GX + 6] = nbv[6];
            }
            node_ratio_lam_g<CS>(kge1, Cs, Ds, an, ad);
            const double mt = node_ratio_t_g<CS>(kge1, Ss, Ds, 1.0);
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
            const GCon<CS> &C = Cs;
            const GStep<CS> &D = Ds;
            const double2 du = ldv(st + W_DU);
            stv(st + W_DU, du.x + alpha * duc[0], du.y + alpha * duc[1]);
            stv(st + W_SL, C.sl[0] + alpha * D.dsl[0], C.sl[1] + alpha * D.dsl[1]);
            stv(st + W_SU, C.su[0] + alpha * D.dsu[0], C.su[1] + alpha * D.dsu[1]);
            double ln[NR], tn[NR];
#pragma unroll
            for (int c = 0; c < NR; c++) {
                const bool on = row_on<CS>(c, kge1);
                ln[c] = on ? fmax(C.lam[c] + alpha * D.dlv[c], o.lam_min) : C.lam[c];
                tn[c] = on ? fmax(C.t[c] + alpha * D.dtv[c], o.t_min) : C.t[c];
            }
#pragma unroll
            for (int c = 0; c < NR; c += 2) { stv(st + W_LAM + c, ln[c], ln[c + 1]); stv(st + W_T + c, tn[c], tn[c + 1]); }
        }
        bsync<NW>();
        // pi and dx wait for the adjoint sweep
        if (sweeper) mmag_adjoint(rec, term, N, l);
        bsync<NW>();
        if (k <= N) {
            if (k < N) {
                double *st = rec + (size_t)k * W_RS;
#pragma unroll
                for (int a = 0; a < 7; a++) st[W_PI + a] += alpha * st[W_PB + a];
            }
            if (kge1) {
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
        const double *gr = P.lin_im + ((size_t)k * Bp + i) * LIMG_STRIDE;
        const double *xbs = (k < N) ? gr + G_XB : term + T_XB;       // linearisation point: from the global record again
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
                    double v = gr[G_UB + jj];
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
            for (int c = 0; c < NR; c++) {
                const bool on = row_on<CS>(c, kge1);
                ATS(P.lamb, k * NR + c) = on ? st[W_LAM + c] : 0.0;
                ATS(P.tb, k * NR + c) = on ? st[W_T + c] : 1.0;
            }
        }
    }
}

template <int NW, int CS> static void launch_one(const Params &P, size_t sm, cudaStream_t s)
{
    static SmemGuard configured;
    if (configured.need(sm)) cudaFuncSetAttribute(qp_mma_g_kernel<NW, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    qp_mma_g_kernel<NW, CS><<<P.B, 32 * NW, sm, s>>>(P);
}
// false: horizon outside the range of this kernel, or no instance-major records on this handle
bool launch_qp_mma_g(const Params &P, cudaStream_t s)
{
    const int N = P.o.N;
    if (N > 127 || !P.lin_im) return false;
    const size_t sm = ((size_t)N * W_RS + T_SIZE) * sizeof(double);
    const bool cs1 = (P.o.con_set == 1);
    if (N <= 31) { if (cs1) launch_one<1, 1>(P, sm, s); else launch_one<1, 0>(P, sm, s); }
    else if (N <= 63) { if (cs1) launch_one<2, 1>(P, sm, s); else launch_one<2, 0>(P, sm, s); }
    else { if (cs1) launch_one<4, 1>(P, sm, s); else launch_one<4, 0>(P, sm, s); }
    return true;
}
