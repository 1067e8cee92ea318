// qp_rw.cu -- feedback phase, v6: ONE WARP PER MPC INSTANCE with the WHOLE solve resident in shared memory (N <= 63).
//
// Per shooting node one record of W_RS doubles holds the stage matrices M = [B | A(:,2:7)], the Riccati outputs
// (K, Guu^-1, P rb, k_ff), the right-hand-side vectors of the sweeps AND the node's slice of the IPM iterate
// (du, dx, pi, lam, t, slacks).  Nothing of the iterate lives in registers across the sweeps (the v4 kernel kept 72
// doubles per lane: 1.6 KB of spills, 20 M local loads per launch); 22 KB per instance at N = 20 -> 10 warps per SM.
//
// Lane roles:
//   * node role   : lane k owns node k (and k + 32 when N > 31): residuals, barrier terms, step lengths, update --
//                   shared memory -> registers -> shared memory with 16-byte accesses, one pass per phase; what a later
//                   phase of the same iteration needs again (affine complementarity products, the final step) stays in
//                   registers of the same lane.
//   * factor role : lane = (qd, c), qd = lane >> 3 a row pair, c = lane & 7 a column (c < 7: column c of M, c = 7: the
//                   vector column rb / p).  W(2qd..2qd+1, c) = P(rows) col_c  ->  columns of W exchanged through shared
//                   memory  ->  G(2qd..2qd+1, c) = M(:, rows)^T W(:, c)  ->  2x2 pivot  ->  Schur complement, every lane
//                   two entries.  Operands needed by many lanes (P rows, M columns, the u-rows of G, the gains) are 16-byte
//                   BROADCAST loads (strides chosen so that the four row pairs hit distinct banks), not shuffles.  The vector
//                   recursion p_k rides in the c = 7 lanes of the same instruction stream.
//   * vector role : corrector backward sweep, the two roll-outs, adjoint sweep: lane <-> row / column of the 7x7 stage
//                   matrix, the 7-vector of the recursion broadcast through a double-buffered slot.
//
// Algorithm: HPIPM-style Mehrotra predictor-corrector IPM on the OCP-structured QP [EXT], replacing
// FULL_CONDENSING_HPIPM (acados_solver_sim_car.c:145,688-693); identical maths to the oracle (oracle/rti_oracle.c
// orc_qp_solve), results differ by rounding only.  The corrector's barrier gradient is formed as predictor value + the
// change caused by the complementarity right-hand side.
#include "common.cuh"
#include "tma.cuh"


// ---- node record (doubles).  The first LIM_STRIDE doubles are the instance-major linearisation record written by the
// preparation kernel (common.cuh LIM_*), pulled in by ONE TMA bulk copy per stage.  7-vectors start at even offsets
// (16-byte loads), the odd slots between them hold scalars.
#define W_M 0       // 42  column c (0,1 = u0,u1 ; 2..6 = x2..x6) at c*6 + r, r < 6
#define W_LB 42     // 7   b_k   (b, q, r: 16 contiguous doubles, read once per IPM iteration)
#define W_LQ 49     // 7   q_k
#define W_LR 56     // 2   r_k
#define W_XB 58     // 7   linearisation point x_k (bounds in delta form, full step of the epilogue)
#define W_UB 65     // 2   linearisation point u_k
#define W_K0 68     // 7   first row of the gain K over the states x0..x6
#define W_GI0 75    //     Guu^-1 (0,0)
#define W_K1 76     // 7   second row
#define W_GI1 83    //     Guu^-1 (0,1)
#define W_RB 84     // 7   dynamics residual ; the corrector roll-out leaves ddx_{k+1} here
#define W_GI2 91    //     Guu^-1 (1,1)
#define W_PB 92     // 7   P_{k+1} rb_k ; the adjoint sweep leaves dpi_k here
#define W_KF0 99    //     k_ff
#define W_GX 100    // 7   rgx0..rgx5, qt6 ; the corrector roll-out leaves the adjoint base vector here
#define W_KF1 107
#define W_BAR 108   // 5   Rt0 Rt1 | rt0 rt1 | Qt6 ; the corrector roll-out leaves its ddu in rt0, rt1
#define W_DD 113    // 3   affine ddu0 ddu1 ddx_k[6]
#define W_DX 116    // 7   iterate: dx_k
#define W_PI 123    // 7   iterate: pi_k   (dx, pi = 14 contiguous doubles from an even offset)
#define W_LAM 130   // 10  iterate: lam
#define W_T 140     // 10  iterate: t
#define W_DU 150    // 2
#define W_SL 152    // 2
#define W_SU 154    // 2
#define W_RS 158    // record stride (even; 16-byte node-parallel accesses are conflict-free: 2*W_RS mod 32 = 28)
// terminal record
#define T_DX 0      // 7
#define T_GX 8      // 7   r_x,N ; later We dx_N + r_x,N (adjoint start)
#define T_LQ 16     // 7   q_N
#define T_XB 24     // 7   x_N of the linearisation point
#define T_SIZE 32
// scratch of one instance.  Row stride 10 doubles (80 B): the four row pairs of a broadcast load fall on distinct banks.
#define XS 10
#define X_P 0       // 80  P_{k+1}, full symmetric, row a at a*XS (row 7 / column 7 padding) ; pads of rows 0, 1: the
                    //     (G[u0][x_a], G[u1][x_a]) pairs of the states x0, x1
#define X_W 80      // 80  W(:, c) at c*XS ; the pad of column c: (G[u0][c], G[u1][c]), (g_u0, g_u1) at c = 7
#define X_HV 80     //     vector sweeps (aliases X_W): double-buffered broadcast of the 7-vector of the recursion
#define X_SIZE 160

#include "qp_node.cuh"

// ---- factor sweep (predictor): Riccati factorisation + affine vector recursion, lanes (qd, c) -----------------------------
__device__ __forceinline__ void rw_factor(const admpc_opts &o, double *rec, double *term, double *xs, int N, int l)
{
    asm volatile("" : "+r"(l));          // per-lane role constants are rebuilt here, not carried through the node role
    const int qd = l >> 3, c = l & 7;
    const double Ts = o.dt, hdt = o.dt;
    const bool vec = (c == 7);
    const int colOff = vec ? W_RB : W_M + 6 * c;
    const double m6c = (c == 1) ? hdt : (c == 6) ? 1.0 : 0.0;
    // diagonal of G at (c, c): from the record for c = 0, 1 (Rt), 6 (Qt6), a constant weight for c = 2..5
    const int doff = (c == 1) ? W_BAR + 1 : (c == 6) ? W_BAR + 4 : W_BAR;
    const bool dsel = (c < 2 || c == 6);
    const double dconst = (c >= 2 && c < 6) ? Ts * sel7w(o.W, c) : 0.0;
    const double dm0 = (2 * qd == c) ? 1.0 : 0.0, dm1 = (2 * qd + 1 == c) ? 1.0 : 0.0;
    const double vm = vec ? 1.0 : 0.0;
    // M(6, d) of this lane's two rows d = 2qd, 2qd + 1: M(6,1) = dt, M(6,6) = 1
    const double m6d0 = (qd == 3) ? 1.0 : 0.0, m6d1 = (qd == 0) ? hdt : 0.0;
    const int goff = (c < 2) ? X_P + c * XS + 8 : X_W + c * XS + 8;
    const bool q0 = (qd == 0);
    const bool st_gi = (qd == 1) && vec, st_w01 = q0 && (c < 2);
    const int kidx0 = vec ? W_KF0 : W_K0 + c, kidx1 = vec ? W_KF1 : W_K1 + c;
    const bool st_row = q0 && (c < 7), st_col = (c >= 2 && c < 7);
    const bool useW = q0 && (c >= 2), useP = q0 && (c < 2);
    const double pd0 = (c == 0) ? Ts * o.W[0] : 0.0, pd1 = (c == 1) ? Ts * o.W[1] : 0.0;
    const int a0 = 2 * qd, a1 = 2 * qd + 1;
    const double *pr0 = xs + X_P + a0 * XS, *pr1 = xs + X_P + a1 * XS;
    double *wst = xs + X_W + c * XS + 2 * qd;
    const double *wld = xs + X_W + c * XS;
    const int mo0 = W_M + 6 * a0, mo1 = W_M + 6 * a1;      // a1 = 7 (qd = 3) reads b_k: a padding entry
    // broadcast pad holding (G[u0][a], G[u1][a]) of row a: the states x0, x1 in the pads of P, the M-columns 2..6 in the pads of W
    const int roff0 = (a0 < 2) ? X_P + a0 * XS + 8 : X_W + a0 * XS + 8;
    const int roff1 = (a1 < 2) ? X_P + a1 * XS + 8 : X_W + ((a1 < 7) ? a1 : 6) * XS + 8;

    // terminal: P_N = diag(We), p_N = r_x,N
    xs[X_P + l] = 0.0; xs[X_P + 32 + l] = 0.0;
    if (l < 16) xs[X_P + 64 + l] = 0.0;
    __syncwarp();
    if (l < 7) xs[X_P + l * XS + l] = sel7w(o.We, l);
    double pv0 = (vec && a0 < 7) ? term[T_GX + a0] : 0.0, pv1 = (vec && a1 < 7) ? term[T_GX + a1] : 0.0;
    __syncwarp();
    double *st = rec + (size_t)(N - 1) * W_RS;
    for (int k = N - 1; k >= 0; k--, st -= W_RS) {
        ADMPC_ASSERT(st == rec + (size_t)k * W_RS && st + mo1 + 5 < rec + (size_t)N * W_RS + T_SIZE + X_SIZE);
        // ---- 1. W(a, c) = P(a, :) col_c for the two rows of this lane ----------------------------------------------------
        const double2 c01 = ldv(st + colOff), c23 = ldv(st + colOff + 2), c45 = ldv(st + colOff + 4);
        const double rb6 = st[W_RB + 6];
        const double col6 = vec ? rb6 : m6c;
        const double2 p01 = ldv(pr0), p23 = ldv(pr0 + 2), p45 = ldv(pr0 + 4), p67 = ldv(pr0 + 6);
        const double2 r01 = ldv(pr1), r23 = ldv(pr1 + 2), r45 = ldv(pr1 + 4), r67 = ldv(pr1 + 6);
        double wq0 = fma(p67.x, col6, dot6v(p01, p23, p45, c01, c23, c45));
        double wq1 = fma(r67.x, col6, dot6v(r01, r23, r45, c01, c23, c45));
        const double pold0 = (c == 0) ? p01.x : p01.y, pold1 = (c == 0) ? r01.x : r01.y;    // P_{k+1}(0..1, c): the (x0, x1) block
        if (vec) {      // P rb is kept for the corrector
            if (qd < 3) stv(st + W_PB + a0, wq0, wq1);
            else st[W_PB + 6] = wq0;
        }
        wq0 = fma(vm, pv0, wq0);          // vector column continues as h = P rb + p
        wq1 = fma(vm, pv1, wq1);
        // ---- 2. columns of W through shared memory: w[0..7] = W(:, c) --------------------------------------------------------
        stv(wst, wq0, wq1);
        __syncwarp();
        const double2 w01 = ldv(wld), w23 = ldv(wld + 2), w45 = ldv(wld + 4), w67 = ldv(wld + 6);
        // ---- 3. G(d, c) = M(:, d)^T W(:, c) for d = 2qd, 2qd + 1 (+ diagonal) ; on the vector column: M(:, d)^T h -----------
        const double dv = st[doff];
        const double dval = dsel ? dv : dconst;
        const double2 rt = ldv(st + W_BAR + 2);
        double G0, G1;
        {
            const double2 m01 = ldv(st + mo0), m23 = ldv(st + mo0 + 2), m45 = ldv(st + mo0 + 4);
            const double2 n01 = ldv(st + mo1), n23 = ldv(st + mo1 + 2), n45 = ldv(st + mo1 + 4);
            G0 = fma(m6d0, w67.x, dot6v(m01, m23, m45, w01, w23, w45));
            G1 = fma(m6d1, w67.x, dot6v(n01, n23, n45, w01, w23, w45));
            G0 = fma(dm0, dval, G0);
            G1 = fma(dm1, dval, G1);
        }
        // ---- 4. the u-rows of G go to the broadcast pads ----------------------------------------------------------------------
        if (q0) stv(xs + X_W + c * XS + 8, fma(vm, rt.x, G0), fma(vm, rt.y, G1));      // vector column: g_u = rt + B^T h
        if (st_w01) { xs[X_P + 8 + c] = w01.x; xs[X_P + XS + 8 + c] = w01.y; }           // G[u_c][x0], G[u_c][x1]
        __syncwarp();
        // ---- 5. 2x2 pivot, gains, Schur complement P_k(a, x_c) = base + G[u][a]^T K(:, c) ; vector column: p_k(a) ---------------
        // (G[u][a] of this lane's two rows comes from the same broadcast pads as G[u][c]: no round trip through the gains)
        const double2 gA = ldv(xs + X_W + 8), gB = ldv(xs + X_W + XS + 8);           // G00 G10 | G01 G11
        const double2 gc = ldv(xs + goff);                                           // (G[u0][.], G[u1][.]) of this lane's column
        const double2 ga0 = ldv(xs + roff0), ga1 = ldv(xs + roff1);                  // ... of this lane's two rows
        const double2 gx = ldv(st + W_GX + a0);
        const double g00 = gA.x + o.reg, g01 = gA.y, g11 = gB.y + o.reg;
        const double idet = rcp_w(g00 * g11 - g01 * g01);
        const double gi00 = g11 * idet, gi01 = -g01 * idet, gi11 = g00 * idet;
        const double K0c = -(gi00 * gc.x + gi01 * gc.y), K1c = -(gi01 * gc.x + gi11 * gc.y);
        if (q0) { st[kidx0] = K0c; st[kidx1] = K1c; }
        if (st_gi) { st[W_GI0] = gi00; st[W_GI1] = gi01; st[W_GI2] = gi11; }
        {
            double b0 = useW ? w01.x : (useP ? pold0 + pd0 : G0);
            double b1 = useW ? w01.y : (useP ? pold1 + pd1 : G1);
            b0 = fma(vm, gx.x, b0);
            b1 = fma(vm, gx.y, b1);
            const double v0 = fma(ga0.y, K1c, fma(ga0.x, K0c, b0));
            const double v1 = fma(ga1.y, K1c, fma(ga1.x, K0c, b1));
            if (st_col) stv(xs + X_P + c * XS + a0, v0, v1);
            if (st_row) { xs[X_P + c] = v0; xs[X_P + XS + c] = v1; }
            pv0 = v0; pv1 = v1;
        }
        __syncwarp();
    }
}

// The vector sweeps need 9..11 lanes.  RW_VEC16 = 1 runs them on the lower half warp only (lanes 16..31 wait at the closing
// __syncwarp): a 16-byte shared load of a half warp costs half the wavefronts of a full-warp one.
#ifndef RW_VEC16
#define RW_VEC16 1
#endif
#ifndef RW_UV
#define RW_UV 2         // unroll factor of the horizon loops of the vector sweeps
#endif
#define RW_PRAGMA_(x) _Pragma(#x)
#define RW_UNROLL(n) RW_PRAGMA_(unroll n)
#if RW_VEC16
#define VMASK 0x0000ffffu
#define VEC_ONLY if (l < 16)
#else
#define VMASK FULL
#define VEC_ONLY
#endif

// ---- corrector backward sweep: vector part only ------------------------------------------------------------------------------
// lane v < 7: M-column v (g_v = base + M(:,v)^T h) ; lanes 7, 8: the states x0, x1.  p_k lives on lanes 7, 8, 2..6 (entries
// x0, x1, x2..x6); h = P rb + p is published through the double-buffered broadcast slot (one __syncwarp per stage).
__device__ __forceinline__ void rw_backward_vec(const admpc_opts &o, double *rec, double *term, double *xs, int N, int l)
{
    asm volatile("" : "+r"(l));
    const double hdt = o.dt;
    const int v = (l < 9) ? l : 0;
    const int gb = (v < 2) ? W_BAR + 2 + v : (v < 7) ? W_GX + v : W_GX + (v - 7);
    const int mOff = W_M + ((v < 7) ? v : 0) * 6;
    const double m6 = (v == 1) ? hdt : (v == 6) ? 1.0 : 0.0;
    const bool isx01 = (v >= 7);
    const int sx = (v >= 2 && v < 7) ? v : (v == 8) ? 1 : 0;    // state index of the p entry this lane carries
    const bool owner = (l >= 2 && l < 9);
    const bool kfj = (l == 10);
    const int gio = kfj ? W_GI1 : W_GI0, gio2 = kfj ? W_GI2 : W_GI1;
    double pown = term[T_GX + sx];                              // p_N = r_x,N
    double *st = rec + (size_t)(N - 1) * W_RS;
    VEC_ONLY
RW_UNROLL(RW_UV)
    for (int k = N - 1; k >= 0; k--, st -= W_RS) {
        double *hb = xs + X_HV + (k & 1) * 8;
        const double hown = st[W_PB + sx] + pown;
        if (owner) hb[sx] = hown;
        __syncwarp(VMASK);
        const double2 h01 = ldv(hb), h23 = ldv(hb + 2), h45 = ldv(hb + 4);
        const double h6 = hb[6];
        const double2 m01 = ldv(st + mOff), m23 = ldv(st + mOff + 2), m45 = ldv(st + mOff + 4);
        const double d = fma(m6, h6, dot6v(m01, m23, m45, h01, h23, h45));
        const double g = st[gb] + (isx01 ? hown : d);
        const double gu0 = __shfl_sync(VMASK, g, 0), gu1 = __shfl_sync(VMASK, g, 1);
        const double c0 = st[gio], c1 = st[gio2];               // lane 9: (gi00, gi01) ; lane 10: (gi01, gi11)
        const double kf = -(c0 * gu0 + c1 * gu1);
        if (l == 9) st[W_KF0] = kf;
        if (l == 10) st[W_KF1] = kf;
        pown = g + st[W_K0 + sx] * gu0 + st[W_K1 + sx] * gu1;
    }
    __syncwarp();
}

// ---- forward roll-out: lanes 0..5 rows of M, lane 6 the delta row, lanes 7 / 8 the gain rows ---------------------------------
// ADJ (corrector): ddu goes to the rt slots of the record, ddx_{k+1} to the rb slot, the adjoint base vector
// Qt_k ddx_k + gt_k to the r_x slot.  Predictor: (ddu, ddx_k[6]) to the DD slot only.
template <bool ADJ>
__device__ __forceinline__ void rw_forward(const admpc_opts &o, double *rec, double *term, double *xs, int N, int l)
{
    asm volatile("" : "+r"(l));
    const double hdt = o.dt, Ts = o.dt;
    const int l7 = (l < 7) ? l : 6;
    const double wq_l = Ts * sel7w(o.W, l7), we_l = sel7w(o.We, l7);
    const double cself = (l < 2 || l == 6) ? 1.0 : 0.0, cdt = (l == 6) ? hdt : 0.0, mB = (l < 6) ? 1.0 : 0.0;
    const double is6 = (l == 6) ? 1.0 : 0.0;
    const int job = (l < 9) ? l : 8;                       // lanes 9..31 shadow lane 8
    const bool isK = (job >= 7);
    const int rbase = isK ? ((job == 7) ? W_K0 : W_K1) : W_M + ((job < 6) ? job : 5);
    const int rstr = isK ? 1 : 6;                          // element c of the lane's row at rbase + c * rstr
    const int offi = isK ? ((job == 7) ? W_KF0 : W_KF1) : W_RB + l7;
    const int o1 = rstr, o2 = 2 * rstr, o3 = 3 * rstr, o4 = 4 * rstr, o5 = 5 * rstr, o6 = 6 * rstr;
    double dxr = 0.0;
    double *st = rec;
    VEC_ONLY
RW_UNROLL(RW_UV)
    for (int k = 0; k < N; k++, st += W_RS) {
        const double *row = st + rbase;
        double *bx = xs + X_HV + (k & 1) * 8;
        if (l < 8) bx[l] = dxr;                        // lane 7 writes the zero pad
        __syncwarp(VMASK);
        const double2 x01 = ldv(bx), x23 = ldv(bx + 2), x45 = ldv(bx + 4);
        const double x6 = bx[6];
        double ta = row[o2] * x23.x, tb = row[o3] * x23.y;
        ta = fma(row[o4], x45.x, ta); tb = fma(row[o5], x45.y, tb);
        ta = fma(row[o6], x6, ta);
        const double e0 = row[0], e1 = row[o1], off = st[offi];
        const double duj = fma(e0, x01.x, off) + fma(e1, x01.y, tb) + ta;      // gain lanes: ddu_j = K_j . x + k_ff,j
        const double du0 = __shfl_sync(VMASK, duj, 7), du1 = __shfl_sync(VMASK, duj, 8);
        if (l == 7) {
            if (ADJ) stv(st + W_BAR + 2, du0, du1);
            else { st[W_DD + 0] = du0; st[W_DD + 1] = du1; st[W_DD + 2] = x6; }
        }
        if (ADJ && k >= 1) {
            const double Qd = fma(is6, st[W_BAR + 4] - wq_l, wq_l);
            const double nb = fma(Qd, dxr, st[W_GX + l7]);
            if (l < 7) st[W_GX + l] = nb;
        }
        const double d = fma(e0, du0, ta) + fma(e1, du1, tb);                  // state lanes: (A ddx)[l] + (B ddu)[l]
        double v = off + fma(cself, dxr, cdt * du1);
        v = fma(mB, d, v);
        dxr = (l < 7) ? v : 0.0;
        if (ADJ && l < 7) st[W_RB + l] = dxr;           // ddx_{k+1}
    }
    __syncwarp();
    if (ADJ && l < 7) term[T_GX + l] = fma(we_l, dxr, term[T_GX + l]);
    __syncwarp();
}

// ---- adjoint sweep: dpi_{k-1} = base_k + A_k^T dpi_k ; leaves dpi_k in the P rb slot -----------------------------------------
__device__ __forceinline__ void rw_adjoint(double *rec, double *term, double *xs, int N, int l)
{
    asm volatile("" : "+r"(l));
    const int l7 = (l < 7) ? l : 6;
    const int lc = (l >= 2 && l < 7) ? l : 2;
    const double cself = (l < 2 || l == 6) ? 1.0 : 0.0, mA = (l >= 2 && l < 7) ? 1.0 : 0.0;
    double dpr = (l < 7) ? term[T_GX + l] : 0.0;             // dpi_{N-1} = We dx_N + r_x,N
    double *st = rec + (size_t)(N - 1) * W_RS;
    VEC_ONLY
RW_UNROLL(RW_UV)
    for (int k = N - 1; k >= 0; k--, st -= W_RS) {
        double *hb = xs + X_HV + (k & 1) * 8;
        if (l < 7) { st[W_PB + l] = dpr; hb[l] = dpr; }
        if (k == 0) break;
        __syncwarp(VMASK);
        const double2 q01 = ldv(hb), q23 = ldv(hb + 2), q45 = ldv(hb + 4);
        const double *mc = st + W_M + lc * 6;
        const double d = dot6v(ldv(mc), ldv(mc + 2), ldv(mc + 4), q01, q23, q45);
        const double v = st[W_GX + l7] + fma(cself, dpr, mA * d);  // A(:,0..1) = e0,e1 ; A[6][6] = 1
        dpr = (l < 7) ? v : 0.0;
    }
    __syncwarp();
}

// NS node slots per lane: lane l owns nodes l, l + 32, ...  (NS = 1: N <= 31, NS = 2: N <= 63)
#ifndef RW_MINB
#define RW_MINB 8      // 255 registers: the node role keeps a whole node in registers between the step-length reduction and the update
#endif
template <int NS>
__global__ void __launch_bounds__(32, (NS == 1) ? RW_MINB : 5) qp_rw_kernel(const Params P)
{
    extern __shared__ __align__(16) double smr[];
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const int l = threadIdx.x;
    const int i = blockIdx.x;                        // one instance per CTA
    double *rec = smr;
    double *term = rec + (size_t)N * W_RS;
    double *xs = term + T_SIZE;
    const double Ts = o.dt, hdt = o.dt;
#ifdef ADMPC_DEBUG
    {
        unsigned dyn;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        ADMPC_ASSERT((size_t)dyn >= ((size_t)N * W_RS + T_SIZE + X_SIZE) * sizeof(double));
        ADMPC_ASSERT(i < P.B && N >= 2 && N <= 32 * NS - 1 && P.lin_im != nullptr);
        ADMPC_ASSERT((((size_t)(P.lin_im + ((size_t)0 * Bp + i) * LIM_STRIDE)) & 15) == 0);
    }
#endif
    const int flag = P.lin_bad[i];                   // 1: NaN/Inf in the linearisation ; 2: finished instance of the SQP loop
    if (flag) {
        if (l == 0 && flag == 1) { P.status[i] = 1; P.qp_status[i] = 0; P.qp_iter[i] = 0; }
        if (P.gat_x) {                               // fused gather: the (untouched) iterate still goes to the root's block
            for (int k = l; k <= N; k += 32) {
                for (int a = 0; a < 7; a++) P.gat_x[((size_t)i * (N + 1) + k) * 7 + a] = ATS(P.xb, k * 7 + a);
                if (k < N) for (int jj = 0; jj < 2; jj++) P.gat_u[((size_t)i * N + k) * 2 + jj] = ATS(P.ub, k * 2 + jj);
            }
            if (l == 0) P.gat_st[i] = (flag == 1) ? 1 : P.status[i];
        }
        return;
    }

    // ---- stage the linearisation: one TMA bulk copy per stage record (M, b, q, r, x, u = 544 B), one mbarrier ---------------
    __shared__ uint64_t bar;
    if (l == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&bar, (uint32_t)(N * LIM_STRIDE * sizeof(double)));
    }
    __syncwarp();
    for (int k = l; k < N; k += 32)
        tma_bulk_g2s(rec + (size_t)k * W_RS, P.lin_im + ((size_t)k * Bp + i) * LIM_STRIDE, LIM_STRIDE * sizeof(double), &bar);
    if (l < 7) {                                     // terminal node: q_N and x_N only
        const double *rn = P.lin_im + ((size_t)N * Bp + i) * LIM_STRIDE;
        term[T_LQ + l] = rn[LIM_Q + l]; term[T_XB + l] = rn[LIM_X + l]; term[T_DX + l] = 0.0;
    }
    double x0v[7];
#pragma unroll
    for (int a = 0; a < 7; a++) x0v[a] = (l == 0) ? ATS(P.x0, a) : 0.0;
    mbar_wait(&bar, 0);
    // ---- cold start ------------------------------------------------------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int k = l + 32 * s;
        if (k >= N) continue;
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
    }
    __syncwarp();

    const double inv_nc = 1.0 / (double)(NC * N - 2);
    int status = 1, iter = 0;
    double res0 = 0, res1 = 0, res2 = 0, res3 = 0;
    for (iter = 0;; iter++) {
        // ================= residuals of the current point + predictor barrier terms (node role) ==============================
        double ng = 0, nb = 0, nd = 0, nm = 0, summ = 0;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int k = l + 32 * s;
            if (k > N) continue;
            if (k == N) {
                const double *prev = rec + (size_t)(N - 1) * W_RS;
#pragma unroll
                for (int a = 0; a < 7; a++) {
                    const double gq = o.We[a] * term[T_DX + a] + term[T_LQ + a] - prev[W_PI + a];
                    term[T_GX + a] = gq;
                    ng = nmx(ng, fabs(gq));
                }
                continue;
            }
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
            double rgu[2], rgx[7], rbv[6];
            {
                const double *dxn = (k + 1 < N) ? st + W_RS + W_DX : term + T_DX;
#pragma unroll
                for (int r = 0; r < 6; r++) rbv[r] = lb[r] - dxn[r] + ((r < 2) ? dx[r] : 0.0);
                const double rb6 = lb[6] - dxn[6] + dx[6] + hdt * C.du[1];
                st[W_RB + 6] = rb6;
                nb = nmx(nb, fabs(rb6));
            }
#pragma unroll
            for (int cc = 0; cc < 7; cc++) {
                const double2 m01 = ldv(st + W_M + cc * 6), m23 = ldv(st + W_M + cc * 6 + 2), m45 = ldv(st + W_M + cc * 6 + 4);
                const double mm[6] = {m01.x, m01.y, m23.x, m23.y, m45.x, m45.y};
                const double xv = (cc < 2) ? C.du[cc] : dx[cc];
                double gq = 0.0;
#pragma unroll
                for (int r = 0; r < 6; r++) { rbv[r] = fma(mm[r], xv, rbv[r]); gq = fma(mm[r], pi[r], gq); }
                if (cc < 2) rgu[cc] = gq; else rgx[cc] = gq;
            }
            stv(st + W_RB, rbv[0], rbv[1]); stv(st + W_RB + 2, rbv[2], rbv[3]); stv(st + W_RB + 4, rbv[4], rbv[5]);
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
        ng = wmax32(ng); nb = wmax32(nb); nd = wmax32(nd); nm = wmax32(nm); summ = wsum32(summ);
        res0 = ng; res1 = nb; res2 = nd; res3 = nm;
        if (!(isfinite(ng) && isfinite(nb) && isfinite(nd) && isfinite(nm))) { status = 3; break; }
        if (ng < o.tol_stat && nb < o.tol_eq && nd < o.tol_ineq && nm < o.tol_comp) { status = 0; break; }
        if (iter >= o.iter_max) { status = 1; break; }
        const double mu = summ * inv_nc;
        __syncwarp();

        // ================= predictor ========================================================================================
        rw_factor(o, rec, term, xs, N, l);
        rw_forward<false>(o, rec, term, xs, N, l);
        // affine step: step length, mu_aff ; the complementarity products and the two linear functionals the corrected
        // barrier gradient needs stay in registers of the node's lane
        double an = 1.0, ad = 1.0, s1 = 0.0, s2 = 0.0;
        double pr[NS][NC], fa[NS][3], fb[NS][3];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int k = l + 32 * s;
#pragma unroll
            for (int c = 0; c < NC; c++) pr[s][c] = 0.0;
#pragma unroll
            for (int c = 0; c < 3; c++) { fa[s][c] = 0.0; fb[s][c] = 0.0; }
            if (k >= N) continue;
            const double *st = rec + (size_t)k * W_RS;
            NCon C; load_ncon(o, st, C);
            NRes R; node_res_w(o, k >= 1, C, R);
            NScal S; node_scal_w(o, C, S);
            double rm[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) rm[c] = ((c == 2 || c == 5) && k == 0) ? 0.0 : C.lam[c] * C.t[c];
            NStep D;
            node_step_w(k >= 1, C, R, S, rm, st[W_DD + 0], st[W_DD + 1], st[W_DD + 2], D);
            node_ratio_w(k >= 1, C, D, an, ad);
            double ea[NC], eb[NC];       // change of g = (rm - lam rd)/t caused by rm -> rm + dlam dt - sigma mu: ea - sigma mu eb
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const bool on = !((c == 2 || c == 5) && k == 0);
                pr[s][c] = on ? D.dlv[c] * D.dtv[c] : 0.0;
                ea[c] = pr[s][c] * S.it[c];
                eb[c] = on ? S.it[c] : 0.0;
                if (on) {
                    s1 += C.lam[c] * D.dtv[c] + C.t[c] * D.dlv[c];
                    s2 += pr[s][c];
                }
            }
#pragma unroll
            for (int jj = 0; jj < 2; jj++) {
                fa[s][jj] = (ea[jj] - S.Sl[jj] * (ea[jj] + ea[6 + jj]) * S.iDl[jj]) - (ea[3 + jj] - S.Su[jj] * (ea[3 + jj] + ea[8 + jj]) * S.iDu[jj]);
                fb[s][jj] = (eb[jj] - S.Sl[jj] * (eb[jj] + eb[6 + jj]) * S.iDl[jj]) - (eb[3 + jj] - S.Su[jj] * (eb[3 + jj] + eb[8 + jj]) * S.iDu[jj]);
            }
            fa[s][2] = ea[2] - ea[5];
            fb[s][2] = eb[2] - eb[5];
        }
        warp_ratio(an, ad);
        s1 = wsum32(s1); s2 = wsum32(s2);
        const double a_aff = an * rcp_w(ad);
        const double mu_aff = (summ + a_aff * s1 + a_aff * a_aff * s2) * inv_nc;
        double sigma = mu_aff * rcp_w(mu);
        sigma = sigma * sigma * sigma;
        const double sigmu = sigma * mu;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int k = l + 32 * s;
            if (k >= N) continue;
            double *st = rec + (size_t)k * W_RS;
            const double2 rt = ldv(st + W_BAR + 2);
            stv(st + W_BAR + 2, rt.x + fma(-sigmu, fb[s][0], fa[s][0]), rt.y + fma(-sigmu, fb[s][1], fa[s][1]));
            if (k >= 1) st[W_GX + 6] += fma(-sigmu, fb[s][2], fa[s][2]);
        }
        __syncwarp();
        // ================= corrector ========================================================================================
        rw_backward_vec(o, rec, term, xs, N, l);
        rw_forward<true>(o, rec, term, xs, N, l);
        // final step: step length, then the update of the constraint part of the iterate from the same registers
        an = 1.0; ad = 1.0;
        NCon Cs[NS];
        NStep Ds[NS];
        double duc[NS][2];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int k = l + 32 * s;
            if (k >= N) continue;
            const double *st = rec + (size_t)k * W_RS;
            load_ncon(o, st, Cs[s]);
            NRes R; node_res_w(o, k >= 1, Cs[s], R);
            NScal S; node_scal_w(o, Cs[s], S);
            double rm[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) rm[c] = ((c == 2 || c == 5) && k == 0) ? 0.0 : Cs[s].lam[c] * Cs[s].t[c] + pr[s][c] - sigmu;
            const double2 du = ldv(st + W_BAR + 2);
            duc[s][0] = du.x; duc[s][1] = du.y;
            const double dx6 = (k >= 1) ? (st - W_RS)[W_RB + 6] : 0.0;
            node_step_w(k >= 1, Cs[s], R, S, rm, du.x, du.y, dx6, Ds[s]);
            node_ratio_w(k >= 1, Cs[s], Ds[s], an, ad);
        }
        warp_ratio(an, ad);
        double alpha = an * rcp_w(ad);
        if (alpha < o.alpha_min) { status = 2; break; }
        if (alpha < 1.0) alpha *= 0.995;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int k = l + 32 * s;
            if (k >= N) continue;
            double *st = rec + (size_t)k * W_RS;
            const NCon &C = Cs[s];
            const NStep &D = Ds[s];
            stv(st + W_DU, C.du[0] + alpha * duc[s][0], C.du[1] + alpha * duc[s][1]);
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
        __syncwarp();
        // pi and dx wait for the adjoint sweep
        rw_adjoint(rec, term, xs, N, l);
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int k = l + 32 * s;
            if (k > N) continue;
            if (k < N) {
                double *st = rec + (size_t)k * W_RS;
#pragma unroll
                for (int a = 0; a < 7; a++) st[W_PI + a] += alpha * st[W_PB + a];
            }
            if (k >= 1) {
                const double *prev = rec + (size_t)(k - 1) * W_RS;      // ddx_k was left in record k-1
                double *dst = (k < N) ? rec + (size_t)k * W_RS + W_DX : term + T_DX;
#pragma unroll
                for (int a = 0; a < 7; a++) dst[a] += alpha * prev[W_RB + a];
            }
        }
        __syncwarp();
    }

    // ---- epilogue: statuses + fused RTI update (full step; duals <- QP duals) --------------------------------------------
    const int qps = (status == 0) ? 0 : ((status == 1) ? 2 : ((status == 2) ? 3 : 1));   // hpipm -> acados numbering
    const int nlp_status = (qps == 0 || qps == 2) ? 0 : 4;
    if (l == 0) {
        P.qp_status[i] = qps; P.qp_iter[i] = iter; P.status[i] = nlp_status;
        ATS(P.res_out, 0) = res0; ATS(P.res_out, 1) = res1; ATS(P.res_out, 2) = res2; ATS(P.res_out, 3) = res3;
    }
    const bool upd = (nlp_status == 0);
    for (int k = l; k <= N; k += 32) {
        ADMPC_ASSERT(k >= 0 && k <= N && soa_at(k * 7 + 6, (N + 1) * 7, i, Bp) < (size_t)(N + 1) * 7 * Bp);
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
bool launch_qp_rw(const Params &P, cudaStream_t s)
{
    const int N = P.o.N;
    if (N > 63 || !P.lin_im) return false;
    const size_t sm = ((size_t)N * W_RS + T_SIZE + X_SIZE) * sizeof(double);
    if (N <= 31) {
        static SmemGuard configured;
        if (configured.need(sm)) cudaFuncSetAttribute(qp_rw_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        qp_rw_kernel<1><<<P.B, 32, sm, s>>>(P);
    } else {
        static SmemGuard configured2;
        if (configured2.need(sm)) cudaFuncSetAttribute(qp_rw_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        qp_rw_kernel<2><<<P.B, 32, sm, s>>>(P);
    }
    return true;
}
