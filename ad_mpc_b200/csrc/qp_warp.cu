// qp_warp.cu -- feedback phase, v4: ONE WARP PER MPC INSTANCE (N <= 31).
//
// Two lane roles alternate inside the interior-point iteration:
//   * stage role  : lane k owns shooting node k (lane N the terminal node).  Its slice of the IPM iterate
//                   (du, dx, pi, lam, t, slacks), the stage's cost gradient / dynamics residual data and the bounds
//                   stay IN REGISTERS for the whole solve; residuals, barrier terms, slack / multiplier steps, step
//                   length and the update are pure register maths + warp shuffles (neighbour nodes, reductions).
//   * matrix role : in the horizon-sequential sweeps the 32 lanes spread over the entries of the 7x8 / 9x9 stage
//                   matrices (P [M|rb], the Gram block M^T P M, gains, Schur complement), operands broadcast from
//                   SHARED MEMORY, where the Riccati working set of the instance lives for the whole solve
//                   (M=[B|A(:,2:7)], rb, K, Guu^-1, P rb, k_ff, barrier diagonals, r_x: 91 doubles per stage).
// No global-memory traffic inside the iteration: HBM is touched once to stage the linearisation in and once to
// write the updated iterate back (fused RTI update).
//
// Algorithm: HPIPM-style Mehrotra predictor-corrector IPM on the OCP-structured QP [EXT], replacing
// FULL_CONDENSING_HPIPM (acados_solver_sim_car.c:145,688-693).  Same maths as qp_ipm.cu; divisions by
// t are done as multiplications by 1/t and the ratio test keeps (num, den) pairs, so results differ from the other
// variants by rounding only.
#include "common.cuh"

#define FULL 0xffffffffu
// unroll factors of the horizon loops of the sweeps (the recursion is serial, but unrolling lets ptxas hoist the next
// stage's operand loads above the current stage's dependent chain): factor sweep / vector sweeps
#ifndef QPW_UB
#define QPW_UB 1
#endif
#ifndef QPW_UVEC
#define QPW_UVEC 2
#endif
#ifndef QPW_UFWD
#define QPW_UFWD 2
#endif
#ifndef QPW_UADJ
#define QPW_UADJ 2
#endif
#ifndef QPW_PADS
#define QPW_PADS 1      // 1: b_k[0..5] and r_k live in the 8 unused doubles of the stage record instead of being re-read from L2 every
                        //    IPM iteration (uncoalesced reads: one sector per lane).  Four more slots for q_k did not pay (3.11 vs 3.10 ms)
#endif
#define QPW_PRAGMA_(x) _Pragma(#x)
#define QPW_UNROLL(n) QPW_PRAGMA_(unroll n)
#define ATS(arr, row) (arr)[(size_t)(row) * Bp + i]      // SoA interface arrays [row][Bp]

// stage record in shared memory (doubles).  Every vector-loaded block starts at an even offset and the stride is
// even, so 16-byte LDS.128 (double2) loads are legal everywhere.
#define R_M 0       // 42: column c (0,1 = u0,u1 ; 2..6 = x2..x6) at c*6 + r, r < 6
#define R_RB 42     // 7   dynamics residual ; corrector roll-out leaves ddx_{k+1} here
#define R_K0 50     // 7   first row of the gain K
#define R_K1 58     // 7   second row
#define R_GI 66     // 3   Guu^-1
#define R_PB 70     // 7   P_{k+1} rb_k ; the adjoint sweep leaves dpi_k here
#define R_KF 78     // 2
#define R_BAR 80    // 5   Rt0 Rt1 Qt6 rt0 rt1
#define R_GX 86     // 7   rgx0..rgx5, qt6 ; corrector roll-out leaves the adjoint base vector here
#define R_DD 94     // 3   ddu0 ddu1 ddx_k[6]
#define R_STRIDE 98
#define R_TERM 8    // terminal record: r_x,N (7) | pad
// scratch after the N stage records and the 8-double terminal record
#define PSS 10      // row stride of P / column stride of W: even (LDS.128) and conflict-free over 7 rows (80 B)
#define X_PS 0      // 72  P_{k+1}, full symmetric, row a at a*PSS
#define X_WS 72     // 80  W = P [M | rb], column c at c*PSS + a
#define X_GS 152    // 32  M^T P M packed lower over the 7 M-indices (+ diagonal terms)
#define X_PV 184    // 8   p_{k+1}
#define X_HV 192    // 8   h = P rb + p
#define X_GV 200    // 12  g vector: [gu0 gu1 gx2..gx6 | gx0 gx1]
#define X_DX 212    // 16  double-buffered broadcast of the roll-out / adjoint state
#define X_SIZE 228

__device__ __forceinline__ constexpr int tri(int a, int b) { return (a >= b) ? (a * (a + 1) / 2 + b) : (b * (b + 1) / 2 + a); }
__device__ __forceinline__ int tri_rt(int a, int b) { return (a >= b) ? (a * (a + 1) / 2 + b) : (b * (b + 1) / 2 + a); }
__device__ __forceinline__ double sel7(const double *a, int idx)
{
    double v = 0.0;
#pragma unroll
    for (int c = 0; c < 7; c++) if (c == idx) v = a[c];
    return v;
}
// 1/x for positive normal x without the out-of-line slow path of the IEEE division (the compiler keeps registers free
// around that call): hardware seed (2^-20) + two Newton steps, <= 1 ulp; NaN propagates, x = 0 gives NaN instead of inf.
__device__ __forceinline__ double rcp_nr(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double wsum(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double nmaxw(double a, double b) { return (a > b || a != a) ? a : b; }   // NaN-propagating
__device__ __forceinline__ double wmax(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = nmaxw(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

struct StageState {            // registers of lane k
    double du[2], dx[7], pi[7], lam[NC], t[NC], sl[2], su[2];
    double lo[2], hi[2], lox, hix;     // bounds around the iterate (delta form)
    double rm[NC], rgu[2], rgx6;
};

// Residual pieces that are cheap functions of the iterate are RECOMPUTED after every sweep instead of being kept live
// in registers across it (the sweeps need those registers for their per-lane operand pointers): rd (bound residuals)
// and the slack stationarity residuals.  The asm barrier keeps the compiler from carrying the values across.
struct StageRes { double rd[NC], rgsl[2], rgsu[2]; };
__device__ __forceinline__ void stage_res(const admpc_opts &o, bool k_ge1, StageState &S, StageRes &R)
{
    const double Ts = o.dt;
#pragma unroll
    for (int c = 0; c < NC; c++) asm volatile("" : "+d"(S.t[c]));
#pragma unroll
    for (int j = 0; j < 2; j++) {
        R.rgsl[j] = Ts * o.zl[j] + Ts * o.Zl[j] * S.sl[j] - S.lam[j] - S.lam[6 + j];
        R.rgsu[j] = Ts * o.zu[j] + Ts * o.Zu[j] * S.su[j] - S.lam[3 + j] - S.lam[8 + j];
        R.rd[j] = S.t[j] - (S.du[j] - S.lo[j] + S.sl[j]);
        R.rd[3 + j] = S.t[3 + j] - (S.hi[j] - S.du[j] + S.su[j]);
        R.rd[6 + j] = S.t[6 + j] - S.sl[j];
        R.rd[8 + j] = S.t[8 + j] - S.su[j];
    }
    if (k_ge1) {
        R.rd[2] = S.t[2] - (S.dx[6] - S.lox);
        R.rd[5] = S.t[5] - (S.hix - S.dx[6]);
    } else {
        R.rd[2] = 0.0; R.rd[5] = 0.0;
    }
}

// barrier-modified Hessian diagonal / gradient of one stage (soft-bound slacks eliminated), it = 1/t
__device__ __forceinline__ void barrier_w(const admpc_opts &o, bool k_ge1, const StageState &S, const StageRes &R,
                                          const double it[NC], double Rt[2], double &Qt6, double rt[2], double &qt6)
{
    const double Ts = o.dt;
    double g[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) g[c] = (S.rm[c] - S.lam[c] * R.rd[c]) * it[c];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const double Sl = S.lam[j] * it[j], Su = S.lam[3 + j] * it[3 + j];
        const double Ssl = S.lam[6 + j] * it[6 + j], Ssu = S.lam[8 + j] * it[8 + j];
        const double iDl = rcp_nr(Ts * o.Zl[j] + Sl + Ssl), iDu = rcp_nr(Ts * o.Zu[j] + Su + Ssu);
        Rt[j] = Ts * o.W[7 + j] + Sl * (1.0 - Sl * iDl) + Su * (1.0 - Su * iDu);
        const double cl = R.rgsl[j] + g[j] + g[6 + j];
        const double cu = R.rgsu[j] + g[3 + j] + g[8 + j];
        rt[j] = S.rgu[j] + (g[j] - Sl * cl * iDl) - (g[3 + j] - Su * cu * iDu);
    }
    if (k_ge1) {
        Qt6 = Ts * o.W[6] + S.lam[2] * it[2] + S.lam[5] * it[5];
        qt6 = S.rgx6 + g[2] - g[5];
    } else {
        Qt6 = Ts * o.W[6];
        qt6 = 0.0;
    }
}

__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
// 7-term dot product a[0..6] . b[0..6] with a, b 16-byte aligned, two accumulation chains, b6 scaled separately
__device__ __forceinline__ double dot6(const double *a, const double *b)
{
    const double2 a0 = ld2(a), a1 = ld2(a + 2), a2 = ld2(a + 4);
    const double2 b0 = ld2(b), b1 = ld2(b + 2), b2 = ld2(b + 4);
    double v = a0.x * b0.x, w = a0.y * b0.y;
    v = fma(a1.x, b1.x, v); w = fma(a1.y, b1.y, w);
    v = fma(a2.x, b2.x, v); w = fma(a2.y, b2.y, w);
    return v + w;
}
// the same with the first operand already in registers (shared by several dot products of the lane)
__device__ __forceinline__ double dot6r(double2 a0, double2 a1, double2 a2, const double *b)
{
    const double2 b0 = ld2(b), b1 = ld2(b + 2), b2 = ld2(b + 4);
    double v = a0.x * b0.x, w = a0.y * b0.y;
    v = fma(a1.x, b1.x, v); w = fma(a1.y, b1.y, w);
    v = fma(a2.x, b2.x, v); w = fma(a2.y, b2.y, w);
    return v + w;
}

// ---- sequential backward sweep (matrix role) -----------------------------------------------------------------------
// Uniform instruction stream: every lane runs the same code on per-lane shared-memory pointers set up once before the
// horizon loop; lanes without a role in a phase shadow a neighbour (same value, same address) or have their store
// predicated off.  Three __syncwarp phases per stage when factorising, two in the vector-only (corrector) sweep.
template <bool FACTOR>
__device__ __forceinline__ void w_backward(const admpc_opts &o, double *sm, double *xs, int N, int l)
{
    asm volatile("" : "+r"(l));      // per-lane role constants are rebuilt per sweep, not kept live across the stage role

    const double Ts = o.dt, hdt = o.dt;
    // phase 1: W = P [M | rb], entry (a, cc) at cc*PSS + a.  Lane l < 28 owns row a = l % 7 of P for the two columns
    // cc0 = l / 7 and cc0 + 4: the row of P is fetched once for both dot products (lanes 28..31 shadow lane 27)
    const int lw = (l < 28) ? l : 27;
    const int cc0 = lw / 7, a0 = lw - cc0 * 7, cc1 = cc0 + 4, a1 = a0;
    const double *p1P0 = xs + X_PS + a0 * PSS;
    const int p1c0 = (cc0 < 7) ? R_M + cc0 * 6 : R_RB, p1c1 = (cc1 < 7) ? R_M + cc1 * 6 : R_RB;
    const double m6c0 = (cc0 == 1) ? hdt : (cc0 == 6) ? 1.0 : 0.0, m6r0 = (cc0 == 7) ? 1.0 : 0.0;
    const double m6c1 = (cc1 == 1) ? hdt : (cc1 == 6) ? 1.0 : 0.0, m6r1 = (cc1 == 7) ? 1.0 : 0.0;
    double *p1o0 = xs + X_WS + cc0 * PSS + a0, *p1o1 = xs + X_WS + cc1 * PSS + a1;
    const bool v71 = (cc1 == 7);
    // phases 2 and 3: packed lower-triangle pair (ta >= tb); lanes 28..31 shadow lane 27
    const int lt = (l < 28) ? l : 27;
    const int ta = (lt >= 21) ? 6 : (lt >= 15) ? 5 : (lt >= 10) ? 4 : (lt >= 6) ? 3 : (lt >= 3) ? 2 : (lt >= 1) ? 1 : 0;
    const int tb = lt - ta * (ta + 1) / 2;
    const int p2M = R_M + ta * 6;
    const double *p2W = xs + X_WS + tb * PSS;
    const double p2m6 = (ta == 1) ? hdt : (ta == 6) ? 1.0 : 0.0;
    const bool dg = (ta == tb);
    const double p2dm = (dg && (ta < 2 || ta == 6)) ? 1.0 : 0.0;                 // diagonal term read from the record
    const double p2dc = (dg && ta >= 2 && ta < 6) ? Ts * sel7(o.W, ta) : 0.0;    // or a constant weight
    const int p2dOff = R_BAR + ((ta == 6) ? 2 : (ta == 1) ? 1 : 0);
    // gradient vector g as a second job of phase 2: entry v < 7 (M-column v) on the lane with (ta, tb) = (v, 0), which
    // already holds M(:, v) in registers for its Gram entry ; v = 7, 8 (x0, x1: no dot product) on lanes 28, 29
    const bool isG = (l < 28) ? (tb == 0) : (l < 30);
    const int gv = (l < 28) ? ta : (l < 30) ? l - 21 : 0;
    const int p3gb = (gv < 2) ? R_BAR + 3 + gv : (gv < 7) ? R_GX + gv : R_GX + (gv - 7);
    const double p3gm = (gv < 7) ? 1.0 : 0.0, p3gh = (gv < 7) ? 0.0 : 1.0;
    const int p3h = (gv < 7) ? 0 : gv - 7;
    // phase 3: gains (lanes 0..13: j = l/7, x = l%7), Schur entry (ta, tb), k_ff (lanes 28, 29), p_k (lanes 0..6)
    const int lk = (l < 14) ? l : 13, kj = lk / 7, kx = lk - 7 * kj;
    const double *p3a0 = (kx < 2) ? xs + X_WS + 0 * PSS + kx : xs + X_GS + tri_rt(kx, 0);
    const double *p3a1 = (kx < 2) ? xs + X_WS + 1 * PSS + kx : xs + X_GS + tri_rt(kx, 1);
    const bool isK = (l < 14);
    const double *p4g = (tb >= 2) ? xs + X_GS + lt : (ta >= 2) ? xs + X_WS + ta * PSS + tb : xs + X_PS + ta * PSS + tb;
    const double p4add = (ta < 2 && dg) ? Ts * sel7(o.W, ta) : 0.0;
    const double *p4a0 = (ta >= 2) ? xs + X_GS + tri_rt(ta, 0) : xs + X_WS + 0 * PSS + ta;   // G[x_ta][u0]
    const double *p4a1 = (ta >= 2) ? xs + X_GS + tri_rt(ta, 1) : xs + X_WS + 1 * PSS + ta;   // G[x_ta][u1]
    const double *p4b0 = (tb >= 2) ? xs + X_GS + tri_rt(tb, 0) : xs + X_WS + 0 * PSS + tb;   // G[u0][x_tb]
    const double *p4b1 = (tb >= 2) ? xs + X_GS + tri_rt(tb, 1) : xs + X_WS + 1 * PSS + tb;   // G[u1][x_tb]
    double *p4o0 = xs + X_PS + ta * PSS + tb, *p4o1 = xs + X_PS + tb * PSS + ta;
    const int l7 = (l < 7) ? l : 6;
    const double *p4gx = xs + X_GV + ((l7 < 2) ? 7 + l7 : l7);
    const bool kfj = (l == 29);

    // terminal: P_N = diag(We), p_N = r_x,N
    if (FACTOR) { xs[X_PS + l] = 0.0; xs[X_PS + 32 + l] = 0.0; if (l < 8) xs[X_PS + 64 + l] = 0.0; }
    if (l < 7) xs[X_PV + l] = sm[N * R_STRIDE + l];
    __syncwarp();
    if (FACTOR && l < 7) xs[X_PS + l * PSS + l] = sel7(o.We, l);
    __syncwarp();
    double *st = sm + (N - 1) * R_STRIDE;
QPW_UNROLL(QPW_UB)
    for (int k = N - 1; k >= 0; k--, st -= R_STRIDE) {
        double gi00, gi01, gi11;
        if (FACTOR) {
            // ---- phase 1: W = P_{k+1} [M | rb] ; P rb ; h = P rb + p ------------------------------------------------------
            const double rb6 = st[R_RB + 6];
            {
                const double2 r01 = ld2(p1P0), r23 = ld2(p1P0 + 2), r45 = ld2(p1P0 + 4);
                const double r6 = p1P0[6];
                const double v = fma(r6, fma(m6r0, rb6, m6c0), dot6r(r01, r23, r45, st + p1c0));
                *p1o0 = v;
                const double w = fma(r6, fma(m6r1, rb6, m6c1), dot6r(r01, r23, r45, st + p1c1));
                *p1o1 = w;
                if (v71) { st[R_PB + a1] = w; xs[X_HV + a1] = w + xs[X_PV + a1]; }
            }
            __syncwarp();
            // ---- phase 2: Gram block G[ta][tb] = M(:,ta)^T W(:,tb) + diagonal ; gradient vector g ---------------------------
            {
                const double2 m01 = ld2(st + p2M), m23 = ld2(st + p2M + 2), m45 = ld2(st + p2M + 4);
                double v = fma(p2m6, p2W[6], dot6r(m01, m23, m45, p2W));
                v += fma(p2dm, st[p2dOff], p2dc);
                xs[X_GS + lt] = v;
                const double d = fma(p2m6, xs[X_HV + 6], dot6r(m01, m23, m45, xs + X_HV));
                const double g = st[p3gb] + fma(p3gm, d, p3gh * xs[X_HV + p3h]);
                if (isG) xs[X_GV + gv] = g;
            }
        }
        __syncwarp();
        // ---- phase 3: 2x2 pivot, gains, Schur complement, k_ff, p_k -----------------------------------------------------------
        const double gu0 = xs[X_GV + 0], gu1 = xs[X_GV + 1];
        if (FACTOR) {
            const double g00 = xs[X_GS + 0] + o.reg, g01 = xs[X_GS + 1], g11 = xs[X_GS + 2] + o.reg;
            const double idet = rcp_nr(g00 * g11 - g01 * g01);
            gi00 = g11 * idet; gi01 = -g01 * idet; gi11 = g00 * idet;
            const double a0v = *p3a0, a1v = *p3a1;       // G[u0][x_kx], G[u1][x_kx] (kx = l on lanes 0..6: reused for p_k below)
            {
                const double c0 = kj ? gi01 : gi00, c1 = kj ? gi11 : gi01;
                const double kval = -(c0 * a0v + c1 * a1v);
                if (isK) st[(kj ? R_K1 : R_K0) + kx] = kval;
            }
            if (l == 31) { st[R_GI + 0] = gi00; st[R_GI + 1] = gi01; st[R_GI + 2] = gi11; }
            {   // P_k[ta][tb] = Gxx + G[x_ta][u] K(:, x_tb)
                const double b0 = *p4b0, b1 = *p4b1;
                const double kb0 = -(gi00 * b0 + gi01 * b1), kb1 = -(gi01 * b0 + gi11 * b1);
                const double pn = (*p4g + p4add) + (*p4a0) * kb0 + (*p4a1) * kb1;
                *p4o0 = pn; *p4o1 = pn;
            }
            {   // p_k[l] = g_x[l] + K(:, x_l) . g_u
                const double k0 = -(gi00 * a0v + gi01 * a1v), k1 = -(gi01 * a0v + gi11 * a1v);
                const double pvv = *p4gx + k0 * gu0 + k1 * gu1;
                if (l < 7) xs[X_PV + l] = pvv;
            }
        }
        {
            const double c0 = kfj ? gi01 : gi00, c1 = kfj ? gi11 : gi01;
            const double kf = -(c0 * gu0 + c1 * gu1);
            if (l == 28 || l == 29) st[R_KF + (l - 28)] = kf;
        }
        __syncwarp();
    }
}

// ---- corrector backward sweep: vector part only, NO shared-memory round trip in the recursion ------------------------
// p_{k+1} is carried in registers (entry c on lane 7, 8, 2..6 for c = 0, 1, 2..6); the owner adds (P rb)[c] and seven
// shuffles hand h = P rb + p to the g-lanes; g_u reaches the p-lanes by two more shuffles.
__device__ __forceinline__ void w_backward_vec(const admpc_opts &o, double *sm, double *xs, int N, int l)
{
    asm volatile("" : "+r"(l));      // per-lane role constants are rebuilt per sweep, not kept live across the stage role

    const double hdt = o.dt;
    const int v = (l < 9) ? l : 0;                              // v < 7: M-column v ; v = 7, 8: x0, x1
    const int gb = (v < 2) ? R_BAR + 3 + v : (v < 7) ? R_GX + v : R_GX + (v - 7);
    const int mOff = R_M + ((v < 7) ? v : 0) * 6;
    const double m6 = (v == 1) ? hdt : (v == 6) ? 1.0 : 0.0;
    const double gm = (v < 7) ? 1.0 : 0.0, gh = (v < 7) ? 0.0 : 1.0;
    const bool h1sel = (v == 8);
    const int sx = (v >= 2 && v < 7) ? v : (v == 8) ? 1 : 0;    // state index of the p entry this lane carries
    const bool kfj = (l == 10);
    const int gio = R_GI + (kfj ? 1 : 0);
    double pown = sm[N * R_STRIDE + sx];                        // p_N = r_x,N
    double *st = sm + (N - 1) * R_STRIDE;
QPW_UNROLL(QPW_UVEC)
    for (int k = N - 1; k >= 0; k--, st -= R_STRIDE) {
        const double hown = st[R_PB + sx] + pown;
        const double h0 = __shfl_sync(FULL, hown, 7), h1 = __shfl_sync(FULL, hown, 8), h2 = __shfl_sync(FULL, hown, 2);
        const double h3 = __shfl_sync(FULL, hown, 3), h4 = __shfl_sync(FULL, hown, 4), h5 = __shfl_sync(FULL, hown, 5);
        const double h6 = __shfl_sync(FULL, hown, 6);
        const double2 m01 = ld2(st + mOff), m23 = ld2(st + mOff + 2), m45 = ld2(st + mOff + 4);
        double d = m6 * h6, d2 = m01.x * h0;
        d = fma(m01.y, h1, d); d2 = fma(m23.x, h2, d2);
        d = fma(m23.y, h3, d); d2 = fma(m45.x, h4, d2);
        d = fma(m45.y, h5, d) + d2;
        const double g = st[gb] + fma(gm, d, gh * (h1sel ? h1 : h0));
        const double gu0 = __shfl_sync(FULL, g, 0), gu1 = __shfl_sync(FULL, g, 1);
        const double c0 = st[gio], c1 = st[gio + 1];             // lane 9: (gi00, gi01) ; lane 10: (gi01, gi11)
        const double kf = -(c0 * gu0 + c1 * gu1);
        if (l == 9 || l == 10) st[R_KF + (l - 9)] = kf;
        pown = g + st[R_K0 + sx] * gu0 + st[R_K1 + sx] * gu1;
    }
    __syncwarp();
}

// ---- sequential forward roll-out (matrix role) -------------------------------------------------------------------------
// One uniform "row . x" stream per stage: lanes 0..5 own row l of M = [B | A(:,2:7)] (ddx_{k+1}[l]), lane 6 the trivial
// delta row, lanes 7 / 8 the gain rows K0 / K1 (ddu_k).  The five state terms c = 2..6 multiply the same x_c on every
// lane; the gain lanes add K(:,0:1) . x_{0:1} + k_ff, two shuffles hand ddu to the state lanes, which add B . ddu.  Each
// lane fetches only ITS row (8 per-lane LDS.64) instead of all of K (the roll-outs were 40 % of the kernel's
// shared-memory wavefronts).  The state is broadcast through a double-buffered 8-double slot (one __syncwarp per stage).
template <bool ADJ>
__device__ __forceinline__ void w_forward(const admpc_opts &o, double *sm, double *xs, int N, int l)
{
    asm volatile("" : "+r"(l));      // per-lane role constants are rebuilt per sweep, not kept live across the stage role

    const double hdt = o.dt, Ts = o.dt;
    const int l7 = (l < 7) ? l : 6;
    const double wq_l = Ts * sel7(o.W, l7), we_l = sel7(o.We, l7);
    const double cself = (l < 2 || l == 6) ? 1.0 : 0.0, cdt = (l == 6) ? hdt : 0.0, mB = (l < 6) ? 1.0 : 0.0;
    const double is6 = (l == 6) ? 1.0 : 0.0;
    const int job = (l < 9) ? l : 8;                       // lanes 9..31 shadow lane 8
    const bool isK = (job >= 7);
    const int rbase = isK ? ((job == 7) ? R_K0 : R_K1) : R_M + ((job < 6) ? job : 5);
    const int rstr = isK ? 1 : 6;                          // element c of the lane's row at rbase + c * rstr
    const double *row = sm + rbase;
    const double *offp = sm + (isK ? R_KF + (job - 7) : R_RB + l7);     // k_ff,j or rb[l]
    const int o1 = rstr, o2 = 2 * rstr, o3 = 3 * rstr, o4 = 4 * rstr, o5 = 5 * rstr, o6 = 6 * rstr;
    double dxr = 0.0;
    double *st = sm;
QPW_UNROLL(QPW_UFWD)
    for (int k = 0; k < N; k++, st += R_STRIDE, row += R_STRIDE, offp += R_STRIDE) {
        double *bx = xs + X_DX + (k & 1) * 8;
        if (l < 8) bx[l] = dxr;                        // lane 7 writes the zero pad
        __syncwarp();
        const double2 x01 = ld2(bx), x23 = ld2(bx + 2), x45 = ld2(bx + 4);
        const double x6 = bx[6];
        // state terms (every lane)
        double ta = row[o2] * x23.x, tb = row[o3] * x23.y;
        ta = fma(row[o4], x45.x, ta); tb = fma(row[o5], x45.y, tb);
        ta = fma(row[o6], x6, ta);
        const double e0 = row[0], e1 = row[o1], off = *offp;
        // gain lanes: ddu_j = K_j . x + k_ff,j
        const double duj = fma(e0, x01.x, off) + fma(e1, x01.y, tb) + ta;
        const double du0 = __shfl_sync(FULL, duj, 7), du1 = __shfl_sync(FULL, duj, 8);
        if (l == 7) { st[R_DD + 0] = du0; st[R_DD + 1] = du1; st[R_DD + 2] = x6; }
        if (ADJ && k >= 1) {
            const double Qd = fma(is6, st[R_BAR + 2] - wq_l, wq_l);
            const double nb = fma(Qd, dxr, st[R_GX + l7]);
            if (l < 7) st[R_GX + l] = nb;
        }
        // state lanes: ddx_{k+1}[l] = rb[l] + (A ddx)[l] + (B ddu)[l]
        const double d = fma(e0, du0, ta) + fma(e1, du1, tb);
        double v = off + fma(cself, dxr, cdt * du1);
        v = fma(mB, d, v);
        dxr = (l < 7) ? v : 0.0;
        if (ADJ && l < 7) st[R_RB + l] = dxr;           // ddx_{k+1}
    }
    if (ADJ && l < 7) sm[N * R_STRIDE + l] = fma(we_l, dxr, sm[N * R_STRIDE + l]);
    __syncwarp();
}

// ---- sequential adjoint sweep: dpi_{k-1} = base_k + A_k^T dpi_k ; leaves dpi_k in the P rb slot ----------------------
// dpi_k stays on lanes 0..6; rows 0..5 reach the M-column lanes by shuffles (the store into the record is off the chain).
__device__ __forceinline__ void w_adjoint(double *sm, double *xs, int N, int l)
{
    asm volatile("" : "+r"(l));      // per-lane role constants are rebuilt per sweep, not kept live across the stage role

    const int l7 = (l < 7) ? l : 6;
    const int lc = (l >= 2 && l < 7) ? l : 2;
    const double cself = (l < 2 || l == 6) ? 1.0 : 0.0, mA = (l >= 2 && l < 7) ? 1.0 : 0.0;
    double dpr = (l < 7) ? sm[N * R_STRIDE + l] : 0.0;       // dpi_{N-1} = We dx_N + r_x,N
    double *st = sm + (N - 1) * R_STRIDE;
QPW_UNROLL(QPW_UADJ)
    for (int k = N - 1; k >= 0; k--, st -= R_STRIDE) {
        if (l < 7) st[R_PB + l] = dpr;
        if (k == 0) break;
        const double q0 = __shfl_sync(FULL, dpr, 0), q1 = __shfl_sync(FULL, dpr, 1), q2 = __shfl_sync(FULL, dpr, 2);
        const double q3 = __shfl_sync(FULL, dpr, 3), q4 = __shfl_sync(FULL, dpr, 4), q5 = __shfl_sync(FULL, dpr, 5);
        const double *mc = st + R_M + lc * 6;
        const double2 a0 = ld2(mc), a1 = ld2(mc + 2), a2 = ld2(mc + 4);
        double d = a0.x * q0, w = a0.y * q1;
        d = fma(a1.x, q2, d); w = fma(a1.y, q3, w);
        d = fma(a2.x, q4, d); w = fma(a2.y, q5, w);
        d += w;
        const double v = st[R_GX + l7] + fma(cself, dpr, mA * d);  // A(:,0..1) = e0,e1 ; A[6][6] = 1
        dpr = (l < 7) ? v : 0.0;
    }
    __syncwarp();
}

// NW = 1: one warp per instance (N <= 31).  NW = 2: two warps per instance (N <= 63): stage k lives in thread k of a
// 64-thread CTA, the horizon-sequential sweeps run on warp 0 only, neighbour exchange and reductions cross the warp
// boundary through a small shared-memory window with CTA barriers.
#define XCH_STRIDE 15   // odd: conflict-free 64-bit accesses over consecutive stages
#define XCH_SIZE (64 * XCH_STRIDE + 48)
template <int NW> __device__ __forceinline__ void cta_sync() { if (NW == 1) __syncwarp(); else __syncthreads(); }

template <int NW>
__global__ void __launch_bounds__(32 * NW, 12 / NW) qp_warp_kernel(const Params P)
{
    extern __shared__ __align__(16) double smw[];
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const int l = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const bool sweeper = (NW == 1) || wrp == 0;
    const int i = blockIdx.x;                    // one instance per CTA
    double *sm = smw;
    double *xs = smw + N * R_STRIDE + R_TERM;
    double *xc = xs + X_SIZE;                    // NW == 2 only: exchange window [64][XCH_STRIDE] + 48 reduction slots
    double *red = xc + 64 * XCH_STRIDE;
    const double Ts = o.dt, hdt = o.dt;
    if (const int flag = P.lin_bad[i]) {         // 1: NaN/Inf in the linearisation: ACADOS_FAILURE, iterate untouched
        if (threadIdx.x == 0 && flag == 1) { P.status[i] = 1; P.qp_status[i] = 0; P.qp_iter[i] = 0; }
        if (P.gat_x) {                           // fused gather: the (untouched) iterate still goes to the root's block
            const int kk = threadIdx.x;
            if (kk <= N) {
                for (int a = 0; a < 7; a++) P.gat_x[((size_t)i * (N + 1) + kk) * 7 + a] = ATS(P.xb, kk * 7 + a);
                if (kk < N) for (int j = 0; j < 2; j++) P.gat_u[((size_t)i * N + kk) * 2 + j] = ATS(P.ub, kk * 2 + j);
            }
            if (kk == 0) P.gat_st[i] = (flag == 1) ? 1 : P.status[i];
        }
        return;                                  // 2: finished instance of the full-SQP loop, nothing to do
    }
    // ---- stage M into shared memory ------------------------------------------------------------------------------------
    // entry w = cc*6 + r of a stage (cc < 2: B(r,cc), else A(r,cc)) on lane w and lane w - 32; the source row of each lane
    // is fixed, the loop over the stages is two loads + two stores with pointer increments (all loads in flight at once)
    {
        const int w0 = l, w1 = (l + 32 < 42) ? l + 32 : 41;
        const int c0 = w0 / 6, r0 = w0 - c0 * 6, c1 = w1 / 6, r1 = w1 - c1 * 6;
        const int s0 = (c0 < 2) ? LIN_B + r0 * 2 + c0 : LIN_A + r0 * 5 + (c0 - 2);
        const int s1 = (c1 < 2) ? LIN_B + r1 * 2 + c1 : LIN_A + r1 * 5 + (c1 - 2);
        const double *g0 = P.lin + (size_t)s0 * Bp + i, *g1 = P.lin + (size_t)s1 * Bp + i;
        const size_t gstep = (size_t)LIN_ROWS * Bp;
        for (int k = wrp; k < N; k += NW) {
            const double v0 = g0[(size_t)k * gstep], v1 = g1[(size_t)k * gstep];
            sm[k * R_STRIDE + R_M + w0] = v0;
            if (l < 10) sm[k * R_STRIDE + R_M + w1] = v1;
        }
    }
    // ---- stage role: load this node's data, cold start ---------------------------------------------------------------------
    const int k = threadIdx.x;
    const bool isst = k < N, isterm = (k == N);
    StageState S;
#pragma unroll
    for (int a = 0; a < 7; a++) { S.dx[a] = 0.0; S.pi[a] = 0.0; }
#pragma unroll
    for (int c = 0; c < NC; c++) { S.lam[c] = 0.0; S.t[c] = 1.0; S.rm[c] = 0.0; }
#pragma unroll
    for (int j = 0; j < 2; j++) { S.du[j] = 0.0; S.sl[j] = 0.0; S.su[j] = 0.0; S.lo[j] = -1.0; S.hi[j] = 1.0; S.rgu[j] = 0.0; }
    S.lox = -1.0; S.hix = 1.0; S.rgx6 = 0.0;
    if (isst || isterm) {
        if (isst) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const double cur = ATS(P.ub, k * 2 + j);
                S.lo[j] = o.lbu[j] - cur; S.hi[j] = o.ubu[j] - cur;
            }
            const double cur6 = ATS(P.xb, k * 7 + 6);
            S.lox = o.lbx - cur6; S.hix = o.ubx - cur6;
        }
        if (k == 0) {
#pragma unroll
            for (int a = 0; a < 7; a++) S.dx[a] = ATS(P.x0, a) - ATS(P.xb, a);     // x0 eliminated (nbxe_0 = 7)
        }
    }
#if QPW_PADS
    if (isst) {     // record pads 49, 57, 65, 69, 77, 85, 93, 97 (never written by the sweeps): b_k[0..5], r_k
        const double *lin = P.lin + (size_t)k * LIN_ROWS * Bp;
        double *stp = sm + k * R_STRIDE;
        stp[49] = ATS(lin, LIN_b + 0); stp[57] = ATS(lin, LIN_b + 1); stp[65] = ATS(lin, LIN_b + 2);
        stp[69] = ATS(lin, LIN_b + 3); stp[77] = ATS(lin, LIN_b + 4); stp[85] = ATS(lin, LIN_b + 5);
        stp[93] = ATS(lin, LIN_r + 0); stp[97] = ATS(lin, LIN_r + 1);
    }
#endif
    if (isst) {
        // cold start: primal 0 pushed thr0 inside its box, t from the box, lam = mu0 / t
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (j == 2 && k == 0) continue;
            const double lo = (j < 2) ? S.lo[j] : S.lox, hi = (j < 2) ? S.hi[j] : S.hix;
            double v = 0.0;
            if (v - lo < o.thr0) {
                if (hi - v < o.thr0) v = 0.5 * (lo + hi);
                else v = lo + o.thr0;
            } else if (hi - v < o.thr0) v = hi - o.thr0;
            if (j < 2) S.du[j] = v; else S.dx[6] = v;
            const double tl = fmax(o.thr0, v - lo), tu = fmax(o.thr0, hi - v);
            S.t[j] = tl; S.t[3 + j] = tu;
            S.lam[j] = o.mu0 / tl; S.lam[3 + j] = o.mu0 / tu;
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            S.t[6 + j] = o.thr0; S.t[8 + j] = o.thr0;
            S.lam[6 + j] = o.mu0 / o.thr0; S.lam[8 + j] = o.mu0 / o.thr0;
        }
    }
    cta_sync<NW>();

    const double inv_nc = 1.0 / (double)(NC * N - 2);
    int status = 1, iter = 0;
    double res0 = 0, res1 = 0, res2 = 0, res3 = 0;
    for (iter = 0;; iter++) {
        // ================= residuals of the current point (stage role) ===================================================
        double pim[7], dxn[7], it[NC];
        // q_k and b_k[6] are re-read each iteration (L2-resident) instead of occupying registers for the whole solve;
        // b_k[0..5] and r_k sit in the unused doubles of the stage record
        double lb[7], lq[7], lr[2];
        {
            const double *lin = P.lin + (size_t)((isst || isterm) ? k : 0) * LIN_ROWS * Bp;
#if QPW_PADS
            const double *stp = sm + (isst ? k : 0) * R_STRIDE;
#pragma unroll
            for (int a = 0; a < 7; a++) lq[a] = ATS(lin, LIN_q + a);
            lb[0] = stp[49]; lb[1] = stp[57]; lb[2] = stp[65]; lb[3] = stp[69]; lb[4] = stp[77]; lb[5] = stp[85];
            lb[6] = ATS(lin, LIN_b + 6);
            lr[0] = stp[93]; lr[1] = stp[97];
#else
#pragma unroll
            for (int a = 0; a < 7; a++) { lq[a] = ATS(lin, LIN_q + a); lb[a] = ATS(lin, LIN_b + a); }
            lr[0] = ATS(lin, LIN_r + 0); lr[1] = ATS(lin, LIN_r + 1);
#endif
        }
        if (NW == 1) {
#pragma unroll
            for (int a = 0; a < 7; a++) {
                const double up = __shfl_up_sync(FULL, S.pi[a], 1);
                pim[a] = (k >= 1) ? up : 0.0;
                dxn[a] = __shfl_down_sync(FULL, S.dx[a], 1);
            }
        } else {
#pragma unroll
            for (int a = 0; a < 7; a++) { xc[k * XCH_STRIDE + a] = S.pi[a]; xc[k * XCH_STRIDE + 7 + a] = S.dx[a]; }
            __syncthreads();
            const int km = (k >= 1) ? k - 1 : 0, kp = (k < 63) ? k + 1 : 63;
#pragma unroll
            for (int a = 0; a < 7; a++) {
                const double up = xc[km * XCH_STRIDE + a];
                pim[a] = (k >= 1) ? up : 0.0;
                dxn[a] = xc[kp * XCH_STRIDE + 7 + a];
            }
        }
        double ng = 0, nb = 0, nd = 0, nm = 0, summ = 0;
        double *st = sm + (isst ? k : 0) * R_STRIDE;
        if (isst) {
#pragma unroll
            for (int c = 0; c < NC; c++) it[c] = rcp_nr(S.t[c]);
            StageRes R;
            stage_res(o, k >= 1, S, R);
#pragma unroll
            for (int j = 0; j < 2; j++) {
                double g = Ts * o.W[7 + j] * S.du[j] + lr[j] - S.lam[j] + S.lam[3 + j];
#pragma unroll
                for (int r = 0; r < 6; r++) g = fma(st[R_M + j * 6 + r], S.pi[r], g);
                if (j == 1) g = fma(hdt, S.pi[6], g);
                S.rgu[j] = g;
                ng = nmaxw(ng, nmaxw(fabs(g), nmaxw(fabs(R.rgsl[j]), fabs(R.rgsu[j]))));
                nd = nmaxw(nd, nmaxw(nmaxw(fabs(R.rd[j]), fabs(R.rd[3 + j])), nmaxw(fabs(R.rd[6 + j]), fabs(R.rd[8 + j]))));
            }
            if (k >= 1) nd = nmaxw(nd, nmaxw(fabs(R.rd[2]), fabs(R.rd[5])));
#pragma unroll
            for (int r = 0; r < 6; r++) {
                double v = lb[r] - dxn[r] + ((r < 2) ? S.dx[r] : 0.0);
                v = fma(st[R_M + 0 * 6 + r], S.du[0], v);
                v = fma(st[R_M + 1 * 6 + r], S.du[1], v);
#pragma unroll
                for (int cc = 0; cc < 5; cc++) v = fma(st[R_M + (2 + cc) * 6 + r], S.dx[2 + cc], v);
                st[R_RB + r] = v;
                nb = nmaxw(nb, fabs(v));
            }
            {
                const double v = lb[6] - dxn[6] + S.dx[6] + hdt * S.du[1];
                st[R_RB + 6] = v;
                nb = nmaxw(nb, fabs(v));
            }
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const bool on = !((c == 2 || c == 5) && k == 0);
                S.rm[c] = on ? S.lam[c] * S.t[c] : 0.0;
                nm = nmaxw(nm, fabs(S.rm[c]));
                summ += S.rm[c];
            }
            S.rgx6 = 0.0;
            if (k >= 1) {
#pragma unroll
                for (int a = 0; a < 7; a++) {
                    double g = Ts * o.W[a] * S.dx[a] + lq[a] - pim[a];
                    if (a < 2) g += S.pi[a];
                    else {
#pragma unroll
                        for (int r = 0; r < 6; r++) g = fma(st[R_M + a * 6 + r], S.pi[r], g);
                        if (a == 6) g += S.pi[6] - S.lam[2] + S.lam[5];
                    }
                    if (a < 6) st[R_GX + a] = g; else S.rgx6 = g;
                    ng = nmaxw(ng, fabs(g));
                }
            } else {
#pragma unroll
                for (int a = 0; a < 6; a++) st[R_GX + a] = 0.0;
            }
            double Rt[2], Qt6, rt[2], qt6;
            barrier_w(o, k >= 1, S, R, it, Rt, Qt6, rt, qt6);
            st[R_BAR + 0] = Rt[0]; st[R_BAR + 1] = Rt[1]; st[R_BAR + 2] = Qt6;
            st[R_BAR + 3] = rt[0]; st[R_BAR + 4] = rt[1]; st[R_GX + 6] = qt6;
        } else if (isterm) {
#pragma unroll
            for (int a = 0; a < 7; a++) {
                const double g = o.We[a] * S.dx[a] + lq[a] - pim[a];
                sm[N * R_STRIDE + a] = g;
                ng = nmaxw(ng, fabs(g));
            }
        }
        ng = wmax(ng); nb = wmax(nb); nd = wmax(nd); nm = wmax(nm); summ = wsum(summ);
        if (NW == 2) {
            if (l == 0) { double *r = red + wrp * 8; r[0] = ng; r[1] = nb; r[2] = nd; r[3] = nm; r[4] = summ; }
            __syncthreads();
            ng = nmaxw(red[0], red[8]); nb = nmaxw(red[1], red[9]); nd = nmaxw(red[2], red[10]);
            nm = nmaxw(red[3], red[11]); summ = red[4] + red[12];
        }
        res0 = ng; res1 = nb; res2 = nd; res3 = nm;
        if (!(isfinite(ng) && isfinite(nb) && isfinite(nd) && isfinite(nm))) { status = 3; break; }
        if (ng < o.tol_stat && nb < o.tol_eq && nd < o.tol_ineq && nm < o.tol_comp) { status = 0; break; }
        if (iter >= o.iter_max) { status = 1; break; }
        const double mu = summ * inv_nc;
        cta_sync<NW>();

        // ================= predictor ===================================================================================
        if (sweeper) {
            w_backward<true>(o, sm, xs, N, l);
            w_forward<false>(o, sm, xs, N, l);
        }
        if (NW == 2) __syncthreads();
        double dsl[2], dsu[2], dtv[NC], dlv[NC];
        double an = 1.0, ad = 1.0;          // step length as a ratio an/ad (<= 1)
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
            // pass 0: affine step -> mu_aff, sigma, corrected rhs ; pass 1: final step -> alpha
            an = 1.0; ad = 1.0; s1 = 0.0; s2 = 0.0;
            StageRes R;
            if (isst) {
                stage_res(o, k >= 1, S, R);
                const double du0 = st[R_DD + 0], du1 = st[R_DD + 1], dx6 = st[R_DD + 2];
                double gq[NC];
#pragma unroll
                for (int c = 0; c < NC; c++) gq[c] = (S.rm[c] - S.lam[c] * R.rd[c]) * it[c];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const double duj = (j == 0) ? du0 : du1;
                    const double Sl = S.lam[j] * it[j], Su = S.lam[3 + j] * it[3 + j];
                    const double Ssl = S.lam[6 + j] * it[6 + j], Ssu = S.lam[8 + j] * it[8 + j];
                    const double iDl = rcp_nr(Ts * o.Zl[j] + Sl + Ssl), iDu = rcp_nr(Ts * o.Zu[j] + Su + Ssu);
                    const double cl = R.rgsl[j] + gq[j] + gq[6 + j];
                    const double cu = R.rgsu[j] + gq[3 + j] + gq[8 + j];
                    dsl[j] = -(cl + Sl * duj) * iDl;
                    dsu[j] = -(cu - Su * duj) * iDu;
                    dtv[j] = duj + dsl[j] - R.rd[j];
                    dtv[3 + j] = -duj + dsu[j] - R.rd[3 + j];
                    dtv[6 + j] = dsl[j] - R.rd[6 + j];
                    dtv[8 + j] = dsu[j] - R.rd[8 + j];
                }
                if (k >= 1) { dtv[2] = dx6 - R.rd[2]; dtv[5] = -dx6 - R.rd[5]; }
                else { dtv[2] = 0.0; dtv[5] = 0.0; }
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    const bool on = !((c == 2 || c == 5) && k == 0);
                    dlv[c] = on ? -(S.rm[c] + S.lam[c] * dtv[c]) * it[c] : 0.0;
                    if (on) {
                        // ratio test without divisions: keep the smallest lam/(-dlam), t/(-dt) as a pair
                        if (dlv[c] < 0.0 && S.lam[c] * ad < an * (-dlv[c])) { an = S.lam[c]; ad = -dlv[c]; }
                        if (dtv[c] < 0.0 && S.t[c] * ad < an * (-dtv[c])) { an = S.t[c]; ad = -dtv[c]; }
                        s1 += S.lam[c] * dtv[c] + S.t[c] * dlv[c];
                        s2 += dlv[c] * dtv[c];
                    }
                }
            }
#pragma unroll
            for (int off = 16; off; off >>= 1) {
                const double bn = __shfl_xor_sync(FULL, an, off), bd = __shfl_xor_sync(FULL, ad, off);
                if (bn * ad < an * bd) { an = bn; ad = bd; }
            }
            an = __shfl_sync(FULL, an, 0); ad = __shfl_sync(FULL, ad, 0);     // one representative pair for all lanes
            if (pass == 0) { s1 = wsum(s1); s2 = wsum(s2); }
            if (NW == 2) {
                double *r = red + 16 + pass * 16;
                if (l == 0) { r[wrp * 4 + 0] = an; r[wrp * 4 + 1] = ad; r[wrp * 4 + 2] = s1; r[wrp * 4 + 3] = s2; }
                __syncthreads();
                an = r[0]; ad = r[1];
                if (r[4] * ad < an * r[5]) { an = r[4]; ad = r[5]; }
                s1 = r[2] + r[6]; s2 = r[3] + r[7];
            }
            if (pass == 0) {
                const double a_aff = an * rcp_nr(ad);
                const double mu_aff = (summ + a_aff * s1 + a_aff * a_aff * s2) * inv_nc;
                double sigma = mu_aff * rcp_nr(mu);
                sigma = sigma * sigma * sigma;
                const double sigmu = sigma * mu;
                if (isst) {
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        const bool on = !((c == 2 || c == 5) && k == 0);
                        S.rm[c] = on ? S.rm[c] + dlv[c] * dtv[c] - sigmu : 0.0;
                    }
                    double Rt[2], Qt6, rt[2], qt6;
                    barrier_w(o, k >= 1, S, R, it, Rt, Qt6, rt, qt6);
                    st[R_BAR + 3] = rt[0]; st[R_BAR + 4] = rt[1]; st[R_GX + 6] = qt6;
                }
                cta_sync<NW>();
                // ================= corrector ===========================================================================
                if (sweeper) {
                    w_backward_vec(o, sm, xs, N, l);
                    w_forward<true>(o, sm, xs, N, l);
                }
                if (NW == 2) __syncthreads();
            }
        }
        double alpha = an * rcp_nr(ad);
        if (alpha < o.alpha_min) { status = 2; break; }
        if (alpha < 1.0) alpha *= 0.995;
        // ================= update (stage role); pi and dx wait for the adjoint sweep ==========================================
        if (isst) {
            const double du0 = st[R_DD + 0], du1 = st[R_DD + 1];
            S.du[0] += alpha * du0; S.du[1] += alpha * du1;
#pragma unroll
            for (int j = 0; j < 2; j++) { S.sl[j] += alpha * dsl[j]; S.su[j] += alpha * dsu[j]; }
#pragma unroll
            for (int c = 0; c < NC; c++) {
                if ((c == 2 || c == 5) && k == 0) continue;
                S.lam[c] = fmax(S.lam[c] + alpha * dlv[c], o.lam_min);
                S.t[c] = fmax(S.t[c] + alpha * dtv[c], o.t_min);
            }
        }
        if (sweeper) w_adjoint(sm, xs, N, l);
        if (NW == 2) __syncthreads();
        if (isst) {
#pragma unroll
            for (int a = 0; a < 7; a++) S.pi[a] += alpha * st[R_PB + a];
        }
        if ((isst || isterm) && k >= 1) {
            const double *prev = sm + (k - 1) * R_STRIDE;      // ddx_k was left in record k-1
#pragma unroll
            for (int a = 0; a < 7; a++) S.dx[a] += alpha * prev[R_RB + a];
        }
        cta_sync<NW>();
    }
    // ---- epilogue: statuses + fused RTI update (full step; duals <- QP duals) --------------------------------------------
    const int qps = (status == 0) ? 0 : ((status == 1) ? 2 : ((status == 2) ? 3 : 1));   // hpipm -> acados numbering
    const int nlp_status = (qps == 0 || qps == 2) ? 0 : 4;
    if (threadIdx.x == 0) {
        P.qp_status[i] = qps; P.qp_iter[i] = iter; P.status[i] = nlp_status;
        ATS(P.res_out, 0) = res0; ATS(P.res_out, 1) = res1; ATS(P.res_out, 2) = res2; ATS(P.res_out, 3) = res3;
    }
    // fused RTI update + fused solution gather (multi-GPU): the new [u | x | status] of the instance also goes,
    // instance-major, to this rank's slice of the root's block -- peer memory over NVLink when the root is another GPU
    const bool upd = (nlp_status == 0);
    if ((upd || P.gat_x) && (isst || isterm)) {
#pragma unroll
        for (int a = 0; a < 7; a++) {
            double v = ATS(P.xb, k * 7 + a);
            if (upd) { v += S.dx[a]; ATS(P.xb, k * 7 + a) = v; }
            if (P.gat_x) P.gat_x[((size_t)i * (N + 1) + k) * 7 + a] = v;
        }
        if (isst) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                double v = ATS(P.ub, k * 2 + j);
                if (upd) { v += S.du[j]; ATS(P.ub, k * 2 + j) = v; }
                if (P.gat_x) P.gat_u[((size_t)i * N + k) * 2 + j] = v;
            }
        }
        if (k == 0 && P.gat_x) P.gat_st[i] = nlp_status;
    }
    if (upd && (isst || isterm)) {
        if (isst) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                ATS(P.slb, k * 2 + j) = S.sl[j];
                ATS(P.sub, k * 2 + j) = S.su[j];
            }
#pragma unroll
            for (int a = 0; a < 7; a++) ATS(P.pib, k * 7 + a) = S.pi[a];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const bool on = !((c == 2 || c == 5) && k == 0);
                ATS(P.lamb, k * NC + c) = on ? S.lam[c] : 0.0;
                ATS(P.tb, k * NC + c) = on ? S.t[c] : 1.0;
            }
        }
    }
}

bool launch_qp_warp(const Params &P, cudaStream_t s)
{
    const int N = P.o.N;
    if (N > 63) return false;
    if (N <= 31) {
        const size_t sm = (size_t)(N * R_STRIDE + R_TERM + X_SIZE) * sizeof(double);
        static SmemGuard configured;
        if (configured.need(sm)) {
            cudaFuncSetAttribute(qp_warp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        }
        qp_warp_kernel<1><<<P.B, 32, sm, s>>>(P);
    } else {
        const size_t sm = (size_t)(N * R_STRIDE + R_TERM + X_SIZE + XCH_SIZE) * sizeof(double);
        static SmemGuard configured2;
        if (configured2.need(sm)) {
            cudaFuncSetAttribute(qp_warp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        }
        qp_warp_kernel<2><<<P.B, 64, sm, s>>>(P);
    }
    return true;
}
