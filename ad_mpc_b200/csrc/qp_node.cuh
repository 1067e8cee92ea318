// qp_node.cuh -- node-role arithmetic of the warp-per-instance IPM kernel (qp_mma.cu): constraint data of one shooting
// node streamed from its shared-memory record, residuals, barrier scalings, slack / t / lambda steps, division-free ratio test,
// warp reductions.  Include AFTER the record layout (W_* offsets) of the kernel has been defined.
#pragma once
#include "common.cuh"

#define FULL 0xffffffffu
#define ATS(arr, row) (arr)[(size_t)(row) * Bp + i]      // SoA interface arrays [row][Bp]

__device__ __forceinline__ double sel7w(const double *a, int idx)
{
    double v = 0.0;
#pragma unroll
    for (int c = 0; c < 7; c++) if (c == idx) v = a[c];
    return v;
}
// 1/x for positive normal x: hardware seed + two Newton steps (<= 1 ulp), no out-of-line slow path
__device__ __forceinline__ double rcp_w(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
// NaN-propagating maximum.  NMX_INT (qp_mma.cu, qp_mma_g.cu): every use is nmx(running maximum, fabs(..)) or a reduction of
// such, i.e. NON-NEGATIVE values; for sign bit 0 the IEEE order is the order of the bit patterns and a NaN is larger than every
// finite pattern, so an integer maximum does it off the FP64 pipe (a DSETP pair per maximum otherwise).  Measured: qp_mma_g<1,1>
// 1.777 -> 1.727 ms; qp_mma<1> 1.500 -> 1.519 ms on the kernel with two separate passes (it started to spill), 1.506 -> 1.484 ms
// together with the rolled two-pass loop (250 registers, no spills).
#ifdef NMX_INT
__device__ __forceinline__ double nmx(double a, double b)
{
    const long long ia = __double_as_longlong(a), ib = __double_as_longlong(b);
    return __longlong_as_double(ia > ib ? ia : ib);
}
#else
__device__ __forceinline__ double nmx(double a, double b) { return (a > b || a != a) ? a : b; }
#endif
__device__ __forceinline__ double wsum32(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double wmax32(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = nmx(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double2 ldv(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void stv(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }
__device__ __forceinline__ double dot6v(double2 a0, double2 a1, double2 a2, double2 b0, double2 b1, double2 b2)
{
    double v = a0.x * b0.x, w = a0.y * b0.y;
    v = fma(a1.x, b1.x, v); w = fma(a1.y, b1.y, w);
    v = fma(a2.x, b2.x, v); w = fma(a2.y, b2.y, w);
    return v + w;
}

// ---- constraint data of one node, streamed from its record ---------------------------------------------------------------
struct NCon {
    double lam[NC], t[NC], sl[2], su[2], du[2], dx6;
    double lo[2], hi[2], lox, hix;
};
#ifdef W_XB      // (kernels that keep the linearisation point in the record; qp_mma_g.cu has its own descriptor-based loader)
__device__ __forceinline__ void load_ncon(const admpc_opts &o, const double *st, NCon &C)
{
    const double ub0 = st[W_UB], ub1 = st[W_UB + 1], xb6 = st[W_XB + 6];
#pragma unroll
    for (int c = 0; c < NC; c += 2) {
        const double2 l = ldv(st + W_LAM + c), t = ldv(st + W_T + c);
        C.lam[c] = l.x; C.lam[c + 1] = l.y; C.t[c] = t.x; C.t[c + 1] = t.y;
    }
    const double2 du = ldv(st + W_DU), sl = ldv(st + W_SL), su = ldv(st + W_SU);
    C.du[0] = du.x; C.du[1] = du.y; C.sl[0] = sl.x; C.sl[1] = sl.y; C.su[0] = su.x; C.su[1] = su.y;
    C.dx6 = st[W_DX + 6];
    C.lo[0] = o.lbu[0] - ub0; C.hi[0] = o.ubu[0] - ub0;
    C.lo[1] = o.lbu[1] - ub1; C.hi[1] = o.ubu[1] - ub1;
    C.lox = o.lbx - xb6; C.hix = o.ubx - xb6;
}
#endif
struct NRes { double rd[NC], rgsl[2], rgsu[2]; };
__device__ __forceinline__ void node_res_w(const admpc_opts &o, bool k_ge1, const NCon &C, NRes &R)
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
        R.rd[5] = C.t[5] - (C.hix - C.dx6);
    } else {
        R.rd[2] = 0.0; R.rd[5] = 0.0;
    }
}
// 1/t, the barrier scalings and the slack-elimination pivots
struct NScal { double it[NC], Sl[2], Su[2], iDl[2], iDu[2]; };
__device__ __forceinline__ void node_scal_w(const admpc_opts &o, const NCon &C, NScal &S)
{
    const double Ts = o.dt;
#pragma unroll
    for (int c = 0; c < NC; c++) S.it[c] = rcp_w(C.t[c]);
#pragma unroll
    for (int j = 0; j < 2; j++) {
        S.Sl[j] = C.lam[j] * S.it[j]; S.Su[j] = C.lam[3 + j] * S.it[3 + j];
        const double Ssl = C.lam[6 + j] * S.it[6 + j], Ssu = C.lam[8 + j] * S.it[8 + j];
        S.iDl[j] = rcp_w(Ts * o.Zl[j] + S.Sl[j] + Ssl);
        S.iDu[j] = rcp_w(Ts * o.Zu[j] + S.Su[j] + Ssu);
    }
}
// slack / t / lambda steps of one node for a given primal step (du, dx6) and complementarity right-hand side rm
struct NStep { double dsl[2], dsu[2], dtv[NC], dlv[NC]; };
__device__ __forceinline__ void node_step_w(bool k_ge1, const NCon &C, const NRes &R, const NScal &S, const double rm[NC],
                                            double du0, double du1, double dx6, NStep &D)
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
__device__ __forceinline__ void node_ratio_w(bool k_ge1, const NCon &C, const NStep &D, double &an, double &ad)
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
// Ratio test through the reciprocals the barrier terms already hold: t / (-dt) = 1 / (-dt it), so the t rows contribute
// max_c(-dt_c it_c) to m and the step bound is 1 / max(1, m): one multiply and one max per row, no serial (num, den) chain.
__device__ __forceinline__ double node_ratio_t(bool k_ge1, const NScal &S, const NStep &D, double m)
{
    double w[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const bool on = !((c == 2 || c == 5) && !k_ge1);
        w[c] = on ? -D.dtv[c] * S.it[c] : 0.0;
    }
    const double a = fmax(fmax(w[0], w[1]), fmax(w[2], w[3])), b = fmax(fmax(w[4], w[5]), fmax(w[6], w[7]));
    return fmax(m, fmax(fmax(a, b), fmax(w[8], w[9])));
}
// Affine step only (rm = lam t): dlam = -lam (1 + dt / t), so lam / (-dlam) = 1 / (1 + dt it) and both halves of the test come
// from w_c = dt_c it_c: m = max_c max(-w_c, 1 + w_c).
__device__ __forceinline__ double node_ratio_aff(bool k_ge1, const NScal &S, const NStep &D, double m)
{
    double w[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const bool on = !((c == 2 || c == 5) && !k_ge1);
        const double q = D.dtv[c] * S.it[c];
        w[c] = on ? fmax(-q, 1.0 + q) : 0.0;
    }
    const double a = fmax(fmax(w[0], w[1]), fmax(w[2], w[3])), b = fmax(fmax(w[4], w[5]), fmax(w[6], w[7]));
    return fmax(m, fmax(fmax(a, b), fmax(w[8], w[9])));
}
// lam rows of the final step: division-free (num, den) pair as in node_ratio_w
__device__ __forceinline__ void node_ratio_lam(bool k_ge1, const NCon &C, const NStep &D, double &an, double &ad)
{
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const bool on = !((c == 2 || c == 5) && !k_ge1);
        if (on && D.dlv[c] < 0.0 && C.lam[c] * ad < an * (-D.dlv[c])) { an = C.lam[c]; ad = -D.dlv[c]; }
    }
}
// Warp maximum / minimum of POSITIVE doubles in two REDUX instructions instead of five shuffle rounds: for sign bit 0 the order of
// the values is the order of the bit patterns, so reduce the high words, then the low words of the lanes that hold the winning
// high word.  (A NaN has the largest pattern: it wins a maximum and loses a minimum.)
__device__ __forceinline__ double wmax_pos(double v)
{
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(FULL, hi);
    const unsigned ml = __reduce_max_sync(FULL, (hi == mh) ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}
__device__ __forceinline__ double wmin_pos(double v)
{
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_min_sync(FULL, hi);
    const unsigned ml = __reduce_min_sync(FULL, (hi == mh) ? lo : 0xffffffffu);
    return __hiloint2double((int)mh, (int)ml);
}
__device__ __forceinline__ double wmaxf32(double v)       // plain max (no NaN propagation needed: NaNs are caught by the residual norms)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ void warp_ratio(double &an, double &ad)
{
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        const double bn = __shfl_xor_sync(FULL, an, off), bd = __shfl_xor_sync(FULL, ad, off);
        if (bn * ad < an * bd) { an = bn; ad = bd; }
    }
    an = __shfl_sync(FULL, an, 0); ad = __shfl_sync(FULL, ad, 0);     // one representative pair for all lanes
}
