// refgen.cu -- batched reference generation: the step immediately before the solve.
//
// Replaces, for a whole batch, `RefTrajectory.get_waypoints` (reference: ad_mpc/ref_traj.py:89-171) plus the glue that
// turns its output into the solver's yref: nodes/gp_ad_mpc_node.py:180-187 (ref = [x,y,psi,v,0,0,0], u_ref = 0),
// ad_mpc/ad_3d_optimizer.py:343-345 (pad to N+1 by repeating the last row) and :420-438 (heading unwrap against the
// current heading).  The interpolation abscissae of get_waypoints depend only on the track (ref_traj.py:128-131), so the
// arc-length tables are built once per track on the host (set_track); per instance the kernel does the closest-point
// search, the Frenet errors, the heading fix relative to the vehicle heading, the 3-point blend from the current pose
// and writes yref straight into the solver's SoA layout -- no host round trip, 8 B x 3 of input per instance.
// HBM-bound: (9N+7) doubles written per instance.
#include <math.h>
#include <vector>

#include "common.cuh"

#define PI_D 3.141592653589793

// numpy / Python float modulo: result has the sign of the divisor
__host__ __device__ static inline double pymod(double a, double b)
{
    double m = fmod(a, b);
    if (m != 0.0) { if ((b < 0.0) != (m < 0.0)) m += b; }
    else m = copysign(0.0, b);
    return m;
}
__host__ __device__ static inline double bound_pi(double a) { return pymod(a + PI_D, 2.0 * PI_D) - PI_D; }   // ref_traj.py:27-28

// one step of numpy.unwrap (period 2 pi): correction to add to the running sum
__host__ __device__ static inline double unwrap_corr(double dd)
{
    double ddmod = pymod(dd + PI_D, 2.0 * PI_D) - PI_D;
    if (ddmod == -PI_D && dd > 0.0) ddmod = PI_D;
    return (fabs(dd) < PI_D) ? 0.0 : ddmod - dd;
}

// numpy.interp for one abscissa (compiled_base.c arr_interp, monotone xp)
static double np_interp1(double x, const std::vector<double> &xp, const std::vector<double> &fp)
{
    const int n = (int)xp.size();
    if (x < xp[0]) return fp[0];
    if (x >= xp[n - 1]) return fp[n - 1];
    int lo = 0, hi = n - 1;               // xp[lo] <= x < xp[hi]
    while (hi - lo > 1) { const int mid = (lo + hi) / 2; if (xp[mid] <= x) lo = mid; else hi = mid; }
    if (xp[lo] == x) return fp[lo];
    const double slope = (fp[lo + 1] - fp[lo]) / (xp[lo + 1] - xp[lo]);
    double r = slope * (x - xp[lo]) + fp[lo];
    if (isnan(r)) { r = slope * (x - xp[lo + 1]) + fp[lo + 1]; if (isnan(r) && fp[lo] == fp[lo + 1]) r = fp[lo]; }
    return r;
}

struct TrackHost {
    std::vector<double> dev;      // packed: x[L] y[L] psi[L] cdist[L] psi_unwrapped[L] vel[L] | tabx[H] taby[H] tabpsi[H] tabv[H]
    int L = 0, H = 0, stop = 0;
};

// instance-independent part of get_waypoints (ref_traj.py:120-150): traj rows are [vel, x, y, psi, cdist, curv]
int refgen_build_track(int L, const double *traj, int H, double dt, TrackHost &T)
{
    if (L < 2 || H < 4 || !traj) return ADMPC_E_ARG;
    std::vector<double> vel(L), x(L), y(L), psi(L), cd(L), psiu(L);
    for (int i = 0; i < L; i++) { vel[i] = traj[i * 6 + 0]; x[i] = traj[i * 6 + 1]; y[i] = traj[i * 6 + 2]; psi[i] = traj[i * 6 + 3]; cd[i] = traj[i * 6 + 4]; }
    double cum = 0.0;                                          // np.unwrap(psi)
    psiu[0] = psi[0];
    for (int i = 1; i < L; i++) { cum += unwrap_corr(psi[i] - psi[i - 1]); psiu[i] = psi[i] + cum; }
    while ((int)vel.size() < H + 1) vel.push_back(0.01);       // ref_traj.py:125-126
    std::vector<double> fit(H);
    fit[0] = dt * vel[0];
    for (int h = 1; h < H; h++) fit[h] = fit[h - 1] + dt * vel[h];
    std::vector<double> tx(H), ty(H), tp(H), tc(H), tv(H);
    for (int h = 0; h < H; h++) { tx[h] = np_interp1(fit[h], cd, x); ty[h] = np_interp1(fit[h], cd, y); tp[h] = np_interp1(fit[h], cd, psiu); tc[h] = np_interp1(fit[h], cd, cd); }
    for (int h = 0; h + 1 < H; h++) tv[h] = (tc[h + 1] - tc[h]) / dt;        // ref_traj.py:148-149
    tv[H - 1] = tv[H - 2];
    T.stop = (tc[H - 1] == cd[L - 1]) ? 1 : 0;                               // ref_traj.py:151-153
    // v_ref after the blend (ref_traj.py:166-167): [v2 v2 v2 | v[2..H-2]]
    std::vector<double> tvo(H);
    for (int h = 0; h < H; h++) tvo[h] = (h < 3) ? tv[2] : tv[h - 1];
    T.L = L; T.H = H;
    T.dev.clear();
    T.dev.insert(T.dev.end(), x.begin(), x.end());
    T.dev.insert(T.dev.end(), y.begin(), y.end());
    T.dev.insert(T.dev.end(), psi.begin(), psi.end());
    T.dev.insert(T.dev.end(), cd.begin(), cd.end());
    T.dev.insert(T.dev.end(), psiu.begin(), psiu.end());
    T.dev.insert(T.dev.end(), vel.begin(), vel.begin() + L);
    T.dev.insert(T.dev.end(), tx.begin(), tx.end());
    T.dev.insert(T.dev.end(), ty.begin(), ty.end());
    T.dev.insert(T.dev.end(), tp.begin(), tp.end());
    T.dev.insert(T.dev.end(), tvo.begin(), tvo.end());
    return 0;
}

#define REF_TILE 1024
#define REF_HMAX 64
// numpy.interp on the device with a forward-moving bracket (abscissae are non-decreasing along the horizon)
__device__ __forceinline__ double interp_fwd(double s, const double *__restrict__ xp, const double *__restrict__ fp, int L, int &j)
{
    if (s < xp[0]) return fp[0];
    if (s >= xp[L - 1]) return fp[L - 1];
    while (j + 1 < L - 1 && xp[j + 1] <= s) j++;
    if (xp[j] == s) return fp[j];
    const double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
    return __dadd_rn(__dmul_rn(slope, s - xp[j]), fp[j]);
}

// one thread per instance; the track polyline is swept through shared memory in tiles (broadcast reads).
// ANCHOR = false: literal get_waypoints (abscissae measured from the start of the track window, ref_traj.py:128-131).
// ANCHOR = true : extension for a shared global track: abscissae start at the closest waypoint of each vehicle
//                 (what the reference gets implicitly because its planner publishes a window that starts at the car).
template <bool ANCHOR>
__global__ void __launch_bounds__(256) refgen_kernel(const Params P, const double *__restrict__ trk, int L, int H, double dt, double *info)
{
    __shared__ double sx[REF_TILE], sy[REF_TILE];
    const int N = P.o.N, Bp = P.Bp;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = i < P.B;
    const double *tx = trk, *ty = trk + L, *tpsi = trk + 2 * L, *tcd = trk + 3 * L;
    const double *tpsiu = trk + 4 * (size_t)L, *tvel = trk + 5 * (size_t)L;
    const double *tabx = trk + 6 * (size_t)L, *taby = tabx + H, *tabp = taby + H, *tabv = tabp + H;
    double lx_[ANCHOR ? REF_HMAX : 1], ly_[ANCHOR ? REF_HMAX : 1], lp_[ANCHOR ? REF_HMAX : 1], lv_[ANCHOR ? REF_HMAX : 1];
    const double X0 = act ? P.x0[(size_t)0 * Bp + i] : 0.0, Y0 = act ? P.x0[(size_t)1 * Bp + i] : 0.0;
    const double psi0 = act ? P.x0[(size_t)2 * Bp + i] : 0.0;
    // (1) closest waypoint: argmin_j sqrt(dx^2 + dy^2), first minimum, exactly as numpy resolves it (ref_traj.py:101).
    //     sqrt is monotone, so min_j sqrt(d2_j) = sqrt(min_j d2_j): pass 1 finds the smallest squared distance without
    //     any square root; pass 2 takes the first waypoint whose ROUNDED distance equals the rounded minimum (waypoints
    //     within a few ulp of the minimum are the only ones that need a sqrt).  No FMA contraction in d2.
    double m2 = INFINITY;
    for (int t0 = 0; t0 < L; t0 += REF_TILE) {
        const int n = min(REF_TILE, L - t0);
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += blockDim.x) { sx[j] = tx[t0 + j]; sy[j] = ty[t0 + j]; }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < n; j++) {
            const double dx = sx[j] - X0, dy = sy[j] - Y0;
            m2 = fmin(m2, __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        }
    }
    const double smin = sqrt(m2), thr = m2 * (1.0 + 2e-15);
    int ci = -1;
    for (int t0 = 0; t0 < L; t0 += REF_TILE) {
        const int n = min(REF_TILE, L - t0);
        if (L > REF_TILE) {
            __syncthreads();
            for (int j = threadIdx.x; j < n; j += blockDim.x) { sx[j] = tx[t0 + j]; sy[j] = ty[t0 + j]; }
            __syncthreads();
        }
        for (int j = 0; j < n; j++) {
            const double dx = sx[j] - X0, dy = sy[j] - Y0;
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            if (ci < 0 && d2 <= thr && sqrt(d2) == smin) ci = t0 + j;
        }
    }
    if (ci < 0) ci = 0;
    if (!act) return;
    // (2) Frenet errors (ref_traj.py:104-117)
    const double pw = tpsi[ci];
    const double ex = X0 - tx[ci], ey = Y0 - ty[ci];
    const double psi_init = bound_pi(psi0);
    if (info) {
        info[(size_t)0 * Bp + i] = tcd[ci];
        info[(size_t)1 * Bp + i] = __dadd_rn(__dmul_rn(-sin(pw), ex), __dmul_rn(cos(pw), ey));
        info[(size_t)2 * Bp + i] = bound_pi(psi_init - pw);
    }
    if (ANCHOR) {
        // per-vehicle arc-length tables: s_h = cdist[ci] + sum_{m<=h} dt * vel[ci+m]   (0.01 beyond the end, :125-126)
        double fit = 0.0, cprev = 0.0;
        int jb = ci;
        for (int hh = 0; hh < H; hh++) {
            const int vi = ci + hh;
            fit = __dadd_rn(fit, __dmul_rn(dt, (vi < L) ? tvel[vi] : 0.01));
            const double sq = tcd[ci] + fit;
            int j1 = jb, j2 = jb, j3 = jb, j4 = jb;
            lx_[hh] = interp_fwd(sq, tcd, tx, L, j1);
            ly_[hh] = interp_fwd(sq, tcd, ty, L, j2);
            lp_[hh] = interp_fwd(sq, tcd, tpsiu, L, j3);
            const double cq = interp_fwd(sq, tcd, tcd, L, j4);
            jb = j1;
            if (hh >= 1) lv_[hh - 1] = (cq - cprev) / dt;
            cprev = cq;
        }
        lv_[H - 1] = lv_[H - 2];
        // blended speed reference: [v2 v2 v2 | v[2..H-2]]  (in place, back to front)
        for (int hh = H - 1; hh >= 3; hh--) lv_[hh] = lv_[hh - 1];
        const double v2 = lv_[2];
        lv_[0] = v2; lv_[1] = v2;
        tabx = lx_; taby = ly_; tabp = lp_; tabv = lv_;
    }
    // (3) references: heading fixed relative to the vehicle heading (ref_traj.py:30-35,143-145), 3-point blend from the
    //     current pose (:157-167), padding to N+1, heading unwrap of run_optimization (ad_3d_optimizer.py:420-438)
    const double x1 = tabx[1], y1 = taby[1];
    const double xm = __dadd_rn(__dmul_rn(1.0, (x1 - X0) / 2.0), X0), ym = __dadd_rn(__dmul_rn(1.0, (y1 - Y0) / 2.0), Y0);   // np.linspace(a, b, 3)[1]
    // row writer: applies run_optimization's unwrap against the raw heading and stores one yref row (SoA)
    auto put = [&](int j, double rx, double ry, double rpsi, double rv) {
        if (j > N) return;
        double ps = rpsi;
        if (psi0 < 0.0) { if (psi0 + PI_D < ps) ps -= 2.0 * PI_D; }
        else if (psi0 > 0.0) { if (psi0 - PI_D > ps) ps += 2.0 * PI_D; }
        double *yr = (double *)P.yref + (size_t)(j * 9) * Bp + i;
        yr[(size_t)0 * Bp] = rx; yr[(size_t)1 * Bp] = ry; yr[(size_t)2 * Bp] = ps; yr[(size_t)3 * Bp] = rv;
        yr[(size_t)4 * Bp] = 0.0; yr[(size_t)5 * Bp] = 0.0; yr[(size_t)6 * Bp] = 0.0;
        if (j < N) { yr[(size_t)7 * Bp] = 0.0; yr[(size_t)8 * Bp] = 0.0; }
    };
    // psi_fix[h] = bound(psi_init + unwrap(bound(tab_psi - psi_init))[h]); output row j uses psi_fix[0] for j < 3 and
    // psi_fix[j-1] for j >= 3, x/y likewise shifted by one (hstack([linspace(.., 3), ref[2:-1]]))
    double prev = bound_pi(tabp[0] - psi_init), cum = 0.0;
    const double pf0 = bound_pi(psi_init + prev);
    put(0, X0, Y0, pf0, tabv[0]);
    put(1, xm, ym, pf0, tabv[1]);
    put(2, x1, y1, pf0, tabv[2]);
    double lx = x1, ly = y1, lp = pf0;
    for (int h = 1; h <= H - 2; h++) {
        const double dh = bound_pi(tabp[h] - psi_init);
        cum += unwrap_corr(dh - prev);
        prev = dh;
        if (h >= 2) {
            lx = tabx[h]; ly = taby[h]; lp = bound_pi(psi_init + (dh + cum));
            put(h + 1, lx, ly, lp, tabv[h + 1]);
        }
    }
    for (int j = H; j <= N; j++) put(j, lx, ly, lp, tabv[H - 1]);       // set_reference_trajectory padding
}

void launch_refgen(const Params &P, const double *trk, int L, int H, double dt, int anchor, double *info, cudaStream_t s)
{
    if (anchor) refgen_kernel<true><<<(P.Bp + 255) / 256, 256, 0, s>>>(P, trk, L, H, dt, info);
    else refgen_kernel<false><<<(P.Bp + 255) / 256, 256, 0, s>>>(P, trk, L, H, dt, info);
}
