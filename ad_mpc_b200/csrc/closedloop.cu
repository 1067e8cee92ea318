// closedloop.cu -- what the reference does right after a solve, batched and on the device:
//   * geometric sanity check of the predicted path against the reference, `is_valid_command`
//     (ad_mpc/ad_3d_optimizer.py:385-394, duplicated at nodes/gp_ad_mpc_node.py:248-257): distances at nodes 0..N-1
//     (entry N stays 0), valid iff mean < 3 m, sample variance (np.cov, ddof 1) < 2 and max < 4 m;
//   * backup control: an invalid prediction re-uses the previous valid sequence shifted by one node
//     (ad_3d_optimizer.py:469-476: w = prev_w[2:...], so the applied pair is the previous u_1);
//   * safety counter: solver status > 0 resets it, the command is released only after `threshold` consecutive
//     successes (gp_ad_mpc_node.py:62,206-216);
//   * plant step: RK4 of the nominal bicycle model (A2) over dt with the applied control (steering rate clipped to
//     its bounds like gp_ad_mpc_node.py:222), so that thousands of closed loops advance without a host round trip.
// One thread per instance; HBM-bound (reads the predicted states and the reference once).
#include "model.cuh"

struct LoopDev {
    double *prev_u;      // [2N][Bp] last valid control sequence
    int *has_prev;       // [Bp]
    int *safe_count;     // [Bp]
    int *valid;          // [Bp] result of is_valid_command for the last solve
    int *cmd_ok;         // [Bp] 1 when the command would be published (counter >= threshold and prediction healthy)
    double *u_apply;     // [2][Bp]
    double *x_next;      // [7][Bp] plant state after the step (also written to x0)
    int threshold;
};

__global__ void __launch_bounds__(128) postsolve_kernel(const Params P, const LoopDev Lp, int advance)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.B) return;
    // ---- is_valid_command ---------------------------------------------------------------------------------------------
    double sum = 0.0, sumsq = 0.0, mx = 0.0;
    for (int k = 0; k < N; k++) {
        const double dx = P.yref[(size_t)(k * 9 + 0) * Bp + i] - P.xb[(size_t)(k * 7 + 0) * Bp + i];
        const double dy = P.yref[(size_t)(k * 9 + 1) * Bp + i] - P.xb[(size_t)(k * 7 + 1) * Bp + i];
        const double d = sqrt(dx * dx + dy * dy);
        sum += d; mx = fmax(mx, d);
    }
    const double n = (double)(N + 1), mean = sum / n;
    for (int k = 0; k < N; k++) {
        const double dx = P.yref[(size_t)(k * 9 + 0) * Bp + i] - P.xb[(size_t)(k * 7 + 0) * Bp + i];
        const double dy = P.yref[(size_t)(k * 9 + 1) * Bp + i] - P.xb[(size_t)(k * 7 + 1) * Bp + i];
        const double d = sqrt(dx * dx + dy * dy) - mean;
        sumsq += d * d;
    }
    sumsq += mean * mean;                                   // the (N+1)-th entry of tmp_dist is 0
    const double var = sumsq / (n - 1.0);                   // np.cov of a 1-D array: ddof = 1
    const bool valid = (mean < 3.0) && (var < 2.0) && (mx < 4.0);
    Lp.valid[i] = valid ? 1 : 0;
    // ---- control selection ----------------------------------------------------------------------------------------------
    double u0 = P.ub[(size_t)0 * Bp + i], u1 = P.ub[(size_t)1 * Bp + i];
    if (valid) {
        for (int r = 0; r < 2 * N; r++) Lp.prev_u[(size_t)r * Bp + i] = P.ub[(size_t)r * Bp + i];
        Lp.has_prev[i] = 1;
    } else if (Lp.has_prev[i]) {
        u0 = Lp.prev_u[(size_t)2 * Bp + i];                 // prev_w[2], prev_w[3]
        u1 = Lp.prev_u[(size_t)3 * Bp + i];
    }
    // ---- safety counter ----------------------------------------------------------------------------------------------------
    const int st = P.status[i];
    const int cnt = (st > 0) ? 0 : Lp.safe_count[i] + 1;
    Lp.safe_count[i] = cnt;
    Lp.cmd_ok[i] = (cnt >= Lp.threshold && valid) ? 1 : 0;
    Lp.u_apply[(size_t)0 * Bp + i] = u0;
    Lp.u_apply[(size_t)1 * Bp + i] = u1;
    if (!advance) return;
    // ---- plant: one RK4 step of the nominal model --------------------------------------------------------------------------
    double x[7], u[2] = {u0, fmin(fmax(u1, o.lbu[1]), o.ubu[1])};
#pragma unroll
    for (int c = 0; c < 7; c++) x[c] = P.x0[(size_t)c * Bp + i];
    const double p = P.p[(size_t)0 * Bp + i], h = o.dt;
    double kx[7], ax[7], gz[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 7; c++) { kx[c] = 0.0; ax[c] = 0.0; }
#pragma unroll 1
    for (int s = 0; s < 4; s++) {
        const double as = (s == 0) ? 0.0 : ((s == 3) ? 1.0 : 0.5), bs = (s == 0 || s == 3) ? (1.0 / 6.0) : (1.0 / 3.0);
        double xs[7], f[7];
#pragma unroll
        for (int c = 0; c < 7; c++) xs[c] = fma(h * as, kx[c], x[c]);
        Jac J;
        model_eval<false>(o, nullptr, 0, 0u, xs, u, p, gz, 0.0, f, J);
#pragma unroll
        for (int c = 0; c < 7; c++) { kx[c] = f[c]; ax[c] = fma(bs, f[c], ax[c]); }
    }
    double vx_next = 0.0;
#pragma unroll
    for (int c = 0; c < 7; c++) {
        const double xn = fma(h, ax[c], x[c]);
        if (c == 3) vx_next = xn;
        Lp.x_next[(size_t)c * Bp + i] = xn;
        ((double *)P.x0)[(size_t)c * Bp + i] = xn;
        ((double *)P.gps)[(size_t)c * Bp + i] = xn;        // gp_state follows the measured state (quad_3d_optimizer.py:549)
    }
    // the reference recomputes the kinematic/dynamic blend from the measured v_x before every solve
    // (ad_3d_optimizer.py:443-450): the next solve (and the next plant step) of this vehicle uses it on all stages
    if (o.blend_max > o.blend_min) {
        const double pn = fmin(fmax((vx_next - o.blend_min) / (o.blend_max - o.blend_min), 0.0), 1.0);
        for (int k = 0; k < N; k++) ((double *)P.p)[(size_t)k * Bp + i] = pn;
    }
}

// p of every stage from the current x0 (pose-only step without a p from the host), ad_3d_optimizer.py:443-450
__global__ void __launch_bounds__(128) blend_kernel(const Params P)
{
    const admpc_opts &o = P.o;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.B) return;
    const double vx = P.x0[(size_t)3 * P.Bp + i];
    const double pn = fmin(fmax((vx - o.blend_min) / (o.blend_max - o.blend_min), 0.0), 1.0);
    for (int k = 0; k < o.N; k++) ((double *)P.p)[(size_t)k * P.Bp + i] = pn;
}
void launch_blend(const Params &P, cudaStream_t s) { blend_kernel<<<(P.B + 127) / 128, 128, 0, s>>>(P); }

void launch_postsolve(const Params &P, double *prev_u, int *ibuf, double *u_apply, double *x_next, int threshold, int advance, cudaStream_t s)
{
    LoopDev L;
    L.prev_u = prev_u;
    L.has_prev = ibuf; L.safe_count = ibuf + P.Bp; L.valid = ibuf + 2 * (size_t)P.Bp; L.cmd_ok = ibuf + 3 * (size_t)P.Bp;
    L.u_apply = u_apply; L.x_next = x_next; L.threshold = threshold;
    postsolve_kernel<<<(P.B + 127) / 128, 128, 0, s>>>(P, L, advance);
}
