// prepare.cu -- preparation phase of the SQP-RTI step: one thread per (instance, shooting interval).
//
// Replaces, for a whole batch, what acados does per stage on the CPU:
//   * ERK4 (4 stages x 1 step) of the bicycle model WITH forward sensitivities -- the reference runs the CasADi
//     generated sim_car_expl_vde_forw (c_generated_code/sim_car_model/sim_car_expl_vde_forw.c:120) 4x per interval;
//     here the Jacobian is hand-derived and only its structurally non-zero entries are touched:
//     columns p_x,p_y of J_x are zero and row delta is trivial, so S keeps [e0 e1 | 6x5 block] + 6x2 input block.
//   * optional GP residual f + B_x mu(z) and its Jacobian (model_fitting/gp.py:117-165,446-460;
//     quad_mpc/quad_3d_optimizer.py:295,315) in TWO passes: gp_sweep_kernel evaluates mean and feature gradient at the
//     four RK4 stage points (training set staged into shared memory once per CTA with a TMA bulk copy, cp.async.bulk +
//     mbarrier; every thread sweeps the M points with broadcast LDS), prepare_kernel<true> then propagates the
//     sensitivities with those terms read back -- the stage points do not depend on the sensitivities.
//   * LINEAR_LS gradient and multiple-shooting residual (acados_solver_sim_car.c:378-485).
// Output: lin[k][58][Bp] (A 6x5, B 6x2, b 7, q 7, r 2), SoA, coalesced.
#include "common.cuh"

#include "tma.cuh"

#include "model.cuh"

// ---- pass 1 of the GP-augmented preparation: the GP sweeps -----------------------------------------------------------------
// One thread per (instance, shooting interval) walks the four RK4 stages of the STATE only (the stage points do not depend
// on the sensitivities) and evaluates the GP mean and its feature gradient at each of them: 4 sweeps over the M training
// points staged in shared memory by one TMA bulk copy per CTA.  The 84-double sensitivity state is not alive here, so the
// kernel runs at 24 warps per SM without pushing it through local memory around every sweep (the fused kernel moved
// 2.8 GB of DRAM traffic per launch that way); the results (10 doubles per stage at d_z = 4, two outputs) go to gpr.
// ENS: GP ensemble -- the cluster model is chosen per instance, so the training-set loads use per-thread addresses; a
// single model keeps warp-uniform addresses (uniform-register LDS), which is measurably cheaper.
// FR: Frenet variant -- the pose rows of f are the curvilinear ones (curvature per node or from the kappa(s) spline at the
// sub-stage's own arc length); the GP terms and rows 3..6 are shared with the Cartesian model.
template <int BLOCK, int MINB, bool ENS, int PREC, bool FR>
__global__ void __launch_bounds__(BLOCK, MINB) gp_sweep_kernel(const Params P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (uint32_t)P.gp.bytes);
        uint32_t off = 0;                      // TMA bulk copies, <= 64 KB each
        while (off < (uint32_t)P.gp.bytes) {
            uint32_t n = min((uint32_t)P.gp.bytes - off, 65536u);
            tma_bulk_g2s(smem_raw + off, (const unsigned char *)P.gp.blob + off, n, &bar);
            off += n;
        }
    }
    const double *gpsm = reinterpret_cast<const double *>(smem_raw);
    const admpc_opts &o = P.o;
    const int Bp = P.Bp;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;                                   // k < N
    const double h = o.dt;
    const bool active = (i < P.B) && (P.lin_bad[i] != 2);      // 2: finished instance of the full-SQP loop (sqp.cu)
    const bool trg = (o.gp_stage0_trigger && k == 0);
    const double trig = trg ? 1.0 : 0.0;
    mbar_wait(&bar, 0);
    if (!active) return;
    const double *gpm = ENS ? gpsm + (size_t)P.gp_sel[i] * P.gp.model_doubles : gpsm;   // this instance's cluster model
    const uint32_t tab = (uint32_t)__cvta_generic_to_shared(gpsm + (size_t)P.gp.n_models * P.gp.model_doubles);
    const int R = gpr_rows(o), dz = o.gp_dz;
    double kx[7];
#pragma unroll
    for (int c = 0; c < 7; c++) kx[c] = 0.0;
    int kap_j = -1;                                             // FR: spline piece of the previous sub-stage
#pragma unroll 1
    for (int s = 0; s < 4; s++) {
        const double as = (s == 0) ? 0.0 : ((s == 3) ? 1.0 : 0.5);
        const double ha = h * as;
        // x, u, p, gp_state are re-read per stage (coalesced, L1/L2 hits) instead of being kept alive across the sweep
        double xs[7], u[2], gpx[7], f[7];
#pragma unroll
        for (int c = 0; c < 7; c++) {
            xs[c] = fma(ha, kx[c], P.xb[(size_t)(k * 7 + c) * Bp + i]);
            gpx[c] = trg ? P.gps[(size_t)c * Bp + i] : 0.0;
        }
        u[0] = P.ub[(size_t)(k * 2 + 0) * Bp + i]; u[1] = P.ub[(size_t)(k * 2 + 1) * Bp + i];
        const double pk = P.p[(size_t)k * Bp + i];
        GpOut G;
        gp_eval<PREC>(o, gpm, P.gp.stride_out, tab, xs, u, gpx, trig, G);
        ADMPC_ASSERT(k < o.N && soa_at((k * 4 + s) * R + R - 1, o.N * 4 * R, i, Bp) < (size_t)o.N * 4 * R * Bp);
        double *out = P.gpr + (size_t)(k * 4 + s) * R * Bp + i;
#pragma unroll
        for (int j = 0; j < ADMPC_GPOUT_MAX; j++) {
            if (j >= o.gp_nout) continue;
            out[(size_t)(j * (1 + dz)) * Bp] = G.m[j];
#pragma unroll
            for (int d = 0; d < ADMPC_DZMAX; d++) if (d < dz) out[(size_t)(j * (1 + dz) + 1 + d) * Bp] = G.g[j][d];
        }
        if (s == 3) break;
        Jac J;
        model_eval<false>(o, nullptr, 0, 0u, xs, u, pk, gpx, trig, f, J);        // nominal f (its Jacobian is dead code here)
        gp_apply(o, trig, G, f, J);
        if (FR) {
            double kap = P.kappa[(size_t)k * Bp + i], dkap;
            if (P.kap_K > 0) kappa_spline(P, i, xs[0], kap, dkap, kap_j);
            frenet_pose_rows(xs, kap, f);
        }
#pragma unroll
        for (int c = 0; c < 7; c++) kx[c] = f[c];
    }
}

// ---- pass 2 / nominal preparation: one RK4 step with forward sensitivities --------------------------------------------------
// Sensitivity state: rows 0..5 x 7 columns [x2..x6 | u0 u1]; row 6 (delta) is analytic: d delta / d delta = 1,
// d delta / d u1 = t.  GPR: the GP mean / gradient of every RK4 stage come from gpr (written by gp_sweep_kernel).
// IM: the linearisation goes to the instance-major records lin_im[k][i][LIM_STRIDE] = {M, b, q, r, x, u} that the
// shared-memory-resident QP kernel pulls in with one TMA bulk copy per stage; otherwise to the SoA rows lin[k][58][Bp].
template <bool GPR, bool IM>
__device__ __forceinline__ void prepare_body(const Params &P, int i, int k, double *lin);

#define LIM_PAD (LIM_STRIDE + 1)      // row stride of the staging tile: odd, so that a column of the tile hits 32 banks
#ifndef PREP_MINB
#define PREP_MINB 2
#endif
template <bool GPR, bool IM>
__global__ void __launch_bounds__(128, PREP_MINB) prepare_kernel(const Params P)
{
    extern __shared__ __align__(16) double tile[];              // IM: [128][LIM_PAD] records of this CTA, written out coalesced
    __shared__ int act[128];
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    const double h = o.dt;
    ADMPC_ASSERT(k <= N);                                     // (threads past the batch are masked by `active` below)
    const bool active = (i < P.B) && (P.lin_bad[i] != 2);      // 2: finished instance of the full-SQP loop (sqp.cu)
    if (IM && k < N) {
        act[threadIdx.x] = active ? 1 : 0;
        if (active) prepare_body<GPR, IM>(P, i, k, tile + (size_t)threadIdx.x * LIM_PAD);
        __syncthreads();
        // the 128 records of this CTA are one contiguous block of lin_im[k]: coalesced 8-byte stores
        double *dst = P.lin_im + ((size_t)k * Bp + (size_t)blockIdx.x * blockDim.x) * LIM_STRIDE;
        const int nrec = min((int)blockDim.x, Bp - (int)(blockIdx.x * blockDim.x));
        for (int e = threadIdx.x; e < nrec * LIM_STRIDE; e += blockDim.x) {
            const int r = e / LIM_STRIDE, c = e - r * LIM_STRIDE;
            if (act[r]) dst[e] = tile[(size_t)r * LIM_PAD + c];
        }
        return;
    }
    if (!active) return;
    prepare_body<GPR, IM>(P, i, k, IM ? P.lin_im + ((size_t)k * Bp + i) * LIM_STRIDE : P.lin + (size_t)k * LIN_ROWS * Bp + i);
}

// body of one (instance, interval): `lin` is the record to fill (IM: LIM_* offsets, contiguous; else SoA rows, stride Bp)
template <bool GPR, bool IM>
__device__ __forceinline__ void prepare_body(const Params &P, int i, int k, double *lin)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const double h = o.dt;

    double x[7], u[2], xn[7], yr[9];
    double pk = 0.0;
#pragma unroll
    for (int c = 0; c < 7; c++) x[c] = P.xb[(size_t)(k * 7 + c) * Bp + i];
    if (k < N) {
#pragma unroll
        for (int c = 0; c < 7; c++) xn[c] = P.xb[(size_t)((k + 1) * 7 + c) * Bp + i];
#pragma unroll
        for (int c = 0; c < 2; c++) u[c] = P.ub[(size_t)(k * 2 + c) * Bp + i];
#pragma unroll
        for (int c = 0; c < 9; c++) yr[c] = P.yref[(size_t)(k * 9 + c) * Bp + i];
        pk = P.p[(size_t)k * Bp + i];
    } else {
#pragma unroll
        for (int c = 0; c < 7; c++) yr[c] = P.yref[(size_t)(N * 9 + c) * Bp + i];
    }
#define LIN_PUT(row, v) do { if (IM) lin[lim_of_row(row)] = (v); else lin[(size_t)(row) * Bp] = (v); } while (0)
    if (IM) {            // linearisation point: the QP kernel forms bounds / the full step from it without touching HBM again
#pragma unroll
        for (int c = 0; c < 7; c++) lin[LIM_X + c] = x[c];
        if (k < N) { lin[LIM_U] = u[0]; lin[LIM_U + 1] = u[1]; }
    }
    if (k == N) {
        // terminal cost gradient, scaling 1
#pragma unroll
        for (int c = 0; c < 7; c++) LIN_PUT(LIN_q + c, o.We[c] * (x[c] - yr[c]));
        return;
    }
    double gpx[7];
#pragma unroll
    for (int c = 0; c < 7; c++) gpx[c] = 0.0;
    const double trig = (GPR && o.gp_stage0_trigger && k == 0) ? 1.0 : 0.0;
    const int R = GPR ? gpr_rows(o) : 0, dz = o.gp_dz;

    // K = current stage derivative of the sensitivity block, acc = weighted sum (rows 0..5 x 7 cols)
    double K[6][7], acc[6][7];
    double kx[7], ax[7];
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = 0; c < 7; c++) { K[r][c] = 0.0; acc[r][c] = 0.0; }
#pragma unroll
    for (int c = 0; c < 7; c++) { kx[c] = 0.0; ax[c] = 0.0; }

    GpOut Gn;
    if (GPR) gpr_load(P, o, (k * 4) * R, dz, i, Gn);
#pragma unroll 1
    for (int s = 0; s < 4; s++) {
        const double as = (s == 0) ? 0.0 : ((s == 3) ? 1.0 : 0.5);
        const double bs = (s == 0 || s == 3) ? (1.0 / 6.0) : (1.0 / 3.0);
        const double ha = h * as;
        double xs[7], f[7];
#pragma unroll
        for (int c = 0; c < 7; c++) xs[c] = fma(ha, kx[c], x[c]);
        Jac J;
        model_eval<false>(o, nullptr, 0, 0u, xs, u, pk, gpx, trig, f, J);
        if (GPR) {
            const GpOut G = Gn;                      // loaded one stage ahead: its latency hides behind the model evaluation
            if (s < 3) gpr_load(P, o, (k * 4 + s + 1) * R, dz, i, Gn);
            gp_apply(o, trig, G, f, J);
        }
#pragma unroll
        for (int c = 0; c < 7; c++) { kx[c] = f[c]; ax[c] = fma(bs, f[c], ax[c]); }
        // sensitivity columns: c = 0..4 <-> x2..x6, c = 5,6 <-> u0,u1
#pragma unroll
        for (int c = 0; c < 7; c++) {
            // stage input S_in[:,c] = S0[:,c] + ha * K[:,c]; S0 = e_(2+c) for state columns, 0 for input columns
            double sv[7];   // rows 0..6
#pragma unroll
            for (int r = 0; r < 6; r++) sv[r] = ha * K[r][c] + ((c < 5 && r == 2 + c) ? 1.0 : 0.0);
            sv[6] = (c == 4) ? 1.0 : ((c == 6) ? ha : 0.0);
            double kn[6];
            kn[0] = fma(J.j0[0], sv[2], fma(J.j0[1], sv[3], J.j0[2] * sv[4]));
            kn[1] = fma(J.j1[0], sv[2], fma(J.j1[1], sv[3], J.j1[2] * sv[4]));
            kn[2] = sv[5];
#pragma unroll
            for (int rr = 0; rr < 3; rr++) {
                double v = (c >= 5) ? J.ju[rr][c - 5] : 0.0;
#pragma unroll
                for (int l = 0; l < 5; l++) v = fma(J.jr[rr][l], sv[2 + l], v);
                kn[3 + rr] = v;
            }
#pragma unroll
            for (int r = 0; r < 6; r++) { K[r][c] = kn[r]; acc[r][c] = fma(bs, kn[r], acc[r][c]); }
        }
    }
    bool bad = false;
    // x+ = x + h * sum b_s k_s ; b = x+ - x_{k+1}
#pragma unroll
    for (int c = 0; c < 7; c++) {
        const double xp = fma(h, ax[c], x[c]);
        bad |= !isfinite(xp);
        LIN_PUT(LIN_b + c, xp - xn[c]);
    }
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int c = 0; c < 5; c++) {
            const double v = h * acc[r][c] + ((r == 2 + c) ? 1.0 : 0.0);
            bad |= !isfinite(v);
            LIN_PUT(LIN_A + r * 5 + c, v);
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const double v = h * acc[r][5 + c];
            bad |= !isfinite(v);
            LIN_PUT(LIN_B + r * 2 + c, v);
        }
    }
    const double Ts = o.dt;
#pragma unroll
    for (int c = 0; c < 7; c++) LIN_PUT(LIN_q + c, Ts * o.W[c] * (x[c] - yr[c]));
#pragma unroll
    for (int c = 0; c < 2; c++) LIN_PUT(LIN_r + c, Ts * o.W[7 + c] * (u[c] - yr[7 + c]));
    if (bad) P.lin_bad[i] = 1;
}

// pass 1 of a GP-augmented preparation (both model variants): the GP sweeps at the four RK4 stage points -> P.gpr
void launch_gp_sweep(const Params &P, cudaStream_t s)
{
    const bool frenet = P.o.model_variant == 1;
    const size_t sm = (size_t)P.gp.bytes;
    const bool ens = P.gp.n_models > 1;
    const int prec = P.o.gp_precision ? 1 : 0;
    // Register budget capped at 80 (24 warps/SM) for small models; models above 36 KB run 4 CTAs x 128 registers, models
    // above 54 KB one CTA per SM whose width is picked below.
#define LAUNCH_GP1(BOUND, MINB, BLK, ENS, PREC, FR)                                                                         \
    do {                                                                                                                \
        static SmemGuard configured;                                                                                   \
        if (configured.need(sm))                                                                                        \
            cudaFuncSetAttribute(gp_sweep_kernel<BOUND, MINB, ENS, PREC, FR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
        dim3 grid((P.Bp + (BLK) - 1) / (BLK), P.o.N);                                                                   \
        gp_sweep_kernel<BOUND, MINB, ENS, PREC, FR><<<grid, (BLK), sm, s>>>(P);                                             \
    } while (0)
#define LAUNCH_GP2(BOUND, MINB, BLK, FR)                                                                                 \
    do {                                                                                                                \
        if (ens) { if (prec) LAUNCH_GP1(BOUND, MINB, BLK, true, 1, FR); else LAUNCH_GP1(BOUND, MINB, BLK, true, 0, FR); } \
        else { if (prec) LAUNCH_GP1(BOUND, MINB, BLK, false, 1, FR); else LAUNCH_GP1(BOUND, MINB, BLK, false, 0, FR); }   \
    } while (0)
#define LAUNCH_GP(BOUND, MINB, BLK)                                                                                      \
    do { if (frenet) LAUNCH_GP2(BOUND, MINB, BLK, true); else LAUNCH_GP2(BOUND, MINB, BLK, false); } while (0)
    if (sm <= 36 * 1024) {
        LAUNCH_GP(128, 6, 128);       // 80 registers, 24 warps/SM
    } else if (sm <= 54 * 1024) {
        LAUNCH_GP(128, 4, 128);
    } else {
        // One CTA per SM by shared memory.  All CTAs of the N grid rows take the same time, so the run time is
        // (number of waves over the device's SMs) x (CTA width): pick the width, in warps, that minimises it; wider CTAs
        // run with a lower register cap (more resident warps, slightly better latency hiding).
        static int sm_count[64] = {};                        // SMs of the device this launch goes to (cached per device)
        int dev = 0;
        cudaGetDevice(&dev);
        dev &= 63;
        if (!sm_count[dev]) cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
        const int nsm = sm_count[dev] > 0 ? sm_count[dev] : 148;
        int best = 512;
        double best_cost = 1e300;
        for (int blk = 1024; blk >= 256; blk -= 32) {
            const long ctas = (long)((P.Bp + blk - 1) / blk) * P.o.N;
            const double cost = (double)((ctas + nsm - 1) / nsm) * blk;
            if (cost < best_cost) { best_cost = cost; best = blk; }
        }
        if (best <= 512) LAUNCH_GP(512, 1, best);
        else if (best <= 640) LAUNCH_GP(640, 1, best);
        else if (best <= 768) LAUNCH_GP(768, 1, best);
        else LAUNCH_GP(1024, 1, best);
    }
}

void launch_prepare(const Params &P, cudaStream_t s)
{
    dim3 gridB((P.Bp + 127) / 128, P.o.N + 1);
    const size_t tile_bytes = (size_t)128 * LIM_PAD * sizeof(double);       // 70.7 KB: needs the opt-in attribute
    if (P.lin_im) {
        static SmemGuard cfg;
        if (cfg.need(tile_bytes)) {
            cudaFuncSetAttribute(prepare_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes);
            cudaFuncSetAttribute(prepare_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes);
        }
    }
    if (!P.o.gp_enabled) {
        if (P.lin_im) prepare_kernel<false, true><<<gridB, 128, tile_bytes, s>>>(P);
        else prepare_kernel<false, false><<<gridB, 128, 0, s>>>(P);
        return;
    }
    launch_gp_sweep(P, s);
    // pass 2: RK4 + forward sensitivities with the stored GP terms
    if (P.lin_im) prepare_kernel<true, true><<<gridB, 128, tile_bytes, s>>>(P);
    else prepare_kernel<true, false><<<gridB, 128, 0, s>>>(P);
}
