// prepare.cu -- preparation phase of the SQP-RTI step: one thread per (instance, shooting interval).
//
// Replaces, for a whole batch, what acados does per stage on the CPU:
//   * ERK4 (4 stages x 1 step) of the bicycle model WITH forward sensitivities -- the reference runs the CasADi
//     generated sim_car_expl_vde_forw (c_generated_code/sim_car_model/sim_car_expl_vde_forw.c:120) 4x per interval;
//     here the Jacobian is hand-derived and only its structurally non-zero entries are touched:
//     columns p_x,p_y of J_x are zero and row delta is trivial, so S keeps [e0 e1 | 6x5 block] + 6x2 input block.
//   * optional GP residual f + B_x mu(z) and its Jacobian (model_fitting/gp.py:117-165,446-460;
//     quad_mpc/quad_3d_optimizer.py:295,315): training set staged into shared memory once per CTA with a TMA bulk
//     copy (cp.async.bulk + mbarrier), every thread then sweeps the M points with broadcast LDS.
//   * LINEAR_LS gradient and multiple-shooting residual (acados_solver_sim_car.c:378-485).
// Output: lin[k][58][Bp] (A 6x5, B 6x2, b 7, q 7, r 2), SoA, coalesced.
#include "common.cuh"

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

#include "model.cuh"

// One RK4 step with forward sensitivities. Sensitivity state: rows 0..5 x 7 columns [x2..x6 | u0 u1];
// row 6 (delta) is analytic: d delta / d delta = 1, d delta / d u1 = t.
// ENS: GP ensemble -- the cluster model is chosen per instance, so the training-set loads use per-thread addresses; a
// single model keeps warp-uniform addresses (uniform-register LDS), which is measurably cheaper.
template <bool GP, int BLOCK, int MINB, bool ENS = false>
__global__ void __launch_bounds__(BLOCK, MINB) prepare_kernel(const Params P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    const double *gpsm = nullptr;
    if (GP) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, (uint32_t)P.gp.bytes);
            // TMA bulk copies, <= 64 KB each
            uint32_t off = 0;
            while (off < (uint32_t)P.gp.bytes) {
                uint32_t n = min((uint32_t)P.gp.bytes - off, 65536u);
                tma_bulk_g2s(smem_raw + off, (const unsigned char *)P.gp.blob + off, n, &bar);
                off += n;
            }
        }
        gpsm = reinterpret_cast<const double *>(smem_raw);
    }
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    const double h = o.dt;
    const bool active = (i < P.B) && (P.lin_bad[i] != 2);      // 2: finished instance of the full-SQP loop (sqp.cu)

    double x[7], u[2], xn[7], yr[9];
    double pk = 0.0;
    if (active) {
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
    }
    double *lin = P.lin + (size_t)k * LIN_ROWS * Bp + i;
    if (k == N) {
        // terminal cost gradient, scaling 1
        if (active) {
#pragma unroll
            for (int c = 0; c < 7; c++) lin[(size_t)(LIN_q + c) * Bp] = o.We[c] * (x[c] - yr[c]);
        }
        if (GP) mbar_wait(&bar, 0);   // do not exit with the bulk copy in flight
        return;
    }
    double gpx[7];
    double trig = 0.0;
    if (GP) {
        trig = (o.gp_stage0_trigger && k == 0) ? 1.0 : 0.0;
#pragma unroll
        for (int c = 0; c < 7; c++) gpx[c] = (active && k == 0 && o.gp_stage0_trigger) ? P.gps[(size_t)c * Bp + i] : 0.0;
        mbar_wait(&bar, 0);
    }
    if (!active) return;
    const double *gpm = (GP && ENS) ? gpsm + (size_t)P.gp_sel[i] * P.gp.model_doubles : gpsm;   // this instance's cluster model
    const uint32_t tab = GP ? (uint32_t)__cvta_generic_to_shared(gpsm + (size_t)P.gp.n_models * P.gp.model_doubles) : 0u;

    // K = current stage derivative of the sensitivity block, acc = weighted sum (rows 0..5 x 7 cols)
    double K[6][7], acc[6][7];
    double kx[7], ax[7];
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = 0; c < 7; c++) { K[r][c] = 0.0; acc[r][c] = 0.0; }
#pragma unroll
    for (int c = 0; c < 7; c++) { kx[c] = 0.0; ax[c] = 0.0; }

#pragma unroll 1
    for (int s = 0; s < 4; s++) {          // not unrolled: 4 copies of the GP sweep would not fit the instruction cache
        const double as = (s == 0) ? 0.0 : ((s == 3) ? 1.0 : 0.5);
        const double bs = (s == 0 || s == 3) ? (1.0 / 6.0) : (1.0 / 3.0);
        const double ha = h * as;
        double xs[7], f[7];
#pragma unroll
        for (int c = 0; c < 7; c++) xs[c] = fma(ha, kx[c], x[c]);
        Jac J;
        model_eval<GP>(o, gpm, P.gp.stride_out, tab, xs, u, pk, gpx, trig, f, J);
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
        lin[(size_t)(LIN_b + c) * Bp] = xp - xn[c];
    }
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int c = 0; c < 5; c++) {
            const double v = h * acc[r][c] + ((r == 2 + c) ? 1.0 : 0.0);
            bad |= !isfinite(v);
            lin[(size_t)(LIN_A + r * 5 + c) * Bp] = v;
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const double v = h * acc[r][5 + c];
            bad |= !isfinite(v);
            lin[(size_t)(LIN_B + r * 2 + c) * Bp] = v;
        }
    }
    const double Ts = o.dt;
#pragma unroll
    for (int c = 0; c < 7; c++) lin[(size_t)(LIN_q + c) * Bp] = Ts * o.W[c] * (x[c] - yr[c]);
#pragma unroll
    for (int c = 0; c < 2; c++) lin[(size_t)(LIN_r + c) * Bp] = Ts * o.W[7 + c] * (u[c] - yr[7 + c]);
    if (bad) P.lin_bad[i] = 1;
}

void launch_prepare(const Params &P, cudaStream_t s)
{
    if (P.o.gp_enabled) {
        // register budget capped at 80 (24 warps/SM) or 128 (16 warps/SM): the GP sweep itself needs ~60 registers, the
        // RK4 sensitivity state is spilled around it (once per RK4 stage, negligible against the M-point loop).
        // Small models: 6 CTAs x 128 threads per SM; large models (one CTA per SM by shared memory): 512 / 768 threads.
        const size_t sm = (size_t)P.gp.bytes;
#define LAUNCH_PREP(BOUND, MINB, BLK)                                                                                    \
    do {                                                                                                                \
        static SmemGuard configured;                                                                                   \
        if (configured.need(sm)) {                                                                                          \
            cudaFuncSetAttribute(prepare_kernel<true, BOUND, MINB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
            cudaFuncSetAttribute(prepare_kernel<true, BOUND, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
        }                                                                                                               \
        dim3 grid((P.Bp + (BLK) - 1) / (BLK), P.o.N + 1);                                                               \
        if (P.gp.n_models > 1) prepare_kernel<true, BOUND, MINB, true><<<grid, (BLK), sm, s>>>(P);                      \
        else prepare_kernel<true, BOUND, MINB, false><<<grid, (BLK), sm, s>>>(P);                                       \
    } while (0)
        if (sm <= 36 * 1024) {
            LAUNCH_PREP(128, 6, 128);       // 80 registers, 24 warps/SM
        } else if (sm <= 54 * 1024) {
            LAUNCH_PREP(128, 4, 128);
        } else {
            // One CTA per SM by shared memory.  All CTAs of the N working grid rows take the same time, so the run time is
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
                const double eff = blk > 768 ? 1.02 : blk > 640 ? 0.96 : blk > 512 ? 0.98 : 1.0;   // measured per-thread cost
                const double cost = (double)((ctas + nsm - 1) / nsm) * blk * eff;
                if (cost < best_cost) { best_cost = cost; best = blk; }
            }
            if (best <= 512) LAUNCH_PREP(512, 1, best);
            else if (best <= 640) LAUNCH_PREP(640, 1, best);
            else if (best <= 768) LAUNCH_PREP(768, 1, best);
            else LAUNCH_PREP(1024, 1, best);
        }
    } else {
        dim3 grid((P.Bp + 127) / 128, P.o.N + 1);
        prepare_kernel<false, 128, 1><<<grid, 128, 0, s>>>(P);
    }
}
