// sqp.cu -- full SQP mode (nlp_solver_type "SQP"): per-instance NLP KKT residuals and termination bookkeeping.
//
// Reference: the point-reference controller asks acados for "SQP" instead of "SQP_RTI"
// (ad_mpc/create_ros_ad_mpc.py:47-51 -> ad_3d_optimizer.py:205); nlp_solver_max_iter 100 and the four NLP tolerances
// 1e-6 are in acados_models/sim_car_acados_ocp.json:868-873.  acados' ocp_nlp_sqp [EXT] repeats
//   linearise -> NLP residuals -> stop if below tolerance -> QP -> full step
// and returns 0 (converged), 2 (max_iter), 4 (QP failure) or 1 (NaN).  The batched loop in api.cu runs the same
// prepare / QP kernels as the RTI step; instances that have finished are marked lin_bad = 2 and skipped by both.
//
// nlp_res_kernel: one thread per instance, stages swept sequentially (coalesced SoA rows); the residual is the IPM
// residual function at a zero step with the iterate's own multipliers and slacks (oracle: orc_nlp_residuals).
#include "common.cuh"

#define ATS(arr, row) (arr)[(size_t)(row) * Bp + i]
__device__ __forceinline__ double nmx(double a, double b) { return (a > b || a != a) ? a : b; }   // NaN-propagating

__global__ void __launch_bounds__(128) nlp_res_kernel(const Params P, int it, double tol0, double tol1, double tol2, double tol3,
                                                      int *active)
{
    const admpc_opts &o = P.o;
    const int N = o.N, Bp = P.Bp;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.B) return;
    const int flag = P.lin_bad[i];
    if (flag == 2) return;                                                         // finished earlier
    // QP failure (4) in iteration it-1: that iteration does not count (the iterate was not updated)
    if (it > 0 && P.status[i] != 0) { P.sqp_status[i] = P.status[i]; P.sqp_iter[i] = it - 1; P.lin_bad[i] = 2; return; }
    if (flag == 1) { P.sqp_status[i] = 1; P.status[i] = 1; P.sqp_iter[i] = it; P.lin_bad[i] = 2; return; }   // NaN linearisation
    const double Ts = o.dt, hdt = o.dt;
    double ng = 0, nb = 0, nd = 0, nm = 0;
    double pim[7];
#pragma unroll
    for (int a = 0; a < 7; a++) { pim[a] = 0.0; nb = nmx(nb, fabs(ATS(P.x0, a) - ATS(P.xb, a))); }
    for (int k = 0; k < N; k++) {
        double pi[7], lam[NC], t[NC];
#pragma unroll
        for (int a = 0; a < 7; a++) pi[a] = ATS(P.pib, k * 7 + a);
#pragma unroll
        for (int c = 0; c < NC; c++) { lam[c] = ATS(P.lamb, k * NC + c); t[c] = ATS(P.tb, k * NC + c); }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            double g = lin_get(P, k, LIN_r + j, i);
#pragma unroll
            for (int r = 0; r < 6; r++) g += lin_get(P, k, LIN_B + r * 2 + j, i) * pi[r];
            if (j == 1) g += hdt * pi[6];
            g += -lam[j] + lam[3 + j];
            const double sl = ATS(P.slb, k * 2 + j), su = ATS(P.sub, k * 2 + j), u = ATS(P.ub, k * 2 + j);
            const double gsl = Ts * o.zl[j] + Ts * o.Zl[j] * sl - lam[j] - lam[6 + j];
            const double gsu = Ts * o.zu[j] + Ts * o.Zu[j] * su - lam[3 + j] - lam[8 + j];
            ng = nmx(ng, nmx(fabs(g), nmx(fabs(gsl), fabs(gsu))));
            const double lo = o.lbu[j] - u, hi = o.ubu[j] - u;
            nd = nmx(nd, fabs(t[j] - (0.0 - lo + sl)));
            nd = nmx(nd, fabs(t[3 + j] - (hi - 0.0 + su)));
            nd = nmx(nd, fabs(t[6 + j] - sl));
            nd = nmx(nd, fabs(t[8 + j] - su));
        }
        if (k >= 1) {
            const double x6 = ATS(P.xb, k * 7 + 6);
            nd = nmx(nd, fabs(t[2] - (0.0 - (o.lbx - x6))));
            nd = nmx(nd, fabs(t[5] - ((o.ubx - x6) - 0.0)));
        }
#pragma unroll
        for (int a = 0; a < 7; a++) nb = nmx(nb, fabs(lin_get(P, k, LIN_b + a, i)));
#pragma unroll
        for (int c = 0; c < NC; c++) {
            if ((c == 2 || c == 5) && k == 0) continue;
            nm = nmx(nm, fabs(lam[c] * t[c]));
        }
        if (k >= 1) {
#pragma unroll
            for (int a = 0; a < 7; a++) {
                double g = lin_get(P, k, LIN_q + a, i) - pim[a];
                if (a < 2) g += pi[a];
                else {
#pragma unroll
                    for (int r = 0; r < 6; r++) g += lin_get(P, k, LIN_A + r * 5 + (a - 2), i) * pi[r];
                    if (a == 6) g += pi[6] - lam[2] + lam[5];
                }
                ng = nmx(ng, fabs(g));
            }
        }
#pragma unroll
        for (int a = 0; a < 7; a++) pim[a] = pi[a];
    }
    {
#pragma unroll
        for (int a = 0; a < 7; a++) ng = nmx(ng, fabs(lin_get(P, N, LIN_q + a, i) - pim[a]));
    }
    ATS(P.nlp_res, 0) = ng; ATS(P.nlp_res, 1) = nb; ATS(P.nlp_res, 2) = nd; ATS(P.nlp_res, 3) = nm;
    if (ng < tol0 && nb < tol1 && nd < tol2 && nm < tol3) {
        P.sqp_status[i] = 0; P.sqp_iter[i] = it; P.status[i] = 0; P.lin_bad[i] = 2;
        return;
    }
    P.sqp_iter[i] = it + 1;
    atomicAdd(active, 1);
}

// after the last iteration: instances still running either failed in their last QP or hit max_iter (ACADOS_MAXITER)
__global__ void sqp_finalize_kernel(const Params P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.B || P.lin_bad[i] == 2) return;
    const int st = P.status[i];
    P.sqp_status[i] = st != 0 ? st : 2;
    if (st != 0) P.sqp_iter[i] -= 1;          // the failed last iteration does not count
    P.status[i] = P.sqp_status[i];
}

void launch_nlp_res(const Params &P, int it, const double tol[4], int *active, cudaStream_t s)
{
    nlp_res_kernel<<<(P.B + 127) / 128, 128, 0, s>>>(P, it, tol[0], tol[1], tol[2], tol[3], active);
}
void launch_sqp_finalize(const Params &P, cudaStream_t s) { sqp_finalize_kernel<<<(P.B + 127) / 128, 128, 0, s>>>(P); }
