// common.cuh -- shared definitions of the B200 batched SQP-RTI solver (device layouts, parameter block).
//
// HBM layout: every per-instance quantity is stored structure-of-arrays, "row-major over instances":
//     value(row, i) at base[row * Bp + i],   i = instance (fastest), Bp = batch padded to a multiple of 32.
// A warp therefore always touches 32 consecutive doubles (256 B, two full 128-B lines) per row.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/admpc.h"

#define NX 7
#define NU 2
#define NC 10

// rows of one stage of the linearisation block (written by prepare, read by the QP kernel)
//   A: rows 0..5 x cols 2..6 of A_k (cols 0,1 are e0,e1 and row 6 is e6 for this model: SURVEY Appendix B)
//   B: rows 0..5 x 2 of B_k (row 6 is [0, dt])
#define LIN_A 0
#define LIN_B 30
#define LIN_b 42
#define LIN_q 49
#define LIN_r 56
#define LIN_ROWS 58
// instance-major twin of the block, lin_im[k][i][LIM_STRIDE] (one 544-byte record per instance and stage, pulled into shared
// memory by one TMA bulk copy): M = [B | A(:,2:7)] column-major c*6 + r, then b, q, r and the linearisation point x_k, u_k
#define LIM_M 0
#define LIM_B 42
#define LIM_Q 49
#define LIM_R 56
#define LIM_X 58
#define LIM_U 65
#define LIM_STRIDE 68
__host__ __device__ __forceinline__ constexpr int lim_of_row(int row)      // SoA row -> offset inside the record
{
    return (row < LIN_B) ? (2 + row % 5) * 6 + row / 5 : (row < LIN_b) ? ((row - LIN_B) % 2) * 6 + (row - LIN_B) / 2 : row;
}

// entries of the 2^(j/GP_TAB) table of the device exp2 (model.cuh), stored behind the last GP output in the blob
#ifndef GP_TAB_BITS
#define GP_TAB_BITS 8
#endif
#define GP_TAB (1 << GP_TAB_BITS)
struct GpDev {
    // packed, 16-byte aligned GP block in HBM, staged to shared memory by one TMA bulk copy per CTA:
    //   per output j: pts[M][dz+2] = {log2e*X_id/ell_d^2 (dz), -0.5*log2e*sum_d X_id^2/ell_d^2, sigma_f*alpha_i},
    //   then w[dz] = 1/ell^2, then y_mean (padded to an even count)
    const double *blob;
    int bytes;          // multiple of 16
    int stride_out;     // doubles per output block
    // ensemble (GPEnsemble, gp.py:536-770): n_models cluster models back to back, model_doubles apart; the table of the
    // device exp2 follows the last model.  centroids [n_models][dz] in HBM for the per-instance nearest-centroid choice.
    int n_models;
    int model_doubles;
    const double *centroids;
};

// inequality rows per stage of the configured constraint set = stride of lam / t (admpc.h con_set)
__host__ __device__ __forceinline__ int con_rows(const admpc_opts &o) { return o.con_set == 1 ? 12 : NC; }

#define DL_ROWS 79
struct Params {
    admpc_opts o;
    GpDev gp;
    int B, Bp;
    // inputs
    const double *x0, *yref, *p, *gps;
    // iterate
    double *xb, *ub, *pib, *lamb, *tb, *slb, *sub;
    // linearisation: SoA rows (lin) or instance-major records (lin_im), exactly one of them per handle
    double *lin;
    double *lin_im;
    double *gpr;         // GP mean / feature gradient at the 4 RK4 stage points of every interval: [N*4*nout*(1+dz)][Bp] (prepare.cu)
    // QP solution (delta form) + workspace
    double *dx, *du, *pi, *lam, *t, *sl, *su;
    double *rgu, *rgx, *rgsl, *rgsu, *rb, *rd, *rm;
    double *K, *Ginv, *P, *Pb, *kf, *pv;
    double *ddu, *ddx, *dpi, *dlam, *dt, *dsl, *dsu;
    int *status, *qp_status, *qp_iter, *lin_bad;
    double *res_out;     // [4][Bp] final residual norms
    // full SQP mode (sqp.cu): lin_bad == 2 marks an instance that has finished and is skipped by prepare / QP kernels
    // Frenet variant (frenet.cu): dense linearisation [k][DL_ROWS][Bp] (A 49, B 14, b 7, q 7, r 2), curvature [N][Bp]
    double *lin_d;
    int skip_lin_d;      // 1: this solve runs the tensor-core QP kernel off the instance-major records, nothing reads the SoA rows
    const double *kappa;
    const double *kap_sp;    // Frenet variant: kappa(s) as piecewise cubics [(K+1) breaks | K x 4 coefficients][Bp], kap_K pieces (0: off)
    int kap_K;
    int *sqp_status, *sqp_iter;
    const int *gp_sel;   // [Bp] cluster model of every instance (GP ensemble; zeros for a single model)
    double *nlp_res;     // [4][Bp] NLP KKT residual norms of the last check
    // fused solution gather (multi-GPU): instance-major u [B][2N], x [B][7(N+1)], status [B] of this rank inside the
    // root's gathered block (peer memory when the root is another GPU); null = off
    double *gat_u, *gat_x;
    int *gat_st;
};

// one entry of the linearisation of stage k, whichever layout the handle uses (kernels off the hot path)
__device__ __forceinline__ double lin_get(const Params &P, int k, int row, int i)
{
    return P.lin_im ? P.lin_im[((size_t)k * P.Bp + i) * LIM_STRIDE + lim_of_row(row)] : P.lin[((size_t)k * LIN_ROWS + row) * P.Bp + i];
}

// -DADMPC_DEBUG: device-side bounds asserts on the SoA / record indexing (compute-sanitizer is closed on the GPU pool, this
// build is its substitute; scripts/debug_build.sh, the parity suite runs under it with ADMPC_LIB=...).  A failed assert
// prints file:line and traps, which the host sees as a CUDA error on the next synchronisation.
#ifdef ADMPC_DEBUG
#include <stdio.h>
#define ADMPC_ASSERT(c) do { if (!(c)) { printf("ADMPC_ASSERT failed %s:%d: %s\n", __FILE__, __LINE__, #c); __trap(); } } while (0)
#else
#define ADMPC_ASSERT(c) ((void)0)
#endif
// checked element of an SoA interface array [rows][Bp]: row within the array's row count, instance within the padded batch
__device__ __forceinline__ size_t soa_at(int row, int rows, int i, int Bp)
{
    ADMPC_ASSERT(row >= 0 && row < rows && i >= 0 && i < Bp);
    return (size_t)row * Bp + i;
}

#define CUDA_CHECK_RET(call)                                                         \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) { admpc_set_error(#call, cudaGetErrorString(e_)); return ADMPC_E_CUDA; } \
    } while (0)

void admpc_set_error(const char *what, const char *msg);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE setting: remember the largest size configured on each
// device of the process (handles on several GPUs in one process are legal).
struct SmemGuard {
    size_t done[64] = {};
    bool need(size_t bytes)
    {
        int d = 0;
        cudaGetDevice(&d);
        d &= 63;
        if (bytes <= done[d]) return false;
        done[d] = bytes;
        return true;
    }
};

struct admpc_batch;
int admpc_batch_set_stream_level(admpc_batch *h, int level);      // api.cu, used by the chunk pipeline (pipe.cu)

// kernel launchers (defined in the .cu files)
void launch_prepare(const Params &P, cudaStream_t s);
void launch_gp_sweep(const Params &P, cudaStream_t s);   // pass 1 of a GP-augmented preparation: GP mean / gradient at the RK4 stage points -> gpr
void launch_qp(const Params &P, cudaStream_t s);
bool launch_qp_warp(const Params &P, cudaStream_t s);   // false: N > 31
bool launch_qp_mma(const Params &P, cudaStream_t s);    // one / two / four warps per instance, whole solve resident in shared memory, sweeps on the FP64 tensor cores (DMMA), false: N > 127
void launch_transpose_in(const double *src, double *dst, int B, int Bp, int F, cudaStream_t s);   // [B][F] -> [F][Bp]
void launch_transpose_out(const double *src, double *dst, int B, int Bp, int F, cudaStream_t s);  // [F][Bp] -> [B][F]
void launch_bcast_rows(const double *src, double *dst, int Bp, int F, cudaStream_t s);            // [Bp] -> [F][Bp]
void launch_update(const Params &P, cudaStream_t s);
void launch_fill(double *dst, size_t n, double v, cudaStream_t s);
void launch_blend(const Params &P, cudaStream_t s);     // p of every stage from x0's v_x (opts.blend_min / blend_max)
double run_fp64_peak(int device, int nint);
void launch_prepare_dense(const Params &P, cudaStream_t s);
void launch_capsule_scatter(const Params &P, const double *in, int with_kappa, int with_iterate, cudaStream_t s);
void launch_capsule_gather(const Params &P, double *out, cudaStream_t s);
void launch_gp_select(const Params &P, const double *xq, const double *uq, int *sel, cudaStream_t s);
void launch_qp_dense(const Params &P, cudaStream_t s);
bool launch_qp_warp_f(const Params &P, cudaStream_t s);   // Frenet structure, false: N > 63
bool launch_qp_mma_g(const Params &P, cudaStream_t s);    // Frenet variant on the FP64 tensor cores (dense column of s, both constraint sets), false: N > 127 or no instance-major records
void launch_nlp_res_dense(const Params &P, int it, const double tol[4], int *active, cudaStream_t s);
void launch_nlp_res(const Params &P, int it, const double tol[4], int *active, cudaStream_t s);
void launch_sqp_finalize(const Params &P, cudaStream_t s);
