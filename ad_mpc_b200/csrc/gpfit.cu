// gpfit.cu -- GP model update on the device: the dense linear algebra of CustomGPRegression.fit /
// _nll (reference: model_fitting/gp.py:283-289,305-311,361-363) for one output dimension:
//     K = sigma_f exp(-1/2 |x_i/l - x_j/l|^2) + sigma_n^2 I        (two-argument kernel call, gp.py:103-105)
//     L = chol(K)
//     alpha = K^-1 y                  (the vector the inference kernels consume; the reference forms inv(K) explicitly)
//     nll = sum log L_ii + 1/2 y^T alpha + M/2 log 2 pi
// The hyper-parameter search stays a host loop (L-BFGS-B like the reference, ad_mpc_b200/gpfit.py); every NLL
// evaluation and the final alpha run here.  FP64 throughout: K has condition ~ sigma_f M / sigma_n^2.
//
// The right-hand side is one more tile row below the matrix (augmented factorisation): the steps that factorise K leave
// z = L^-1 y in it, the NLL needs no substitution (sum log L_ii + z^T z / 2), alpha one backward substitution.
// Blocked right-looking Cholesky on 32x32 tiles of the padded matrix (row-major, lower triangle):
//   potrf32 (one CTA)  ->  trsm32 (one CTA per tile below the diagonal)  ->  syrk32 (one CTA per trailing tile pair).
// The trailing update is the dense GEMM of this path (M^3/3 flops, 2.7 GFLOP at M = 2000); B200 has no FP64 tcgen05
// kind and its FP64 DMMA peak equals the DFMA peak, so the tile GEMM is register-tiled DFMA (2x2 outputs per thread).
#include <math.h>
#include <vector>

#include "common.cuh"

#define TB 32

// K build: one thread per (i, j) of the padded Mp x Mp matrix; padding = identity
__global__ void gpfit_build_kernel(const double *__restrict__ Xs, int M, int Mp, int ld, int dz, double sigma_f, double sn2,
                                   double *__restrict__ A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= Mp || j >= Mp) return;
    double v;
    if (i < M && j < M) {
        double d2 = 0.0;
        for (int d = 0; d < dz; d++) { const double t = Xs[(size_t)i * dz + d] - Xs[(size_t)j * dz + d]; d2 = fma(t, t, d2); }
        v = sigma_f * exp(-0.5 * d2) + ((i == j) ? sn2 : 0.0);
    } else {
        v = (i == j) ? 1.0 : 0.0;
    }
    A[(size_t)i * ld + j] = v;
}
// the right-hand side as one more row below the matrix (row Mp of the augmented factorisation): the same trsm32 / syrk32 steps
// that factorise K leave z = L^-1 y in it -- the forward substitution comes for free
__global__ void gpfit_aug_row_kernel(double *__restrict__ A, int ld, int Mp, int M, const double *__restrict__ y)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < Mp) A[(size_t)Mp * ld + j] = (j < M) ? y[j] : 0.0;
}

// Cholesky of the diagonal tile (j, j): ONE WARP, lane r owns row r of the tile in registers; the pivot column is
// published through a 32-double shared-memory slot and read back as broadcast LDS.128 (no block barriers: the 32
// column steps are the serial part of every block step of the factorisation, so their latency is what matters).
__global__ void __launch_bounds__(32) gpfit_potrf32(double *A, int Mp, int jb, int *fail)
{
    __shared__ __align__(16) double col[2][TB];      // double-buffered: step k reads buffer k&1 while step k+1 fills the other
    __shared__ double pv[2];
    const int r = threadIdx.x;
    double *Ajj = A + (size_t)(jb * TB) * Mp + jb * TB;
    double row[TB];
#pragma unroll
    for (int c = 0; c < TB; c++) row[c] = Ajj[(size_t)r * Mp + c];
    bool bad = false;
#pragma unroll
    for (int k = 0; k < TB; k++) {
        if (r == k) pv[k & 1] = row[k];
        __syncwarp();
        const double piv = pv[k & 1];
        bad |= !(piv > 0.0);                       // not positive definite (np.linalg.LinAlgError in the reference)
        const double d = sqrt(piv);
        const double lrk = (r == k) ? d : row[k] / d;          // column k of L (rows >= k)
        row[k] = lrk;
        if (r > k) col[k & 1][r] = lrk;
        __syncwarp();
#pragma unroll
        for (int c = k + 1; c < TB; c++) {
            const double lck = col[k & 1][c];
            if (r >= c) row[c] = fma(-lrk, lck, row[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < TB; c++) Ajj[(size_t)r * Mp + c] = (c <= r) ? row[c] : 0.0;
    if (bad && r == 0) *fail = 1;
}

// X = A_ij L_jj^-T for the tiles below the diagonal: one thread per row of the tile (32 rows), 4 tiles per CTA
__global__ void __launch_bounds__(128) gpfit_trsm32(double *A, int Mp, int jb, int nb)
{
    __shared__ double Lt[TB][TB + 1];
    for (int e = threadIdx.x; e < TB * TB; e += blockDim.x) {
        const int r = e / TB, c = e % TB;
        Lt[r][c] = A[(size_t)(jb * TB + r) * Mp + jb * TB + c];
    }
    __syncthreads();
    const int tile = jb + 1 + blockIdx.x * 4 + (threadIdx.x >> 5);
    if (tile >= nb) return;
    const int r = threadIdx.x & 31;
    double *row = A + (size_t)(tile * TB + r) * Mp + jb * TB;
    double x[TB];
#pragma unroll
    for (int c = 0; c < TB; c++) x[c] = row[c];
#pragma unroll
    for (int c = 0; c < TB; c++) {
        double v = x[c];
#pragma unroll
        for (int m = 0; m < c; m++) v = fma(-x[m], Lt[c][m], v);
        x[c] = v / Lt[c][c];
    }
#pragma unroll
    for (int c = 0; c < TB; c++) row[c] = x[c];
}

// trailing update C_ik -= A_ij A_kj^T for i >= k > j: one CTA (16x16 threads, 2x2 outputs each) per tile pair
__global__ void __launch_bounds__(256) gpfit_syrk32(double *A, int Mp, int jb, int nb)
{
    const int ti = jb + 1 + blockIdx.y, tk = jb + 1 + blockIdx.x;
    if (ti >= nb || tk > ti) return;
    __shared__ double Ai[TB][TB + 1], Ak[TB][TB + 1];
    const int tx = threadIdx.x, ty = threadIdx.y, t = ty * 16 + tx;
    for (int e = t; e < TB * TB; e += 256) {
        const int r = e / TB, c = e % TB;
        Ai[r][c] = A[(size_t)(ti * TB + r) * Mp + jb * TB + c];
        Ak[r][c] = A[(size_t)(tk * TB + r) * Mp + jb * TB + c];
    }
    __syncthreads();
    double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
    const int r0 = ty * 2, k0 = tx * 2;
#pragma unroll 8
    for (int m = 0; m < TB; m++) {
        const double a0 = Ai[r0][m], a1 = Ai[r0 + 1][m], b0 = Ak[k0][m], b1 = Ak[k0 + 1][m];
        c00 = fma(a0, b0, c00); c01 = fma(a0, b1, c01); c10 = fma(a1, b0, c10); c11 = fma(a1, b1, c11);
    }
    double *C = A + (size_t)(ti * TB + r0) * Mp + tk * TB + k0;
    C[0] -= c00; C[1] -= c01; C[Mp] -= c10; C[Mp + 1] -= c11;
}

// y <- L^-1 y, then y <- L^-T y (blocked by 32, single CTA of 32 warps, left-looking), plus log-determinant and
// y0^T alpha.  Forward: warp w owns row w of the current block row and sweeps its already-solved columns with coalesced
// 256-byte reads.  Backward: lane c owns column c of the current block column, the warps stride over the rows below it
// (again one coalesced 256-byte read per row), partial sums meet in shared memory.  The 32x32 diagonal solves stay on
// warp 0.  (the Mp x Mp factor may sit in the top-left corner of a larger matrix: row stride ldA)
// do_fwd / do_bwd select the two substitutions (admpc_gp_fit gets z from the augmented factorisation and needs the backward one only,
// and only when alpha is asked for)
__global__ void __launch_bounds__(1024) gpfit_solve_kernel(const double *__restrict__ A, int M, int Mp, int ldA, double *y,
                                                            const double *__restrict__ y0, double *out2, int do_fwd, int do_bwd)
{
    __shared__ double xb[TB];
    __shared__ double part[TB][TB + 1];
    __shared__ double red[32];
    const int nb = Mp / TB, t = threadIdx.x, w = t >> 5, l = t & 31;
    // forward: L z = y
    for (int b = 0; b < (do_fwd ? nb : 0); b++) {
        {
            const double *row = A + (size_t)(b * TB + w) * ldA;
            double s = 0.0;
            for (int c = l; c < b * TB; c += 32) s = fma(row[c], y[c], s);
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (l == 0) xb[w] = y[b * TB + w] - s;
        }
        __syncthreads();
        if (t < 32) {
            // diagonal block: row t in registers first (all 32 loads in flight at once), then 32 shuffle steps
            double Lr[TB];
#pragma unroll
            for (int c = 0; c < TB; c++) Lr[c] = A[(size_t)(b * TB + t) * ldA + b * TB + c];
            double dinv = 0.0;
#pragma unroll
            for (int c = 0; c < TB; c++) if (t == c) dinv = 1.0 / Lr[c];
            double v = xb[t];
#pragma unroll
            for (int c = 0; c < TB; c++) {
                const double xc = __shfl_sync(0xffffffffu, v * dinv, c);
                if (t == c) v = xc;
                else if (t > c) v = fma(-Lr[c], xc, v);
            }
            y[b * TB + t] = v;
        }
        __syncthreads();
    }
    // backward: L^T alpha = z
    for (int b = (do_bwd ? nb - 1 : -1); b >= 0; b--) {
        {
            double s = 0.0;
            for (int i = (b + 1) * TB + w; i < Mp; i += 32) s = fma(A[(size_t)i * ldA + b * TB + l], y[i], s);
            part[w][l] = s;
        }
        __syncthreads();
        if (t < 32) {
            double s = 0.0;
#pragma unroll 8
            for (int q = 0; q < TB; q++) s += part[q][t];
            double v = y[b * TB + t] - s;
            double Lc[TB];                       // column t of the diagonal block (coalesced across lanes)
#pragma unroll
            for (int c = 0; c < TB; c++) Lc[c] = A[(size_t)(b * TB + c) * ldA + b * TB + t];
            double dinv = 0.0;
#pragma unroll
            for (int c = 0; c < TB; c++) if (t == c) dinv = 1.0 / Lc[c];
#pragma unroll
            for (int c = TB - 1; c >= 0; c--) {
                const double xc = __shfl_sync(0xffffffffu, v * dinv, c);
                if (t == c) v = xc;
                else if (t < c) v = fma(-Lc[c], xc, v);
            }
            y[b * TB + t] = v;
        }
        __syncthreads();
    }
    // reductions: sum_i log L_ii (i < M) and y0^T alpha
    double ld = 0.0, qa = 0.0;
    for (int i = t; i < M; i += blockDim.x) { ld += log(A[(size_t)i * ldA + i]); qa = fma(y0[i], y[i], qa); }
    for (int pass = 0; pass < 2; pass++) {
        double v = pass ? qa : ld;
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((t & 31) == 0) red[t >> 5] = v;
        __syncthreads();
        if (t < 32) {
            v = red[t];
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (t == 0) out2[pass] = v;
        }
        __syncthreads();
    }
}

extern "C" int admpc_gp_fit(int device, int M, int dz, const double *X, const double *y, const double *ell,
                            double sigma_f, double sigma_n, double *alpha_out, double *nll_out, float *ms_out)
{
    if (M < 1 || dz < 1 || dz > ADMPC_DZMAX || !X || !y || !ell) { admpc_set_error("admpc_gp_fit", "bad argument"); return ADMPC_E_ARG; }
    int ndev = 0;
    CUDA_CHECK_RET(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { admpc_set_error("admpc_gp_fit", "no such CUDA device"); return ADMPC_E_CUDA; }
    CUDA_CHECK_RET(cudaSetDevice(device));
    const int Mp = (M + TB - 1) / TB * TB, nb = Mp / TB;
    const int Mt = Mp + TB, nbt = Mt / TB;          // one more tile row: the right-hand side (augmented factorisation)
    std::vector<double> Xs((size_t)M * dz), yp(Mp, 0.0);
    for (int i = 0; i < M; i++)
        for (int d = 0; d < dz; d++) Xs[(size_t)i * dz + d] = X[(size_t)i * dz + d] / ell[d];     // cdist(x/l, x/l), gp.py:103
    for (int i = 0; i < M; i++) yp[i] = y[i];
    double *dXs = nullptr, *dA = nullptr, *dy = nullptr, *dy0 = nullptr, *dout = nullptr;
    int *dfail = nullptr;
    cudaStream_t s;
    CUDA_CHECK_RET(cudaStreamCreate(&s));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rc = 0;
    do {
#define GP_TRY(call) { cudaError_t e_ = (call); if (e_ != cudaSuccess) { admpc_set_error(#call, cudaGetErrorString(e_)); rc = ADMPC_E_CUDA; break; } }
        GP_TRY(cudaMalloc(&dXs, Xs.size() * sizeof(double)));
        GP_TRY(cudaMalloc(&dA, (size_t)Mt * Mt * sizeof(double)));
        GP_TRY(cudaMalloc(&dy, Mp * sizeof(double)));
        GP_TRY(cudaMalloc(&dy0, Mp * sizeof(double)));
        GP_TRY(cudaMalloc(&dout, 2 * sizeof(double)));
        GP_TRY(cudaMalloc(&dfail, sizeof(int)));
        GP_TRY(cudaMemcpyAsync(dXs, Xs.data(), Xs.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        GP_TRY(cudaMemcpyAsync(dy, yp.data(), Mp * sizeof(double), cudaMemcpyHostToDevice, s));
        GP_TRY(cudaMemcpyAsync(dy0, yp.data(), Mp * sizeof(double), cudaMemcpyHostToDevice, s));
        GP_TRY(cudaMemsetAsync(dfail, 0, sizeof(int), s));
        GP_TRY(cudaEventRecord(e0, s));
        dim3 bb(16, 16), bg((Mp + 15) / 16, (Mp + 15) / 16);
        GP_TRY(cudaMemsetAsync(dA + (size_t)Mp * Mt, 0, (size_t)TB * Mt * sizeof(double), s));     // the extra tile row
        gpfit_build_kernel<<<bg, bb, 0, s>>>(dXs, M, Mp, Mt, dz, sigma_f, sigma_n * sigma_n, dA);
        gpfit_aug_row_kernel<<<(Mp + 255) / 256, 256, 0, s>>>(dA, Mt, Mp, M, dy0);
        for (int j = 0; j < nb; j++) {                       // the first Mp / 32 block columns: K = L L^T, row Mp <- (L^-1 y)^T
            gpfit_potrf32<<<1, 32, 0, s>>>(dA, Mt, j, dfail);
            const int rem = nbt - j - 1;
            gpfit_trsm32<<<(rem + 3) / 4, 128, 0, s>>>(dA, Mt, j, nbt);
            gpfit_syrk32<<<dim3(rem, rem), dim3(16, 16), 0, s>>>(dA, Mt, j, nbt);
        }
        // z = L^-1 y sits in row Mp; nll needs sum log L_ii + z^T z / 2, alpha (only when asked for) the backward substitution
        GP_TRY(cudaMemcpyAsync(dy, dA + (size_t)Mp * Mt, (size_t)Mp * sizeof(double), cudaMemcpyDeviceToDevice, s));
        gpfit_solve_kernel<<<1, 1024, 0, s>>>(dA, M, Mp, Mt, dy, alpha_out ? dy0 : dy, dout, 0, alpha_out ? 1 : 0);
        GP_TRY(cudaEventRecord(e1, s));
        GP_TRY(cudaGetLastError());
        double out2[2];
        int fail = 0;
        GP_TRY(cudaMemcpyAsync(out2, dout, sizeof out2, cudaMemcpyDeviceToHost, s));
        GP_TRY(cudaMemcpyAsync(&fail, dfail, sizeof fail, cudaMemcpyDeviceToHost, s));
        if (alpha_out) GP_TRY(cudaMemcpyAsync(alpha_out, dy, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, s));
        GP_TRY(cudaStreamSynchronize(s));
        if (fail) { admpc_set_error("admpc_gp_fit", "kernel matrix is not positive definite"); rc = ADMPC_E_ARG; break; }
        if (nll_out) *nll_out = out2[0] + 0.5 * out2[1] + 0.5 * M * log(2.0 * 3.141592653589793);
        if (ms_out) cudaEventElapsedTime(ms_out, e0, e1);
#undef GP_TRY
    } while (0);
    cudaFree(dXs); cudaFree(dA); cudaFree(dy); cudaFree(dy0); cudaFree(dout); cudaFree(dfail);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaStreamDestroy(s);
    return rc;
}


// ---------------------------------------------------------------------------------------------- posterior variance ---
// GP posterior at n test points (reference: CustomGPRegression.predict(x, return_cov=True), gp.py:402-441):
//     mu  = k_s K^-1 y + y_mean ,   cov = k(x*, x*) + 1e-8 I - k_s K^-1 k_s^T
// computed WITHOUT forming K^-1: the blocked Cholesky above is run on the augmented matrix
//     [ K      k_s^T        ]
//     [ k_s    k_ss + 1e-8 I ]
// for the first Mp/32 block columns only; what is left in the trailing n x n block is the Schur complement = cov
// (same potrf32 / trsm32 / syrk32 kernels, the test rows are just more tiles below the diagonal).
__global__ void gpfit_build_aug_kernel(const double *__restrict__ Xs, const double *__restrict__ Xt, int M, int Mp, int n,
                                       int Mt, int dz, double sigma_f, double sn2, double *__restrict__ A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= Mt || j >= Mt) return;
    const bool itr = i < M, jtr = j < M, its = i >= Mp && i < Mp + n, jts = j >= Mp && j < Mp + n;
    double v = (i == j) ? 1.0 : 0.0;                       // padding rows / columns: identity
    if ((itr || its) && (jtr || jts)) {
        const double *pi = itr ? Xs + (size_t)i * dz : Xt + (size_t)(i - Mp) * dz;
        const double *pj = jtr ? Xs + (size_t)j * dz : Xt + (size_t)(j - Mp) * dz;
        double d2 = 0.0;
        for (int d = 0; d < dz; d++) { const double t = pi[d] - pj[d]; d2 = fma(t, t, d2); }
        v = sigma_f * exp(-0.5 * d2);
        if (i == j) v += itr ? sn2 : 1e-8;
    }
    A[(size_t)i * Mt + j] = v;
}

// mu_i = sum_j k(x*_i, x_j) alpha_j + y_mean : one warp per test point
__global__ void gpfit_mean_kernel(const double *__restrict__ Xs, const double *__restrict__ Xt, const double *__restrict__ alpha,
                                  int M, int n, int dz, double sigma_f, double y_mean, double *__restrict__ mu)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, l = threadIdx.x & 31;
    if (w >= n) return;
    double s = 0.0;
    for (int j = l; j < M; j += 32) {
        double d2 = 0.0;
        for (int d = 0; d < dz; d++) { const double t = Xt[(size_t)w * dz + d] - Xs[(size_t)j * dz + d]; d2 = fma(t, t, d2); }
        s = fma(sigma_f * exp(-0.5 * d2), alpha[j], s);
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (l == 0) mu[w] = s + y_mean;
}

// variance = diagonal of the Schur complement; optional full covariance (lower triangle mirrored)
__global__ void gpfit_cov_out_kernel(const double *__restrict__ A, int Mp, int n, int Mt, double *__restrict__ var,
                                     double *__restrict__ cov)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= n || j >= n) return;
    const int r = i > j ? i : j, c = i > j ? j : i;
    const double v = A[(size_t)(Mp + r) * Mt + Mp + c];
    if (cov) cov[(size_t)i * n + j] = v;
    if (i == j) var[i] = v;
}

extern "C" int admpc_gp_predict(int device, int M, int dz, const double *X, const double *y, const double *ell, double sigma_f,
                                double sigma_n, double y_mean, int n, const double *Xtest, double *mu_out, double *var_out,
                                double *cov_out)
{
    if (M < 1 || n < 1 || dz < 1 || dz > ADMPC_DZMAX || !X || !y || !ell || !Xtest) { admpc_set_error("admpc_gp_predict", "bad argument"); return ADMPC_E_ARG; }
    int ndev = 0;
    CUDA_CHECK_RET(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { admpc_set_error("admpc_gp_predict", "no such CUDA device"); return ADMPC_E_CUDA; }
    CUDA_CHECK_RET(cudaSetDevice(device));
    const int Mp = (M + TB - 1) / TB * TB, np_ = (n + TB - 1) / TB * TB, Mt = Mp + np_, nbM = Mp / TB, nb = Mt / TB;
    std::vector<double> Xs((size_t)M * dz), Xt((size_t)n * dz), yp(Mp, 0.0);
    for (int i = 0; i < M; i++)
        for (int d = 0; d < dz; d++) Xs[(size_t)i * dz + d] = X[(size_t)i * dz + d] / ell[d];
    for (int i = 0; i < n; i++)
        for (int d = 0; d < dz; d++) Xt[(size_t)i * dz + d] = Xtest[(size_t)i * dz + d] / ell[d];
    for (int i = 0; i < M; i++) yp[i] = y[i];
    double *dXs = nullptr, *dXt = nullptr, *dA = nullptr, *dy = nullptr, *dy0 = nullptr, *dout = nullptr, *dmu = nullptr, *dvar = nullptr, *dcov = nullptr;
    int *dfail = nullptr;
    cudaStream_t s;
    CUDA_CHECK_RET(cudaStreamCreate(&s));
    int rc = 0;
    do {
#define GP_TRY(call) { cudaError_t e_ = (call); if (e_ != cudaSuccess) { admpc_set_error(#call, cudaGetErrorString(e_)); rc = ADMPC_E_CUDA; break; } }
        GP_TRY(cudaMalloc(&dXs, Xs.size() * sizeof(double)));
        GP_TRY(cudaMalloc(&dXt, Xt.size() * sizeof(double)));
        GP_TRY(cudaMalloc(&dA, (size_t)Mt * Mt * sizeof(double)));
        GP_TRY(cudaMalloc(&dy, Mp * sizeof(double)));
        GP_TRY(cudaMalloc(&dy0, Mp * sizeof(double)));
        GP_TRY(cudaMalloc(&dout, 2 * sizeof(double)));
        GP_TRY(cudaMalloc(&dmu, n * sizeof(double)));
        GP_TRY(cudaMalloc(&dvar, n * sizeof(double)));
        if (cov_out) GP_TRY(cudaMalloc(&dcov, (size_t)n * n * sizeof(double)));
        GP_TRY(cudaMalloc(&dfail, sizeof(int)));
        GP_TRY(cudaMemcpyAsync(dXs, Xs.data(), Xs.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        GP_TRY(cudaMemcpyAsync(dXt, Xt.data(), Xt.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        GP_TRY(cudaMemcpyAsync(dy, yp.data(), Mp * sizeof(double), cudaMemcpyHostToDevice, s));
        GP_TRY(cudaMemcpyAsync(dy0, yp.data(), Mp * sizeof(double), cudaMemcpyHostToDevice, s));
        GP_TRY(cudaMemsetAsync(dfail, 0, sizeof(int), s));
        dim3 bb(16, 16), bg((Mt + 15) / 16, (Mt + 15) / 16);
        gpfit_build_aug_kernel<<<bg, bb, 0, s>>>(dXs, dXt, M, Mp, n, Mt, dz, sigma_f, sigma_n * sigma_n, dA);
        for (int j = 0; j < nbM; j++) {                      // eliminate the training block only
            gpfit_potrf32<<<1, 32, 0, s>>>(dA, Mt, j, dfail);
            const int rem = nb - j - 1;
            gpfit_trsm32<<<(rem + 3) / 4, 128, 0, s>>>(dA, Mt, j, nb);
            gpfit_syrk32<<<dim3(rem, rem), dim3(16, 16), 0, s>>>(dA, Mt, j, nb);
        }
        gpfit_solve_kernel<<<1, 1024, 0, s>>>(dA, M, Mp, Mt, dy, dy0, dout, 1, 1);  // alpha = K^-1 y
        gpfit_mean_kernel<<<(n * 32 + 127) / 128, 128, 0, s>>>(dXs, dXt, dy, M, n, dz, sigma_f, y_mean, dmu);
        dim3 cg((n + 15) / 16, (n + 15) / 16);
        gpfit_cov_out_kernel<<<cg, bb, 0, s>>>(dA, Mp, n, Mt, dvar, dcov);
        GP_TRY(cudaGetLastError());
        int fail = 0;
        GP_TRY(cudaMemcpyAsync(&fail, dfail, sizeof fail, cudaMemcpyDeviceToHost, s));
        if (mu_out) GP_TRY(cudaMemcpyAsync(mu_out, dmu, n * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (var_out) GP_TRY(cudaMemcpyAsync(var_out, dvar, n * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (cov_out) GP_TRY(cudaMemcpyAsync(cov_out, dcov, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, s));
        GP_TRY(cudaStreamSynchronize(s));
        if (fail) { admpc_set_error("admpc_gp_predict", "kernel matrix is not positive definite"); rc = ADMPC_E_ARG; break; }
#undef GP_TRY
    } while (0);
    cudaFree(dXs); cudaFree(dXt); cudaFree(dA); cudaFree(dy); cudaFree(dy0); cudaFree(dout); cudaFree(dmu); cudaFree(dvar); cudaFree(dcov); cudaFree(dfail);
    cudaStreamDestroy(s);
    return rc;
}
