// util.cu -- layout kernels (instance-major host layout <-> SoA device layout) and the FP64 peak probe.
#include "common.cuh"

// [B][F] (instance-major, what the host hands over) -> [F][Bp] (SoA). 32x32 tiles through shared memory so both
// sides are coalesced.  HBM-bound byte shuffling; grid = tiles.
__global__ void transpose_in_kernel(const double *__restrict__ src, double *__restrict__ dst, int B, int Bp, int F)
{
    __shared__ double tile[32][33];
    const int i0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, f = f0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < B && f < F) ? src[(size_t)i * F + f] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int f = f0 + r, i = i0 + threadIdx.x;
        if (f < F && i < Bp) dst[(size_t)f * Bp + i] = tile[threadIdx.x][r];
    }
}

__global__ void transpose_out_kernel(const double *__restrict__ src, double *__restrict__ dst, int B, int Bp, int F)
{
    __shared__ double tile[32][33];
    const int i0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int f = f0 + r, i = i0 + threadIdx.x;
        tile[r][threadIdx.x] = (f < F && i < Bp) ? src[(size_t)f * Bp + i] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, f = f0 + threadIdx.x;
        if (i < B && f < F) dst[(size_t)i * F + f] = tile[threadIdx.x][r];
    }
}

__global__ void bcast_rows_kernel(const double *__restrict__ src, double *__restrict__ dst, int Bp, int F)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Bp) return;
    const double v = src[i];
    for (int f = blockIdx.y; f < F; f += gridDim.y) dst[(size_t)f * Bp + i] = v;
}

__global__ void fill_kernel(double *dst, size_t n, double v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = v;
}

void launch_transpose_in(const double *src, double *dst, int B, int Bp, int F, cudaStream_t s)
{
    dim3 grid((Bp + 31) / 32, (F + 31) / 32), block(32, 8);
    transpose_in_kernel<<<grid, block, 0, s>>>(src, dst, B, Bp, F);
}
void launch_transpose_out(const double *src, double *dst, int B, int Bp, int F, cudaStream_t s)
{
    dim3 grid((Bp + 31) / 32, (F + 31) / 32), block(32, 8);
    transpose_out_kernel<<<grid, block, 0, s>>>(src, dst, B, Bp, F);
}
void launch_bcast_rows(const double *src, double *dst, int Bp, int F, cudaStream_t s)
{
    dim3 grid((Bp + 255) / 256, F < 64 ? F : 64);
    bcast_rows_kernel<<<grid, 256, 0, s>>>(src, dst, Bp, F);
}
void launch_fill(double *dst, size_t n, double v, cudaStream_t s)
{
    fill_kernel<<<148 * 8, 256, 0, s>>>(dst, n, v);
}

// ---- single-instance fast path of the acados shim: one packed block in, one packed block out ------------------------
// in : [x0 7 | yref 9N+7 | p N | kappa N | x (N+1)*7 | u 2N]   (kappa / iterate sections used on demand)
// out: [x (N+1)*7 | u 2N | res 4 | status qp_status qp_iter (as doubles)]
__global__ void capsule_scatter_kernel(const Params P, const double *__restrict__ in, int with_kappa, int with_iterate)
{
    const int N = P.o.N, Bp = P.Bp;
    const int nyr = 9 * N + 7, nx = (N + 1) * 7, nu = 2 * N;
    const int tot = 7 + nyr + N + N + nx + nu;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += gridDim.x * blockDim.x) {
        int r = e;
        if (r < 7) { ((double *)P.x0)[(size_t)r * Bp] = in[e]; continue; }
        r -= 7;
        if (r < nyr) { ((double *)P.yref)[(size_t)r * Bp] = in[e]; continue; }
        r -= nyr;
        if (r < N) { ((double *)P.p)[(size_t)r * Bp] = in[e]; continue; }
        r -= N;
        if (r < N) { if (with_kappa) ((double *)P.kappa)[(size_t)r * Bp] = in[e]; continue; }
        r -= N;
        if (!with_iterate) continue;
        if (r < nx) { P.xb[(size_t)r * Bp] = in[e]; continue; }
        r -= nx;
        P.ub[(size_t)r * Bp] = in[e];
    }
}
__global__ void capsule_gather_kernel(const Params P, double *__restrict__ out)
{
    const int N = P.o.N, Bp = P.Bp;
    const int nx = (N + 1) * 7, nu = 2 * N;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nx + nu + 7; e += gridDim.x * blockDim.x) {
        int r = e;
        if (r < nx) { out[e] = P.xb[(size_t)r * Bp]; continue; }
        r -= nx;
        if (r < nu) { out[e] = P.ub[(size_t)r * Bp]; continue; }
        r -= nu;
        if (r < 4) { out[e] = P.res_out[(size_t)r * Bp]; continue; }
        r -= 4;
        out[e] = (double)((r == 0) ? P.status[0] : (r == 1) ? P.qp_status[0] : P.qp_iter[0]);
    }
}
void launch_capsule_scatter(const Params &P, const double *in, int with_kappa, int with_iterate, cudaStream_t s)
{
    capsule_scatter_kernel<<<4, 256, 0, s>>>(P, in, with_kappa, with_iterate);
}
void launch_capsule_gather(const Params &P, double *out, cudaStream_t s) { capsule_gather_kernel<<<2, 256, 0, s>>>(P, out); }

// ---- GP ensemble: nearest-centroid model choice per instance (GPEnsemble.select_gp, model_fitting/gp.py:738-770) -----
// z = B_z [x; u] from the query state / input (SoA rows), distance sqrt(sum (z - c)^2) like the reference, first minimum
__global__ void gp_select_kernel(const Params P, const double *__restrict__ xq, const double *__restrict__ uq, int *sel)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.B) return;
    const int dz = P.o.gp_dz, K = P.gp.n_models, Bp = P.Bp;
    double z[ADMPC_DZMAX];
    for (int d = 0; d < dz; d++) {
        const int f = P.o.gp_feat[d];
        z[d] = (f < 7) ? xq[(size_t)f * Bp + i] : (uq ? uq[(size_t)(f - 7) * Bp + i] : 0.0);
    }
    int best = 0;
    double bd = 0.0;
    for (int c = 0; c < K; c++) {
        double d2 = 0.0;
        for (int d = 0; d < dz; d++) { const double t = z[d] - P.gp.centroids[c * dz + d]; d2 = d2 + t * t; }
        const double dist = sqrt(d2);
        if (c == 0 || dist < bd) { bd = dist; best = c; }
    }
    sel[i] = best;
}
void launch_gp_select(const Params &P, const double *xq, const double *uq, int *sel, cudaStream_t s)
{
    if (P.gp.n_models <= 1) { cudaMemsetAsync(sel, 0, (size_t)P.Bp * sizeof(int), s); return; }
    gp_select_kernel<<<(P.B + 127) / 128, 128, 0, s>>>(P, xq, uq, sel);
}

// ---- FP64 peak probe: 8 independent DFMA chains per thread, full occupancy -------------------------------
// NINT > 0 adds NINT independent integer multiply-adds per 8 DFMAs (issue-slot pressure probe: how much FP64
// throughput survives when the scheduler also has address / index arithmetic to issue).
template <int NINT>
__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b, int ia)
{
    double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    int q[8];
#pragma unroll
    for (int j = 0; j < 8; j++) q[j] = threadIdx.x + j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
            r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
#pragma unroll
            for (int j = 0; j < NINT; j++) q[j] = q[j] * ia + u;
        }
    }
    int qs = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) qs += q[j];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7 + (double)qs;
}

// three distinct register operands per DFMA (the shape of the kernels' inner loops: operand-fetch pressure probe)
__global__ void __launch_bounds__(256) dfma3_kernel(double *out, int iters, double a, double b)
{
    double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    double s[8], t[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { s[j] = a + 1e-9 * (threadIdx.x + j); t[j] = b * (1 + j + threadIdx.x); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            r0 = fma(r0, s[(u + 0) & 7], t[(u + 1) & 7]); r1 = fma(r1, s[(u + 1) & 7], t[(u + 2) & 7]);
            r2 = fma(r2, s[(u + 2) & 7], t[(u + 3) & 7]); r3 = fma(r3, s[(u + 3) & 7], t[(u + 4) & 7]);
            r4 = fma(r4, s[(u + 4) & 7], t[(u + 5) & 7]); r5 = fma(r5, s[(u + 5) & 7], t[(u + 6) & 7]);
            r6 = fma(r6, s[(u + 6) & 7], t[(u + 7) & 7]); r7 = fma(r7, s[(u + 7) & 7], t[(u + 0) & 7]);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
}

double run_fp64_peak(int device, int nint)
{
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    double *out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        if (nint == 16) dfma3_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        else if (nint <= 0) dfma_kernel<0><<<blocks, threads>>>(out, iters, 0.999999, 1e-9, 3);
        else if (nint <= 2) dfma_kernel<2><<<blocks, threads>>>(out, iters, 0.999999, 1e-9, 3);
        else if (nint <= 4) dfma_kernel<4><<<blocks, threads>>>(out, iters, 0.999999, 1e-9, 3);
        else dfma_kernel<8><<<blocks, threads>>>(out, iters, 0.999999, 1e-9, 3);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return best;
}
