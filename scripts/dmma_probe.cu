// dmma_probe.cu -- microbenchmark: latency / throughput of the FP64 tensor-core instruction mma.sync.m8n8k4.f64 (SASS DMMA.884)
// on sm_100a, next to DFMA, SHFL and a shared-memory round trip.  Decides whether the Riccati sweeps of the QP kernel can live
// in MMA fragments (profiles/r02_dmma_probe.md).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/_bin/dmma_probe scripts/dmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b, double c0, double c1)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// one warp, chain of dependent DMMAs through the accumulator
__global__ void lat_c(double *out, long long *cyc, int n)
{
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9, c0 = 0, c1 = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) dmma(c0, c1, a, b, c0, c1);
    long long t1 = clock64();
    out[threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// chain through the A operand (the result of one product is the left operand of the next: the vector sweeps)
__global__ void lat_a(double *out, long long *cyc, int n)
{
    double a = 1e-3 * threadIdx.x, b = 0.125, c0, c1;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) { dmma(c0, c1, a, b, 0.0, 0.0); a = c0; }
    long long t1 = clock64();
    out[threadIdx.x] = a + c1;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_fma(double *out, long long *cyc, int n)
{
    double a = 1.0 + threadIdx.x * 1e-9, c = 0;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) c = fma(a, c, a);
    long long t1 = clock64();
    out[threadIdx.x] = c;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_shfl(double *out, long long *cyc, int n)
{
    double c = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) c = __shfl_xor_sync(0xffffffffu, c, 1);
    long long t1 = clock64();
    out[threadIdx.x] = c;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_smem(double *out, long long *cyc, int n)
{
    __shared__ double s[64];
    double c = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < n; i++) { s[threadIdx.x] = c; __syncwarp(); c = s[threadIdx.x ^ 1]; __syncwarp(); }
    long long t1 = clock64();
    out[threadIdx.x] = c;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput: W warps per CTA, CH independent chains per warp
template <int CH>
__global__ void thr(double *out, int n)
{
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c0[CH], c1[CH];
#pragma unroll
    for (int j = 0; j < CH; j++) { c0[j] = j; c1[j] = -j; }
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < CH; j++) dmma(c0[j], c1[j], a, b, c0[j], c1[j]);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < CH; j++) s += c0[j] + c1[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void thr_fma(double *out, int n)
{
    double a = 1.0 + threadIdx.x * 1e-9;
    double c[CH];
#pragma unroll
    for (int j = 0; j < CH; j++) c[j] = j;
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < CH; j++) c[j] = fma(a, c[j], a);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < CH; j++) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// DMMA and DFMA streams interleaved in one warp: do they share the pipe?
__global__ void thr_mix(double *out, int n)
{
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c0[4], c1[4], f[8];
#pragma unroll
    for (int j = 0; j < 4; j++) { c0[j] = j; c1[j] = -j; }
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] = j;
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) dmma(c0[j], c1[j], a, b, c0[j], c1[j]);
#pragma unroll
        for (int j = 0; j < 8; j++) f[j] = fma(a, f[j], a);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) s += c0[j] + c1[j];
#pragma unroll
    for (int j = 0; j < 8; j++) s += f[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float timeit(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, clock %d kHz\n", pr.name, sms, khz);
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * 1024 * 2048);
    cudaMalloc(&cyc, 8);
    const int n = 4096;
    lat_c<<<1, 32>>>(out, cyc, n); lat_c<<<1, 32>>>(out, cyc, n);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("DMMA.884 dependent through C : %.2f cycles\n", (double)h / n);
    lat_a<<<1, 32>>>(out, cyc, n); lat_a<<<1, 32>>>(out, cyc, n);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("DMMA.884 dependent through A : %.2f cycles\n", (double)h / n);
    lat_fma<<<1, 32>>>(out, cyc, n); lat_fma<<<1, 32>>>(out, cyc, n);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("DFMA dependent               : %.2f cycles\n", (double)h / n);
    lat_shfl<<<1, 32>>>(out, cyc, n); lat_shfl<<<1, 32>>>(out, cyc, n);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("SHFL (64-bit = 2 x 32) dep.  : %.2f cycles\n", (double)h / n);
    lat_smem<<<1, 32>>>(out, cyc, n); lat_smem<<<1, 32>>>(out, cyc, n);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("STS + syncwarp + LDS + sync  : %.2f cycles\n", (double)h / n);

    const int it = 20000;
    for (int warps = 1; warps <= 16; warps *= 2) {
        float ms1 = timeit([&] { thr<1><<<sms, 32 * warps>>>(out, it); });
        float ms4 = timeit([&] { thr<4><<<sms, 32 * warps>>>(out, it); });
        float ms8 = timeit([&] { thr<8><<<sms, 32 * warps>>>(out, it); });
        double clk = khz * 1e3;
        printf("DMMA %2d warps/SM: 1 chain %.3f  4 chains %.3f  8 chains %.3f DMMA/clk/SM  (8 chains: %.2f TFLOP/s)\n", warps,
               (double)it * warps / (ms1 * 1e-3 * clk), 4.0 * it * warps / (ms4 * 1e-3 * clk), 8.0 * it * warps / (ms8 * 1e-3 * clk),
               8.0 * it * warps * sms * 512.0 / (ms8 * 1e-3) * 1e-12);
    }
    for (int warps = 1; warps <= 16; warps *= 2) {
        float ms8 = timeit([&] { thr_fma<8><<<sms, 32 * warps>>>(out, it); });
        double clk = khz * 1e3;
        printf("DFMA %2d warps/SM: 8 chains %.3f warp-DFMA/clk/SM (%.2f TFLOP/s)\n", warps, 8.0 * it * warps / (ms8 * 1e-3 * clk),
               8.0 * it * warps * sms * 64.0 / (ms8 * 1e-3) * 1e-12);
    }
    for (int warps = 4; warps <= 16; warps *= 2) {
        float ms = timeit([&] { thr_mix<<<sms, 32 * warps>>>(out, it); });
        printf("mix  %2d warps/SM: 4 DMMA + 8 DFMA per iteration: %.2f TFLOP/s (DMMA part %.2f, DFMA part %.2f)\n", warps,
               (double)it * warps * sms * (4 * 512.0 + 8 * 64.0) / (ms * 1e-3) * 1e-12,
               (double)it * warps * sms * (4 * 512.0) / (ms * 1e-3) * 1e-12, (double)it * warps * sms * (8 * 64.0) / (ms * 1e-3) * 1e-12);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
