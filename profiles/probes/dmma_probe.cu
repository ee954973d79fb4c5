// FP64 tensor-core (DMMA m8n8k4) latency / throughput probe for sm_100a.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void k_dmma(double* out, int iters, long long* cyc)
{
    double c[CH][2];
    for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i) dmma(c[i][0], c[i][1], a, b);
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
__global__ void k_dfma(double* out, int iters, long long* cyc)
{
    double c[CH];
    for (int i = 0; i < CH; ++i) c[i] = threadIdx.x + i;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i) c[i] = fma(c[i], a, b);
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CH; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <class K>
void run(const char* name, K kern, int grid, int block, int iters, double flop_per_thread_iter)
{
    double* out; long long* cyc; cudaMalloc(&out, 8 * grid * block); cudaMalloc(&cyc, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, block>>>(out, 10, cyc); cudaDeviceSynchronize();
    cudaEventRecord(e0); kern<<<grid, block>>>(out, iters, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1); long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s grid %5d block %4d: %8.3f ms, %7.2f cycles/iter (thread 0), %8.2f TFLOP/s  [%s]\n", name, grid, block, ms, (double)c / iters,
           flop_per_thread_iter * iters * grid * block / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    // one warp: latency of a dependent chain (CH = 1) and issue rate of independent chains
    run("dmma 1 warp, 1 chain", k_dmma<1>, 1, 32, 4096, 2.0 * 256 / 32);
    run("dmma 1 warp, 4 chains", k_dmma<4>, 1, 32, 4096, 4 * 2.0 * 256 / 32);
    run("dmma 1 warp, 8 chains", k_dmma<8>, 1, 32, 4096, 8 * 2.0 * 256 / 32);
    run("dfma 1 warp, 1 chain", k_dfma<1>, 1, 32, 4096, 2.0);
    run("dfma 1 warp, 8 chains", k_dfma<8>, 1, 32, 4096, 16.0);
    // whole GPU
    run("dmma full, 8 chains", k_dmma<8>, 148 * 8, 256, 20000, 8 * 2.0 * 256 / 32);
    run("dfma full, 8 chains", k_dfma<8>, 148 * 8, 256, 20000, 16.0);
    run("dmma 4 warps/SM, 4 chains", k_dmma<4>, 148, 128, 20000, 4 * 2.0 * 256 / 32);
    run("dmma 21 warps/SM, 2 chains", k_dmma<2>, 148, 672, 20000, 2 * 2.0 * 256 / 32);
    return 0;
}
