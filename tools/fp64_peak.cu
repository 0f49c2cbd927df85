// fp64_peak.cu -- developer microbenchmark: DFMA vs DMMA.8x8x4 throughput on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters, double a, double b)
{
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dfma_chain_kernel(double* out, int iters, double a, double b)
{
    double x = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x = fma(x, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void dmma_kernel(double* out, int iters, double a, double b)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class K> float timeit(K k)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k(); cudaDeviceSynchronize();
    cudaEventRecord(e0); k(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 4096;
    for (int warps = 4; warps <= 32; warps *= 2) {
        int threads = warps * 32, blocks = sms * 2;
        float ms = timeit([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double fl = 2.0 * 16 * iters * (double)threads * blocks;
        printf("DFMA  16 chains  %2d warps/CTA x2 CTA/SM: %.2f TFLOP/s\n", warps, fl / ms / 1e9);
        ms = timeit([&] { dfma_chain_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = 2.0 * 16 * iters * (double)threads * blocks;
        printf("DFMA  1 chain    %2d warps/CTA x2 CTA/SM: %.2f TFLOP/s\n", warps, fl / ms / 1e9);
        ms = timeit([&] { dmma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = 2.0 * 256 * 8 * iters * (double)warps * blocks;
        printf("DMMA  8 chains   %2d warps/CTA x2 CTA/SM: %.2f TFLOP/s\n", warps, fl / ms / 1e9);
    }
    return 0;
}
