// dmma_probe.cu -- developer microbenchmark: what keeps a 32-accumulator DMMA.8x8x4 loop below the
// FP64 tensor peak?  One CTA of 8 warps per SM (2 warps per scheduler, like the library's kernels).
//   V0  32 DMMAs per step on fixed operands
//   V1  + 12 DADDs (centring) per step
//   V2  + 12 LDS.64 fragment loads per step (conflict-free pitch 132), no DADD
//   V3  loads + DADDs (the library's consumer step without barriers)
//   V4  V3 with the loads of step k+1 issued before the DMMAs of step k
//   V5  V3 with 16 warps (4 per scheduler) of 16 x 64 outputs (16 DMMAs per step, 2 + 8 loads)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int LD = 132, ROWS = 64;

template <int V>
__global__ void __launch_bounds__(256, 1) probe(double* out, int iters, double cv0)
{
    extern __shared__ double sm[];
    for (int e = threadIdx.x; e < 2 * ROWS * LD; e += blockDim.x) sm[e] = 1.0 + 1e-9 * e;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane & 3, fc = lane >> 2;
    const int ib = (warp >> 1) * 32, jb = (warp & 1) * 64;
    const double* sA = sm;
    const double* sB = sm + ROWS * LD;
    double c[4][8][2];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 8; ++q) c[p][q][0] = c[p][q][1] = 0.0;
    double a[4], b[8], a2[4], b2[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) a[q] = a2[q] = 1.0 + q;
#pragma unroll
    for (int q = 0; q < 8; ++q) b[q] = b2[q] = 0.5 + q;
    double cv = cv0;
    if (V == 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) a2[q] = sA[fr * LD + ib + q * 8 + fc];
#pragma unroll
        for (int q = 0; q < 8; ++q) b2[q] = sB[fr * LD + jb + q * 8 + fc];
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
            const int kr = ((it & 1) * 8 + k4) * 4 + fr;
            if (V == 2 || V == 3) {
#pragma unroll
                for (int q = 0; q < 4; ++q) a[q] = sA[kr * LD + ib + q * 8 + fc];
#pragma unroll
                for (int q = 0; q < 8; ++q) b[q] = sB[kr * LD + jb + q * 8 + fc];
            }
            if (V == 4) {
#pragma unroll
                for (int q = 0; q < 4; ++q) { a[q] = a2[q]; }
#pragma unroll
                for (int q = 0; q < 8; ++q) { b[q] = b2[q]; }
                const int kn = (kr + 4) & (ROWS - 1);
#pragma unroll
                for (int q = 0; q < 4; ++q) a2[q] = sA[kn * LD + ib + q * 8 + fc];
#pragma unroll
                for (int q = 0; q < 8; ++q) b2[q] = sB[kn * LD + jb + q * 8 + fc];
            }
            if (V == 1 || V == 3 || V == 4) {
                if (V == 1) cv += 1e-12;
                else cv = sm[kr];
#pragma unroll
                for (int q = 0; q < 4; ++q) a[q] -= cv;
#pragma unroll
                for (int q = 0; q < 8; ++q) b[q] -= cv;
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 8; ++q) dmma(c[p][q][0], c[p][q][1], a[p], b[q]);
        }
    }
    double s = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 8; ++q) s += c[p][q][0] + c[p][q][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 16 warps, 16 x 64 outputs per warp
__global__ void __launch_bounds__(512, 1) probe16(double* out, int iters, double cv0)
{
    extern __shared__ double sm[];
    for (int e = threadIdx.x; e < 2 * ROWS * LD; e += blockDim.x) sm[e] = 1.0 + 1e-9 * e;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane & 3, fc = lane >> 2;
    const int ib = (warp >> 1) * 16, jb = (warp & 1) * 64;
    const double* sA = sm;
    const double* sB = sm + ROWS * LD;
    double c[2][8][2];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 8; ++q) c[p][q][0] = c[p][q][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
            const int kr = ((it & 1) * 8 + k4) * 4 + fr;
            double a[2], b[8];
#pragma unroll
            for (int q = 0; q < 2; ++q) a[q] = sA[kr * LD + ib + q * 8 + fc];
#pragma unroll
            for (int q = 0; q < 8; ++q) b[q] = sB[kr * LD + jb + q * 8 + fc];
            const double cv = sm[kr];
#pragma unroll
            for (int q = 0; q < 2; ++q) a[q] -= cv;
#pragma unroll
            for (int q = 0; q < 8; ++q) b[q] -= cv;
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int q = 0; q < 8; ++q) dmma(c[p][q][0], c[p][q][1], a[p], b[q]);
        }
    }
    double s = 0;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 8; ++q) s += c[p][q][0] + c[p][q][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K> float timeit(K k)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k(); cudaDeviceSynchronize();
    cudaEventRecord(e0); k(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

template <int V> void run(const char* name, double* out, int sms, int iters)
{
    const size_t smem = sizeof(double) * 2 * ROWS * LD;
    cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float ms = timeit([&] { probe<V><<<sms, 256, smem>>>(out, iters, 1e-3); });
    double fl = 2.0 * 256 * 32 * 8 * (double)iters * 8 * sms;
    printf("%-58s %8.3f ms  %6.2f TFLOP/s\n", name, ms, fl / ms / 1e9);
}

int main(int argc, char** argv)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
    for (int iters : {2000, 40000}) {
        printf("-- %d iterations of 8 k-steps\n", iters);
        run<0>("V0 32 DMMA / step, fixed operands", out, sms, iters);
        run<1>("V1 + 12 DADD", out, sms, iters);
        run<2>("V2 + 12 LDS.64 (no DADD)", out, sms, iters);
        run<3>("V3 LDS + DADD (consumer step)", out, sms, iters);
        run<4>("V4 V3, loads one step ahead", out, sms, iters);
        const size_t smem = sizeof(double) * 2 * ROWS * LD;
        cudaFuncSetAttribute(probe16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        float ms = timeit([&] { probe16<<<sms, 512, smem>>>(out, iters, 1e-3); });
        double fl = 2.0 * 256 * 16 * 8 * (double)iters * 16 * sms;
        printf("%-58s %8.3f ms  %6.2f TFLOP/s\n", "V5 16 warps x (16 x 64), LDS + DADD", ms, fl / ms / 1e9);
    }
    return 0;
}
