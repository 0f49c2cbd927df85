// synth.cu -- deterministic synthetic snapshot generator (DESIGN.md "Synthetic workload").
// Bit-identical to the CPU twin in oracle/csrc/oracle.c: every operation is one IEEE rounding
// (the library is compiled with -fmad=false).
#include "common.cuh"
#include "../../include/omb200.h"

namespace omb {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(uint64_t seed, uint64_t i, uint64_t j)
{
    return (double)(splitmix64(seed ^ (i * 0x9E3779B97F4A7C15ULL + j)) >> 11) * 0x1.0p-53;
}

constexpr int SY_THREADS = 128;   // one thread per row (cell) of a feature block

// dynamic smem: H[K*m] (mode coefficients per snapshot), amp[K], theta[K], dec_eps[m]
__global__ void __launch_bounds__(SY_THREADS)
synth_kernel(double* __restrict__ X, int64_t n_cells, int64_t cell0, int64_t ncell_loc, int m, int K,
             uint64_t seed, const double* __restrict__ amp, const double* __restrict__ dec, double eps)
{
    extern __shared__ double sm[];
    double* H = sm;                 // K*m
    double* s_amp = H + (size_t)K * m;
    double* s_theta = s_amp + K;
    double* s_e = s_theta + K;      // m: eps * dec[j]
    const int f = blockIdx.y;
    for (int t = threadIdx.x; t < K * m; t += SY_THREADS) {
        int k = t / m, j = t - k * m;
        H[t] = 2.0 * u01(seed + 1, (uint64_t)k, (uint64_t)j) - 1.0;
    }
    for (int k = threadIdx.x; k < K; k += SY_THREADS) {
        s_amp[k] = amp[k];
        s_theta[k] = u01(seed + 2, (uint64_t)k, (uint64_t)f);
    }
    for (int j = threadIdx.x; j < m; j += SY_THREADS) s_e[j] = eps * dec[j];
    __syncthreads();

    const double mu = ldexp(1.0, f) * (1.0 + (double)f / 8.0);
    for (int64_t cl = (int64_t)blockIdx.x * SY_THREADS + threadIdx.x; cl < ncell_loc;
         cl += (int64_t)gridDim.x * SY_THREADS) {
        const int64_t c = cell0 + cl;
        const double sc = ((double)c + 0.5) / (double)n_cells;
        const uint64_t irow = (uint64_t)((int64_t)f * n_cells + c);
        double* row = X + ((int64_t)f * ncell_loc + cl) * m;
        // the K spatial factors are recomputed per snapshot chunk to bound registers
        for (int j = 0; j < m; ++j) {
            double acc = 0.0;
            for (int k = 0; k < K; ++k) {
                double omega = (double)(k + 1) * 0.6180339887498949 + 0.5;
                double t = omega * sc;
                t = t + s_theta[k];
                t = t - floor(t);
                double tri = 4.0 * fabs(t - 0.5) - 1.0;
                double g = s_amp[k] * tri;
                double p = g * H[k * m + j];
                acc = acc + p;
            }
            double nz = 2.0 * u01(seed + 3, irow, (uint64_t)j) - 1.0;
            double e = s_e[j] * nz;
            double v = 0.25 * acc;
            v = v + e;
            v = 1.0 + v;
            row[j] = mu * v;
        }
    }
}

}  // namespace omb

extern "C" int omb_synth_fill(double* d_X, int64_t F, int64_t n_cells, int64_t cell0,
                              int64_t ncell_loc, int64_t m, int64_t K, uint64_t seed,
                              const double* d_amp, const double* d_dec, double eps, void* stream)
{
    using namespace omb;
    OMB_CHECK_ARG(d_X && d_amp && d_dec, "null pointer");
    OMB_CHECK_ARG(F > 0 && n_cells > 0 && ncell_loc > 0 && m > 0 && K > 0, "non-positive size");
    size_t smem = sizeof(double) * ((size_t)K * m + 2 * K + m);
    OMB_CHECK_ARG(smem <= 200 * 1024, "K*m too large for the generator's shared memory");
    OMB_CUDA(cudaFuncSetAttribute(synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t gx = ceil_div(ncell_loc, SY_THREADS);
    if (gx > 148 * 16) gx = 148 * 16;
    dim3 grid((unsigned)gx, (unsigned)F);
    synth_kernel<<<grid, SY_THREADS, smem, (cudaStream_t)stream>>>(d_X, n_cells, cell0, ncell_loc, (int)m,
                                                                   (int)K, seed, d_amp, d_dec, eps);
    return check_launch("synth_kernel");
}
