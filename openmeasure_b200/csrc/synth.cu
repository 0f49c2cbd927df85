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

constexpr int SY_THREADS = 256;   // 8 warps, one row per warp at a time

// H[k][j] = 2 u(seed+1, k, j) - 1 (mode coefficient per snapshot), theta[k][f] = u(seed+2, k, f)
__global__ void synth_tables_kernel(double* __restrict__ H, double* __restrict__ theta, int m, int K, int F,
                                    uint64_t seed)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < K * m; t += gridDim.x * blockDim.x) {
        int k = t / m, j = t - k * m;
        H[t] = 2.0 * u01(seed + 1, (uint64_t)k, (uint64_t)j) - 1.0;
    }
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < K * F; t += gridDim.x * blockDim.x) {
        int k = t / F, f = t - k * F;
        theta[t] = u01(seed + 2, (uint64_t)k, (uint64_t)f);
    }
}

// dynamic smem: g[8][K] (spatial factors of the row each warp is working on)
__global__ void __launch_bounds__(SY_THREADS)
synth_kernel(double* __restrict__ X, int64_t n_cells, int64_t cell0, int64_t ncell_loc, int m, int K, int F,
             uint64_t seed, const double* __restrict__ amp, const double* __restrict__ dec, double eps,
             const double* __restrict__ H, const double* __restrict__ theta)
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* g = sm + (size_t)warp * K;
    const int f = blockIdx.y;
    const double mu = ldexp(1.0, f) * (1.0 + (double)f / 8.0);
    for (int64_t cl = (int64_t)blockIdx.x * (SY_THREADS / 32) + warp; cl < ncell_loc;
         cl += (int64_t)gridDim.x * (SY_THREADS / 32)) {
        const int64_t c = cell0 + cl;
        const double sc = ((double)c + 0.5) / (double)n_cells;
        const uint64_t irow = (uint64_t)((int64_t)f * n_cells + c);
        __syncwarp();
        for (int k = lane; k < K; k += 32) {
            double omega = (double)(k + 1) * 0.6180339887498949 + 0.5;
            double t = omega * sc;
            t = t + theta[k * F + f];
            t = t - floor(t);
            double tri = 4.0 * fabs(t - 0.5) - 1.0;
            g[k] = amp[k] * tri;
        }
        __syncwarp();
        double* row = X + ((int64_t)f * ncell_loc + cl) * m;
        for (int j = lane; j < m; j += 32) {
            double acc = 0.0;
            for (int k = 0; k < K; ++k) {
                double p = g[k] * H[k * m + j];
                acc = acc + p;
            }
            double nz = 2.0 * u01(seed + 3, irow, (uint64_t)j) - 1.0;
            double e = eps * dec[j];
            e = e * nz;
            double v = 0.25 * acc;
            v = v + e;
            v = 1.0 + v;
            row[j] = mu * v;
        }
    }
}

}  // namespace omb

extern "C" int64_t omb_synth_ws_bytes(int64_t F, int64_t m, int64_t K) { return (int64_t)sizeof(double) * K * (m + F); }

extern "C" int omb_synth_fill(double* d_X, int64_t F, int64_t n_cells, int64_t cell0, int64_t ncell_loc,
                              int64_t m, int64_t K, uint64_t seed, const double* d_amp, const double* d_dec,
                              double eps, void* d_ws, void* stream)
{
    using namespace omb;
    OMB_CHECK_ARG(d_X && d_amp && d_dec && d_ws, "null pointer");
    OMB_CHECK_ARG(F > 0 && n_cells > 0 && ncell_loc > 0 && m > 0 && K > 0 && K <= 2048, "bad size");
    double* H = (double*)d_ws;
    double* theta = H + K * m;
    cudaStream_t st = (cudaStream_t)stream;
    synth_tables_kernel<<<64, 256, 0, st>>>(H, theta, (int)m, (int)K, (int)F, seed);
    int rc = check_launch("synth_tables_kernel");
    if (rc) return rc;
    size_t smem = sizeof(double) * (size_t)K * (SY_THREADS / 32);
    int64_t gx = ceil_div(ncell_loc, SY_THREADS / 32);
    if (gx > 148 * 8) gx = 148 * 8;
    dim3 grid((unsigned)gx, (unsigned)F);
    synth_kernel<<<grid, SY_THREADS, smem, st>>>(d_X, n_cells, cell0, ncell_loc, (int)m, (int)K, (int)F, seed, d_amp,
                                                 d_dec, eps, H, theta);
    return check_launch("synth_kernel");
}
